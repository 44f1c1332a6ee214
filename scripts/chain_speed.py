"""Quick throughput check of the chaining engine against the compiled reference / the oracle on the host."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import chaingen, chain_harness as ch
from gmap_2024_b200 import Engine

nd = int(sys.argv[1]) if len(sys.argv) > 1 else 128
tile = int(sys.argv[2]) if len(sys.argv) > 2 else 16
t = time.time()
rng = np.random.default_rng(99)
base = [chaingen.make_problem(rng, glen=int(rng.integers(50000, 200000)), nexons=int(rng.integers(2, 12)), exon_len=(80, 400),
                              err=0.01, k=8, window="full" if i % 2 else "2000", max_nalignments=10) for i in range(nd)]
print("generated %d problems in %.1f s; mean L %.0f, mean hits %.0f" % (nd, time.time() - t, np.mean([p["querylength"] for p in base]),
      np.mean([len(p["positions"]) for p in base])), flush=True)
cpu = ch.RefChain() if ch.have_ref() else ch.OracleChain()
t = time.time()
exp = [cpu.paths(p) for p in base]
cpu_s = time.time() - t
print("host (%s, 1 thread): %.3f s for %d problems = %.1f problems/s" % (type(cpu).__name__, cpu_s, nd, nd / cpu_s), flush=True)
e = Engine(0); e.chain_setup(**ch.SETUP)
b = e.chain_batch()
ids = [b.add(p) for _ in range(tile) for p in base]
b.upload()
for _ in range(2): b.run_resident()
ms = [b.run_resident() for _ in range(3)]
b.download()
n = nd * tile
print("device: %.3f ms per batch of %d problems = %.0f problems/s, %.1f M hits/s" % (np.mean(ms), n, n / (np.mean(ms) / 1e3), b.nhits() / (np.mean(ms) / 1e3) / 1e6))
t = time.time(); b.run(); print("e2e run(): %.1f ms" % ((time.time() - t) * 1e3))
ok = all(len(b.paths(ids[i])) == len(exp[i % nd]) and all(np.array_equal(a[1], x) for a, x in zip(b.paths(ids[i]), exp[i % nd])) for i in range(0, n, max(1, n // 200)))
print("spot check vs host:", ok)
