"""Generates tests/golden/dp_golden.json.gz from the COMPILED, UNMODIFIED reference.

Run in the build container only (needs oracle/_ref, i.e. `make -C oracle ref`, which compiles
/root/reference/src in place):   python tests/golden/make_golden.py

Every record is a self-contained box (inputs exactly as the reference saw them: query chars, the
genomic segments it fetched with Genome_get_segment_*, its MaxEnt probabilities) plus the
reference's outputs (out-parameters and the full pair list).  Boxes whose outputs change under a
different MALLOC_PERTURB_ value would be reference UB; the generator runs itself
twice in subprocesses with MALLOC_PERTURB_ = 85 / 170 and keeps only boxes that agree.
Also records the fill-level known answers (tests/golden/fill_golden.npz).
"""
import gzip
import json
import os
import random
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import dpgen  # noqa: E402
from harness import Ref  # noqa: E402

WORLDS = [  # (seed, n, mode, rmin, rmax, edge)
    (1, 160, None, 4, 140, False), (2, 120, None, 4, 140, True), (3, 40, "single", 20, 130, False),
    (4, 40, "genome", 20, 260, False), (5, 30, "cdna", 10, 120, False), (6, 30, "end5", 4, 200, True),
    (7, 30, "end3", 4, 200, True), (8, 12, None, 300, 900, False),
]


def jsonable(res):
    n, iout, dout, pairs = res
    return {"n": n, "iout": iout, "dout": dout,
            "pairs": [[p[0], p[1], p[2].decode("latin1"), p[3].decode("latin1"), p[4].decode("latin1"),
                       p[5].decode("latin1")] + list(p[6:]) for p in pairs]}


def generate():
    ref = Ref()
    out = []
    for seed, n, mode, rmin, rmax, edge in WORLDS:
        boxes, _ = dpgen.ref_boxes(ref, seed, n, mode, rmin, rmax, edge)
        for b in boxes:
            res = ref.run(b)
            rec = {k: v for k, v in b.items() if k != "world"}
            rec["expect"] = jsonable(res)
            out.append(rec)
    return out


def fills():
    import numpy as np
    ref = Ref()
    rng = random.Random(99)
    recs = {}
    for it in range(60):
        kind = it % 3
        bits = 8 if it % 2 else 16
        rl = rng.randrange(1, 90 if bits == 8 else 200)
        gl = rng.randrange(1, 100 if bits == 8 else 220)
        g = dpgen.rand_dna(rng, gl, 0.01)
        alt = bytearray(g)
        if rng.random() < 0.5:
            alt[rng.randrange(gl)] = rng.choice(b"ACGT")
        q = dpgen.decorate_query(rng, (dpgen.mutate(rng, g, rng.choice([0, 0.05, 0.2])) + dpgen.rand_dna(rng, rl, 0))[:rl])
        mt, op, ex = rng.randrange(4), rng.randrange(-10, -5), rng.randrange(-3, 0)
        extra = rng.randrange(0, 20)
        lb, ub = (extra, gl - rl + extra) if gl >= rl else (rl - gl + extra, extra)
        late, revp = rng.randrange(2), rng.randrange(2)
        H, dN, dE, dF, dNraw = ref.fill(kind, bits, q, g, bytes(alt), mt, op, ex, lb, ub, late, revp)
        recs["p%d" % it] = np.array([kind, bits, mt, op, ex, lb, ub, late, revp], dtype=np.int32)
        recs["q%d" % it] = np.frombuffer(q, dtype=np.uint8)
        recs["g%d" % it] = np.frombuffer(g, dtype=np.uint8)
        recs["a%d" % it] = np.frombuffer(bytes(alt), dtype=np.uint8)
        recs["H%d" % it] = H
        recs["N%d" % it] = dNraw
        recs["E%d" % it] = dE
        recs["F%d" % it] = dF
    np.savez_compressed(os.path.join(HERE, "fill_golden.npz"), **recs)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        json.dump(generate(), sys.stdout)
        sys.exit(0)
    runs = []
    for perturb in ("85", "170"):
        env = dict(os.environ, MALLOC_PERTURB_=perturb)
        runs.append(json.loads(subprocess.check_output([sys.executable, __file__, "--child"], env=env)))
    keep = [a for a, b in zip(*runs) if a == b]
    print("boxes: %d generated, %d stable under MALLOC_PERTURB_ (dropped %d)" % (len(runs[0]), len(keep), len(runs[0]) - len(keep)))
    with gzip.open(os.path.join(HERE, "dp_golden.json.gz"), "wt") as f:
        json.dump(keep, f)
    fills()
    print("wrote dp_golden.json.gz and fill_golden.npz")
