"""Host-only speed of the pair-list replay (the part of `e2e` that is not the device): a benchmark batch is queued, every
device box gets a synthetic but realistic edit script (diagonal runs of ~60 with 1-3 nt gaps, the whole box traced), and
GmapDP_batch_complete is timed.  No GPU.   python scripts/replay_speed.py [--boxes N] [--threads T] [--lib path]"""
import argparse
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np


class NoDevice:
    def __init__(self, lib):
        self.lib, self.ctx = lib, C.c_void_p()


def script_for(rng, r, c, runlen=60.0):
    ops = []
    while r > 0 and c > 0:
        n = min(int(rng.geometric(1.0 / runlen)), r, c)
        ops.append(n << 2); r -= n; c -= n
        if r > 0 and c > 0:
            g = int(rng.integers(1, 4))
            if rng.random() < 0.5:
                g = min(g, c); ops.append((g << 2) | 1); c -= g
            else:
                g = min(g, r); ops.append((g << 2) | 2); r -= g
    return ops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--boxes", type=int, default=40000)
    ap.add_argument("--threads", default="1,8")
    ap.add_argument("--lib", default=None)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--decorated", action="store_true", help="the decorated stratum of the generator (N, IUPAC, lower case: not plain A C G T)")
    ap.add_argument("--runlen", type=float, default=60.0, help="mean length of a diagonal run between gaps")
    a = ap.parse_args()
    import benchgen
    from gmap_2024_b200.engine import Batch, Box, DeviceResult, load_library
    lib = load_library(a.lib)
    b = Batch(NoDevice(lib), 2000, 2030)
    benchgen.fill_batch(b, 20241018, 0, a.boxes, decor=a.decorated)
    ptr, n, _, _, _, _ = b.device_view()
    rng = np.random.default_rng(1)
    results = (DeviceResult * n)()
    script = []
    for k in range(n):
        x = Box.from_address(ptr.value + C.sizeof(Box) * k)
        r = results[k]
        r.status = 0
        r.script_off = len(script)
        if x.mode == 0:
            s = script_for(rng, x.rlenL, x.glenL, a.runlen); r.script_lenA = len(s); script += s
        elif x.mode in (3, 4):
            r.bestrL, r.bestcL = x.rlenL, min(x.glenL, x.rlenL + 3)
            s = script_for(rng, r.bestrL, r.bestcL, a.runlen); r.script_lenA = len(s); script += s
        else:
            hl = x.rlenL // 2 if x.mode == 1 else x.rlenL
            hr = x.rlenL - hl if x.mode == 1 else x.rlenR
            r.bestrL, r.bestcL, r.bestrR, r.bestcR = hl, min(x.glenL - 2, hl + 2), hr, min(x.glenR - 2, hr + 2)
            if x.mode == 2:
                r.bestcL = max(1, min(r.bestcL, x.glenL // 2 - 8)); r.bestcR = max(1, min(r.bestcR, x.glenL // 2 - 8))
            sa, sb = script_for(rng, r.bestrR, r.bestcR, a.runlen), script_for(rng, r.bestrL, r.bestcL, a.runlen)
            r.script_lenA, r.script_lenB = len(sa), len(sb); script += sa + sb
    sc = np.asarray(script, dtype=np.uint32)
    scp = sc.ctypes.data_as(C.POINTER(C.c_uint32))
    npairs = int(sum(int(v) >> 2 for v in script))
    print("%d calls, %d device boxes, %d script ops, ~%.1f M pairs" % (b.ncalls(), n, len(script), npairs / 1e6))
    for t in [int(v) for v in a.threads.split(",")]:
        os.environ["GMAPDP_REPLAY_THREADS"] = str(t)
        best = 1e9
        for _ in range(a.reps + 1):
            b.rewind()
            t0 = time.perf_counter()
            lib.GmapDP_batch_complete(b.h, results, scp)
            best = min(best, time.perf_counter() - t0)
        print("threads %2d: %.1f ms, %.0f M pairs/s (%.0f M pairs/s per thread)" % (t, best * 1e3, npairs / best / 1e6, npairs / best / 1e6 / t))
    b.free()


if __name__ == "__main__":
    main()
