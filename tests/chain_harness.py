"""ctypes front ends for the chaining checkers: the compiled reference (oracle/_ref/ref_stage2.so, when present)
and our scalar restatement (oracle/liboracle.so).  Test infrastructure."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "ref_stage2.so")
ORC_SO = os.path.join(ROOT, "oracle", "liboracle.so")

SETUP = dict(splicingp=1, cross_species_p=0, sufflookback=60, nsufflookback=5, maxintronlen=500000)   # gmap.c:269,270,347

u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def canon_labels(a):
    """fwd_tracei is only compared for equality: relabel by first appearance, keep -1 and 0"""
    out = np.empty_like(a)
    seen = {}
    for i, v in enumerate(a.tolist()):
        if v <= 0:
            out[i] = v
        else:
            out[i] = seen.setdefault(v, len(seen) + 1)
    return out


class _Chain:
    prefix = None

    def __init__(self, path):
        self.lib = C.CDLL(path)
        f = getattr(self.lib, self.prefix + "_setup")
        f.argtypes = [C.c_int] * 5
        f(SETUP["splicingp"], SETUP["cross_species_p"], SETUP["sufflookback"], SETUP["nsufflookback"], SETUP["maxintronlen"])
        self._s, self._p = {}, {}
        for fwd, suffix in ((False, ""), (True, "_fwd")):
            self._bind(fwd, suffix)

    def _bind(self, fwd, suffix):
        s = getattr(self.lib, self.prefix + "_scores" + suffix)
        s.argtypes = [u32p, i32p, C.c_int, C.c_int, u32p, u32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                      i32p, i32p, i32p, C.c_int]
        s.restype = C.c_int
        p = getattr(self.lib, self.prefix + "_paths" + suffix)
        p.argtypes = [u32p, i32p, C.c_int, C.c_int, u32p, u32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                      C.c_char_p, C.c_char_p, i32p, C.c_int, i32p, C.c_int]
        p.restype = C.c_int
        self._s[fwd], self._p[fwd] = s, p

    def scores(self, pb, fwd=False):
        tot = len(pb["positions"])
        pos = pb["positions"] if tot else np.zeros(1, dtype=np.uint32)
        links = np.zeros(5 * max(tot, 1), dtype=np.int32)
        sc = np.zeros(max(tot, 1), dtype=np.int32)
        cells = np.zeros(5 * max(tot, 1), dtype=np.int32)
        n = self._s[fwd](pos, pb["npositions"], pb["querylength"], tot, pb["minactive"], pb["maxactive"], pb["querystart"],
                    pb["queryend"], pb["indexsize"], pb["localp"], pb["skip_repetitive_p"], pb["favor_right_p"], pb["middlep"],
                    links, sc, cells, max(tot, 1))
        links = links[:5 * tot].reshape(-1, 5).copy()
        if tot:
            links[:, 4] = canon_labels(links[:, 4])
        return links, sc[:tot].copy(), cells[:5 * n].reshape(-1, 5).copy()

    def paths(self, pb, fwd=False):
        tot = len(pb["positions"])
        pos = pb["positions"] if tot else np.zeros(1, dtype=np.uint32)
        cap = 4 * pb["querylength"] + 16
        plen = np.zeros(tot + 1, dtype=np.int32)
        while True:
            pairs = np.zeros(2 * cap, dtype=np.int32)
            n = self._p[fwd](pos, pb["npositions"], pb["querylength"], tot, pb["minactive"], pb["maxactive"], pb["querystart"],
                        pb["queryend"], pb["indexsize"], pb["localp"], pb["skip_repetitive_p"], pb["favor_right_p"], pb["middlep"],
                        pb["max_nalignments"], pb["queryseq"], pb["queryseq"], plen, tot + 1, pairs, cap)
            if n >= 0:
                break
            cap = -n + 16       # ties at the best score are all traced (stage2.c:4476): can be many
        out, o = [], 0
        for k in range(n):
            out.append(pairs[2 * o:2 * (o + plen[k])].reshape(-1, 2).copy())
            o += int(plen[k])
        return out


class RefChain(_Chain):
    prefix = "refs2"

    def __init__(self):
        super().__init__(REF_SO)


class OracleChain(_Chain):
    prefix = "orcs2"

    def __init__(self):
        super().__init__(ORC_SO)


def have_ref():
    return os.path.exists(REF_SO)
