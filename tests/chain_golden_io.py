"""Reader of tests/golden/chain_golden.npz (written by tests/golden/make_chain_golden.py from the compiled reference)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SCALARS = ("querylength", "querystart", "queryend", "indexsize", "localp", "skip_repetitive_p", "favor_right_p", "middlep", "max_nalignments")


def load(fwd=False):
    """[(problem, links, scores, cells, paths)] of the lookback (default) or lookforward run of the reference"""
    f = "f_" if fwd else ""
    z = np.load(os.path.join(HERE, "golden", "chain_golden.npz"))
    out = []
    for i in range(int(z["n"])):
        pb = {k: np.ascontiguousarray(z["p%d_%s" % (i, k)]) for k in ("positions", "npositions", "minactive", "maxactive")}
        for k, v in zip(SCALARS, z["p%d_scalars" % i].tolist()):
            pb[k] = int(v)
        pb["queryseq"] = z["p%d_queryseq" % i].tobytes()
        plen = z["p%d_%spathlen" % (i, f)]
        pairs = z["p%d_%spairs" % (i, f)]
        paths, o = [], 0
        for n in plen.tolist():
            paths.append(pairs[o:o + n])
            o += n
        out.append((pb, z["p%d_%slinks" % (i, f)], z["p%d_%sscores" % (i, f)], z["p%d_%scells" % (i, f)], paths))
    return out
