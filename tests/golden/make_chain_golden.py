"""Golden vectors for the stage-2 chaining (row A15): runs the COMPILED REFERENCE (oracle/_ref/ref_stage2.so =
the reference's stage2.c compiled in place) on seeded synthetic problems and stores inputs and outputs in
tests/golden/chain_golden.npz.  Build-container only (needs /root/reference to have built oracle/_ref)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import chaingen  # noqa: E402
import chain_harness as ch  # noqa: E402

KEYS_IN = ("positions", "npositions", "minactive", "maxactive")
SCALARS = ("querylength", "querystart", "queryend", "indexsize", "localp", "skip_repetitive_p", "favor_right_p", "middlep", "max_nalignments")


def main():
    ref = ch.RefChain()
    problems = chaingen.make_set(20241018, 48, small=True) + chaingen.make_set(77, 8, small=False)
    out = {"n": np.array(len(problems))}
    for i, pb in enumerate(problems):
        links, scores, cells = ref.scores(pb)
        paths = ref.paths(pb)
        flinks, fscores, fcells = ref.scores(pb, fwd=True)        # align_compute_scores_lookforward
        fpaths = ref.paths(pb, fwd=True)
        for k in KEYS_IN:
            out["p%d_%s" % (i, k)] = pb[k]
        out["p%d_scalars" % i] = np.array([pb[k] for k in SCALARS], dtype=np.int64)
        out["p%d_queryseq" % i] = np.frombuffer(pb["queryseq"], dtype=np.uint8)
        out["p%d_links" % i] = links
        out["p%d_scores" % i] = scores
        out["p%d_cells" % i] = cells
        out["p%d_pathlen" % i] = np.array([len(p) for p in paths], dtype=np.int32)
        out["p%d_pairs" % i] = np.concatenate(paths) if paths else np.zeros((0, 2), dtype=np.int32)
        out["p%d_f_links" % i] = flinks
        out["p%d_f_scores" % i] = fscores
        out["p%d_f_cells" % i] = fcells
        out["p%d_f_pathlen" % i] = np.array([len(p) for p in fpaths], dtype=np.int32)
        out["p%d_f_pairs" % i] = np.concatenate(fpaths) if fpaths else np.zeros((0, 2), dtype=np.int32)
    path = os.path.join(HERE, "chain_golden.npz")
    np.savez_compressed(path, **out)
    print(len(problems), "problems ->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
