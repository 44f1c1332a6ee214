"""Stage-2 chaining: the scalar restatement oracle/chain_oracle.c against
(a) the committed reference outputs (tests/golden/chain_golden.npz) and (b), when oracle/_ref is built,
the compiled reference itself on fresh seeded problems: links (consecutive, rootposition, predecessor,
trace label up to relabelling), scores, ranked cells and the traced paths must all be identical."""
import numpy as np
import pytest

import chain_golden_io
import chain_harness as ch
import chaingen


def same(a, b):
    la, sa, ca = a
    lb, sb, cb = b
    return np.array_equal(la, lb) and np.array_equal(sa, sb) and np.array_equal(ca, cb)


@pytest.mark.parametrize("fwd", [False, True])
def test_chain_oracle_matches_reference_golden(fwd):
    orc = ch.OracleChain()
    gold = chain_golden_io.load(fwd)
    assert len(gold) >= 50
    npaths = 0
    for pb, links, scores, cells, paths in gold:
        assert same(orc.scores(pb, fwd), (links, scores, cells))
        got = orc.paths(pb, fwd)
        assert len(got) == len(paths)
        for x, y in zip(got, paths):
            assert np.array_equal(x, y)
        npaths += len(paths)
    assert npaths > 40


@pytest.mark.skipif(not ch.have_ref(), reason="oracle/_ref/ref_stage2.so not built (needs /root/reference)")
@pytest.mark.parametrize("fwd", [False, True])
def test_chain_oracle_matches_compiled_reference(fwd):
    ref, orc = ch.RefChain(), ch.OracleChain()
    for seed, n, small in ((5, 40, True), (6, 12, False)):
        for pb in chaingen.make_set(seed, n, small=small):
            assert same(ref.scores(pb, fwd), orc.scores(pb, fwd))
            a, b = ref.paths(pb, fwd), orc.paths(pb, fwd)
            assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))


def test_chain_oracle_edge_cases():
    orc = ch.OracleChain()
    rng = np.random.default_rng(3)
    pb = chaingen.make_problem(rng, glen=4000, nexons=2, exon_len=(50, 80), err=0.0)
    # no hits at all
    empty = dict(pb, positions=np.zeros(0, dtype=np.uint32), npositions=np.zeros(pb["querylength"], dtype=np.int32))
    links, scores, cells = orc.scores(empty)
    assert len(cells) == 0 and orc.paths(empty) == []
    # window that excludes every hit
    shut = dict(pb, minactive=np.full(pb["querylength"], 4000000, dtype=np.uint32), maxactive=np.full(pb["querylength"], 4000001, dtype=np.uint32))
    links, scores, cells = orc.scores(shut)
    # the first querypos with hits is initialised before the window is applied (stage2.c:3793-3812)
    assert (scores > 0).sum() == pb["npositions"][np.nonzero(pb["npositions"] > 0)[0][0]]
    # error-free transcript of two exons: the best path covers (almost) every k-mer
    paths = orc.paths(pb)
    assert len(paths) >= 1 and len(paths[0]) >= pb["querylength"] - 2 * pb["indexsize"] - 8
    q = paths[0][:, 0]
    assert (np.diff(q) > 0).all()
    # lookforward: the same chain found from the other end, list order = highest querypos first (stage2.c:5062)
    fpaths = orc.paths(pb, fwd=True)
    assert len(fpaths) >= 1 and (np.diff(fpaths[0][:, 0]) < 0).all()
    assert len(set(map(tuple, fpaths[0].tolist())) & set(map(tuple, paths[0].tolist()))) >= len(paths[0]) - 16
