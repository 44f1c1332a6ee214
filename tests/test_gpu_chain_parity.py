"""GPU parity of the stage-2 chaining engine (row A15) through the C ABI (gmapchain_b200.h): paths, their cells,
the full link matrix and score matrix against the oracle on seeded problems, and against the committed outputs of
the compiled reference (tests/golden/chain_golden.npz)."""
import numpy as np
import pytest

import chain_golden_io
import chain_harness as ch
import chaingen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from gmap_2024_b200 import Engine
    e = Engine(0)
    e.chain_setup(**ch.SETUP)
    yield e
    e.close()


def check_against(engine, problems, expected, fwd=False):
    """expected[i] = (links, scores, cells, paths) with the oracle's / reference's conventions"""
    b = engine.chain_batch()
    ids = [b.add(pb, forward=fwd) for pb in problems]
    b.run()
    links, scores = b.links()
    off = 0
    for cid, pb, (elinks, escores, ecells, epaths) in zip(ids, problems, expected):
        tot = len(pb["positions"])
        assert np.array_equal(scores[off:off + tot], escores), "fwd_scores differ in problem %d" % cid
        assert np.array_equal(links[off:off + tot, :4], elinks[:, :4]), "links differ in problem %d" % cid
        off += tot
        got = b.paths(cid)
        assert len(got) == len(epaths), "number of paths differs in problem %d: %d vs %d" % (cid, len(got), len(epaths))
        for k, ((cell, pairs), ep) in enumerate(zip(got, epaths)):
            assert np.array_equal(pairs, ep), "path %d of problem %d differs" % (k, cid)
            assert cell == ecells[k].tolist(), "cell %d of problem %d differs" % (k, cid)
    b.free()


@pytest.mark.parametrize("fwd", [False, True])
def test_chain_matches_reference_golden(engine, fwd):
    gold = chain_golden_io.load(fwd)
    check_against(engine, [g[0] for g in gold], [(g[1], g[2], g[3], g[4]) for g in gold], fwd)


@pytest.mark.parametrize("fwd", [False, True])
@pytest.mark.parametrize("seed,n,small", [(11, 120, True), (12, 40, False)])
def test_chain_matches_oracle(engine, seed, n, small, fwd):
    orc = ch.OracleChain()
    problems = chaingen.make_set(seed, n, small=small)
    exp = []
    for pb in problems:
        links, scores, cells = orc.scores(pb, fwd)
        exp.append((links, scores, cells, orc.paths(pb, fwd)))
    check_against(engine, problems, exp, fwd)


def test_chain_mixed_directions_in_one_batch(engine):
    orc = ch.OracleChain()
    problems = chaingen.make_set(21, 30, small=True)
    b = engine.chain_batch()
    ids = [b.add(pb, forward=bool(i & 1)) for i, pb in enumerate(problems)]
    b.run()
    for i, (cid, pb) in enumerate(zip(ids, problems)):
        want = orc.paths(pb, bool(i & 1))
        got = b.paths(cid)
        assert len(got) == len(want) and all(np.array_equal(g[1], w) for g, w in zip(got, want))
    b.free()


def test_chain_edge_cases_and_batch_independence(engine):
    orc = ch.OracleChain()
    rng = np.random.default_rng(5)
    pb = chaingen.make_problem(rng, glen=4000, nexons=2, exon_len=(50, 80), err=0.01)
    empty = dict(pb, positions=np.zeros(0, dtype=np.uint32), npositions=np.zeros(pb["querylength"], dtype=np.int32))
    shut = dict(pb, minactive=np.full(pb["querylength"], 4000000, dtype=np.uint32),
                maxactive=np.full(pb["querylength"], 4000001, dtype=np.uint32))
    one = dict(pb, querystart=pb["querylength"] - 1, queryend=pb["querylength"] - 1)
    inverted = dict(pb, querystart=40, queryend=20)
    problems = [pb, empty, shut, one, inverted, pb]
    exp = []
    for p in problems:
        links, scores, cells = orc.scores(p)
        exp.append((links, scores, cells, orc.paths(p)))
    check_against(engine, problems, exp)
    # the same call alone and inside a batch gives the same answer
    b = engine.chain_batch()
    cid = b.add(pb)
    b.run()
    alone = b.paths(cid)
    b.free()
    assert len(alone) == len(exp[0][3]) and all(np.array_equal(a[1], e) for a, e in zip(alone, exp[0][3]))
    # use_canonical_p is refused loudly
    from gmap_2024_b200 import EngineError
    b = engine.chain_batch()
    with pytest.raises(EngineError):
        b.add(pb, use_canonical_p=1)
    b.free()
