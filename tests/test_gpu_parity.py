"""GPU parity: the CUDA path (through the C ABI / shim) against the CPU oracle, bit for bit.

Every out-parameter and the complete pair list of each call must be equal.  Runs only on a box
with a B200 (`-m gpu`); the oracle is the checker, never the thing under test.
"""
import collections

import pytest

import dpgen
from harness import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from gmap_2024_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def run_and_compare(engine, oracle, boxes, label):
    batch = engine.batch()
    ids = [batch.add(b) for b in boxes]
    batch.run()
    stats = collections.Counter()
    bad = []
    for b, cid in zip(boxes, ids):
        got = batch.result(cid, b["mode"])
        want = oracle.run(b)
        stats[b["mode"]] += 1
        if got != want:
            bad.append((b, got, want))
    batch.free()
    if bad:
        b, got, want = bad[0]
        desc = {k: v for k, v in b.items() if k not in ("queryseq", "left_probs", "right_probs", "world") and "gseg" not in k}
        first = next((i for i, (x, y) in enumerate(zip(got[3], want[3])) if x != y), None)
        pytest.fail("%s: %d / %d boxes differ; first: %r\n got  n=%s iout=%s dout=%s\n want n=%s iout=%s dout=%s\n first differing pair %s: %s vs %s"
                    % (label, len(bad), len(boxes), desc, got[0], got[1], got[2], want[0], want[1], want[2], first,
                       got[3][first] if first is not None and first < len(got[3]) else None,
                       want[3][first] if first is not None and first < len(want[3]) else None))
    return stats


@pytest.mark.parametrize("mode", ["end3", "end5", "single", "genome", "cdna"])
def test_small_boxes_per_mode(engine, oracle, mode):
    boxes = dpgen.synth_boxes(seed=100 + len(mode), n=600, mode=mode, rmin=4, rmax=150)
    run_and_compare(engine, oracle, boxes, mode)


def test_mixed_batch(engine, oracle):
    boxes = dpgen.synth_boxes(seed=7, n=3000, rmin=10, rmax=200)
    stats = run_and_compare(engine, oracle, boxes, "mixed")
    assert len(stats) == 5


def test_medium_boxes(engine, oracle):
    """16-bit paths, multi-stripe fills, wide bands"""
    boxes = dpgen.synth_boxes(seed=11, n=250, rmin=150, rmax=700)
    run_and_compare(engine, oracle, boxes, "medium")


def test_large_single_and_end(engine, oracle):
    boxes = []
    for mode in ("single", "end3", "end5", "genome"):
        boxes += dpgen.synth_boxes(seed=13 + len(boxes), n=6, mode=mode, rmin=1200, rmax=1990)
    run_and_compare(engine, oracle, boxes, "large")


def test_batch_composition_independence(engine, oracle):
    """a box's result must not depend on what else is in the batch (threading)"""
    boxes = dpgen.synth_boxes(seed=21, n=300)
    b1 = engine.batch()
    ids = [b1.add(b) for b in boxes]
    b1.run()
    full = [b1.result(i, b["mode"]) for i, b in zip(ids, boxes)]
    b1.free()
    for k in (0, 17, 150, 299):
        b2 = engine.batch()
        cid = b2.add(boxes[k])
        b2.run()
        assert b2.result(cid, boxes[k]["mode"]) == full[k]
        b2.free()


def test_empty_and_degenerate(engine, oracle):
    e = engine.batch()
    e.run()                      # empty batch is legal
    assert e.ncalls() == 0
    e.free()
    boxes = dpgen.synth_boxes(seed=5, n=40, rmin=2, rmax=6)
    run_and_compare(engine, oracle, boxes, "tiny")


@pytest.mark.parametrize("mode", ["genome", "cdna", "end3", "end5", "single"])
def test_low_complexity_ties(engine, oracle, mode):
    """repeats and quantised probabilities: long runs of equal scores exercise every tie rule -- the bridge's
    "score, then probability sum, then scan order" (tie lists and their overflow path), the cDNA bridge's
    positional tie-break across columns, the endpoint search and the tracebacks"""
    boxes = dpgen.lowcomplexity_boxes(seed=900 + len(mode), n=400, mode=mode, rmin=8, rmax=260)
    run_and_compare(engine, oracle, boxes, "low-complexity " + mode)


def test_low_complexity_large_genome(engine, oracle):
    boxes = dpgen.lowcomplexity_boxes(seed=77, n=40, mode="genome", rmin=600, rmax=1600)
    run_and_compare(engine, oracle, boxes, "low-complexity large genome")


def test_tie_list_overflow_path(oracle):
    """the genome bridge keeps at most GEN_TIECAP tied candidates per lane inside the fills and repeats the fills
    when a winning lane's list overflowed; a build with one-entry lists takes that path on ordinary boxes"""
    import os
    from gmap_2024_b200 import Engine
    from gmap_2024_b200.build import LIB_TIECAP1
    assert os.path.exists(LIB_TIECAP1), "run __graft_entry__.build()"
    eng = Engine(0, lib_path=LIB_TIECAP1)
    try:
        boxes = dpgen.synth_boxes(seed=41, n=500, mode="genome", rmin=8, rmax=400)
        boxes += dpgen.lowcomplexity_boxes(seed=42, n=300, mode="genome", rmin=8, rmax=300)
        boxes += dpgen.synth_boxes(seed=43, n=12, mode="genome", rmin=900, rmax=1900)
        run_and_compare(eng, oracle, boxes, "one-entry tie lists")
    finally:
        eng.close()


@pytest.mark.parametrize("user_open,user_extend", [(-12, -4), (-5, -1), (-30, -10), (0, -2), (-3, 0)])
def test_user_dynprog(engine, oracle, user_open, user_extend):
    """--indel-open / --indel-extend (user_dynprog_p: dynprog_single.c:470, dynprog_genome.c:3367, dynprog_end.c:1334,1964)"""
    boxes = dpgen.synth_boxes(seed=300 - user_open, n=500, rmin=8, rmax=300)
    oracle.set_user_dynprog(user_open, user_extend)
    try:
        want = [oracle.run(b) for b in boxes]
    finally:
        oracle.set_user_dynprog(0, 0, False)
    batch = engine.batch()
    batch.set_user_dynprog(user_open, user_extend)
    ids = [batch.add(b) for b in boxes]
    batch.run()
    got = [batch.result(cid, b["mode"]) for b, cid in zip(boxes, ids)]
    batch.free()
    bad = [k for k in range(len(boxes)) if got[k] != want[k]]
    assert not bad, "%d of %d boxes differ with user penalties %d/%d; first mode %s" % (len(bad), len(boxes), user_open, user_extend, boxes[bad[0]]["mode"])


def test_resident_run_refused_after_chunked_run(engine):
    """gmapdp_run_resident must not silently recompute only the first chunk of a pipelined plan"""
    import os
    from gmap_2024_b200 import Engine, EngineError
    os.environ["GMAPDP_CHUNK_MB"] = "1"
    try:
        eng = Engine(0)
    finally:
        del os.environ["GMAPDP_CHUNK_MB"]
    try:
        boxes = dpgen.synth_boxes(seed=88, n=3000, rmin=100, rmax=400)
        b = eng.batch()
        for x in boxes:
            b.add(x)
        b.run_device()                       # several chunks at 1 MB
        with pytest.raises(EngineError):
            b.run_resident()
        b.upload()                           # a fresh single-chunk plan makes it legal again
        b.run_resident()
        b.finish()
        b.free()
    finally:
        eng.close()


def test_long_cdna_gaps(engine):
    """cDNA gaps up to the production limit rL = rR = g + 8 <= 660 (dynprog_cdna.c:869-892, stage3.c:9276): the device
    bridge is a re-derivation (prefix-best tables, O(g band)) of the reference's O(g^2 band^2) scan, so it is checked
    where the two could drift apart -- long segments, many columns, ties between columns far apart.  The oracle runs
    the quartic scan; all host cores share the work."""
    import parallel_oracle
    sets = [("synth", 501, 160, (200, 652)), ("lowcomplexity", 502, 60, (200, 652)), ("synth", 503, 24, (640, 652))]
    for kind, seed, n, grange in sets:
        dpgen.CDNA_GRANGE = grange
        try:
            boxes = dpgen.lowcomplexity_boxes(seed, n, "cdna", 15, 150) if kind == "lowcomplexity" else dpgen.synth_boxes(seed, n, "cdna", 15, 150)
        finally:
            dpgen.CDNA_GRANGE = None
        assert all(grange[0] <= b["glength"] <= grange[1] for b in boxes)
        want = parallel_oracle.dpgen_digests(kind, seed, n, "cdna", 15, 150, grange)
        batch = engine.batch()
        ids = [batch.add(b) for b in boxes]
        batch.run()
        got = parallel_oracle.gpu_digests(batch, ids, [b["mode"] for b in boxes])
        batch.free()
        bad = [k for k in range(n) if got[k] != want[k]]
        assert not bad, "%s seed %d: %d of %d long cdna boxes differ, first glength %d" % (kind, seed, len(bad), n, boxes[bad[0]]["glength"])


def test_genome_gaps_with_real_maxent(engine, oracle):
    """splice-site probabilities from the reference's own MaxEnt model (oracle/_ref, where it travelled) instead of the
    synthetic arrays: real probability landscapes decide the bridge's tie-breaks and the 0.85 / 0.90 thresholds"""
    from harness import Ref, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref not present on this machine")
    ref = Ref()
    boxes = []
    for seed in (41, 42, 43):
        bx, _ = dpgen.ref_boxes(ref, seed, 250, mode="genome", rmin=20, rmax=400)
        boxes += bx
    run_and_compare(engine, oracle, boxes, "genome gaps, real MaxEnt")
    assert any(max(b["left_probs"]) > 0.9 for b in boxes)
