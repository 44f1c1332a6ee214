"""Parity ON THE BENCHMARK WORKLOAD ITSELF (BASELINE.json config 2, full sizes up to 2000 bp per side):
boxes of tests/benchgen (the generator bench.py uses) through the CUDA path vs the CPU oracle, and
size-independent properties of a larger slice: repeatability (checksum of all device results and
scripts), independence of batch composition / chunking, traceback consistency with the fill."""
import pytest

import benchgen
from harness import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from gmap_2024_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def test_benchmark_boxes_bit_exact_vs_oracle(engine):
    o = Oracle()
    boxes = [benchgen.make(20241018, i, small=False) for i in range(300)]
    batch = engine.batch()
    ids = [batch.add(b) for b in boxes]
    batch.run()
    big = 0
    for b, cid in zip(boxes, ids):
        assert batch.result(cid, b["mode"]) == o.run(b), (b["mode"], b.get("rlength"), b.get("glength"))
        big += b.get("rlength", 0) > 1000
    assert big > 50
    batch.free()


def test_resident_form_of_the_workload_bit_exact_vs_oracle(engine):
    """bench.py's default form: the boxes' segments in one synthetic genome resident on the device, boxes queued as
    coordinates, MaxEnt (synthetic tables) evaluated on the device; the oracle gets the same boxes with characters and
    the same probabilities as arrays"""
    o = Oracle()
    seed, n = 20241018, 400
    batch = engine.batch()
    try:
        benchgen.resident_begin(batch, seed, n * 2600 + 1000000)
        assert benchgen.fill_batch(batch, seed, 0, n, 1, False) == n
        gbytes = benchgen.resident_attach(engine)
        assert batch.h2d_bytes() < 2 * gbytes          # queries and box records only
        batch.run()
        hot = 0
        for i in range(n):
            box = benchgen.make(seed, i, small=False)
            assert batch.result(i, box["mode"]) == o.run(box), (i, box["mode"])
            hot += box["mode"] == "genome" and max(box["left_probs"]) > 0.9
        assert hot > 10
    finally:
        benchgen.resident_end()
        batch.free()


@pytest.mark.parametrize("resident", [False, True])
def test_decorated_stratum_bit_exact_vs_oracle(engine, resident):
    """SURVEY.md section 8d's input decorations on the benchmark generator: N in the genomic segments, lower-case and
    IUPAC characters in the query, tandem repeats (8-bit stripe artefact F11 on the small boxes); both forms"""
    o = Oracle()
    seed = 99
    for small, n in ((False, 300), (True, 600)):
        batch = engine.batch()
        try:
            if resident:
                benchgen.resident_begin(batch, seed, n * 2600 + 1000000)
            assert benchgen.fill_batch(batch, seed, 0, n, 1, small, 31, decor=True) == n
            if resident:
                benchgen.resident_attach(engine)
            batch.run()
            nN = 0
            for i in range(n):
                box = benchgen.make(seed, i, small=small, decor=True)
                assert batch.result(i, box["mode"]) == o.run(box), (i, box["mode"], small)
                nN += "N" in box.get("gseg", box.get("gsegL", ""))
            assert nN > n // 20
        finally:
            benchgen.resident_end()
            batch.free()


def test_checksum_repeatable_and_chunking_independent(engine):
    import os
    from gmap_2024_b200 import Engine
    seed, n = 7, 6000
    os.environ["GMAPDP_CHUNK_MB"] = "2"     # ~24 MB of inputs -> a dozen pipelined chunks
    chunky = Engine(0)
    del os.environ["GMAPDP_CHUNK_MB"]

    def digest_of(lo, hi, eng=None):
        b = (eng or engine).batch()
        benchgen.fill_batch(b, seed, lo, hi - lo, 1, False)
        b.run_device()                      # host-buffer path (chunked for large batches)
        d1 = b.digest()
        b.upload()
        b.run_resident()                    # resident path (single launch pair)
        b.download()
        d2 = b.digest()
        cells = b.cells()
        b.free()
        return d1, d2, cells
    d1, d2, cells = digest_of(0, n)
    assert d1 == d2                          # host-buffer path == resident path
    e1, e2, _ = digest_of(0, n)
    assert (e1, e2) == (d1, d2)              # repeatable
    launches0 = chunky.launch_count()
    c1, c2, _ = digest_of(0, n, chunky)
    assert chunky.launch_count() - launches0 > 10      # really ran as many chunks
    assert (c1, c2) == (d1, d2)              # chunked == unchunked
    chunky.close()
    assert cells > 10 ** 8


def test_device_counts_match_host_replay(engine):
    """the device tracebacks' score / match / mismatch / open / indel counts are recomputed by the host
    replay of the edit script for every call; GmapDP_batch_finish fails if any differs"""
    b = engine.batch()
    benchgen.fill_batch(b, 11, 0, 3000, 1, False)
    b.run()                                  # raises EngineError on a count mismatch
    assert b.nboxes() > 2000
    b.free()


def test_benchmark_100k_boxes_bit_exact_vs_oracle(engine):
    """BASELINE config 2 says "bit-exact score and traceback check": the first 40 000 - 100 000 boxes of the benchmark's own
    generator and seed (what bench.py times), every out-parameter and every pair record, against the oracle run on all
    host cores (digests cross the process boundary, the comparison is of the full results)."""
    import os
    import parallel_oracle
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n = min(100000, 2500 * cores)            # ~90 s of oracle time whatever the box has: 40 000 boxes on 16 cores, 80 000 on 32
    seed = 20241018
    b = engine.batch()
    ids, modes = [], []
    box = benchgen.Box()
    lib = benchgen.lib()
    import ctypes as C
    for i in range(n):
        lib.benchgen_make(C.c_uint64(seed), C.c_long(i), C.byref(box), 0)
        ids.append(lib.benchgen_add(b.h, C.byref(box)))
        modes.append(benchgen.MODES[box.mode])
    b.run()
    got = parallel_oracle.gpu_digests(b, ids, modes)
    b.free()
    want = parallel_oracle.bench_digests(seed, n)
    bad = [i for i in range(n) if got[i] != want[i]]
    assert not bad, "%d of %d benchmark boxes differ from the oracle; first: box %d (%s)" % (len(bad), n, bad[0], modes[bad[0]])
