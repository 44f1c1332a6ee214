"""The CPU oracle on all host cores (test infrastructure): worker processes regenerate the boxes from their seeds, run
the oracle and return a digest per call, so that only 16 bytes per box cross the process boundary.  The digest covers
every out-parameter and the raw pair records -- the same bytes the GPU side hashes (digest_of_result)."""
import ctypes as C
import hashlib
import os

import harness


def digest_of_result(n, iout, dout, pair_buf):
    """n = return value (pairs or -1), iout / dout sequences, pair_buf = ctypes array of 48-byte Pair records"""
    h = hashlib.blake2b(digest_size=16)
    h.update(repr((int(n), [int(x) for x in iout], [float(x) for x in dout])).encode())
    if n > 0:
        h.update(C.string_at(C.addressof(pair_buf), int(n) * 48))
    return h.digest()


def _oracle_digest(o, box):
    n, iout, dout, _ = _run_raw(o, box)
    return digest_of_result(n, iout, dout, o.buf)


def _run_raw(o, box):
    # Oracle.run builds Python tuples of every pair; here only the raw buffer is wanted
    saved = harness.pairs_to_list
    harness.pairs_to_list = lambda buf, n: []
    try:
        return o.run(box)
    finally:
        harness.pairs_to_list = saved


def _bench_worker(args):
    seed, lo, hi, small = args
    import benchgen
    o = harness.Oracle()
    return [_oracle_digest(o, benchgen.make(seed, i, small)) for i in range(lo, hi)]


def _dpgen_worker(args):
    kind, seed, n, mode, rmin, rmax, grange, lo, hi = args
    import dpgen
    dpgen.CDNA_GRANGE = grange
    boxes = dpgen.lowcomplexity_boxes(seed, n, mode, rmin, rmax) if kind == "lowcomplexity" else dpgen.synth_boxes(seed, n, mode, rmin, rmax)
    o = harness.Oracle()
    return [_oracle_digest(o, b) for b in boxes[lo:hi]]


def _pool():
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return mp.get_context("spawn").Pool(cores), cores


def bench_digests(seed, n, small=False, block=250):
    """oracle digests of boxes [0, n) of tests/benchgen"""
    pool, _ = _pool()
    with pool:
        parts = pool.map(_bench_worker, [(seed, lo, min(n, lo + block), small) for lo in range(0, n, block)])
    return [d for p in parts for d in p]


def dpgen_digests(kind, seed, n, mode, rmin, rmax, grange=None, block=8):
    """oracle digests of dpgen.synth_boxes / lowcomplexity_boxes(seed, n, ...) (every worker regenerates the list: the
    generator is sequential in its random stream)"""
    pool, _ = _pool()
    with pool:
        parts = pool.map(_dpgen_worker, [(kind, seed, n, mode, rmin, rmax, grange, lo, min(n, lo + block)) for lo in range(0, n, block)])
    return [d for p in parts for d in p]


def gpu_digests(batch, ids, modes):
    """the same digests from a completed shim batch"""
    lib = batch.lib
    out = []
    for cid, mode in zip(ids, modes):
        ni = 10 if mode == "genome" else (3 if mode == "cdna" else 6)
        iout = (C.c_int * ni)()
        dout = (C.c_double * 2)()
        n = lib.GmapDP_result(batch.h, cid, iout, dout, batch.buf, batch.MAXPAIRS)
        assert -1 <= n <= batch.MAXPAIRS, n
        out.append(digest_of_result(n, list(iout), list(dout) if mode == "genome" else [], batch.buf))
    return out
