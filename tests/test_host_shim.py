"""CPU-side checks of the product library: it loads, exports every symbol the headers declare,
resolves the reference's no-fill exits on the host exactly like the oracle, and FAILS LOUDLY when a
device box would need a GPU that is not there (no CPU fallback)."""
import collections
import ctypes as C
import os
import re

import pytest

import dpgen
from harness import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from gmap_2024_b200.build import build_native
    from gmap_2024_b200.engine import load_library
    build_native()
    return load_library()


def declared_symbols():
    names = []
    for h in ("gmapdp_b200.h", "gmapdp_shim.h", "gmapdp_stream.h", "gmapchain_b200.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"\b((?:gmapdp|GmapDP|gmapchain|GmapChain)_\w+)\s*\(", text)
    return sorted(set(names))


def test_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) >= 55 and "gmapchain_run_batch" in names and "GmapChain_lookback" in names and "gmapdp_stream_submit" in names
    for n in names:
        assert hasattr(lib, n), "missing export " + n


def test_struct_sizes(lib):
    from gmap_2024_b200.engine import DeviceResult, Pair
    assert C.sizeof(DeviceResult) == 64
    assert C.sizeof(Pair) == 48


class _NoDevice:
    def __init__(self, lib):
        self.lib = lib
        self.ctx = C.c_void_p()


def test_host_resolved_calls_match_oracle_and_device_calls_fail_loudly(lib):
    from gmap_2024_b200.engine import Batch, EngineError
    o = Oracle()
    boxes = dpgen.synth_boxes(seed=9, n=1500, rmin=2, rmax=120)
    b = Batch(_NoDevice(lib), 2000, 2030)
    ids = [b.add(x) for x in boxes]
    assert b.ncalls() == 1500 and 0 < b.nboxes() < 1500 and b.cells() > 0
    st = collections.Counter()
    for x, cid in zip(boxes, ids):
        if lib.GmapDP_result(b.h, cid, None, None, None, 0) == -3:      # waits for the device
            st["device"] += 1
            continue
        st[x["mode"]] += 1
        assert b.result(cid, x["mode"]) == o.run(x)
    assert st["device"] == b.nboxes()
    assert st["single"] and st["genome"] and st["end5"] and st["end3"]
    with pytest.raises(EngineError, match="no CPU fallback"):
        b.run()
    b.free()


def test_engine_creation_fails_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gmap_2024_b200 import Engine, EngineError
    with pytest.raises(EngineError):
        Engine(0)


def test_maxlengths(lib):
    mr, mg = C.c_int(), C.c_int()
    lib.GmapDP_maxlengths(C.byref(mr), C.byref(mg), 600, 20, 60, 10, 8)
    assert (mr.value, mg.value) == (660, 2000)          # SURVEY.md section 5 / dynprog.c:602-627
    lib.GmapDP_maxlengths(C.byref(mr), C.byref(mg), 1940, 20, 60, 10, 8)
    assert (mr.value, mg.value) == (2000, 2030)


def test_stream_creation_fails_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gmap_2024_b200 import Stream, EngineError
    with pytest.raises(EngineError, match="no CPU fallback|CUDA"):
        Stream(devices=(0,))


def test_limits_and_user_penalties_are_validated(lib):
    """a side longer than 32767 would overflow the kernels' 16-bit row/column fields; user penalties are in [-127, 0]"""
    lib.GmapDP_batch_new.restype = C.c_void_p
    assert lib.GmapDP_batch_new(None, 40000, 2000) is None
    assert lib.GmapDP_batch_new(None, 2000, 32768) is None
    h = C.c_void_p(lib.GmapDP_batch_new(None, 2000, 2030))
    assert h
    assert lib.GmapDP_batch_user_dynprog(h, -12, -4, 1) == 0
    assert lib.GmapDP_batch_user_dynprog(h, 3, -4, 1) != 0
    assert lib.GmapDP_batch_user_dynprog(h, -200, -4, 1) != 0
    lib.GmapDP_batch_free(h)


def test_user_penalties_reach_the_box(lib):
    """host-resolved calls do not depend on the penalties; device boxes carry them (cdna gaps keep CDNA_OPEN/EXTEND)"""
    from gmap_2024_b200.engine import Batch
    b = Batch(_NoDevice(lib), 2000, 2030)
    b.set_user_dynprog(-12, -4)
    boxes = dpgen.synth_boxes(seed=10, n=400, rmin=20, rmax=120)
    for x in boxes:
        b.add(x)
    ptr, n, _, _, _, _ = b.device_view()

    from gmap_2024_b200.engine import Box
    assert C.sizeof(Box) == 96
    seen = set()
    for k in range(n):
        x = Box.from_address(ptr.value + C.sizeof(Box) * k)
        seen.add(x.mode)
        if x.mode == 2:
            assert (x.open, x.extend) == (-10, -7)
        else:
            assert (x.open, x.extend) == (-12, -4)
    assert seen == {0, 1, 2, 3, 4}
    b.free()


def test_cell_counts_of_shim_and_oracle_agree(lib):
    """the product's cell counter (GCUPS numerator of bench.py) against the oracle's independent count"""
    from gmap_2024_b200.engine import Batch
    o = Oracle()
    boxes = dpgen.synth_boxes(seed=12, n=1200, rmin=2, rmax=300)
    b = Batch(_NoDevice(lib), 2000, 2030)
    total = 0
    for x in boxes:
        before = b.cells()
        b.add(x)
        mine = b.cells() - before
        assert mine == o.count_cells(x), x["mode"]
        total += mine
    assert total > 1000000
    b.free()


def _random_script(rng, r, c):
    """a random but well-formed edit script from cell (r, c): ops (len << 2 | kind), kind 0 diagonal run, 1 genome skip,
    2 query skip, in traceback order"""
    ops = []
    while r > 0 and c > 0 and len(ops) < 64:
        u = rng.random()
        if u < 0.7:
            n = rng.randint(1, min(r, c)); ops.append(n << 2); r -= n; c -= n
        elif u < 0.85:
            n = rng.randint(1, min(c, 4)); ops.append((n << 2) | 1); c -= n
        else:
            n = rng.randint(1, min(r, 4)); ops.append((n << 2) | 2); r -= n
    return ops


def test_replay_vector_path_equals_scalar_path(lib, tmp_path):
    """The replay writes eight 8-byte pair records per iteration with SSE2 (positions relative to the call's sides, reversed
    sides written backwards).  The same shim compiled without that path (-DGMAPDP_NO_NT_STORES: one record at a time) must
    hand out identical lists for the same device results -- here random well-formed edit scripts for every device box
    of a mixed batch, completed through GmapDP_batch_complete (no GPU involved)."""
    import random
    import subprocess
    from gmap_2024_b200.build import CSRC
    from gmap_2024_b200.engine import Batch, Box, DeviceResult, load_library
    scalar = str(tmp_path / "libshim_scalar.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-msse2", "-DGMAPDP_NO_NT_STORES", "-o", scalar,
                           os.path.join(CSRC, "gmapdp_shim.cpp"), "-L" + CSRC, "-lgmapdp_b200", "-Wl,-rpath," + CSRC, "-lpthread"])
    boxes = [x for x in dpgen.synth_boxes(seed=21, n=900, rmin=20, rmax=400) if x["mode"] != "cdna"]
    outs = []
    for L in (lib, load_library(scalar)):
        b = Batch(_NoDevice(L), 2000, 2030)
        ids = [b.add(x) for x in boxes]
        ptr, n, _, _, _, _ = b.device_view()
        assert n > 300
        rng = random.Random(5)
        results = (DeviceResult * n)()
        script = []
        for k in range(n):
            x = Box.from_address(ptr.value + C.sizeof(Box) * k)
            r = results[k]
            r.status = 0
            r.script_off = len(script)
            if x.mode == 0:
                a = _random_script(rng, x.rlenL, x.glenL)
                r.script_lenA, r.script_lenB = len(a), 0
                script += a
            elif x.mode in (3, 4):
                r.bestrL, r.bestcL = rng.randint(1, x.rlenL), rng.randint(1, x.glenL)
                a = _random_script(rng, r.bestrL, r.bestcL)
                r.script_lenA, r.script_lenB = len(a), 0
                script += a
            else:
                r.bestrL, r.bestcL = rng.randint(1, x.rlenL - 1), rng.randint(1, x.glenL - 2)
                r.bestrR, r.bestcR = x.rlenL - r.bestrL, rng.randint(1, x.glenR - 2)
                a, bb = _random_script(rng, r.bestrR, r.bestcR), _random_script(rng, r.bestrL, r.bestcL)
                r.script_lenA, r.script_lenB = len(a), len(bb)
                script += a + bb
        sc = (C.c_uint32 * max(len(script), 1))(*script)
        L.GmapDP_batch_complete(b.h, results, sc)          # the counts of a random script differ from the device's: not an error here
        outs.append([b.result(cid, x["mode"]) for cid, x in zip(ids, boxes)])
        b.free()
    assert len(outs[0]) == len(outs[1])
    npairs = 0
    for got, want in zip(*outs):
        assert got == want
        npairs += max(got[0], 0)
    assert npairs > 20000


def test_replay_reads_identical_streams_once(lib):
    """Where the query is its own upper case the replay reads it once (and likewise the alt segment where there is no alt
    genome).  The same calls with the query in lower case -- two distinct streams -- must give the same lists, the cdna
    character in lower case being the only difference."""
    import random
    from gmap_2024_b200.engine import Batch, Box, DeviceResult
    base = [x for x in dpgen.synth_boxes(seed=33, n=600, rmin=20, rmax=300) if x["mode"] != "cdna"]
    upper = [dict(x, queryseq=x["queryseq"].upper()) for x in base]
    lower = [dict(x, queryseq=x["queryseq"].lower()) for x in base]
    outs = []
    for boxes in (upper, lower):
        b = Batch(_NoDevice(lib), 2000, 2030)
        ids = [b.add(x) for x in boxes]
        ptr, n, _, _, _, _ = b.device_view()
        rng = random.Random(7)
        results = (DeviceResult * n)()
        script = []
        for k in range(n):
            x = Box.from_address(ptr.value + C.sizeof(Box) * k)
            r = results[k]
            r.status = 0
            r.script_off = len(script)
            if x.mode == 0:
                a = _random_script(rng, x.rlenL, x.glenL)
                r.script_lenA, r.script_lenB = len(a), 0
                script += a
            elif x.mode in (3, 4):
                r.bestrL, r.bestcL = rng.randint(1, x.rlenL), rng.randint(1, x.glenL)
                a = _random_script(rng, r.bestrL, r.bestcL)
                r.script_lenA, r.script_lenB = len(a), 0
                script += a
            else:
                r.bestrL, r.bestcL = rng.randint(1, x.rlenL - 1), rng.randint(1, x.glenL - 2)
                r.bestrR, r.bestcR = x.rlenL - r.bestrL, rng.randint(1, x.glenR - 2)
                a, bb = _random_script(rng, r.bestrR, r.bestcR), _random_script(rng, r.bestrL, r.bestcL)
                r.script_lenA, r.script_lenB = len(a), len(bb)
                script += a + bb
        sc = (C.c_uint32 * max(len(script), 1))(*script)
        lib.GmapDP_batch_complete(b.h, results, sc)
        outs.append((n, [b.result(cid, x["mode"]) for cid, x in zip(ids, boxes)]))
        b.free()
    assert outs[0][0] == outs[1][0]                      # the same calls reach the device either way
    npairs = 0
    for (nu, iu, du, pu), (nl, il, dl, pl) in zip(outs[0][1], outs[1][1]):
        assert (nu, iu, du) == (nl, il, dl)
        for a, bb in zip(pu, pl):
            assert a[:2] == bb[:2] and a[3:] == bb[3:] and a[2].upper() == bb[2].upper()
        npairs += max(nu, 0)
    assert npairs > 10000
