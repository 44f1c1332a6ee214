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
