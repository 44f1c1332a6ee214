"""N > 1 host logic on CPU: two gloo ranks shard the benchmark box stream; their cell counts must add
up to the single-process count and their blocks must tile the index range without overlap."""
import ctypes as C
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _count_cells(i0, i1):
    import benchgen
    from gmap_2024_b200.engine import Batch, load_library

    class _NoDev:
        pass
    nd = _NoDev()
    nd.lib, nd.ctx = load_library(), C.c_void_p()
    b = Batch(nd, 2000, 2030)
    benchgen.fill_batch(b, 7, i0, i1 - i0, 1, True)
    out = (b.ncalls(), b.nboxes(), b.cells())
    b.free()
    return out


def _worker(rank, world, port, per_rank, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from gmap_2024_b200.sharding import reduce_scalars, shard_range
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    i0, i1 = shard_range(rank, world, per_rank)
    calls, boxes, cells = _count_cells(i0, i1)
    tot = reduce_scalars([calls, boxes, cells], "sum")
    mx = reduce_scalars([i1], "max")
    q.put((rank, i0, i1, calls, boxes, cells, tot, mx))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    import torch.multiprocessing as mp
    from gmap_2024_b200.build import build_native
    import benchgen
    build_native()
    benchgen.build()
    world, per_rank, port = 2, 400, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, per_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, a0, b0, c0, x0, s0, tot0, mx0), (r1, a1, b1, c1, x1, s1, tot1, mx1) = res
    assert (a0, b0, a1, b1) == (0, 400, 400, 800)                   # contiguous, disjoint, complete
    assert tot0 == tot1 == [c0 + c1, x0 + x1, s0 + s1] and mx0 == [800.0]
    whole = _count_cells(0, 800)
    assert list(whole) == [int(v) for v in tot0]                      # shards add up to the whole


def test_split_evenly():
    from gmap_2024_b200.sharding import shard_range, split_evenly
    assert split_evenly(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert split_evenly(3, 8)[-1] == (3, 3)
    assert shard_range(3, 8, 1000000) == (3000000, 4000000)
    with pytest.raises(ValueError):
        shard_range(8, 8, 10)
