"""The batching runtime (include/gmapdp_stream.h) under many submitting threads: every call -- prepared in a private
one-call batch, run in a flight shared with whatever the other threads submitted at that moment, replayed by its own
thread -- must give exactly what the oracle gives, whatever the grouping."""
import threading

import pytest

import dpgen
from harness import Oracle

pytestmark = pytest.mark.gpu


def worker(stream, boxes, out, err):
    try:
        b = stream.private_batch()
        for k, box in boxes:
            b.clear()
            cid = b.add(box)
            stream.call(b)
            out[k] = b.result(cid, box["mode"])
        b.free()
    except Exception as e:  # noqa: BLE001
        err.append(repr(e))


@pytest.mark.parametrize("nthreads,max_boxes", [(1, 0), (24, 0), (48, 16)])
def test_stream_matches_oracle(nthreads, max_boxes):
    from gmap_2024_b200 import Stream
    oracle = Oracle()
    boxes = dpgen.synth_boxes(seed=61 + nthreads, n=1200, rmin=8, rmax=260)
    boxes += dpgen.synth_boxes(seed=62, n=8, mode="single", rmin=900, rmax=1900)
    want = [oracle.run(b) for b in boxes]
    stream = Stream(devices=(0,), max_boxes=max_boxes)       # max_boxes=16: flights fill up, submitters wait for the next one
    try:
        out, err = [None] * len(boxes), []
        parts = [[(k, boxes[k]) for k in range(t, len(boxes), nthreads)] for t in range(nthreads)]
        th = [threading.Thread(target=worker, args=(stream, p, out, err)) for p in parts]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert not err, err[:3]
        st = stream.stats()
    finally:
        stream.close()
    bad = [k for k in range(len(boxes)) if out[k] != want[k]]
    assert not bad, "%d of %d calls differ from the oracle, first: %s" % (len(bad), len(boxes), boxes[bad[0]]["mode"])
    assert st["boxes"] > 0 and st["flights"] > 0 and st["kernel_launches"] >= st["flights"]
    if max_boxes:
        assert st["largest_flight"] <= max_boxes
    if nthreads > 1:
        assert st["flights"] < st["boxes"], "no two boxes ever shared a flight: %r" % st


def test_stream_user_penalties_and_two_lanes():
    """--indel-open / --indel-extend through the private batches; two lanes on the same device behave like two GPUs"""
    from gmap_2024_b200 import Stream
    oracle = Oracle()
    oracle.set_user_dynprog(-12, -4)
    try:
        boxes = dpgen.synth_boxes(seed=71, n=600, rmin=8, rmax=200)
        want = [oracle.run(b) for b in boxes]
    finally:
        oracle.set_user_dynprog(0, 0, False)
    stream = Stream(devices=(0, 0))
    try:
        out, err = [None] * len(boxes), []

        def work(t):
            try:
                b = stream.private_batch()
                b.set_user_dynprog(-12, -4)
                for k in range(t, len(boxes), 8):
                    b.clear()
                    cid = b.add(boxes[k])
                    stream.call(b)
                    out[k] = b.result(cid, boxes[k]["mode"])
                b.free()
            except Exception as e:  # noqa: BLE001
                err.append(repr(e))
        th = [threading.Thread(target=work, args=(t,)) for t in range(8)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert not err, err[:3]
    finally:
        stream.close()
    assert out == want
