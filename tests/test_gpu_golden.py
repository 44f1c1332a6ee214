"""GPU parity against the committed golden vectors of the compiled reference."""
import pytest

from golden_io import load_golden

pytestmark = pytest.mark.gpu


def test_cuda_path_matches_reference_golden():
    from gmap_2024_b200 import Engine
    e = Engine(0)
    recs = load_golden()
    batch = e.batch()
    ids = [batch.add(b) for b, _ in recs]
    batch.run()
    assert batch.nboxes() > 200
    for (b, want), cid in zip(recs, ids):
        assert batch.result(cid, b["mode"]) == want, b["mode"]
    batch.free()
    e.close()
