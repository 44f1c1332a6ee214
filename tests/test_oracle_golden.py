"""The CPU oracle against the committed golden vectors (generated from the compiled reference by
tests/golden/make_golden.py).  Runs everywhere, no GPU, no /root/reference."""
import os

import numpy as np

from golden_io import load_golden
from harness import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))


def test_entry_points_match_reference_golden():
    o = Oracle()
    recs = load_golden()
    assert len(recs) > 400
    modes = set()
    for box, want in recs:
        got = o.run(box)
        modes.add(box["mode"])
        assert got == want, "oracle differs from the reference on a %s box" % box["mode"]
    assert modes == {"single", "genome", "cdna", "end5", "end3"}


def test_fills_match_reference_golden():
    o = Oracle()
    z = np.load(os.path.join(HERE, "golden", "fill_golden.npz"))
    n = len([k for k in z.files if k.startswith("p")])
    assert n == 60
    cells = 0
    for it in range(n):
        kind, bits, mt, op, ex, lb, ub, late, revp = [int(v) for v in z["p%d" % it]]
        q, g, a = z["q%d" % it].tobytes(), z["g%d" % it].tobytes(), z["a%d" % it].tobytes()
        H, dN, dE, dF, dNraw = o.fill(kind, bits, q, g, a, mt, op, ex, lb, ub, late, revp)
        rl, gl = len(q), len(g)
        R, C = np.meshgrid(np.arange(rl + 1), np.arange(gl + 1), indexing="ij")
        if kind == 0:
            mask = (R >= C - ub) & (R <= C + lb)
        elif kind == 1:
            mask = (C >= R) & (C <= R + ub)
        else:
            mask = (R >= C) & (R <= C + lb)
        cells += int(mask.sum())
        assert o.cells(kind, rl, gl, lb, ub) == int(mask.sum())
        assert np.array_equal(H[mask], z["H%d" % it][mask])
        assert np.array_equal((dNraw != 0)[mask], (z["N%d" % it] != 0)[mask])
        assert np.array_equal(dE[mask], z["E%d" % it][mask])
        if kind == 0:
            assert np.array_equal(dF[mask], z["F%d" % it][mask])
            # VERT vs HORIZ in directions_nogap must agree too
            assert np.array_equal(dNraw[mask], z["N%d" % it][mask])
    assert cells > 100000


def test_tables_known_answers():
    """spot values of pairdistance_array / use8p_size (dynprog.c:1022-1142, dynprog.h:43-49)"""
    o = Oracle()
    assert [o.use8p_size(m) for m in range(4)] == [41, 63, 127, 24]
    A, C_, N, R, star = ord("A"), ord("C"), ord("N"), ord("R"), ord("*")
    assert o.pairdistance(0, A, A) == 3 and o.pairdistance(0, A, C_) == -3
    assert o.pairdistance(1, A, C_) == -2 and o.pairdistance(2, A, C_) == -1 and o.pairdistance(3, A, C_) == -5
    assert o.pairdistance(0, N, N) == 3 and o.pairdistance(0, N, A) == 3
    assert o.pairdistance(0, R, A) == 1 and o.pairdistance(0, R, C_) == -3
    assert o.pairdistance(0, ord("a"), A) == 3
    assert o.pairdistance(0, A, star) == 0 and o.pairdistance(0, A, 4) == 0
    assert o.pairdistance(0, A, ord("z")) == 0 and o.pairdistance(0, A, ord("Z")) == -3
    assert o.consistent(R, A) == 1 and o.consistent(R, C_) == 0
