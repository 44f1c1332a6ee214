"""Pins the CPU oracle against the compiled reference itself, live (only where oracle/_ref exists,
i.e. in the build container; the committed golden vectors cover the other machines)."""
import random

import numpy as np
import pytest

import dpgen
from harness import Oracle, Ref, ref_available

pytestmark = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module")
def pair():
    return Oracle(), Ref()


def test_tables_identical(pair):
    o, r = pair
    for mt in range(4):
        assert o.use8p_size(mt) == r.use8p_size(mt)
        for a in range(128):
            for b in range(128):
                assert o.pairdistance(mt, a, b) == r.pairdistance(mt, a, b)
    for a in range(128):
        for b in range(128):
            assert o.consistent(a, b) == r.consistent(a, b)


def test_fills_cell_for_cell(pair):
    o, r = pair
    rng = random.Random(5)
    cells = 0
    for it in range(150):
        kind, bits = rng.randrange(3), rng.choice([8, 16])
        rl = rng.randrange(1, 100 if bits == 8 else 260)
        gl = rng.randrange(1, 110 if bits == 8 else 280)
        g = dpgen.rand_dna(rng, gl, 0.005)
        alt = bytearray(g)
        for _ in range(gl // 100 + (1 if rng.random() < 0.3 else 0)):
            alt[rng.randrange(gl)] = rng.choice(b"ACGT")
        alt = bytes(alt)
        q = dpgen.decorate_query(rng, dpgen.mutate(rng, g, rng.choice([0, 0.03, 0.1, 0.3])))
        q = (q + dpgen.rand_dna(rng, rl, 0))[:rl]
        mt, op, ex = rng.randrange(4), rng.randrange(-10, -5), rng.randrange(-3, 0)
        extra, wide = rng.randrange(0, 20), rng.randrange(2)
        if wide:
            lb, ub = (extra, gl - rl + extra) if gl >= rl else (rl - gl + extra, extra)
        else:
            lb = ub = extra
        late, revp = rng.randrange(2), rng.randrange(2)
        A = o.fill(kind, bits, q, g, alt, mt, op, ex, lb, ub, late, revp)
        B = r.fill(kind, bits, q, g, alt, mt, op, ex, lb, ub, late, revp)
        R, C = np.meshgrid(np.arange(rl + 1), np.arange(gl + 1), indexing="ij")
        mask = ((R >= C - ub) & (R <= C + lb)) if kind == 0 else (((C >= R) & (C <= R + ub)) if kind == 1 else ((R >= C) & (R <= C + lb)))
        cells += int(mask.sum())
        for k in range(4 if kind == 0 else 3):
            assert np.array_equal(A[k][mask], B[k][mask]), (it, kind, bits, k)
    assert cells > 200000


@pytest.mark.parametrize("seed,edge", [(11, False), (12, True), (13, False)])
def test_entry_points_identical(pair, seed, edge):
    o, r = pair
    boxes, _ = dpgen.ref_boxes(r, seed, 250, edge=edge)
    seen = set()
    for b in boxes:
        assert o.run(b) == r.run(b), b["mode"]
        seen.add(b["mode"])
    assert len(seen) == 5


def test_entry_points_identical_low_complexity(pair):
    """repeats (homopolymers, short tandem units): long runs of equal scores, every tie rule of the entry points"""
    o, r = pair
    dpgen.LOW_COMPLEXITY = True
    try:
        boxes, _ = dpgen.ref_boxes(r, 31, 300, rmin=8, rmax=260)
    finally:
        dpgen.LOW_COMPLEXITY = False
    seen = set()
    for b in boxes:
        assert o.run(b) == r.run(b), b["mode"]
        seen.add(b["mode"])
    assert len(seen) == 5


@pytest.mark.parametrize("user_open,user_extend", [(-12, -4), (-5, -1), (-30, -10), (-60, -3), (-3, 0), (0, -2), (0, 0)])
def test_entry_points_user_dynprog(pair, user_open, user_extend):
    """--indel-open / --indel-extend (user_dynprog_p: dynprog_single.c:470, dynprog_genome.c:3367, dynprog_end.c:1334,1964):
    gmap.c:5425-5445 accepts anything in [-127, 0].  Not tested: penalties so large that whole stripes saturate at
    -128 (e.g. -127/-127): with jump_late_p the reference's single-gap traceback then follows ties out of the band
    into direction cells no fill wrote (23 of 2400 such calls came out differently; the fills themselves are identical
    cell for cell), i.e. the reference's own output is not a function of its inputs there."""
    o, r = pair
    o.set_user_dynprog(user_open, user_extend)
    r.set_user_dynprog(user_open, user_extend)
    try:
        boxes, _ = dpgen.ref_boxes(r, 1000 - user_open * 3 - user_extend, 120)
        for b in boxes:
            assert o.run(b) == r.run(b), (b["mode"], user_open, user_extend)
    finally:
        o.set_user_dynprog(0, 0, False)
        r.set_user_dynprog(0, 0, False)


def test_noflush_variant_gives_the_same_results(pair):
    """the fairness build (per-column _mm_clflush of dynprog_simd.c defined away) is the same program"""
    import subprocess, sys, os, json
    from harness import ref_available
    if not ref_available(noflush=True):
        pytest.skip("oracle/_ref/ref_driver_noflush.so not built")
    # the two builds export the same symbols: run the variant in its own process
    code = ("import sys, json; sys.path.insert(0, %r); import dpgen; from harness import Ref; r = Ref(noflush=True); "
            "boxes, _ = dpgen.ref_boxes(r, 17, 150); print(json.dumps([repr(r.run(b)) for b in boxes]))") % os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, check=True).stdout
    got = json.loads(out)
    o, r = pair
    boxes, _ = dpgen.ref_boxes(r, 17, 150)
    assert got == [repr(r.run(b)) for b in boxes]
