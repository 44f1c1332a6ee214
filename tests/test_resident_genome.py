"""Row A13 of SURVEY.md section 8: the reference's compressed genome read in place and its MaxEnt splice-site model,
on the host (CPU tests: the shared arithmetic of csrc/gmapdp_genome.h against the compiled reference) and on the device
(GPU tests: gmapdp_maxent_eval bit-exact against Maxent_hr_*_prob; boxes sent as genome COORDINATES give the results
of the same boxes sent as characters and probability arrays, and of the oracle)."""
import ctypes as C
import random

import pytest

import dpgen
from harness import Ref, ref_available

pytestmark = pytest.mark.skipif(not ref_available(), reason="oracle/_ref (the compiled reference) is not present")


def _tables_struct(ref):
    from gmap_2024_b200.engine import MaxentTables
    mt = MaxentTables()
    for k, a in enumerate(ref.maxent_tables()):
        mt.t[k] = a
    return mt


def _world(ref, seed, n=120, mode=None, edge=False):
    boxes, world = dpgen.ref_boxes(ref, seed, n, mode=mode, rmin=15, rmax=200, edge=edge)
    return boxes, world


def test_host_decode_and_maxent_match_the_reference():
    from gmap_2024_b200.engine import load_library
    lib = load_library()
    lib.gmapdp_maxent_host_prob.restype = C.c_double
    ref = Ref()
    mt = _tables_struct(ref)
    rng = random.Random(5)
    for seed in (3, 4):
        _, world = _world(ref, seed, edge=(seed == 4))
        blocks, nwords = ref.genome_blocks(world["handle"])
        genome = world["genome"]
        # every character of the genome (N included: flags word)
        for pos in range(len(genome)):
            assert lib.gmapdp_genome_host_char(C.c_void_p(blocks), C.c_size_t(nwords), C.c_uint32(pos)) == genome[pos]
        # the four probabilities on every position of the chromosome, incl. the left margin where they are 0.0
        co, ch = world["chroffset"], world["chrhigh"]
        for pos in list(range(co, min(ch, co + 400))) + [rng.randrange(co, ch) for _ in range(600)]:
            for kind in range(4):
                want = ref.maxent(world["handle"], kind, pos, co)
                got = lib.gmapdp_maxent_host_prob(C.c_void_p(blocks), C.c_size_t(nwords), C.byref(mt), kind, C.c_uint32(pos), C.c_uint32(co))
                assert got == want, (seed, kind, pos, got, want)      # bit-exact doubles


@pytest.fixture(scope="module")
def engine():
    from gmap_2024_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.gpu
def test_device_maxent_is_bit_exact(engine):
    ref = Ref()
    _, world = _world(ref, 9)
    blocks, nwords = ref.genome_blocks(world["handle"])
    engine.genome_attach(blocks, nwords, ref.maxent_tables())
    co, ch = world["chroffset"], world["chrhigh"]
    rng = random.Random(1)
    pos = list(range(co, min(ch, co + 64))) + [rng.randrange(co, ch) for _ in range(4000)]
    kinds = [rng.randrange(4) for _ in pos]
    got = engine.maxent_eval(kinds, pos, co)
    want = [ref.maxent(world["handle"], k, p, co) for k, p in zip(kinds, pos)]
    assert got == want
    assert max(want) > 0.5 and min(want) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("seed,edge", [(21, False), (22, True), (23, False)])
def test_coordinate_boxes_equal_character_boxes_and_the_oracle(engine, seed, edge):
    """all five modes, both strands, chromosome edges ('*' padding): the device decodes the segments and evaluates MaxEnt
    itself; results must equal those of the uploaded characters / arrays and the oracle's"""
    from harness import Oracle
    ref, orc = Ref(), Oracle()
    boxes, world = _world(ref, seed, n=400, edge=edge)
    blocks, nwords = ref.genome_blocks(world["handle"])
    engine.genome_attach(blocks, nwords, ref.maxent_tables())
    plain = engine.batch()
    ids = [plain.add(b) for b in boxes]
    plain.run()
    coord = engine.batch()
    cboxes = [dict(b, coords=dpgen.coords_of(b)) for b in boxes]
    cids = [coord.add(b) for b in cboxes]
    assert coord.h2d_bytes() < 0.6 * plain.h2d_bytes()
    coord.run()
    nbad = 0
    for b, i, j in zip(boxes, ids, cids):
        a, c = plain.result(i, b["mode"]), coord.result(j, b["mode"])
        assert a == c, (b["mode"], b["world"]["watsonp"], a[:3], c[:3])
        if c != orc.run(b):
            nbad += 1
    assert nbad == 0
    assert {b["mode"] for b in boxes} == {"single", "genome", "cdna", "end5", "end3"}
    plain.free()
    coord.free()
