"""Python face of tests/benchgen/benchgen.cpp (bench / test infrastructure): the BASELINE.json config-2
box generator.  `fill_batch` queues boxes straight into a shim batch at C++ speed; `make` returns one
box as a dpgen-style dict so the CPU checkers (oracle port, compiled reference) can run the same box."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libbenchgen.so")
MAXSEQ = 2304
MODES = ["single", "genome", "cdna", "end5", "end3"]


class Box(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("mode", "dynprogindex", "jump_late_p", "extraband", "widebandp", "endalign",
                                       "require_pos_score_p", "cdna_direction", "finalp", "halfp", "maxpeelback",
                                       "rlength", "rlengthR", "glength", "glengthR", "roffset", "rev_roffsetR",
                                       "goffset", "rev_goffsetR", "querylength")] + \
               [("defect_rate", C.c_double), ("query", C.c_char * (2 * MAXSEQ)), ("queryuc", C.c_char * (2 * MAXSEQ)),
                ("gsegL", C.c_char * MAXSEQ), ("gsegR", C.c_char * MAXSEQ),
                ("left_probs", C.c_double * MAXSEQ), ("right_probs", C.c_double * MAXSEQ)]


def build(force=False):
    src = os.path.join(HERE, "benchgen.cpp")
    lib = os.path.join(ROOT, "gmap_2024_b200", "csrc", "libgmapdp_b200.so")
    if force or not os.path.exists(SO) or os.path.getmtime(src) > os.path.getmtime(SO) or os.path.getmtime(lib) > os.path.getmtime(SO):
        tmp = "%s.%d.tmp" % (SO, os.getpid())           # atomic: several processes may find the library stale at once
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", tmp, src,
                               "-L" + os.path.dirname(lib), "-lgmapdp_b200",
                               "-Wl,-rpath,$ORIGIN/../../gmap_2024_b200/csrc"])
        os.replace(tmp, SO)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.benchgen_fill_batch.restype = C.c_long
        assert _lib.benchgen_box_size() == C.sizeof(Box)
    return _lib


def resident_begin(batch, seed, capacity_nt):
    """Starts the resident-genome form of the workload on `batch`: fill_batch then queues coordinates into one synthetic
    genome; resident_attach uploads it afterwards."""
    L = lib()
    if L.benchgen_resident_begin(C.c_uint64(seed), C.c_uint64(capacity_nt)) != 0:
        raise MemoryError("benchgen_resident_begin")
    L.benchgen_resident_blocks.restype = C.c_void_p
    L.benchgen_resident_tables.restype = C.c_void_p
    used, cap = C.c_uint64(), C.c_uint64()
    blocks = L.benchgen_resident_blocks(C.byref(used), C.byref(cap))
    batch.lib.GmapDP_batch_genome(batch.h, C.c_void_p(blocks), C.c_size_t(cap.value), C.c_void_p(L.benchgen_resident_tables()))


def resident_attach(engine):
    """uploads the genome laid out by fill_batch (and the synthetic MaxEnt tables) to the engine's device"""
    from gmap_2024_b200.engine import MaxentTables
    L = lib()
    used, cap = C.c_uint64(), C.c_uint64()
    blocks = L.benchgen_resident_blocks(C.byref(used), C.byref(cap))
    mt = MaxentTables.from_address(L.benchgen_resident_tables())
    engine.genome_attach(blocks, used.value, [mt.t[k] for k in range(16)])
    return used.value * 4


def resident_end():
    """back to the plain form (make() returns the u^3 probability arrays again)"""
    lib().benchgen_resident_end()


def resident_tables_only(seed):
    """CPU workers: benchgen.make fills left_probs / right_probs with the resident form's MaxEnt values"""
    lib().benchgen_resident_tables_only(C.c_uint64(seed))


def fill_batch(batch, seed, i0, n, stride=1, small=False, modemask=31, decor=False):
    """decor: the decorated stratum (N in the genome, lower-case / IUPAC query characters, tandem repeats)"""
    return lib().benchgen_fill_batch(batch.h, C.c_uint64(seed), C.c_long(i0), C.c_long(n), C.c_long(stride), int(bool(small)) | (2 if decor else 0),
                                     int(modemask))


def make(seed, i, small=False, max_r=2000, max_g=2030, decor=False):
    b = Box()
    lib().benchgen_make(C.c_uint64(seed), C.c_long(i), C.byref(b), int(bool(small)) | (2 if decor else 0))
    m = MODES[b.mode]
    q = b.query[:b.querylength] if isinstance(b.query, bytes) else bytes(b.query)[:b.querylength]
    d = {"mode": m, "queryseq": q.decode("latin1"), "defect_rate": b.defect_rate, "jump_late_p": b.jump_late_p,
         "dynprogindex": b.dynprogindex, "max_rlength": max_r, "max_glength": max_g, "extraband": b.extraband}
    gL = bytes(b.gsegL)[:b.glength].decode("latin1")
    if m == "single":
        d.update(rlength=b.rlength, glength=b.glength, roffset=b.roffset, goffset=b.goffset, gseg=gL, gseg_alt=gL,
                 widebandp=b.widebandp)
    elif m in ("end5", "end3"):
        d.update(rlength=b.rlength, glength=b.glength, rlength_orig=b.rlength, glength_orig=b.glength, roffset=b.roffset,
                 goffset=b.goffset, gseg=gL, gseg_alt=gL, endalign=b.endalign, require_pos_score_p=b.require_pos_score_p)
    elif m == "genome":
        gR = bytes(b.gsegR)[:b.glengthR].decode("latin1")
        d.update(rlength=b.rlength, glengthL=b.glength, glengthR=b.glengthR, roffset=b.roffset, goffsetL=b.goffset,
                 rev_goffsetR=b.rev_goffsetR, gsegL=gL, gsegL_alt=gL, gsegR=gR, gsegR_alt=gR,
                 left_probs=list(b.left_probs[:b.glength - 1]), right_probs=list(b.right_probs[:b.glengthR - 1]),
                 cdna_direction=b.cdna_direction, maxpeelback=b.maxpeelback, halfp=b.halfp, finalp=b.finalp)
    else:
        d.update(rlengthL=b.rlength, rlengthR=b.rlengthR, glength=b.glength, roffsetL=b.roffset,
                 rev_roffsetR=b.rev_roffsetR, goffset=b.goffset, gseg=gL, gseg_alt=gL, rev_gseg=gL, rev_gseg_alt=gL)
    return d


def attach_ref_world(ref, box):
    """Wraps a generated box in a tiny in-memory genome so the compiled reference can run it (it
    fetches the segments itself and computes its own MaxEnt probabilities)."""
    import random
    rng = random.Random(len(box["queryseq"]))
    junk = lambda n: bytes(rng.choices(b"ACGT", k=n))
    m = box["mode"]
    if m == "genome":
        gl = box["glengthL"]
        mid = box["rev_goffsetR"] - box["goffsetL"] + 1 - 2 * gl
        g = junk(box["goffsetL"]) + box["gsegL"].encode() + junk(mid) + box["gsegR"].encode() + junk(64)
    elif m == "end5":
        g = junk(box["goffset"] - box["glength"] + 1) + box["gseg"].encode() + junk(64)
    else:
        g = junk(box["goffset"]) + box["gseg"].encode() + junk(64)
    b = dict(box)
    b["world"] = {"handle": ref.genome_new(g), "chroffset": 0, "chrhigh": len(g), "watsonp": 1}
    return b
