"""ctypes binding of libgmapdp_b200.so: :class:`Engine` = one device context (``gmapdp_ctx``),
:class:`Batch` = the batch form of the five reference entry points (``gmapdp_shim.h``)."""
import ctypes as C
import os

from .build import LIB

MODE_IDS = {"single": 0, "genome": 1, "cdna": 2, "end5": 3, "end3": 4}


class EngineError(RuntimeError):
    pass


class Pair(C.Structure):
    _fields_ = [("querypos", C.c_int), ("genomepos", C.c_int), ("queryjump", C.c_int),
                ("genomejump", C.c_int), ("dynprogindex", C.c_int), ("introntype", C.c_int),
                ("gapp", C.c_int), ("cdna", C.c_char), ("comp", C.c_char), ("genome", C.c_char),
                ("genomealt", C.c_char), ("donor_prob", C.c_double), ("acceptor_prob", C.c_double)]


class Box(C.Structure):
    """gmapdp_box (include/gmapdp_b200.h)"""
    _fields_ = [("mode", C.c_int32), ("flags", C.c_int32), ("rlenL", C.c_int32), ("rlenR", C.c_int32), ("glenL", C.c_int32),
                ("glenR", C.c_int32), ("mismatchtype", C.c_int8), ("open", C.c_int8), ("extend", C.c_int8), ("cdna_direction", C.c_int8),
                ("lbandL", C.c_int16), ("ubandL", C.c_int16), ("lbandR", C.c_int16), ("ubandR", C.c_int16),
                ("qL_off", C.c_uint32), ("qR_off", C.c_uint32), ("gL_off", C.c_uint32), ("gLalt_off", C.c_uint32),
                ("gR_off", C.c_uint32), ("gRalt_off", C.c_uint32), ("probL_off", C.c_uint32), ("probR_off", C.c_uint32),
                ("offdiff", C.c_int32), ("revmask", C.c_int32),
                ("chroffset", C.c_uint32), ("chrhigh", C.c_uint32), ("gflags", C.c_uint16), ("probkindL", C.c_uint8), ("probkindR", C.c_uint8),
                ("probposL", C.c_uint32), ("probposR", C.c_uint32)]


class Coords(C.Structure):
    """gmapdp_coords (include/gmapdp_shim.h)"""
    _fields_ = [("chroffset", C.c_uint32), ("chrhigh", C.c_uint32), ("gposL", C.c_uint32), ("gposR", C.c_uint32),
                ("negL", C.c_int), ("negR", C.c_int), ("leftL", C.c_int), ("leftR", C.c_int), ("probs", C.c_int),
                ("probposL", C.c_uint32), ("probposR", C.c_uint32), ("probnegL", C.c_int), ("probnegR", C.c_int),
                ("probkindL", C.c_int), ("probkindR", C.c_int)]


class MaxentTables(C.Structure):
    """gmapdp_maxent_tables (include/gmapdp_b200.h): sixteen table pointers"""
    _fields_ = [("t", C.c_void_p * 16)]


class DeviceResult(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("status", "finalscore", "bestrL", "bestcL", "bestrR", "bestcR", "tb_score",
                                         "nmatches", "nmismatches", "nopens", "nindels", "script_off", "script_lenA",
                                         "script_lenB", "cells", "reserved")]


_libs = {}


def load_library(path=None):
    """Loads the native library; raises if it has not been built (``__graft_entry__.build()``).
    ``path`` (or the environment variable GMAPDP_LIB) selects another build of the same sources: tuning builds
    (scripts/variants.sh) and the test build with one-entry tie lists (build.LIB_TIECAP1)."""
    path = path or os.environ.get("GMAPDP_LIB") or LIB
    if path not in _libs:
        if not os.path.exists(path):
            raise EngineError("native library %s is missing: run __graft_entry__.build() (nvcc, sm_100a)" % path)
        lib = C.CDLL(path)
        lib.gmapdp_last_error.restype = C.c_char_p
        lib.gmapdp_launch_count.restype = C.c_long
        lib.GmapDP_batch_new.restype = C.c_void_p
        lib.GmapDP_batch_error.restype = C.c_char_p
        lib.GmapDP_batch_cells.restype = C.c_long
        lib.GmapDP_batch_cells8.restype = C.c_long
        lib.GmapDP_batch_cells_full.restype = C.c_long
        lib.GmapDP_batch_cells8_full.restype = C.c_long
        lib.GmapDP_batch_digest.restype = C.c_ulonglong
        lib.GmapDP_batch_h2d_bytes.restype = C.c_size_t
        lib.GmapDP_batch_d2h_bytes.restype = C.c_size_t
        lib.GmapDP_device_result.restype = C.POINTER(DeviceResult)
        lib.gmapdp_stream_error.restype = C.c_char_p
        lib.GmapChain_batch_new.restype = C.c_void_p
        lib.GmapChain_batch_error.restype = C.c_char_p
        lib.GmapChain_batch_nhits.restype = C.c_long
        lib.GmapChain_batch_h2d_bytes.restype = C.c_long
        lib.GmapChain_batch_d2h_bytes.restype = C.c_long
        lib.GmapChain_batch_digest.restype = C.c_ulonglong
        for fn in ("gmapdp_destroy", "gmapdp_last_error", "gmapdp_launch_count", "gmapdp_device_info"):
            getattr(lib, fn).argtypes = None
        _libs[path] = lib
    return _libs[path]


def _b(s):
    return s if isinstance(s, bytes) else s.encode("latin1")


class Engine:
    def __init__(self, device=0, lib_path=None):
        self.lib = load_library(lib_path)
        self.ctx = C.c_void_p()
        rc = self.lib.gmapdp_create(C.byref(self.ctx), int(device))
        if rc != 0:
            msg = self.lib.gmapdp_last_error(self.ctx).decode() if self.ctx else "gmapdp_create failed"
            raise EngineError(msg)
        self.device = device

    def close(self):
        if self.ctx:
            self.lib.gmapdp_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def launch_count(self):
        return self.lib.gmapdp_launch_count(self.ctx)

    def last_kernel_ms(self):
        """(full-fill kernel ms, E-only kernel ms) of the last resident run"""
        a, b = C.c_float(), C.c_float()
        self.lib.gmapdp_last_kernel_ms(self.ctx, C.byref(a), C.byref(b))
        return a.value, b.value

    def last_kernel_ms4(self):
        """per-kind kernel ms of the last resident run: single, end, genome, cdna"""
        ms = (C.c_float * 4)()
        self.lib.gmapdp_last_kernel_ms4(self.ctx, ms)
        return [float(x) for x in ms]

    def device_info(self):
        sm, grid, bt = C.c_int(), C.c_int(), C.c_int()
        self.lib.gmapdp_device_info(self.ctx, C.byref(sm), C.byref(grid), C.byref(bt))
        return {"sm_count": sm.value, "grid_blocks": grid.value, "block_threads": bt.value}

    def batch(self, max_rlength=2000, max_glength=2030):
        b = Batch(self, max_rlength, max_glength)
        if getattr(self, "_genome", None):
            b.set_genome(*self._genome_host)
        return b

    def genome_attach(self, blocks, nwords, tables):
        """Resident genome: ``blocks`` = address of the reference's Genomecomp_T words (high, low, flags per 32 nt),
        ``tables`` = sixteen addresses of the MaxEnt tables in the order of gmapdp_maxent_tables.  The caller keeps both alive."""
        mt = MaxentTables()
        for k in range(16):
            mt.t[k] = tables[k]
        g = C.c_void_p()
        self.lib.gmapdp_device_info  # (library loaded)
        rc = self.lib.gmapdp_genome_create(C.byref(g), self.device, C.c_void_p(blocks), C.c_size_t(nwords), C.byref(mt))
        if rc != 0:
            raise EngineError("gmapdp_genome_create failed (%d)" % rc)
        if self.lib.gmapdp_genome_attach(self.ctx, g) != 0:
            raise EngineError(self.lib.gmapdp_last_error(self.ctx).decode())
        old = getattr(self, "_genome", None)
        self._genome, self._genome_host = g, (blocks, nwords, mt)
        if old:
            self.lib.gmapdp_genome_destroy(old)

    def maxent_eval(self, kinds, positions, chroffset):
        """Maxent_hr_*_prob on the device for (kind, position) pairs: list of floats"""
        n = len(kinds)
        ka, pa, out = (C.c_int * n)(*kinds), (C.c_uint32 * n)(*positions), (C.c_double * n)()
        if self.lib.gmapdp_maxent_eval(self.ctx, ka, pa, C.c_uint32(chroffset), n, out) != 0:
            raise EngineError(self.lib.gmapdp_last_error(self.ctx).decode())
        return list(out)

    def chain_setup(self, splicingp=1, cross_species_p=0, sufflookback=60, nsufflookback=5, maxintronlen=500000):
        """Stage2_setup (stage2.c:129) for the chaining engine; defaults are gmap's (gmap.c:269,270,347)."""
        rc = self.lib.gmapchain_setup(self.ctx, int(splicingp), int(cross_species_p), int(sufflookback), int(nsufflookback),
                                      int(maxintronlen))
        if rc != 0:
            raise EngineError(self.lib.gmapdp_last_error(self.ctx).decode())

    def chain_batch(self):
        return ChainBatch(self)


class Batch:
    """Queue calls with add(box) (box dicts as produced by tests/dpgen.py), run(), then result(id)."""
    MAXPAIRS = 12000

    def __init__(self, engine, max_rlength, max_glength):
        self.e = engine
        self.lib = engine.lib
        self.h = C.c_void_p(self.lib.GmapDP_batch_new(engine.ctx, max_rlength, max_glength))
        self.buf = (Pair * self.MAXPAIRS)()
        self._keep = []

    def free(self):
        if self.h:
            self.lib.GmapDP_batch_free(self.h)
            self.h = C.c_void_p()

    def clear(self):
        self.lib.GmapDP_batch_clear(self.h)
        self._keep = []

    def set_genome(self, blocks, nwords, tables):
        self._genome_keep = tables
        self.lib.GmapDP_batch_genome(self.h, C.c_void_p(blocks), C.c_size_t(nwords), C.byref(tables))

    def add(self, box):
        m = box["mode"]
        co = box.get("coords")
        if co is not None:              # resident-genome form of the call (tests/dpgen.coords_of)
            cs = Coords(**co)
            self.lib.GmapDP_batch_next_coords(self.h, C.byref(cs))
        q = _b(box["queryseq"])
        quc = q.upper()
        qb, qucb = C.create_string_buffer(q, len(q) + 1), C.create_string_buffer(quc, len(quc) + 1)
        base, baseuc = C.addressof(qb), C.addressof(qucb)
        self._keep.append((qb, qucb))
        dr = C.c_double(box["defect_rate"])
        if m == "single":
            return self.lib.GmapDP_single_gap(self.h, box["dynprogindex"], C.c_void_p(base + box["roffset"]),
                                              C.c_void_p(baseuc + box["roffset"]), box["rlength"], box["glength"],
                                              box["roffset"], box["goffset"], _b(box["gseg"]), _b(box["gseg_alt"]),
                                              box["jump_late_p"], box["extraband"], box["widebandp"], dr)
        if m in ("end5", "end3"):
            fn = self.lib.GmapDP_end5_gap if m == "end5" else self.lib.GmapDP_end3_gap
            return fn(self.h, box["dynprogindex"], C.c_void_p(base + box["roffset"]), C.c_void_p(baseuc + box["roffset"]),
                      box["rlength_orig"], box["glength_orig"], box["roffset"], box["goffset"], _b(box["gseg"]),
                      _b(box["gseg_alt"]), box["jump_late_p"], box["extraband"], dr, box["endalign"],
                      box["require_pos_score_p"])
        if m == "genome":
            if co is not None and co.get("probs"):
                lp = rp = None          # MaxEnt on the device
            else:
                lp = (C.c_double * len(box["left_probs"]))(*box["left_probs"])
                rp = (C.c_double * len(box["right_probs"]))(*box["right_probs"])
            return self.lib.GmapDP_genome_gap(self.h, box["dynprogindex"], C.c_void_p(base + box["roffset"]),
                                              C.c_void_p(baseuc + box["roffset"]), box["rlength"], box["glengthL"],
                                              box["glengthR"], box["roffset"], box["goffsetL"], box["rev_goffsetR"],
                                              _b(box["gsegL"]), _b(box["gsegL_alt"]), _b(box["gsegR"]), _b(box["gsegR_alt"]),
                                              lp, rp, box["cdna_direction"], box["jump_late_p"], box["extraband"], dr,
                                              box["maxpeelback"], box["halfp"], box["finalp"])
        if m == "cdna":
            return self.lib.GmapDP_cdna_gap(self.h, box["dynprogindex"], C.c_void_p(base + box["roffsetL"]),
                                            C.c_void_p(baseuc + box["roffsetL"]), C.c_void_p(base + box["rev_roffsetR"]),
                                            C.c_void_p(baseuc + box["rev_roffsetR"]), box["rlengthL"], box["rlengthR"],
                                            box["glength"], box["roffsetL"], box["rev_roffsetR"], box["goffset"],
                                            _b(box["gseg"]), _b(box["gseg_alt"]), _b(box["rev_gseg"]), _b(box["rev_gseg_alt"]),
                                            box["jump_late_p"], box["extraband"], dr)
        raise ValueError(m)

    def _check(self, rc):
        if rc != 0:
            raise EngineError(self.lib.GmapDP_batch_error(self.h).decode())

    def set_user_dynprog(self, user_open, user_extend, enabled=True):
        """--indel-open / --indel-extend (Dynprog_*_setup's user_open, user_extend, user_dynprog_p)"""
        self._check(self.lib.GmapDP_batch_user_dynprog(self.h, int(user_open), int(user_extend), int(bool(enabled))))

    def device_view(self):
        """(boxes*, nboxes, seq*, seqbytes, probs*, nprobs) of the queued device boxes, as ctypes values"""
        boxes, seq, probs = C.c_void_p(), C.c_void_p(), C.c_void_p()
        n, sb, npb = C.c_int(), C.c_size_t(), C.c_size_t()
        self._check(self.lib.GmapDP_batch_device_view(self.h, C.byref(boxes), C.byref(n), C.byref(seq), C.byref(sb),
                                                      C.byref(probs), C.byref(npb)))
        return boxes, n.value, seq, sb.value, probs, npb.value

    def complete(self, results, script):
        self._check(self.lib.GmapDP_batch_complete(self.h, results, script))

    def run(self):
        self._check(self.lib.GmapDP_batch_run(self.h))

    def rewind(self):
        """device calls back to their queued state (run() can then be repeated and timed)"""
        self.lib.GmapDP_batch_rewind(self.h)

    def upload(self):
        self._check(self.lib.GmapDP_batch_upload(self.h))

    def run_resident(self):
        ms = C.c_float()
        self._check(self.lib.GmapDP_batch_run_resident(self.h, C.byref(ms)))
        return ms.value

    def finish(self):
        self._check(self.lib.GmapDP_batch_finish(self.h))

    def run_device(self):
        """gmapdp_run_batch from host buffers: H2D + kernel + D2H (no pair-list replay)"""
        self._check(self.lib.GmapDP_batch_run_device(self.h))

    def download(self):
        self._check(self.lib.GmapDP_batch_download(self.h))

    def digest(self):
        return self.lib.GmapDP_batch_digest(self.h)

    def cells8(self):
        return self.lib.GmapDP_batch_cells8(self.h)

    def cells_full(self):
        """(cells, 8-bit cells) of the single-gap boxes, i.e. of the full-fill kernel"""
        return self.lib.GmapDP_batch_cells_full(self.h), self.lib.GmapDP_batch_cells8_full(self.h)

    def ncalls(self):
        return self.lib.GmapDP_batch_ncalls(self.h)

    def nboxes(self):
        return self.lib.GmapDP_batch_nboxes(self.h)

    def cells(self):
        return self.lib.GmapDP_batch_cells(self.h)

    def h2d_bytes(self):
        return self.lib.GmapDP_batch_h2d_bytes(self.h)

    def d2h_bytes(self):
        return self.lib.GmapDP_batch_d2h_bytes(self.h)

    def result(self, cid, mode):
        ni = 10 if mode == "genome" else (3 if mode == "cdna" else 6)
        iout = (C.c_int * ni)()
        dout = (C.c_double * 2)()
        n = self.lib.GmapDP_result(self.h, cid, iout, dout, self.buf, self.MAXPAIRS)
        if n < -1:
            raise EngineError("GmapDP_result(%d) = %d" % (cid, n))
        pairs = []
        for i in range(max(n, 0)):
            p = self.buf[i]
            pairs.append((p.querypos, p.genomepos, p.cdna, p.comp, p.genome, p.genomealt, p.gapp,
                          p.queryjump, p.genomejump, p.introntype, p.dynprogindex, p.donor_prob, p.acceptor_prob))
        return n, list(iout), (list(dout) if mode == "genome" else []), pairs

    def device_result(self, cid):
        p = self.lib.GmapDP_device_result(self.h, cid)
        return p.contents if p else None


class Mailbox(C.Structure):
    pass


Mailbox._fields_ = [("result", DeviceResult), ("ops", C.POINTER(C.c_uint32)), ("ops_cap", C.c_size_t), ("state", C.c_int), ("rc", C.c_int),
                    ("wake", C.POINTER(Mailbox) * 2), ("t_submit", C.c_double)]


class Stream:
    """The batching runtime (gmapdp_stream.h): many host threads submit single boxes, flights carry them to the
    device(s).  ``call(batch)`` runs the ONE call queued in a private batch made with ``private_batch()``."""

    def __init__(self, devices=(0,), max_boxes=0, lib_path=None):
        self.lib = load_library(lib_path)
        self.h = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices)
        rc = self.lib.gmapdp_stream_create(C.byref(self.h), devs, len(devices), int(max_boxes))
        if rc != 0:
            msg = self.lib.gmapdp_stream_error(self.h).decode() if self.h else "gmapdp_stream_create failed"
            if self.h:
                self.lib.gmapdp_stream_destroy(self.h)
                self.h = C.c_void_p()
            raise EngineError(msg)

    def close(self):
        if self.h:
            self.lib.gmapdp_stream_destroy(self.h)
            self.h = C.c_void_p()

    def private_batch(self, max_rlength=2000, max_glength=2030):
        class _NoCtx:
            pass
        e = _NoCtx()
        e.lib, e.ctx = self.lib, C.c_void_p()
        b = Batch(e, max_rlength, max_glength)
        cap = 2 * (max_rlength + max_glength) + 64
        b._ops = (C.c_uint32 * cap)()
        b._mailbox = Mailbox()
        b._mailbox.ops = C.cast(b._ops, C.POINTER(C.c_uint32))
        b._mailbox.ops_cap = cap
        return b

    def call(self, batch):
        """submit + wait + complete for the single call queued in ``batch`` (made by private_batch)"""
        boxes, n, seq, sb, probs, npb = batch.device_view()
        if n == 0:
            return
        assert n == 1
        mb = batch._mailbox
        if self.lib.gmapdp_stream_submit(self.h, boxes, seq, C.c_size_t(sb), probs, C.c_size_t(npb), C.byref(mb)) != 0 or \
           self.lib.gmapdp_stream_wait(self.h, C.byref(mb)) != 0:
            raise EngineError(self.lib.gmapdp_stream_error(self.h).decode())
        batch.complete(C.byref(mb.result), mb.ops)

    def stats(self):
        out = (C.c_double * 9)()
        self.lib.gmapdp_stream_stats(self.h, out)
        keys = ("boxes", "flights", "largest_flight", "device_seconds", "flight_seconds", "wait_seconds", "kernel_launches",
                "h2d_bytes", "d2h_bytes")
        return dict(zip(keys, [float(x) for x in out]))


class ChainBatch:
    """Batch form of align_compute_lookback (stage2.c:4402): add(problem) with the dicts of tests/chaingen.py
    (positions CSR, npositions, minactive, maxactive, ...), run(), then paths(id)."""

    def __init__(self, engine):
        import numpy as np
        self.np = np
        self.e = engine
        self.lib = engine.lib
        self.h = C.c_void_p(self.lib.GmapChain_batch_new(engine.ctx))
        self._keep = []

    def free(self):
        if self.h:
            self.lib.GmapChain_batch_free(self.h)
            self.h = C.c_void_p()

    def clear(self):
        self.lib.GmapChain_batch_clear(self.h)
        self._keep = []

    def _check(self, rc):
        if rc < 0:
            raise EngineError(self.lib.GmapChain_batch_error(self.h).decode() or "gmapchain error %d" % rc)
        return rc

    def add(self, pb, use_canonical_p=0, forward=False):
        np = self.np
        L = int(pb["querylength"])
        pos = np.ascontiguousarray(pb["positions"], dtype=np.uint32)
        npos = np.ascontiguousarray(pb["npositions"], dtype=np.int32)
        mina = np.ascontiguousarray(pb["minactive"], dtype=np.uint32)
        maxa = np.ascontiguousarray(pb["maxactive"], dtype=np.uint32)
        if len(pos) == 0:
            pos = np.zeros(1, dtype=np.uint32)
        base = pos.ctypes.data
        cum = np.concatenate([[0], np.cumsum(np.maximum(npos, 0))[:-1]]).astype(np.int64) if L else np.zeros(0, dtype=np.int64)
        ptrs = (np.uint64(base) + np.uint64(4) * cum.astype(np.uint64)) if L else np.zeros(1, dtype=np.uint64)   # Chrpos_T *mappings[q]
        rc = (self.lib.GmapChain_lookforward if forward else self.lib.GmapChain_lookback)(
            self.h, ptrs.ctypes.data_as(C.c_void_p), npos.ctypes.data_as(C.c_void_p), C.c_int(int(np.maximum(npos, 0).sum())),
            mina.ctypes.data_as(C.c_void_p), maxa.ctypes.data_as(C.c_void_p), C.c_int(L), C.c_int(int(pb["querystart"])),
            C.c_int(int(pb["queryend"])), C.c_int(int(pb["indexsize"])), C.c_int(int(pb["localp"])),
            C.c_int(int(pb["skip_repetitive_p"])), C.c_int(int(use_canonical_p)), C.c_int(4), C.c_int(int(pb["favor_right_p"])),
            C.c_int(int(pb["middlep"])), C.c_int(int(pb["max_nalignments"])))
        return self._check(rc)

    def run(self):
        self._check(self.lib.GmapChain_batch_run(self.h))

    def upload(self):
        self._check(self.lib.GmapChain_batch_upload(self.h))

    def run_resident(self):
        ms = C.c_float()
        self._check(self.lib.GmapChain_batch_run_resident(self.h, C.byref(ms)))
        return ms.value

    def download(self):
        self._check(self.lib.GmapChain_batch_download(self.h))

    def ncalls(self):
        return self.lib.GmapChain_batch_ncalls(self.h)

    def nhits(self):
        return self.lib.GmapChain_batch_nhits(self.h)

    def h2d_bytes(self):
        return self.lib.GmapChain_batch_h2d_bytes(self.h)

    def d2h_bytes(self):
        return self.lib.GmapChain_batch_d2h_bytes(self.h)

    def digest(self):
        return self.lib.GmapChain_batch_digest(self.h)

    def paths(self, cid):
        """[(cell[5], pairs (n,2) int32: querypos, position)] in the reference's order"""
        np = self.np
        n = self._check(self.lib.GmapChain_npaths(self.h, cid))
        out = []
        cell = (C.c_int * 5)()
        cap = 4096
        for k in range(n):
            while True:
                q = np.zeros(cap, dtype=np.int32)
                p = np.zeros(cap, dtype=np.uint32)
                m = self.lib.GmapChain_path(self.h, cid, k, cell, q.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), cap)
                if m >= 0:
                    break
                if m > -16:     # an error code, not a size
                    self._check(m)
                cap = -m
            out.append((list(cell), np.stack([q[:m], p[:m].astype(np.int32)], axis=1)))
        return out

    def links(self):
        """(links (tot,5) with the trace label zeroed, scores (tot,)) of the last run, in batch hit order"""
        np = self.np
        tot = self.nhits()
        links = np.zeros((max(tot, 1), 5), dtype=np.int32)
        scores = np.zeros(max(tot, 1), dtype=np.int32)
        rc = self.lib.gmapchain_download_links(self.e.ctx, links.ctypes.data_as(C.c_void_p), scores.ctypes.data_as(C.c_void_p),
                                               C.c_size_t(tot))
        if rc != 0:
            raise EngineError(self.lib.gmapdp_last_error(self.e.ctx).decode())
        return links[:tot], scores[:tot]
