"""ctypes front ends for the two CPU checkers (test infrastructure only).

* ``Oracle``  -> oracle/liboracle.so   : our scalar C restatement (always available; built by
  ``make -C oracle port`` or ``__graft_entry__.build()``).
* ``Ref``     -> oracle/_ref/ref_driver.so : the UNMODIFIED reference compiled from
  /root/reference/src (only where ``make -C oracle ref`` has been run).

Both take the same self-contained *box* dict (see ``dpgen.py``) and return
``(npairs_or_minus1, iout, dout, pairs)`` so results can be compared with ``==``.
"""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

MAXPAIRS = 12000


class Pair(C.Structure):
    _fields_ = [("querypos", C.c_int), ("genomepos", C.c_int), ("queryjump", C.c_int),
                ("genomejump", C.c_int), ("dynprogindex", C.c_int), ("introntype", C.c_int),
                ("gapp", C.c_int), ("cdna", C.c_char), ("comp", C.c_char), ("genome", C.c_char),
                ("genomealt", C.c_char), ("donor_prob", C.c_double), ("acceptor_prob", C.c_double)]


def pairs_to_list(buf, n):
    out = []
    for i in range(max(n, 0)):
        p = buf[i]
        out.append((p.querypos, p.genomepos, p.cdna, p.comp, p.genome, p.genomealt, p.gapp,
                    p.queryjump, p.genomejump, p.introntype, p.dynprogindex,
                    p.donor_prob, p.acceptor_prob))
    return out


def _b(s):
    return s if isinstance(s, bytes) else s.encode("latin1")


def queryuc_of(q):
    return _b(q).upper()


def build_oracle():
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = [os.path.join(ORACLE_DIR, f) for f in ("dp_oracle.c", "dp_oracle.h")]
    if (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "port"], stdout=subprocess.DEVNULL)
    return so


def ref_available(noflush=False):
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "ref_driver_noflush.so" if noflush else "ref_driver.so"))


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        self.lib.orc_init()
        self.lib.orc_cells.restype = C.c_long
        self.buf = (Pair * MAXPAIRS)()

    def set_user_dynprog(self, user_open, user_extend, enabled=True):
        """--indel-open / --indel-extend"""
        self.lib.orc_set_user_dynprog(int(user_open), int(user_extend), int(bool(enabled)))

    def count_cells(self, box):
        """algorithmic in-band cells of the fills this call would run (0 if an entry-point shortcut resolves it):
        the entry point's own control flow up to the fills, nothing filled"""
        self.lib.orc_cells_filled.restype = C.c_long
        self.lib.orc_cells_reset()
        self.lib.orc_set_count_only(1)
        try:
            self.run(box)
        finally:
            self.lib.orc_set_count_only(0)
        return self.lib.orc_cells_filled()

    # ---- tables / fills ----------------------------------------------------------------------
    def pairdistance(self, mt, a, b):
        return self.lib.orc_pairdistance(mt, a, b)

    def consistent(self, a, b):
        return self.lib.orc_consistent(a, b)

    def use8p_size(self, mt):
        return self.lib.orc_use8p_size(mt)

    def cells(self, kind, r, g, lband, uband):
        return self.lib.orc_cells(kind, r, g, lband, uband)

    def fill(self, kind, bits, rseq, gseq, galt, mt, open_, extend, lband, uband, late, revp):
        return _fill(self.lib.orc_fill, kind, bits, rseq, gseq, galt, mt, open_, extend, lband, uband, late, revp)

    # ---- entry points --------------------------------------------------------------------------
    def run(self, box):
        m = box["mode"]
        q, quc = _b(box["queryseq"]), queryuc_of(box["queryseq"])
        mr, mg = box["max_rlength"], box["max_glength"]
        if m == "single":
            iout = (C.c_int * 6)(box["dynprogindex"], -999, -999, -999, -999, -999)
            n = self.lib.orc_single_gap(iout, q, quc, box["rlength"], box["glength"], box["roffset"], box["goffset"],
                                        _b(box["gseg"]), _b(box["gseg_alt"]), box["jump_late_p"],
                                        box["extraband"], box["widebandp"], C.c_double(box["defect_rate"]),
                                        mr, mg, self.buf, MAXPAIRS)
            return n, list(iout), [], pairs_to_list(self.buf, n)
        if m in ("end5", "end3"):
            iout = (C.c_int * 6)(box["dynprogindex"], -999, -999, -999, -999, -999)
            n = self.lib.orc_end_gap(1 if m == "end5" else 0, iout, q, quc, box["rlength"], box["glength"],
                                     box["roffset"], box["goffset"], _b(box["gseg"]), _b(box["gseg_alt"]),
                                     box["jump_late_p"], box["extraband"], C.c_double(box["defect_rate"]),
                                     box["endalign"], box["require_pos_score_p"], mr, mg, self.buf, MAXPAIRS)
            return n, list(iout), [], pairs_to_list(self.buf, n)
        if m == "genome":
            iout = (C.c_int * 10)(box["dynprogindex"], *([-999] * 9))
            dout = (C.c_double * 2)(-9.0, -9.0)
            lp = (C.c_double * len(box["left_probs"]))(*box["left_probs"])
            rp = (C.c_double * len(box["right_probs"]))(*box["right_probs"])
            n = self.lib.orc_genome_gap(iout, dout, q, quc, box["rlength"], box["glengthL"], box["glengthR"],
                                        box["roffset"], box["goffsetL"], box["rev_goffsetR"],
                                        _b(box["gsegL"]), _b(box["gsegL_alt"]), _b(box["gsegR"]), _b(box["gsegR_alt"]),
                                        lp, rp, box["cdna_direction"], box["jump_late_p"], box["extraband"],
                                        C.c_double(box["defect_rate"]), box["maxpeelback"], box["halfp"], box["finalp"],
                                        mr, mg, self.buf, MAXPAIRS)
            return n, list(iout), list(dout), pairs_to_list(self.buf, n)
        if m == "cdna":
            iout = (C.c_int * 3)(box["dynprogindex"], -999, 0)
            n = self.lib.orc_cdna_gap(iout, q, quc, box["rlengthL"], box["rlengthR"], box["glength"],
                                      box["roffsetL"], box["rev_roffsetR"], box["goffset"],
                                      _b(box["gseg"]), _b(box["gseg_alt"]), _b(box["rev_gseg"]), _b(box["rev_gseg_alt"]),
                                      box["jump_late_p"], box["extraband"], C.c_double(box["defect_rate"]),
                                      mr, mg, self.buf, MAXPAIRS)
            return n, list(iout), [], pairs_to_list(self.buf, n)
        raise ValueError(m)


def _fill(fn, kind, bits, rseq, gseq, galt, mt, open_, extend, lband, uband, late, revp):
    import numpy as np
    r, g = len(rseq), len(gseq)
    n = (r + 1) * (g + 1)
    H = np.full(n, -30000, dtype=np.int16)
    dN = np.full(n, 9, dtype=np.int8)
    dE = np.full(n, 9, dtype=np.int8)
    dF = np.zeros(n, dtype=np.int8)
    fn(kind, bits, _b(rseq), _b(gseq), _b(galt), r, g, mt, open_, extend, lband, uband, int(late), int(revp),
       H.ctypes.data_as(C.c_void_p), dN.ctypes.data_as(C.c_void_p), dE.ctypes.data_as(C.c_void_p),
       dF.ctypes.data_as(C.c_void_p))
    sh = (r + 1, g + 1)
    return H.reshape(sh), (dN.reshape(sh) != 0) & (dN.reshape(sh) != 9), (dE.reshape(sh) != 0) & (dE.reshape(sh) != 9), dF.reshape(sh) != 0, dN.reshape(sh)


class Ref:
    """The compiled reference.  A *world* is one in-memory genome (one chromosome occupying
    [chroffset, chrhigh) of it); boxes carry world coordinates under box['world']."""

    def __init__(self, maxlookback=1940, extraquerygap=20, maxpeelback=60, extramaterial_end=10, extramaterial_paired=8, noflush=False):
        # noflush: the same sources with dynprog_simd.c's per-column _mm_clflush defined away (oracle/refbuild/noflush.h)
        self.lib = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "ref_driver_noflush.so" if noflush else "ref_driver.so"))
        self.lib.refdrv_init(maxlookback, extraquerygap, maxpeelback, extramaterial_end, extramaterial_paired)
        self.lib.refdrv_genome_new.restype = C.c_void_p
        self.lib.refdrv_maxent.restype = C.c_double
        self.buf = (Pair * MAXPAIRS)()
        self.max_rlength = self.lib.refdrv_max_rlength()
        self.max_glength = self.lib.refdrv_max_glength()

    def set_user_dynprog(self, user_open, user_extend, enabled=True):
        self.lib.refdrv_set_user_dynprog(int(user_open), int(user_extend), int(bool(enabled)))

    def pairdistance(self, mt, a, b):
        return self.lib.refdrv_pairdistance(mt, a, b)

    def consistent(self, a, b):
        return self.lib.refdrv_consistent(a, b)

    def use8p_size(self, mt):
        return self.lib.refdrv_use8p_size(mt)

    def genome_new(self, chars):
        chars = _b(chars)
        return C.c_void_p(self.lib.refdrv_genome_new(chars, len(chars)))

    def get_segment(self, gh, leftp, coord, length, bound, revcomp):
        seg = C.create_string_buffer(length + 1)
        alt = C.create_string_buffer(length + 1)
        self.lib.refdrv_get_segment(gh, int(leftp), C.c_uint(coord), length, C.c_uint(bound), int(revcomp), seg, alt)
        return seg.raw[:length], alt.raw[:length]

    def genome_blocks(self, gh):
        """(address, nwords) of the reference's compressed genome (Genomecomp_T words)"""
        n = C.c_ulong()
        self.lib.refdrv_genome_blocks.restype = C.c_void_p
        p = self.lib.refdrv_genome_blocks(gh, C.byref(n))
        return p, n.value

    def maxent_tables(self):
        """sixteen addresses of the reference's MaxEnt tables (oracle/_ref/ref_maxent.so: maxent_hr.c compiled in place)"""
        if not hasattr(self, "_me"):
            self._me = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "ref_maxent.so"))
        t = (C.c_void_p * 16)()
        self._me.refme_tables(t)
        return list(t)

    def maxent(self, gh, which, pos, chroffset):
        return self.lib.refdrv_maxent(gh, which, C.c_uint(pos), C.c_uint(chroffset))

    def fill(self, kind, bits, rseq, gseq, galt, mt, open_, extend, lband, uband, late, revp):
        return _fill(self.lib.refdrv_fill, kind, bits, rseq, gseq, galt, mt, open_, extend, lband, uband, late, revp)

    def run(self, box):
        w = box["world"]
        gh, co, ch, wp = w["handle"], w["chroffset"], w["chrhigh"], w["watsonp"]
        m = box["mode"]
        q, quc = _b(box["queryseq"]), queryuc_of(box["queryseq"])
        if m == "single":
            iout = (C.c_int * 6)(box["dynprogindex"], -999, -999, -999, -999, -999)
            n = self.lib.refdrv_single_gap(gh, iout, q, quc, box["rlength"], box["glength"], box["roffset"], box["goffset"],
                                           C.c_uint(co), C.c_uint(ch), wp, box["jump_late_p"], box["extraband"],
                                           box["widebandp"], C.c_double(box["defect_rate"]), self.buf, MAXPAIRS)
            return n, list(iout), [], pairs_to_list(self.buf, n)
        if m in ("end5", "end3"):
            iout = (C.c_int * 6)(box["dynprogindex"], -999, -999, -999, -999, -999)
            n = self.lib.refdrv_end_gap(gh, 1 if m == "end5" else 0, iout, q, quc, box["rlength_orig"], box["glength_orig"],
                                        box["roffset"], box["goffset"], C.c_uint(co), C.c_uint(ch), wp, box["jump_late_p"],
                                        box["extraband"], C.c_double(box["defect_rate"]), box["endalign"],
                                        box["require_pos_score_p"], self.buf, MAXPAIRS)
            return n, list(iout), [], pairs_to_list(self.buf, n)
        if m == "genome":
            iout = (C.c_int * 10)(box["dynprogindex"], *([-999] * 9))
            dout = (C.c_double * 2)(-9.0, -9.0)
            n = self.lib.refdrv_genome_gap(gh, iout, dout, q, quc, box["rlength"], box["glengthL"], box["glengthR"],
                                           box["roffset"], box["goffsetL"], box["rev_goffsetR"], C.c_uint(co), C.c_uint(ch),
                                           box["cdna_direction"], wp, box["jump_late_p"], box["extraband"],
                                           C.c_double(box["defect_rate"]), box["maxpeelback"], box["halfp"], box["finalp"],
                                           self.buf, MAXPAIRS)
            return n, list(iout), list(dout), pairs_to_list(self.buf, n)
        if m == "cdna":
            iout = (C.c_int * 3)(box["dynprogindex"], -999, 0)
            n = self.lib.refdrv_cdna_gap(gh, iout, q, quc, box["rlengthL"], box["rlengthR"], box["glength"],
                                         box["roffsetL"], box["rev_roffsetR"], box["goffset"], C.c_uint(co), C.c_uint(ch),
                                         wp, box["jump_late_p"], box["extraband"], C.c_double(box["defect_rate"]),
                                         self.buf, MAXPAIRS)
            return n, list(iout), [], pairs_to_list(self.buf, n)
        raise ValueError(m)
