"""Synthetic DP-box generator (test / bench infrastructure).

A *box* is one call of one of the five reference entry points, as a self-contained dict: the
query chars, the genomic segment chars exactly as the reference would fetch them
(Genome_get_segment_right/left, genome.c:11023/11079), the scalar arguments, and for genome
boxes the MaxEnt probability arrays (dynprog_genome.c:970-1061).

Two finalisers share one spec generator:
* ``finalize_synthetic``  -- pure python, segments cut from the locus text, pseudo-random splice
  probabilities (what the GPU tests and bench.py use; no reference needed);
* ``build_world`` + ``finalize_with_ref`` -- place the loci on a chromosome of an in-memory
  reference genome and fetch segments / MaxEnt probabilities with the compiled reference itself
  (what tests/golden/make_golden.py and tests/test_oracle_vs_ref.py use).

Shape envelope: genome glengthL = glengthR >= rlength + 1 (production rlength + 8),
end glength >= rlength - 1 (production rlength + 10), cdna rlengthL = rlengthR = glength + 8,
halfp = 0.
"""
import random

COMP = bytes.maketrans(b"ACGTNacgtn", b"TGCANtgcan")
IUPAC = b"RYWSMKHBVDN"


def revcomp(s):
    return s.translate(COMP)[::-1]


LOW_COMPLEXITY = False      # set by lowcomplexity_boxes(): repeats of short units, quantised probabilities


def rand_dna(rng, n, nfrac=0.003):
    if LOW_COMPLEXITY and n > 0:
        # homopolymers / di- and trinucleotide repeats with rare breaks: long runs of equal scores, the
        # stress case for every tie rule (bridge candidates, endpoint search, traceback)
        unit = bytes(rng.choices(b"ACGT", k=rng.choice([1, 1, 2, 2, 3, 5])))
        out = bytearray((unit * (n // len(unit) + 1))[:n])
        for _ in range(n // 40):
            out[rng.randrange(n)] = rng.choice(b"ACGT")
        return bytes(out)
    out = bytearray(rng.choices(b"ACGT", k=n))
    if nfrac > 0:
        for _ in range(int(n * nfrac)):
            out[rng.randrange(n)] = ord("N")
    return bytes(out)


def mutate(rng, s, e):
    """60% substitution / 20% deletion / 20% insertion at per-base rate e."""
    out = bytearray()
    for ch in s:
        if rng.random() < e:
            k = rng.random()
            if k < 0.6:
                out.append(rng.choice(b"ACGT"))
            elif k < 0.8:
                continue
            else:
                out.append(ch)
                out.append(rng.choice(b"ACGT"))
        else:
            out.append(ch)
    return bytes(out)


def decorate_query(rng, q):
    """a small share of IUPAC and lower-case query chars"""
    q = bytearray(q)
    if rng.random() < 0.15 and len(q):
        for _ in range(1 + len(q) // 60):
            q[rng.randrange(len(q))] = rng.choice(IUPAC)
    if rng.random() < 0.15 and len(q):
        a = rng.randrange(len(q))
        b = min(len(q), a + rng.randrange(1, 12))
        q[a:b] = bytes(q[a:b]).lower()
    return bytes(q)


ERR_LEVELS = [(0.0, 0.001), (0.001, 0.001), (0.01, 0.01), (0.05, 0.05), (0.15, 0.15)]


CDNA_GRANGE = None      # (lo, hi): overrides the genomic length of cdna boxes


def gen_spec(rng, mode=None, rmin=15, rmax=150, large=False):
    """Returns a spec: everything about a box except the fetched segments."""
    mode = mode or rng.choice(["single", "genome", "cdna", "end5", "end3"])
    e, defect = rng.choice(ERR_LEVELS)
    s = {"mode": mode, "defect_rate": defect, "jump_late_p": rng.randrange(2),
         "dynprogindex": rng.choice([1, 2, 7, -1, -3]), "watsonp": rng.randrange(2),
         "prefix": rng.randrange(0, 30), "suffix": rng.randrange(0, 30)}
    FL = 40  # locus flank
    if mode == "single":
        g = rng.randrange(rmin, rmax)
        core = rand_dna(rng, g)
        kind = rng.random()
        if kind < 0.15:
            # tandem-repeat stratum: matching diagonal just below the band
            unit = rand_dna(rng, rng.randrange(16, 25), 0)
            core = (unit * (g // len(unit) + 2))[:g]
            rot = rng.randrange(7, 10)
            qseg = mutate(rng, (core + core)[rot:rot + g], e * 0.3)
        elif kind < 0.3:
            qseg = bytearray(core)      # equal lengths, 0-2 substitutions -> single_gap_simple
            for _ in range(rng.randrange(0, 3)):
                qseg[rng.randrange(g)] = rng.choice(b"ACGT")
            qseg = bytes(qseg)
        else:
            qseg = mutate(rng, core, e)
            if rng.random() < 0.3 and len(qseg) > 10:   # a longer indel
                a = rng.randrange(1, len(qseg) - 1)
                n = rng.randrange(1, 12)
                qseg = qseg[:a] + (rand_dna(rng, n, 0) if rng.random() < 0.5 else b"") + qseg[a + (0 if rng.random() < 0.5 else n):]
        if large:
            qseg = qseg[:rng.randrange(max(1, len(qseg) // 3), len(qseg) + 1)]
        if not qseg:
            qseg = b"A"
        r = len(qseg)
        band = rng.choice([6, max(6, abs(r - g)), rng.randrange(0, 20)])
        # every stage3 caller passes widebandp=true (stage3.c:9081,9261,9510); without it the corner
        # (rlength, glength) must still lie inside the band or the reference traces stale memory
        wide = 1 if band < abs(r - g) else rng.choice([1, 1, 0])
        s.update(T=rand_dna(rng, FL) + core + rand_dna(rng, FL), x=FL, qmid=qseg, rlength=r, glength=g,
                 extraband=band, widebandp=wide)
    elif mode in ("end5", "end3"):
        r = rng.randrange(max(2, rmin // 2), rmax)
        g = r + rng.choice([10, 10, 10, 10, -1, 0, 3, 12])
        core = rand_dna(rng, g)
        m = mutate(rng, core, e)
        if rng.random() < 0.3:      # diverging far end
            k = rng.randrange(0, len(m) + 1)
            m = (m[:k] + rand_dna(rng, len(m) - k + 20, 0)) if mode == "end3" else (rand_dna(rng, 20 + k, 0) + m[k:])
        while len(m) < r:
            m = m + rand_dna(rng, r, 0) if mode == "end3" else rand_dna(rng, r, 0) + m
        qseg = m[:r] if mode == "end3" else m[-r:]
        s.update(T=rand_dna(rng, FL) + core + rand_dna(rng, FL), x=FL, qmid=qseg, rlength=r, glength=g,
                 extraband=rng.choice([6, 6, 6, 3, 10]), endalign=rng.choice([0, 1, 2, 3, 0, 1]),
                 require_pos_score_p=1 if rng.random() < 0.1 else 0)
    elif mode == "genome":
        a = rng.randrange(max(2, rmin // 2), max(3, rmax // 2))
        b = rng.randrange(max(2, rmin // 2), max(3, rmax // 2))
        A, B = rand_dna(rng, a), rand_dna(rng, b)
        k = rng.random()
        ilen = rng.randrange(20, 400)
        I = bytearray(rand_dna(rng, ilen))
        if k < 0.65:
            I[:2], I[-2:] = b"GT", b"AG"
        elif k < 0.77:
            I[:2], I[-2:] = b"GC", b"AG"
        elif k < 0.89:
            I[:2], I[-2:] = b"CT", b"AC"
        qseg = mutate(rng, A, e) + mutate(rng, B, e)
        if len(qseg) < 2:
            qseg = b"ACGT"
        r = len(qseg)
        gl = r + rng.choice([8, 8, 8, 8, 8, 2, 5, 12])
        T = rand_dna(rng, FL) + A + bytes(I) + B + rand_dna(rng, FL + 20)
        s.update(T=T, x=FL, xR=FL + a + ilen + b - 1, qmid=qseg, rlength=r, glengthL=gl, glengthR=gl,
                 cdna_direction=rng.choice([1, 1, -1, 0]), extraband=rng.choice([14, 14, 14, 16, 20]),
                 maxpeelback=60, halfp=0, finalp=1 if rng.random() < 0.2 else 0, junction=(a, b))
    elif mode == "cdna":
        # the cDNA bridge is O(g^2 band^2) on the CPU (dynprog_cdna.c:149-375): keep g small
        glo = min(max(4, rmin // 2), 100)
        g = rng.randrange(glo, max(glo + 1, min(rmax, 140)))
        if CDNA_GRANGE is not None:     # long cDNA gaps: production allows rL = rR = g + 8 <= 660 (dynprog_cdna.c:869-892, stage3.c:9276)
            g = rng.randrange(CDNA_GRANGE[0], CDNA_GRANGE[1] + 1)
        core = rand_dna(rng, g)
        k = rng.randrange(1, g)
        ins = rand_dna(rng, rng.randrange(10, 40), 0)
        qseg = mutate(rng, core[:k], e) + ins + mutate(rng, core[k:], e)
        while len(qseg) < g + 10:
            qseg = qseg[:k] + rand_dna(rng, 4, 0) + qseg[k:]
        s.update(T=rand_dna(rng, FL) + core + rand_dna(rng, FL), x=FL, qmid=qseg, glength=g,
                 rlengthL=g + 8, rlengthR=g + 8, extraband=rng.choice([14, 14, 16]))
    s["qmid"] = decorate_query(rng, s["qmid"])
    return s


def _query(rng, s):
    pre, suf = rand_dna(rng, s["prefix"], 0), rand_dna(rng, s["suffix"], 0)
    return pre + s["qmid"] + suf, len(pre)


def _common(s, q, max_r, max_g):
    return {"mode": s["mode"], "queryseq": q.decode("latin1"), "defect_rate": s["defect_rate"],
            "jump_late_p": s["jump_late_p"], "dynprogindex": s["dynprogindex"],
            "max_rlength": max_r, "max_glength": max_g}


def _chop(s, max_r, max_g):
    r, g = s["rlength"], s["glength"]
    if s["endalign"] != 2:
        r, g = min(r, max_r), min(g, max_g)
    return r, g


def synthetic_probs(rng, n, hot=()):
    if LOW_COMPLEXITY:
        return [rng.choice([0.0, 0.0, 0.25, 0.5, 0.9, 0.95]) for _ in range(n)]
    p = [rng.random() ** 3 for _ in range(n)]
    for h in hot:
        if 0 <= h < n:
            p[h] = 0.9 + 0.0999 * rng.random()
    return p


def _alt_of(rng, seg):
    """an alt-genome (SNP) twin of a segment: a few differing ACGT positions"""
    a = bytearray(seg.encode("latin1"))
    for _ in range(1 + len(a) // 80):
        k = rng.randrange(len(a))
        if a[k] in b"ACGT":
            a[k] = rng.choice(b"ACGT")
    return a.decode("latin1")


def finalize_synthetic(rng, s, max_r=2000, max_g=2030):
    """Segments cut straight from the locus text (box orientation)."""
    q, qoff = _query(rng, s)
    b = _common(s, q, max_r, max_g)
    T, x, m = s["T"], s["x"], s["mode"]
    b = _finalize_synthetic(rng, s, b, qoff, T, x, m, max_r, max_g)
    if rng.random() < 0.15:        # exercise the max(std, alt) genome paths
        for k in ("gseg", "gsegL", "gsegR", "rev_gseg"):
            if k in b:
                b[k + "_alt"] = _alt_of(rng, b[k])
        if m == "cdna":
            b["rev_gseg_alt"] = b["gseg_alt"]
    return b


def _finalize_synthetic(rng, s, b, qoff, T, x, m, max_r, max_g):

    def fwd(off, n):
        return T[off:off + n].decode("latin1")

    def rev(end, n):
        return T[end - n + 1:end + 1].decode("latin1")

    if m == "single":
        seg = fwd(x, s["glength"])
        b.update(rlength=s["rlength"], glength=s["glength"], roffset=qoff, goffset=1000 + x, gseg=seg, gseg_alt=seg,
                 extraband=s["extraband"], widebandp=s["widebandp"])
    elif m == "end3":
        r, g = _chop(s, max_r, max_g)
        seg = fwd(x, g)
        b.update(rlength=r, glength=g, rlength_orig=s["rlength"], glength_orig=s["glength"], roffset=qoff, goffset=1000 + x,
                 gseg=seg, gseg_alt=seg, extraband=s["extraband"], endalign=s["endalign"],
                 require_pos_score_p=s["require_pos_score_p"])
    elif m == "end5":
        r, g = _chop(s, max_r, max_g)
        seg = rev(x + s["glength"] - 1, g)
        b.update(rlength=r, glength=g, rlength_orig=s["rlength"], glength_orig=s["glength"],
                 roffset=qoff + s["rlength"] - 1, goffset=1000 + x + s["glength"] - 1,
                 gseg=seg, gseg_alt=seg, extraband=s["extraband"], endalign=s["endalign"],
                 require_pos_score_p=s["require_pos_score_p"])
    elif m == "genome":
        gL, gR = s["glengthL"], s["glengthR"]
        sL, sR = fwd(x, gL), rev(s["xR"], gR)
        a, bb = s["junction"]
        b.update(rlength=s["rlength"], glengthL=gL, glengthR=gR, roffset=qoff, goffsetL=1000 + x, rev_goffsetR=1000 + s["xR"],
                 gsegL=sL, gsegL_alt=sL, gsegR=sR, gsegR_alt=sR,
                 left_probs=synthetic_probs(rng, gL - 1, (a, a - 1, a + 1)), right_probs=synthetic_probs(rng, gR - 1, (bb, bb - 1, bb + 1)),
                 cdna_direction=s["cdna_direction"], extraband=s["extraband"], maxpeelback=s["maxpeelback"],
                 halfp=s["halfp"], finalp=s["finalp"])
    elif m == "cdna":
        g = s["glength"]
        seg = fwd(x, g)
        b.update(rlengthL=s["rlengthL"], rlengthR=s["rlengthR"], glength=g, roffsetL=qoff, rev_roffsetR=qoff + len(s["qmid"]) - 1,
                 goffset=1000 + x, gseg=seg, gseg_alt=seg, rev_gseg=seg, rev_gseg_alt=seg, extraband=s["extraband"])
    return b


def synth_boxes(seed, n, mode=None, rmin=15, rmax=150, max_r=2000, max_g=2030):
    rng = random.Random(seed)
    return [finalize_synthetic(rng, gen_spec(rng, mode, rmin, rmax), max_r, max_g) for _ in range(n)]


def lowcomplexity_boxes(seed, n, mode=None, rmin=15, rmax=150):
    """synth_boxes over low-complexity sequence with quantised splice probabilities (many exact ties)"""
    global LOW_COMPLEXITY
    LOW_COMPLEXITY = True
    try:
        return synth_boxes(seed, n, mode, rmin, rmax)
    finally:
        LOW_COMPLEXITY = False


# ---------------------------------------------------------------------------------------------
# reference-backed finaliser
# ---------------------------------------------------------------------------------------------
def build_world(ref, rng, specs, chroffset=500, edge=False):
    """Lay the loci out on one chromosome occupying [chroffset, chrhigh) of an in-memory genome.
    With edge=True the first and last loci sit flush against the chromosome ends, so fetched
    segments run into '*' padding."""
    pieces, pos = [], 0
    for i, s in enumerate(specs):
        T = s["T"] if s["watsonp"] else revcomp(s["T"])
        if not (edge and i == 0):
            sp = rand_dna(rng, rng.randrange(5, 40))
            pieces.append(sp)
            pos += len(sp)
        s["a"] = pos
        pieces.append(T)
        pos += len(T)
    if not edge:
        pieces.append(rand_dna(rng, 50))
    chrom = b"".join(pieces)
    Lc = len(chrom)
    genome = rand_dna(rng, chroffset) + chrom + rand_dna(rng, 300)
    world = {"handle": ref.genome_new(genome), "chroffset": chroffset, "chrhigh": chroffset + Lc, "Lc": Lc,
             "genome": genome}
    return world


def _boxcoord(world, s, off):
    """box-orientation chromosome coordinate of locus offset `off`"""
    if s["watsonp"]:
        return s["a"] + off
    return world["Lc"] - (s["a"] + len(s["T"]) - 1 - off)


def fetch_fwd(ref, world, watsonp, goffset, glength):
    co, ch = world["chroffset"], world["chrhigh"]
    if watsonp:
        return ref.get_segment(world["handle"], 0, co + goffset, glength, ch, 0)
    return ref.get_segment(world["handle"], 1, ch - goffset + 1, glength, co, 1)


def fetch_rev(ref, world, watsonp, rev_goffset, glength):
    co, ch = world["chroffset"], world["chrhigh"]
    if watsonp:
        return ref.get_segment(world["handle"], 1, co + rev_goffset + 1, glength, co, 0)
    return ref.get_segment(world["handle"], 0, ch - rev_goffset, glength, ch, 1)


def maxent_arrays(ref, world, watsonp, cdna_direction, leftoffset, rightoffset, glengthL, glengthR):
    co, ch, gh = world["chroffset"], world["chrhigh"], world["handle"]
    if watsonp:
        wl, wr = (0, 1) if cdna_direction > 0 else (3, 2)
        left = [ref.maxent(gh, wl, co + leftoffset + c, co) for c in range(glengthL - 1)]
        right = [ref.maxent(gh, wr, co + rightoffset - c + 1, co) for c in range(glengthR - 1)]
    else:
        wl, wr = (2, 3) if cdna_direction > 0 else (1, 0)
        left = [ref.maxent(gh, wl, ch - leftoffset - c + 1, co) for c in range(glengthL - 1)]
        right = [ref.maxent(gh, wr, ch - rightoffset + c, co) for c in range(glengthR - 1)]
    return left, right


def finalize_with_ref(ref, rng, world, s):
    max_r, max_g = ref.max_rlength, ref.max_glength
    q, qoff = _query(rng, s)
    b = _common(s, q, max_r, max_g)
    wp, m = s["watsonp"], s["mode"]
    b["world"] = {"handle": world["handle"], "chroffset": world["chroffset"], "chrhigh": world["chrhigh"], "watsonp": wp}
    d = lambda x: x.decode("latin1")
    if m == "single":
        go = _boxcoord(world, s, s["x"])
        seg, alt = fetch_fwd(ref, world, wp, go, s["glength"])
        b.update(rlength=s["rlength"], glength=s["glength"], roffset=qoff, goffset=go, gseg=d(seg), gseg_alt=d(alt),
                 extraband=s["extraband"], widebandp=s["widebandp"])
    elif m == "end3":
        r, g = _chop(s, max_r, max_g)
        go = _boxcoord(world, s, s["x"])
        seg, alt = fetch_fwd(ref, world, wp, go, g)
        b.update(rlength=r, glength=g, rlength_orig=s["rlength"], glength_orig=s["glength"], roffset=qoff, goffset=go,
                 gseg=d(seg), gseg_alt=d(alt), extraband=s["extraband"], endalign=s["endalign"],
                 require_pos_score_p=s["require_pos_score_p"])
    elif m == "end5":
        r, g = _chop(s, max_r, max_g)
        go = _boxcoord(world, s, s["x"] + s["glength"] - 1)
        seg, alt = fetch_rev(ref, world, wp, go, g)
        b.update(rlength=r, glength=g, rlength_orig=s["rlength"], glength_orig=s["glength"],
                 roffset=qoff + s["rlength"] - 1, goffset=go, gseg=d(seg), gseg_alt=d(alt),
                 extraband=s["extraband"], endalign=s["endalign"], require_pos_score_p=s["require_pos_score_p"])
    elif m == "genome":
        gL, gR = s["glengthL"], s["glengthR"]
        goL, goR = _boxcoord(world, s, s["x"]), _boxcoord(world, s, s["xR"])
        sL, aL = fetch_fwd(ref, world, wp, goL, gL)
        sR, aR = fetch_rev(ref, world, wp, goR, gR)
        lp, rp = maxent_arrays(ref, world, wp, s["cdna_direction"], goL, goR, gL, gR)
        b.update(rlength=s["rlength"], glengthL=gL, glengthR=gR, roffset=qoff, goffsetL=goL, rev_goffsetR=goR,
                 gsegL=d(sL), gsegL_alt=d(aL), gsegR=d(sR), gsegR_alt=d(aR), left_probs=lp, right_probs=rp,
                 cdna_direction=s["cdna_direction"], extraband=s["extraband"], maxpeelback=s["maxpeelback"],
                 halfp=s["halfp"], finalp=s["finalp"])
    elif m == "cdna":
        g = s["glength"]
        go = _boxcoord(world, s, s["x"])
        seg, alt = fetch_fwd(ref, world, wp, go, g)
        rseg, ralt = fetch_rev(ref, world, wp, go + g - 1, g)
        b.update(rlengthL=s["rlengthL"], rlengthR=s["rlengthR"], glength=g, roffsetL=qoff,
                 rev_roffsetR=qoff + len(s["qmid"]) - 1, goffset=go, gseg=d(seg), gseg_alt=d(alt),
                 rev_gseg=d(rseg), rev_gseg_alt=d(ralt), extraband=s["extraband"])
    return b


def coords_of(box, probs=True):
    """The resident-genome form of a ref_boxes() box (gmapdp_coords, include/gmapdp_shim.h): where element 0 of each
    segment array lies in the genome, its direction and which fetch produced it -- the fetches of fetch_fwd / fetch_rev
    above, i.e. of the reference's entry points -- and for genome gaps the coordinates of maxent_arrays()."""
    w = box["world"]
    co, ch, wp, m = w["chroffset"], w["chrhigh"], w["watsonp"], box["mode"]

    def fwd(goffset):                   # fetch_fwd: (gpos, neg, left)
        return (co + goffset, 0, 0) if wp else (ch - goffset, 1, 1)

    def rev(rev_goffset, g):            # fetch_rev
        return (co + rev_goffset + 1 - g, 0, 1) if wp else (ch - rev_goffset + g - 1, 1, 0)

    c = dict(chroffset=co, chrhigh=ch, gposL=0, gposR=0, negL=0, negR=0, leftL=0, leftR=0, probs=0,
             probposL=0, probposR=0, probnegL=0, probnegR=0, probkindL=0, probkindR=0)
    if m in ("single", "end3", "cdna"):
        c["gposL"], c["negL"], c["leftL"] = fwd(box["goffset"])
    elif m == "end5":
        c["gposL"], c["negL"], c["leftL"] = rev(box["goffset"], box["glength"])
    else:
        c["gposL"], c["negL"], c["leftL"] = fwd(box["goffsetL"])
        c["gposR"], c["negR"], c["leftR"] = rev(box["rev_goffsetR"], box["glengthR"])
        if probs:
            d = box["cdna_direction"]
            c["probs"] = 1
            if wp:
                c["probkindL"], c["probkindR"] = (0, 1) if d > 0 else (3, 2)
                c["probposL"], c["probnegL"] = co + box["goffsetL"], 0
                c["probposR"], c["probnegR"] = co + box["rev_goffsetR"] + 1, 1
            else:
                c["probkindL"], c["probkindR"] = (2, 3) if d > 0 else (1, 0)
                c["probposL"], c["probnegL"] = ch - box["goffsetL"] + 1, 1
                c["probposR"], c["probnegR"] = ch - box["rev_goffsetR"], 0
    for k in ("gposL", "gposR", "probposL", "probposR"):
        c[k] &= 0xFFFFFFFF
    return c


def ref_boxes(ref, seed, n, mode=None, rmin=15, rmax=150, edge=False):
    rng = random.Random(seed)
    specs = [gen_spec(rng, mode, rmin, rmax) for _ in range(n)]
    world = build_world(ref, rng, specs, edge=edge)
    return [finalize_with_ref(ref, rng, world, s) for s in specs], world
