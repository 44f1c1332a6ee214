"""Synthetic whole-program workloads of the shapes BASELINE.json names (configs 3, 4, 5): a random genome and
spliced cDNAs cut from it, both seeded and reproducible, plus the recipe that builds a gmap_build index with the
reference's own pipeline (oracle/_ref/bin, made by oracle/Makefile from the unmodified sources).

    config 3   cDNAs 1-3 kb, 2-12 exons, 1 % error, introns 80 nt - 20 kb            (genome 100 Mb in the full size)
    config 5   cDNAs 5-20 kb, 6-40 exons, 5 % error, introns up to ~100-400 kb        (long, divergent, large introns)

Introns are planted with canonical GT...AG (or CT...AC on the minus strand) in 85 % of the cases, GC...AG in 10 %,
unmarked otherwise -- what makes Dynprog_genome_gap's splice-site scoring matter.  Errors: 60 % substitutions,
20 % deletions, 20 % insertions (the mix SURVEY.md section 8d uses).  Test infrastructure: nothing here is product code.
"""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFBIN = os.path.join(ROOT, "oracle", "_ref", "bin")
AVX2 = os.path.join(ROOT, "oracle", "_ref", "gmap.avx2")
SM100 = os.path.join(ROOT, "integration", "_build", "gmap.sm100")

CONFIGS = {
    # name: (len range, exon range, error, intron range)
    "config3": dict(length=(1000, 3000), exons=(2, 12), error=0.01, intron=(80, 20000)),
    "config5": dict(length=(5000, 20000), exons=(6, 40), error=0.05, intron=(200, 300000)),
}

_COMP = np.zeros(256, dtype=np.uint8)
for a, b in zip(b"ACGTN", b"TGCAN"):
    _COMP[a] = b


def revcomp(a):
    return _COMP[a[::-1]]


def make_genome(rng, total_bp, nchrom):
    """list of (name, uint8 array of ACGT)"""
    sizes = np.full(nchrom, total_bp // nchrom)
    return [("chr%d" % (i + 1), np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(n))].copy())
            for i, n in enumerate(sizes)]


def plant_gene(rng, chrom, cfg):
    """chooses exon coordinates on `chrom`, plants splice-site dinucleotides into the introns (modifies chrom),
    returns the spliced transcript (plus-strand sense) and its exon list"""
    L = int(rng.integers(cfg["length"][0], cfg["length"][1] + 1))
    ne = int(rng.integers(cfg["exons"][0], cfg["exons"][1] + 1))
    ne = max(1, min(ne, L // 60))
    cuts = np.sort(rng.choice(np.arange(30, L - 30), size=ne - 1, replace=False)) if ne > 1 else np.array([], dtype=int)
    # exons of at least 25 nt
    bounds = [0] + [int(c) for c in cuts] + [L]
    exlens = [bounds[i + 1] - bounds[i] for i in range(ne)]
    exlens = [max(25, x) for x in exlens]
    lo, hi = cfg["intron"]
    introns = [int(np.exp(rng.uniform(np.log(lo), np.log(hi)))) for _ in range(ne - 1)]
    span = sum(exlens) + sum(introns)
    if span + 2000 >= len(chrom):
        scale = (len(chrom) - 2000 - sum(exlens)) / max(1, sum(introns))
        introns = [max(lo, int(x * scale * 0.9)) for x in introns]
        span = sum(exlens) + sum(introns)
    start = int(rng.integers(1000, len(chrom) - span - 1000))
    exons, pos = [], start
    for i in range(ne):
        exons.append((pos, pos + exlens[i]))
        pos += exlens[i]
        if i < ne - 1:
            k = rng.random()
            a, b = pos, pos + introns[i]
            if k < 0.85:
                chrom[a:a + 2] = np.frombuffer(b"GT", dtype=np.uint8)
                chrom[b - 2:b] = np.frombuffer(b"AG", dtype=np.uint8)
            elif k < 0.95:
                chrom[a:a + 2] = np.frombuffer(b"GC", dtype=np.uint8)
                chrom[b - 2:b] = np.frombuffer(b"AG", dtype=np.uint8)
            pos = b
    return exons


def mutate(rng, s, e):
    n = len(s)
    r = rng.random(n)
    kind = rng.random(n)
    sub = r < e * 0.6
    dele = (r >= e * 0.6) & (r < e * 0.8)
    ins = (r >= e * 0.8) & (r < e)
    out = s.copy()
    out[sub] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(sub.sum()))]
    keep = ~dele
    reps = np.where(ins, 2, 1)[keep]
    out = np.repeat(out[keep], reps)
    # the duplicated base of an insertion becomes a random base
    dup = np.zeros(len(out), dtype=bool)
    idx = np.cumsum(reps) - 1
    dup[idx[reps == 2]] = True
    out[dup] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(dup.sum()))]
    return out


def write_fasta(path, records, width=60):
    with open(path, "wb") as f:
        for name, seq in records:
            f.write(b">" + name.encode() + b"\n")
            n = len(seq)
            full = (n // width) * width
            if full:
                body = seq[:full].reshape(-1, width)
                nl = np.full((body.shape[0], 1), 10, dtype=np.uint8)
                f.write(np.hstack([body, nl]).tobytes())
            if n > full:
                f.write(seq[full:].tobytes() + b"\n")


def generate(outdir, config="config3", genome_bp=4_000_000, nchrom=2, ncdna=200, seed=20241018):
    """writes genome.fa and cdna.fa into outdir; returns their paths"""
    cfg = CONFIGS[config]
    rng = np.random.default_rng(seed)
    os.makedirs(outdir, exist_ok=True)
    genome = make_genome(rng, genome_bp, nchrom)
    cdnas = []
    for k in range(ncdna):
        ci = int(rng.integers(0, nchrom))
        chrom = genome[ci][1]
        exons = plant_gene(rng, chrom, cfg)
        cdnas.append((k, ci, exons, bool(rng.random() < 0.5)))
    records = []
    for k, ci, exons, minus in cdnas:         # cut after every gene has been planted (later genes may overwrite earlier introns)
        chrom = genome[ci][1]
        t = np.concatenate([chrom[a:b] for a, b in exons])
        if minus:
            t = revcomp(t)
        records.append(("cdna%d" % k, mutate(rng, t, cfg["error"])))
    gpath, cpath = os.path.join(outdir, "genome.fa"), os.path.join(outdir, "cdna.fa")
    write_fasta(gpath, genome)
    write_fasta(cpath, records)
    return gpath, cpath


def build_index(genome_fa, dbdir, name, log=None):
    """the reference's own gmap_build (--local=0 skips the GSNAP-only localdb)"""
    os.makedirs(dbdir, exist_ok=True)
    cmd = ["perl", os.path.join(REFBIN, "gmap_build"), "--local=0", "-w", "0", "-B", REFBIN, "-D", dbdir, "-d", name, genome_fa]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if log:
        open(log, "wb").write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("gmap_build failed:\n" + r.stdout.decode()[-3000:])


def normalise(text):
    """drops the line that carries argv[0]"""
    return b"\n".join(l for l in text.split(b"\n") if not l.startswith(b"# Generated by GMAP")) + b"\n"
