"""Whole-program cases beyond the bundled `-g` runs: BASELINE config 1 through a gmap_build index, scaled-down
configs 3 and 5 (synthetic genome + spliced cDNAs, tests/progdata.py), and user indel penalties.  Shared by the
digest generator (tests/golden/make_program_digests.py), the GPU test and bench.py's whole-program leg."""
import hashlib
import os
import subprocess

import progdata

ROOT = progdata.ROOT
DATA = os.path.join(ROOT, "tests", "data")
AVX2, SM100 = progdata.AVX2, progdata.SM100

# name: (kind, generator arguments, gmap arguments after the database / genome selection)
CASES = {
    "cfg1_index_A": ("index", dict(genome=os.path.join(DATA, "genetest2.fa"), cdna=os.path.join(DATA, "cdna2.fa")), ["-A"]),
    "cfg1_index_gff3": ("index", dict(genome=os.path.join(DATA, "genetest2.fa"), cdna=os.path.join(DATA, "cdna2.fa")), ["-f", "gff3_gene"]),
    "cfg3_A": ("synthetic", dict(config="config3", genome_bp=4_000_000, nchrom=2, ncdna=200), ["-A"]),
    "cfg3_gff3": ("synthetic", dict(config="config3", genome_bp=4_000_000, nchrom=2, ncdna=200), ["-f", "gff3_gene"]),
    "cfg5_A": ("synthetic", dict(config="config5", genome_bp=8_000_000, nchrom=2, ncdna=40), ["-A"]),
    "cfg5_gff3": ("synthetic", dict(config="config5", genome_bp=8_000_000, nchrom=2, ncdna=40), ["-f", "gff3_gene"]),
    "her2_mut_userpen_A": ("dashg", dict(genome=os.path.join(DATA, "ss.chr17test"), cdna=os.path.join(DATA, "her2_mutated.fa")),
                           ["--indel-open=-12", "--indel-extend=-4", "-A"]),
}

_prepared = {}


def prepare(name, workdir):
    """generates inputs / builds the index once per (workdir, data set); returns the argument list"""
    kind, gen, args = CASES[name]
    if kind == "dashg":
        return ["-g", gen["genome"]] + args + [gen["cdna"]]
    key = (workdir, kind, tuple(sorted((k, str(v)) for k, v in gen.items())))
    if key not in _prepared:
        d = os.path.join(workdir, "set%d" % len(_prepared))
        if kind == "index":
            genome, cdna = gen["genome"], gen["cdna"]
        else:
            genome, cdna = progdata.generate(d, **gen)
        progdata.build_index(genome, os.path.join(d, "db"), "syn", os.path.join(d, "build.log") if os.path.isdir(d) else None)
        _prepared[key] = (os.path.join(d, "db"), cdna)
    db, cdna = _prepared[key]
    return ["-D", db, "-d", "syn"] + args + [cdna]


def run(exe, case_args, threads=4, env=None, timeout=1800):
    r = subprocess.run([exe, "-t", str(threads), "-O"] + case_args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("%s failed (%d): %s" % (exe, r.returncode, r.stderr.decode()[-2000:]))
    return progdata.normalise(r.stdout), r.stderr.decode()


def digest(text):
    return hashlib.sha256(text).hexdigest()
