"""Whole-program parity on the configurations beyond the bundled `-g` runs (BASELINE.json configs 1, 3, 5):
gmap.sm100 (the unmodified reference program with its DP entry points and stage-2 chaining drivers served by the
B200 engines) must print the same bytes as the stock gmap.avx2 -- both run here, on the GPU box -- with `-A` and
`-f gff3_gene`, through gmap_build indexes made by the reference's own pipeline (oracle/_ref/bin).  The stock
program's output is also checked against the digests committed from the build container
(tests/golden/program_digests.json, make_program_digests.py), so "the same bytes" means the bytes seen there."""
import json
import os
import re

import pytest

import program_cases

pytestmark = pytest.mark.gpu

HAVE = all(os.path.exists(p) for p in (program_cases.AVX2, program_cases.SM100, os.path.join(program_cases.progdata.REFBIN, "gmapindex")))
DIGESTS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "program_digests.json")))


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("progcases"))


def device_stats(err):
    calls = re.search(r"gmap.sm100: (\d+) DP calls in (\d+) device batches", err)
    s2 = re.search(r"stage 2: (\d+) chaining calls on the device in (\d+) batches, (\d+) handed to the reference body", err)
    return calls, s2


@pytest.mark.skipif(not HAVE, reason="needs oracle/_ref (gmap.avx2, gmap_build tools) and integration/_build/gmap.sm100")
@pytest.mark.parametrize("name", sorted(program_cases.CASES))
def test_configs_byte_identical(workdir, name):
    case = program_cases.prepare(name, workdir)
    want, _ = program_cases.run(program_cases.AVX2, case, threads=8)
    assert program_cases.digest(want) == DIGESTS[name], "the stock program prints something else here than in the build container"
    env = dict(os.environ, GMAP_SM100_STATS="1")
    got, err = program_cases.run(program_cases.SM100, case, threads=32, env=env)
    assert got == want, "gmap.sm100 output differs from stock gmap.avx2 for " + name
    calls, s2 = device_stats(err)
    assert calls and int(calls.group(1)) > 0, err[-1500:]
    assert s2 and int(s2.group(1)) > 0 and int(s2.group(3)) == 0, "stage 2: no device calls, or calls handed to the reference body: %r" % (s2 and s2.group(0))


@pytest.mark.skipif(not HAVE, reason="needs oracle/_ref and integration/_build/gmap.sm100")
def test_unsupported_options_are_refused_loudly(workdir):
    """no silent CPU fallback, no silently different alignments: options the device path does not implement stop the program"""
    import subprocess
    data = program_cases.DATA
    for flag, what in (("--homopolymer", "homopolymer"), ("--cross-species", "cross-species"), ("--mode=cmet-stranded", "mode")):
        r = subprocess.run([program_cases.SM100, flag, "-A", "-g", os.path.join(data, "ss.chr17test"), os.path.join(data, "ss.her2")],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
        assert r.returncode != 0, flag
        assert "gmap.sm100:" in r.stderr.decode() and what in r.stderr.decode(), r.stderr.decode()[-500:]
