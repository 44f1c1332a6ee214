"""Digests of the STOCK reference program's output (oracle/_ref/gmap.avx2, compiled from the unmodified
/root/reference/src) on the whole-program cases of tests/program_cases.py: BASELINE config 1 through a gmap_build
index, scaled configs 3 and 5, user indel penalties.  Build-container only; writes tests/golden/program_digests.json."""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import program_cases  # noqa: E402


def main():
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for name in program_cases.CASES:
            case = program_cases.prepare(name, td)
            text, _ = program_cases.run(program_cases.AVX2, case, threads=4)
            out[name] = program_cases.digest(text)
            print(name, out[name][:16], len(text), "bytes")
    json.dump(out, open(os.path.join(HERE, "program_digests.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
