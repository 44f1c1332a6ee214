import gzip
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def load_golden():
    with gzip.open(os.path.join(HERE, "golden", "dp_golden.json.gz"), "rt") as f:
        recs = json.load(f)
    out = []
    for r in recs:
        e = r.pop("expect")
        pairs = [(p[0], p[1], p[2].encode("latin1"), p[3].encode("latin1"), p[4].encode("latin1"),
                  p[5].encode("latin1")) + tuple(p[6:]) for p in e["pairs"]]
        out.append((r, (e["n"], e["iout"], e["dout"], pairs)))
    return out
