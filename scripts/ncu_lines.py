"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[2]
iL, iA = hdr.index("Line No"), hdr.index("Address")
iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
src = {}
agg = collections.defaultdict(lambda: [0, 0, 0])
cur = None
for r in rows[3:]:
    if len(r) <= iI:
        continue
    if r[iL]:
        try:
            cur = int(r[iL]); src[cur] = r[iL + 1]
        except ValueError:
            cur = None
            continue
    if r[iA] and cur is not None:
        try:
            agg[cur][0] += int(r[iS]); agg[cur][1] += int(r[iI]); agg[cur][2] += 1
        except ValueError:
            pass
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
print("total samples %d, warp instructions %d" % (ts, ti))
top = sorted(agg.items(), key=lambda kv: -kv[1][0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for ln, (s, i, n) in sorted(top):
    print("%5d  samp %5.1f%%  inst %5.1f%%  sass %3d | %s" % (ln, 100.0 * s / ts, 100.0 * i / ti, n, src.get(ln, "").strip()[:110]))
