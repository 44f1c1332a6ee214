"""Prints the SASS instructions of an `ncu --page source --csv --print-source cuda,sass` dump whose execution
count is at least argv[2] (fraction) of the maximum, in address order, with the CUDA line they belong to."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[2]
iL, iA, iSass = hdr.index("Line No"), hdr.index("Address"), 3
iS, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
ins = []
cur = None
for r in rows[3:]:
    if len(r) <= iI: continue
    if r[iL]:
        try: cur = int(r[iL])
        except ValueError: cur = None
        continue
    if r[iA].startswith("0x"):
        ins.append((int(r[iA], 16), cur, r[iSass].strip(), int(r[iS]), int(r[iI]), r[iT]))
ins.sort()
mx = max(i[4] for i in ins)
tot = sum(i[4] for i in ins)
sel = [i for i in ins if i[4] >= frac * mx]
print("max exec %d, total %d, selected %d instr carrying %.1f%% of executed" % (mx, tot, len(sel), 100.0 * sum(i[4] for i in sel) / tot))
base = ins[0][0]
for a, ln, s, smp, ex, thr in sel:
    print("%6x L%-4s ex %5.2f smp %5d thr %2s | %s" % (a - base, ln, ex / mx, smp, thr, s[:90]))
