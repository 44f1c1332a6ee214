import sys, collections
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import dpgen
from harness import Oracle
from gmap_2024_b200 import Engine
mode = sys.argv[1]; seed = int(sys.argv[2]); n = int(sys.argv[3]); rmin = int(sys.argv[4]); rmax = int(sys.argv[5])
e = Engine(0); o = Oracle()
boxes = dpgen.synth_boxes(seed=seed, n=n, mode=None if mode == 'all' else mode, rmin=rmin, rmax=rmax)
b = e.batch(); ids = [b.add(x) for x in boxes]; b.run()
st = collections.Counter(); shown = 0
for x, cid in zip(boxes, ids):
    got = b.result(cid, x['mode']); want = o.run(x)
    dr = b.device_result(cid)
    key = x['mode'] + ('_dev' if dr is not None else '_host')
    st[key] += 1
    if got != want:
        st[key + '_BAD'] += 1
        if shown < 12:
            shown += 1
            desc = {k: v for k, v in x.items() if k not in ("queryseq", "left_probs", "right_probs", "world") and "gseg" not in k}
            print("BAD", desc)
            if dr is not None:
                print("   dev: status", dr.status, "final", dr.finalscore, "L", dr.bestrL, dr.bestcL, "R", dr.bestrR, dr.bestcR, "tb", dr.tb_score, dr.nmatches, dr.nmismatches, dr.nopens, dr.nindels, "script", dr.script_lenA, dr.script_lenB)
            print("   got ", got[0], got[1], got[2]); print("   want", want[0], want[1], want[2])
            if got[3] and want[3]: print("   first pairs", got[3][0], want[3][0], "last", got[3][-1], want[3][-1])
print(sorted(st.items()))
