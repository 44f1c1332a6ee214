# host-side timeline of the device batches of the whole program (GMAPDP_TRACE): what one rendezvous batch costs
cd tests/data
GMAPDP_TRACE=1 ../../integration/_build/gmap.sm100 -t 64 -O -A -g ss.chr17test her2_mutated.fa > /dev/null 2> ../../gpurun_out/trace_program.err
cd ../..
grep -c "box scan" gpurun_out/trace_program.err
python - <<'PY'
import re,collections
laps=collections.defaultdict(list)
prev=None
for l in open('gpurun_out/trace_program.err'):
    m=re.match(r'gmapdp_run_batch: (.+?)\s+([\d.]+) ms',l)
    if not m: continue
    laps[m.group(1).strip()].append(float(m.group(2)))
for k,v in laps.items():
    v2=sorted(v); print('%-24s n=%d median %.3f ms  p90 %.3f  mean %.3f'%(k,len(v),v2[len(v2)//2],v2[int(len(v2)*0.9)],sum(v)/len(v)))
PY
grep "^gmap.sm100" gpurun_out/trace_program.err | tail -4
