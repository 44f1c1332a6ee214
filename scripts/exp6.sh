run() { python bench.py --no-cpu-baseline --chain-problems 0 "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('boxes',d['config']['boxes_per_gpu'],'ms %.2f'%d['ms_per_step'],'GCUPS %.1f'%d['value'],'full %.1f others'%r['kernel_ms'],['%.1f'%x for x in r['other_kernels']['ms']],'e2e ms %.2f'%d['e2e']['ms_per_step'])"; }
for v in default tri6 tri8; do
  echo "E-only kinds, $v"; if [ $v = default ]; then run --modemask 30 --boxes 500000; else GMAPDP_LIB=build/variants/lib_$v.so run --modemask 30 --boxes 500000; fi
done
GMAPDP_TRACE=1 python bench.py --no-cpu-baseline --chain-problems 0 --steps 1 --warmup 1 2>&1 | grep "gmapdp_run_batch" | tail -8
