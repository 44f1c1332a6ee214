run() { python bench.py --no-cpu-baseline --chain-problems 0 "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('boxes',d['config']['boxes_per_gpu'],'ms %.2f'%d['ms_per_step'],'GCUPS %.1f'%d['value'],'full %.1f others'%r['kernel_ms'],['%.1f'%x for x in r['other_kernels']['ms']],'e2e ms %.2f'%d['e2e']['ms_per_step'])"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in default bndsmem full5 full6; do
  if [ $v = default ]; then L=""; else L="build/variants/lib_$v.so"; fi
  echo "single only 500k, $v"; GMAPDP_LIB=$L run --modemask 1 --boxes 500000
done
echo "default 1M"; run
