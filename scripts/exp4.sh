run() { python bench.py --no-cpu-baseline --chain-problems 0 "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('boxes',d['config']['boxes_per_gpu'],'ms %.2f'%d['ms_per_step'],'GCUPS %.1f'%d['value'],'full %.1f others'%r['kernel_ms'],['%.1f'%x for x in r['other_kernels']['ms']],'e2e ms %.2f'%d['e2e']['ms_per_step'])"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in default u1n u2n u1i; do
  echo "end only, $v"; if [ $v = default ]; then run --modemask 24 --boxes 500000; else GMAPDP_LIB=build/variants/lib_$v.so run --modemask 24 --boxes 500000; fi
done
for mb in 96 384; do echo "chunk $mb"; GMAPDP_CHUNK_MB=$mb run; done
