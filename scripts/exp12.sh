python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run() { python bench.py --no-cpu-baseline --chain-problems 0 "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('boxes',d['config']['boxes_per_gpu'],'ms %.2f'%d['ms_per_step'],'GCUPS %.1f'%d['value'],'full %.1f others'%r['kernel_ms'],['%.1f'%x for x in r['other_kernels']['ms']],'e2e ms %.2f'%d['e2e']['ms_per_step'], 'digest', d['digest'])"; }
echo "small"; run --small
echo "large"; run
echo "== small trace"; GMAPDP_TRACE=1 python bench.py --no-cpu-baseline --chain-problems 0 --steps 1 --warmup 1 --small 2>&1 | grep "gmapdp_run_batch" | tail -7
