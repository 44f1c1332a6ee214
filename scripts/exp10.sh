run() { python bench.py --no-cpu-baseline --chain-problems 0 "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('boxes',d['config']['boxes_per_gpu'],'cells %.3g'%d['config']['cells_per_step'],'dev boxes',d['config']['device_boxes'],'ms %.2f'%d['ms_per_step'],'GCUPS %.1f'%d['value'],'full %.1f others'%r['kernel_ms'],['%.1f'%x for x in r['other_kernels']['ms']],'e2e ms %.2f'%d['e2e']['ms_per_step'])"; }
echo "small boxes (15-150 bp), 1M"; run --small
for m in 1 2 4 24; do echo "small mask $m"; run --small --modemask $m; done
python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-400
