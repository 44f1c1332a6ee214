for m in 1 2 4 8 16; do
  python bench.py --boxes 100000 --steps 2 --warmup 1 --no-cpu-baseline --modemask $m 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('mask',d['config']['modemask'],'boxes',d['config']['device_boxes'],'cells %.3g'%d['config']['cells_per_step'],'ms %.2f'%d['ms_per_step'],'GCUPS %.1f'%d['value'],'e2e ms %.2f'%d['e2e']['ms_per_step'])"
done
