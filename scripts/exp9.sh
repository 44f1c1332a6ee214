# two ranks on two GPUs of one box, launched as the driver does
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --boxes 400000 --chain-problems 4096 2>gpurun_out/mgpu.err | tail -1 > gpurun_out/mgpu2.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/mgpu2.json').read())
print('n_gpus',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'],'scaling',d['scaling'])
PY
tail -3 gpurun_out/mgpu.err
