#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stream.py tests/test_gpu_gmap_program.py -x -q > gpurun_out/r2_tests2.log 2>&1; tail -3 gpurun_out/r2_tests2.log
export GMAPDP_STREAM_TIMING=1
timeout 600 python scripts/gmap_throughput.py --queries 2000 --threads 16,128,512 > gpurun_out/r2_throughput2.json 2> gpurun_out/r2_throughput2.err
python - <<'PY'
import json
for l in open("gpurun_out/r2_throughput2.err"):
    try: r = json.loads(l)
    except Exception: continue
    print(r["threads"], r["cdnas_per_s"], r["identical_output"])
    for s in r["stats"]: print("   ", s)
PY
tail -c 300 gpurun_out/r2_throughput2.json
