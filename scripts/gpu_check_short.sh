#!/bin/bash
# parity tests that read pair lists back, then the short bench (no CPU legs)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_workload.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/cs_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/cs_tests.log
timeout 300 python bench.py --no-cpu-baseline --chain-problems 0 --program-cdnas 0 > gpurun_out/cs_bench.json 2> gpurun_out/cs_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/cs_bench.json").read().strip().splitlines()[-1])
    r=d["roofline"]; p=d["strata"]["production"]
    print("ms",round(d["ms_per_step"],2),"GCUPS",round(d["value"],1),"single",round(r["kernel_ms"],2),"e2e_device",round(d["e2e_device"]["ms_per_step"],1),"e2e",round(d["e2e"]["ms_per_step"],1),round(d["e2e"]["value"],1),"digest",d["digest"],"prod",round(p["ms_per_step"],2),"prod e2e",round(p["e2e"]["ms_per_step"],1),"decorated e2e",round(d["strata"]["decorated"]["e2e"]["ms_per_step"],1),d["strata"]["decorated"]["digest"])
except Exception as e: print("unreadable",e)
PY
