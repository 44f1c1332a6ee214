#!/bin/bash
# staged interior steps (GMAPDP_STAGE): parity tests, then the bench against the same sources built with GMAPDP_STAGE=0
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/st_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/st_tests.log
B="--no-cpu-baseline --chain-problems 0 --program-cdnas 0 --decorated-boxes 0"
timeout 600 python bench.py $B > gpurun_out/st_bench.json 2> gpurun_out/st_bench.err; echo "bench rc=$?"
GMAPDP_LIB=build/variants/lib_nostage.so timeout 600 python bench.py $B > gpurun_out/st_bench_nostage.json 2> gpurun_out/st_bench_nostage.err; echo "bench nostage rc=$?"
python - <<'PY'
import json
for f in ("st_bench","st_bench_nostage"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        r=d["roofline"]; p=d["strata"]["production"]
        print(f,"ms",round(d["ms_per_step"],2),"single",round(r["kernel_ms"],2),"end/genome/cdna",[round(x,2) for x in r["other_kernels"]["ms"]],"e2e",round(d["e2e"]["ms_per_step"],1),"digest",d["digest"],"prod",round(p["ms_per_step"],2),p["kernel_ms"])
    except Exception as e: print(f,"unreadable",e)
PY
