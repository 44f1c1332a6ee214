#!/bin/bash
# full GPU test suite, then the short bench (no CPU legs)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/ck_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/ck_tests.log
B="--no-cpu-baseline --chain-problems 0 --program-cdnas 0"
timeout 600 python bench.py $B > gpurun_out/ck_bench.json 2> gpurun_out/ck_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("ck_bench",):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        r=d["roofline"]; p=d["strata"]["production"]
        print(f,"ms",round(d["ms_per_step"],2),"GCUPS",round(d["value"],1),"single",round(r["kernel_ms"],2),"end/genome/cdna",[round(x,2) for x in r["other_kernels"]["ms"]],"e2e",round(d["e2e"]["ms_per_step"],1),"digest",d["digest"],"prod",round(p["ms_per_step"],2),{k:round(v,2) for k,v in p["kernel_ms"].items()},"prod e2e",round(p["e2e"]["ms_per_step"],1), "decorated", d["strata"]["decorated"]["digest"], round(d["strata"]["decorated"]["value"],1))
    except Exception as e: print(f,"unreadable",e)
PY
