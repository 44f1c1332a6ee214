#!/bin/bash
# in-step refinements at 500 k boxes: variant sub-key of the launch order on / off, barriers in front of the passes
mkdir -p gpurun_out
B="--no-cpu-baseline --chain-problems 0 --program-cdnas 0 --decorated-boxes 0 --boxes 500000 --stratum-boxes 500000 --steps 2 --warmup 2"
run () { v=$1; shift
  env "$@" timeout 300 python bench.py $B > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/var_%s.json"%v).read().strip().splitlines()[-1])
    r=d["roofline"]; p=d["strata"]["production"]
    print(v,"ms",round(d["ms_per_step"],2),"single",round(r["kernel_ms"],2),"end/genome/cdna",[round(x,2) for x in r["other_kernels"]["ms"]],"digest",d["digest"],"prod",round(p["ms_per_step"],2),{k:round(x,2) for k,x in p["kernel_ms"].items()})
except Exception as e: print(v,"unreadable",e)
PY
}
run sortvar GMAPDP_SORT_VARIANT=1
run nosortvar GMAPDP_SORT_VARIANT=0
run passsync GMAPDP_LIB=build/variants/lib_passsync.so
