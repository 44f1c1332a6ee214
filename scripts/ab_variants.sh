#!/bin/bash
# A/B of tuning builds (scripts/variants.sh puts them into build/variants/lib_<name>.so) on 500 k boxes:
#   bash scripts/ab_variants.sh g5 g8 ...      # the default build first, then every named variant
# (used for the occupancy, unroll and inlining knobs of the E-only kernels once the boxes of a block ran in step)
mkdir -p gpurun_out
B="--no-cpu-baseline --chain-problems 0 --program-cdnas 0 --decorated-boxes 0 --boxes 500000 --stratum-boxes 500000 --steps 2 --warmup 2"
for v in default "$@"; do
  if [ $v = default ]; then unset GMAPDP_LIB; else export GMAPDP_LIB=build/variants/lib_$v.so; fi
  timeout 300 python bench.py $B > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/var_%s.json"%v).read().strip().splitlines()[-1])
    r=d["roofline"]; p=d["strata"]["production"]
    print(v,"ms",round(d["ms_per_step"],2),"single",round(r["kernel_ms"],2),"end/genome/cdna",[round(x,2) for x in r["other_kernels"]["ms"]],"digest",d["digest"],"prod",round(p["ms_per_step"],2),{k:round(x,2) for k,x in p["kernel_ms"].items()})
except Exception as e: print(v,"unreadable",e)
PY
done
