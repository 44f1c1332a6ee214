#!/bin/bash
# what bounds the end-to-end step: record stores (NT / ordinary) and chunk granularity, 200 k boxes
mkdir -p gpurun_out
B="--no-cpu-baseline --chain-problems 0 --program-cdnas 0 --decorated-boxes 0 --stratum-boxes 0 --boxes 200000 --steps 2 --warmup 1"
run () { v=$1; shift
  env "$@" timeout 200 python bench.py $B > gpurun_out/e2e_$v.json 2> gpurun_out/e2e_$v.err
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/e2e_%s.json"%v).read().strip().splitlines()[-1])
    print(v,"dev ms",round(d["ms_per_step"],2),"e2e_device",round(d["e2e_device"]["ms_per_step"],1),"e2e",round(d["e2e"]["ms_per_step"],1))
except Exception as e: print(v,"unreadable",e)
PY
}
run default X=1
run geometric GMAPDP_EQUAL_CHUNKS=0
run plainstores GMAPDP_NT_STORES=0
run geo_plain GMAPDP_EQUAL_CHUNKS=0 GMAPDP_NT_STORES=0
