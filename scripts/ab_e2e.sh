#!/bin/bash
# A/B of the end-to-end step (GmapDP_batch_run incl. pair lists) on 200 k boxes under run-time switches: chunk granularity
# (GMAPDP_EQUAL_CHUNKS) and replay threads (GMAPDP_REPLAY_THREADS): does the thread that hands the chunks over need a core?
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
run geo16 X=1
run eq15 GMAPDP_EQUAL_CHUNKS=1 GMAPDP_REPLAY_THREADS=15
run geo15 GMAPDP_REPLAY_THREADS=15
run eq14 GMAPDP_EQUAL_CHUNKS=1 GMAPDP_REPLAY_THREADS=14
