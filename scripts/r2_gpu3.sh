#!/bin/bash
# runtime experiments
mkdir -p gpurun_out
export GMAPDP_STREAM_TIMING=1
run () {  # label, env...
  label=$1; shift
  env "$@" python scripts/gmap_throughput.py --queries 2000 --threads $THREADS $REFFLAG > gpurun_out/exp_$label.json 2> gpurun_out/exp_$label.err
  python - "$label" <<'PY'
import json, sys
label = sys.argv[1]
for l in open("gpurun_out/exp_%s.err" % label):
    try: r = json.loads(l)
    except Exception: continue
    lanes = [s.split("lane ")[1] for s in r["stats"] if " lane " in s]
    rt = [s for s in r["stats"] if "runtime:" in s and "stage 2" not in s]
    print(label, "t=%d" % r["threads"], r["cdnas_per_s"], r["identical_output"], "cpu", r.get("cpu_user_sys_s"), "|", lanes[0] if lanes else "")
    if rt: print("      ", rt[0].split("largest")[1][:120])
PY
}
REFFLAG=""
THREADS=16,48,96,192,384
run base A=1
grep -o '"reference": {[^}]*}' gpurun_out/exp_base.json
REFFLAG="--skip-ref"
THREADS=96,192
run spin GMAPDP_STREAM_SPIN_US=300
run groups1 GMAP_SM100_CHAIN_GROUPS=32
