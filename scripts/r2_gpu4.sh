#!/bin/bash
mkdir -p gpurun_out
export GMAPDP_STREAM_TIMING=1
show () {
python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1], errors="replace"):
    try: r = json.loads(l)
    except Exception:
        if "failed" in l or "gmap.sm100:" in l: print(l[:300])
        continue
    print(sys.argv[1][-12:], "t=%d" % r["threads"], r["cdnas_per_s"], r["identical_output"], "cpu", r.get("cpu_user_sys_s"))
    for s in r["stats"]:
        if "lane 0" in s or "stage 2 runtime" in s: print("    ", s[:360])
PY
}
run () { label=$1; shift; env "$@" python scripts/gmap_throughput.py --queries 2000 --threads 96,192 --skip-ref > gpurun_out/exp_$label.json 2> gpurun_out/exp_$label.err; show gpurun_out/exp_$label.err; }
run A CUDA_DEVICE_MAX_CONNECTIONS=32
run B CUDA_DEVICE_MAX_CONNECTIONS=32 GMAPDP_STREAM_DEPTH=4
run C GMAPDP_STREAM_DEPTH=1
run D CUDA_DEVICE_MAX_CONNECTIONS=32 GMAP_SM100_CHAIN_GROUPS=4 GMAPDP_STREAM_DEPTH=4
run E CUDA_DEVICE_MAX_CONNECTIONS=32 GMAPDP_STREAM_DEPTH=2
