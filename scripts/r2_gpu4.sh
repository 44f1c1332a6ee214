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
        if "lane 0" in s or "start-up" in s or "runtime:" in s: print("    ", s[:360])
PY
}
python scripts/gmap_throughput.py --queries 2000 --threads 1,32,192 --skip-ref > gpurun_out/exp_a.json 2> gpurun_out/exp_a.err; show gpurun_out/exp_a.err
