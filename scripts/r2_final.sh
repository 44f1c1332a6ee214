#!/bin/bash
# last session of round 2: the whole GPU suite, then the round profile (default bench line + ncu captures of the same commands)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
bash scripts/round_profile.sh > gpurun_out/round_profile.log 2>&1; echo "profile rc=$?"
tail -c 600 gpurun_out/bench_full.json
