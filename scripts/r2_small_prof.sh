#!/bin/bash
# ncu captures of the E-only kernels on the production-size stratum (15-150 bp boxes)
mkdir -p gpurun_out
for spec in "2 genome" "4 cdna" "24 end" "1 single"; do
  set -- $spec
  CMD="python bench.py --small --boxes 300000 --steps 1 --warmup 0 --no-cpu-baseline --chain-problems 0 --program-cdnas 0 --stratum-boxes 0 --decorated-boxes 0 --modemask $1"
  $CMD > gpurun_out/plain_small_$2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gmapdp_dp_kernel -c 1 -o gpurun_out/prof_small_$2 $CMD > gpurun_out/ncu_small_$2.log 2>&1
  tail -1 gpurun_out/ncu_small_$2.log
done
ls -la gpurun_out/prof_small_*
