# ncu evidence for the chaining kernel (B200_PROFILING.md recipe): plain run first, then the launch list and one
# --set full capture of gmapchain_kernel on the same command.
set -x
CMD="python scripts/chain_speed.py 128 64"
$CMD > gpurun_out/chain_plain.log 2>&1 || exit 1
tail -4 gpurun_out/chain_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gmapchain -c 12 --csv --log-file gpurun_out/chain_launches.csv $CMD > gpurun_out/chain_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gmapchain -s 2 -c 1 -o gpurun_out/prof_chain $CMD > gpurun_out/chain_ncu_full.log 2>&1
ls -la gpurun_out | tail -8
