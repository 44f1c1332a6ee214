# Round profile: bench line, ncu launch list of the SAME command, DRAM bytes of one full-size launch,
# and one --set full capture of the DP kernel on a 100k-box batch (B200_PROFILING.md recipe).
set -x
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -c 300 gpurun_out/bench_full.err
CMD="python bench.py --no-cpu-baseline"
$CMD > gpurun_out/plain_launch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 50 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
CMD1="python bench.py --no-cpu-baseline --steps 1 --warmup 0"
$CMD1 > gpurun_out/plain_dram.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 1 --csv --log-file gpurun_out/dram_1m.csv $CMD1 > gpurun_out/ncu_dram.log 2>&1
CMD2="python bench.py --boxes 100000 --steps 1 --warmup 0 --no-cpu-baseline"
$CMD2 > gpurun_out/plain_full.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gmapdp -c 1 -o gpurun_out/prof_full $CMD2 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
