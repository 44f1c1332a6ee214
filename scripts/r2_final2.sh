#!/bin/bash
# Round profile whose output fits gpurun's 64 MiB: the default bench line, the ncu launch list of the same command, DRAM
# bytes of one full-size launch of the dominant kernel, --set full captures of the five kernels on 100 k-box batches
# (B200_PROFILING.md recipe: the ncu passes follow plain runs of the same program that exited 0 -- the default bench and
# the launch-list command just before them), summarised on the box
# by scripts/collect_profiles.py; the reports themselves are deleted.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
CMD="python bench.py --no-cpu-baseline --chain-problems 0 --program-cdnas 0 --stratum-boxes 0 --decorated-boxes 0"
$CMD > gpurun_out/plain_launch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 32 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
CMD1="python bench.py --no-cpu-baseline --chain-problems 0 --program-cdnas 0 --stratum-boxes 0 --decorated-boxes 0 --steps 1 --warmup 0"
ncu --kernel-name-base mangled -k regex:ILi0E --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 1 --csv --log-file gpurun_out/dram_1m.csv $CMD1 > gpurun_out/ncu_dram.log 2>&1
for spec in "1 full" "2 genome" "4 cdna" "24 end"; do
  set -- $spec
  CMD2="python bench.py --boxes 100000 --steps 1 --warmup 0 --no-cpu-baseline --chain-problems 0 --program-cdnas 0 --stratum-boxes 0 --decorated-boxes 0 --modemask $1"
  ncu --set full --clock-control none --import-source on -k regex:gmapdp_dp_kernel -c 1 -o gpurun_out/prof_$2 $CMD2 > gpurun_out/ncu_$2.log 2>&1
done
COLLECT_OUT=gpurun_out/profiles_out python scripts/collect_profiles.py r02s4 2>&1 | tail -2
rm -f gpurun_out/*.ncu-rep gpurun_out/*_src.csv
du -sh gpurun_out; ls gpurun_out/profiles_out | head -30
tail -c 400 gpurun_out/bench_full.json
