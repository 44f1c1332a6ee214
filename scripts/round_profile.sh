# Round profile: bench line, ncu launch list of the SAME command, DRAM bytes of one full-size launch of the dominant
# kernel (gmapdp_dp_kernel<0>, the single-gap full fills), and --set full captures on 100k-box batches
# (B200_PROFILING.md recipe: every ncu pass only after the same command has exited 0 without ncu).
set -x
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -c 300 gpurun_out/bench_full.err
CMD="python bench.py --no-cpu-baseline --chain-problems 0 --program-cdnas 0 --stratum-boxes 0 --decorated-boxes 0"
$CMD > gpurun_out/plain_launch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
CMD1="python bench.py --no-cpu-baseline --chain-problems 0 --program-cdnas 0 --stratum-boxes 0 --decorated-boxes 0 --steps 1 --warmup 0"
$CMD1 > gpurun_out/plain_dram.log 2>&1 && ncu --kernel-name-base mangled -k regex:ILi0E --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 1 --csv --log-file gpurun_out/dram_1m.csv $CMD1 > gpurun_out/ncu_dram.log 2>&1
for spec in "1 full" "2 genome" "4 cdna" "24 end"; do
  set -- $spec
  CMD2="python bench.py --boxes 100000 --steps 1 --warmup 0 --no-cpu-baseline --chain-problems 0 --program-cdnas 0 --stratum-boxes 0 --decorated-boxes 0 --modemask $1"
  $CMD2 > gpurun_out/plain_$2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gmapdp_dp_kernel -c 1 -o gpurun_out/prof_$2 $CMD2 > gpurun_out/ncu_$2.log 2>&1
done
# the MaxEnt kernel (resident genome) of the genome-only batch
CMD3="python bench.py --boxes 100000 --steps 1 --warmup 0 --no-cpu-baseline --chain-problems 0 --program-cdnas 0 --stratum-boxes 0 --decorated-boxes 0 --modemask 2"
ncu --set full --clock-control none --import-source on -k regex:maxent_kernel_batch -c 1 -o gpurun_out/prof_mepass $CMD3 > gpurun_out/ncu_mepass.log 2>&1
# production-size stratum (15-150 bp boxes): the genome and end kernels again, now that the boxes of a block run in step
for spec in "2 genome" "24 end"; do
  set -- $spec
  CMD4="python bench.py --small --boxes 300000 --steps 1 --warmup 0 --no-cpu-baseline --chain-problems 0 --program-cdnas 0 --stratum-boxes 0 --decorated-boxes 0 --modemask $1"
  $CMD4 > gpurun_out/plain_small2_$2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gmapdp_dp_kernel -c 1 -o gpurun_out/prof_small2_$2 $CMD4 > gpurun_out/ncu_small2_$2.log 2>&1
done
# the overflow path of the bridge's tie lists: the same parity tests against a build with one-entry lists
python -m pytest tests/test_gpu_parity.py -x -q -k tie_list_overflow 2>&1 | tail -3 > gpurun_out/tiecap1_tests.log
cat gpurun_out/tiecap1_tests.log
ls -la gpurun_out
