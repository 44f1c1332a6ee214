# ncu --set full captures of the E-only kernels per mode (genome / cdna / end) on 100k-index slices
for spec in "2 genome" "4 cdna"; do
  set -- $spec
  CMD="python bench.py --boxes 100000 --steps 1 --warmup 0 --no-cpu-baseline --modemask $1 --chain-problems 0"
  $CMD > gpurun_out/plain_$2.log 2>&1 || exit 1
  ncu --set full --clock-control none --import-source on -k regex:gmapdp -c 1 -o gpurun_out/prof_$2 $CMD > gpurun_out/ncu_$2.log 2>&1
done
ls -la gpurun_out/
