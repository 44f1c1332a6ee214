# ncu --set full capture of the E-only kernel (genome / cdna / end modes) on a 100k-index slice
set -x
CMD="python bench.py --boxes 100000 --steps 1 --warmup 0 --no-cpu-baseline --modemask 30 --chain-problems 0"
$CMD > gpurun_out/tri_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gmapdp -c 1 -o gpurun_out/prof_tri $CMD > gpurun_out/tri_ncu.log 2>&1
ls -la gpurun_out/prof_tri.ncu-rep
