# usage: prof_mode.sh <modemask> <boxes> <outname> <kernel regex>
CMD="python bench.py --boxes $2 --steps 1 --warmup 0 --no-cpu-baseline --modemask $1"
$CMD > gpurun_out/plain_$3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$4 -c 1 -o gpurun_out/$3 $CMD > gpurun_out/ncu_$3.log 2>&1
tail -2 gpurun_out/ncu_$3.log
