run() { python bench.py --no-cpu-baseline --chain-problems 0 "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('boxes',d['config']['boxes_per_gpu'],'ms %.2f'%d['ms_per_step'],'GCUPS %.1f'%d['value'],'full %.1f others'%r['kernel_ms'],['%.1f'%x for x in r['other_kernels']['ms']],'e2e ms %.2f'%d['e2e']['ms_per_step'])"; }
M="smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__inst_executed.sum,gpu__time_duration.sum"
for v in default nomul; do
  if [ $v = default ]; then L=""; else L="build/variants/lib_$v.so"; fi
  echo "single only 500k, $v"; GMAPDP_LIB=$L run --modemask 1 --boxes 500000
  GMAPDP_LIB=$L ncu --metrics $M --clock-control none -k regex:gmapdp -c 1 --csv --log-file gpurun_out/m_$v.csv python bench.py --boxes 100000 --steps 1 --warmup 0 --no-cpu-baseline --chain-problems 0 --modemask 1 > /dev/null 2>&1
  grep gmapdp gpurun_out/m_$v.csv | cut -d, -f13,15 | tr '\n' ' '; echo
done
