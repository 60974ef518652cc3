#!/bin/bash
# r02 profiling call 1: what bounds the collapse z pass?  (a) passbench variants: classic / no epilogue, y pass
# with 64-byte tiles and 2-3 resident blocks; (b) the real pipeline at 1024^3, classic vs tabulated epilogue, live
# timers and an ncu counter pass (light metric list) of each.
mkdir -p gpurun_out; O=gpurun_out
for v in base noepi; do echo "== passbench_$v z"; timeout 120 ./tools/passbench_$v z; done 2>&1 | tee $O/r02_passbench_z.txt
for v in base ytk4m2 ytk4m3; do echo "== passbench_$v y"; timeout 120 ./tools/passbench_$v y; done 2>&1 | grep -v timeline | tee $O/r02_passbench_y.txt
timeout 200 python scripts/prof_step.py 1024 classic 2 2 0 | tee $O/r02_prof_live.txt
timeout 200 python scripts/prof_step.py 1024 tab 2 2 0 | tee -a $O/r02_prof_live.txt
M=gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio,smsp__average_warps_issue_stalled_drain_per_issue_active.ratio,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers
for mode in classic tab; do
  timeout 400 ncu --metrics $M --clock-control none -k regex:"zpass_collapse|xpass_kernel|ypass_kernel" -s 3 -c 3 --csv --log-file $O/r02_ncu_counters_$mode.csv \
    python scripts/prof_step.py 1024 $mode 2 1 0 > $O/r02_ncu_$mode.log 2>&1
  tail -1 $O/r02_ncu_$mode.log
done
ls -la $O
