#!/bin/bash
# r02 call 23: the two test files written / changed after the final evidence run; the collapse pass of a radius that
# does NOT store the Hessian (the counter pass and the full-set capture of call 22 caught the last radius)
mkdir -p gpurun_out; O=gpurun_out
timeout 400 python -m pytest tests/test_zgpu_10_scaledep_gm.py tests/test_zgpu_11_loader_variants.py -m gpu -q -rA > $O/r02_final_pytest_gpu_rerun.log 2>&1
echo "pytest rc=$?" >> $O/r02_final_pytest_gpu_rerun.log; grep -E "passed|failed|^FAILED|^ERROR|rc=|t_b200" $O/r02_final_pytest_gpu_rerun.log | tail -8
M=gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers
timeout 200 ncu --metrics $M --clock-control none -k regex:zpass_collapse -s 0 -c 1 --csv --log-file $O/r02_final_ncu_counters_zcollapse.csv \
  python scripts/prof_step.py 1024 classic 2 1 0 > $O/r02_final_ncu_counters_zcollapse.log 2>&1
python tools/ncu_to_traffic.py $O/r02_final_ncu_counters_zcollapse.csv 1024 1 $O/r02_final_traffic_zcollapse.json | head -30
timeout 240 ncu --set full --clock-control none --import-source on -k regex:zpass_collapse -s 0 -c 1 -f -o $O/r02_final_ncu_zc \
  python scripts/prof_step.py 1024 classic 2 1 0 > $O/r02_final_ncu_zc.log 2>&1
timeout 100 python tools/ncu_summary.py $O/r02_final_ncu_zc.ncu-rep $O/r02_final_ncu_full_zpass_collapse.csv > /dev/null 2>&1
rm -f $O/r02_final_ncu_zc.ncu-rep; head -12 $O/r02_final_ncu_full_zpass_collapse.csv
