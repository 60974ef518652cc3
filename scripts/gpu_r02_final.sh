#!/bin/bash
# Final GPU call of round 2 (1 x B200): the evidence the judge reads, from the final tree, each step under its own
# timeout.  Outputs: gpurun_out/r02_final_*; the summaries are copied to profiles/ afterwards.
mkdir -p gpurun_out; O=gpurun_out
# 1. the whole -m gpu suite as the driver runs it, but without -x: every failure is listed
timeout 900 python -m pytest tests -m gpu -q -rA --durations=10 > $O/r02_final_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r02_final_pytest_gpu.log
grep -E "passed|failed|^FAILED|^ERROR|rc=" $O/r02_final_pytest_gpu.log | tail -8
# 2. smoke() exactly as the driver calls it
timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' > $O/r02_final_smoke.log 2>&1
echo "smoke rc=$?" >> $O/r02_final_smoke.log; tail -2 $O/r02_final_smoke.log
# 3. both bench arms as the driver runs them
timeout 900 python bench.py > $O/r02_final_bench_1gpu.json 2> $O/r02_final_bench_1gpu.err
echo "bench rc=$?"; tail -c 600 $O/r02_final_bench_1gpu.json
timeout 420 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_final_bench_reference_arm.json 2> $O/r02_final_bench_reference_arm.err
echo "reference arm rc=$?"; tail -c 400 $O/r02_final_bench_reference_arm.json
# 4. launch list of the bench command (shares of the step)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r02_final_launches_1024.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-handoff --no-scaledep > $O/r02_final_launches.log 2>&1
python tools/launch_summary.py $O/r02_final_launches_1024.csv > $O/r02_final_launches_1024_summary.csv 2>/dev/null
head -14 $O/r02_final_launches_1024_summary.csv
# 5. counter pass of the first x / y / collapse launches of a sweep (DRAM bytes, instructions, issue, stalls)
M=gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers
timeout 300 ncu --metrics $M --clock-control none -k regex:"zpass_collapse|xpass_kernel|ypass_kernel|zpass_out|zpass_r2c|sources_kernel" -s 3 -c 14 --csv --log-file $O/r02_final_ncu_counters.csv \
  python scripts/prof_step.py 1024 classic 2 1 1 > $O/r02_final_ncu_counters.log 2>&1
tail -1 $O/r02_final_ncu_counters.log
python tools/ncu_to_traffic.py $O/r02_final_ncu_counters.csv 1024 1 $O/r02_final_traffic.json > /dev/null 2>&1
# 6. full-set captures, sources imported, one launch each: the collapse pass and the two TMA-fed strided passes
for K in zpass_collapse xpass_kernel ypass_kernel; do
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o $O/r02_final_ncu_$K \
    python scripts/prof_step.py 1024 classic 2 1 0 > $O/r02_final_ncu_$K.log 2>&1
  timeout 100 python tools/ncu_summary.py $O/r02_final_ncu_$K.ncu-rep $O/r02_final_ncu_full_$K.csv > /dev/null 2>&1
  rm -f $O/r02_final_ncu_$K.ncu-rep
done
ls -la $O | grep r02_final
