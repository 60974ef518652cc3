#!/bin/bash
# r02 call 17: (a) quick parity of the TMA tile loader (default on) and of the split variant; (b) the 1024^3 kernels through
# the product launchers with and without TMA (tools/slabbench 1024 1); (c) the 2048^3 slab kernels on one GPU: the
# all-local unsplit x pass of the pipelined sweep (XCfg LOCAL) against the scattering kernel, with a value check;
# (d) an ncu counter pass of the slab's x (both), y and collapse kernels.
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_zgpu_8_split_variant.py -m gpu -q -x -k "not 256 and not large_grid and not 512 and not 1024" --durations=3 > $O/r02_pytest_call17.log 2>&1
echo "pytest rc=$?" >> $O/r02_pytest_call17.log; grep -E "passed|failed|FAILED|ERROR|rc=" $O/r02_pytest_call17.log | tail -5
{
echo "== PINB200_TMA=1 slabbench 1024 1"; PINB200_TMA=1 timeout 200 ./tools/slabbench 1024 1 3
echo "== PINB200_TMA=0 slabbench 1024 1"; PINB200_TMA=0 timeout 200 ./tools/slabbench 1024 1 3
echo "== PINB200_TMA=1 slabbench 2048 8"; PINB200_TMA=1 timeout 200 ./tools/slabbench 2048 8 3
echo "== PINB200_TMA=0 slabbench 2048 8"; PINB200_TMA=0 timeout 200 ./tools/slabbench 2048 8 3
} 2>&1 | tee $O/r02_slabbench.txt
M=gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,launch__block_size,launch__grid_size
timeout 500 ncu --metrics $M --clock-control none -k regex:"zpass_collapse|xpass_kernel|xpass_local|ypass_kernel" --csv --log-file $O/r02_ncu_counters_slab2048.csv \
  ./tools/slabbench 2048 8 0 > $O/r02_ncu_slab.log 2>&1
tail -3 $O/r02_ncu_slab.log
python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/r02_ncu_counters_slab2048.csv')))
hi=next(i for i,r in enumerate(rows) if r and r[0]=='ID'); col={h:i for i,h in enumerate(rows[hi])}
per={}
for r in rows[hi+1:]:
    if len(r)>col['Metric Value']: per.setdefault((r[col['ID']], r[col['Kernel Name']][:60]),{})[r[col['Metric Name']]]=r[col['Metric Value']]
for k,m in per.items():
    print(k, {a.split('__')[-1][:40]:b for a,b in m.items() if any(s in a for s in ('duration','dram','issue_active','registers','warps_active','inst_executed'))})
P
