#!/bin/bash
# Last GPU call of round 1 (about 100 s of budget left): parity of the new collapse arithmetic, of the
# scale-dependent displacement kernel and of recompute_sd; then per-kernel device times at 1024^3 from
# the engine's own CUDA-event timers; then (if time is left) instruction counters of the collapse kernel.
# No torch import (ctypes + numpy only) to keep start-up short.
mkdir -p gpurun_out
timeout 45 python -m pytest tests/test_gpu_parity.py tests/test_zgpu_1_scaledep.py -x -q \
  -k "collapse_cells or reference_code_golden or (fmax_and_displacements and 64) or against_reference_golden or (against_oracle and 64) or recompute_sd" \
  > gpurun_out/final_parity.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final_parity.log
tail -4 gpurun_out/final_parity.log
timeout 40 python scripts/gpu_time_kernels.py 1024 2 > gpurun_out/final_timing.log 2>&1
tail -3 gpurun_out/final_timing.log
timeout 60 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_executed_pipe_fp64.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread \
  --clock-control none -k regex:zpass_collapse -c 1 --csv --log-file gpurun_out/final_ncu_zcollapse.csv \
  python scripts/gpu_time_kernels.py 1024 0 > gpurun_out/final_ncu.log 2>&1
tail -12 gpurun_out/final_ncu_zcollapse.csv
