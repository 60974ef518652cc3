#!/bin/bash
mkdir -p gpurun_out
{
echo "=== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "=== bench 1024"; timeout 900 python bench.py --grid 1024 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e
echo "=== ncu launch list 1024"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_1024.csv python bench.py --grid 1024 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
} > gpurun_out/run2.log 2>&1
tail -40 gpurun_out/run2.log
