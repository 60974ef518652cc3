#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_1024b.csv python bench.py --grid 1024 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1
tail -1 gpurun_out/ncu_bench2.log | cut -c1-100
