#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --grid 1024 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:xpass_kernel -s 3 -c 1 -o gpurun_out/prof3_xpass -f $CMD > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
