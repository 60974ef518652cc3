#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --grid 1024 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:zpass_collapse -s 7 -c 1 -o gpurun_out/prof_zcollapse -f $CMD > gpurun_out/ncu1.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:xpass_kernel -s 3 -c 1 -o gpurun_out/prof_xpass -f $CMD >> gpurun_out/ncu1.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ypass_kernel -s 3 -c 1 -o gpurun_out/prof_ypass -f $CMD >> gpurun_out/ncu1.log 2>&1
tail -5 gpurun_out/ncu1.log; ls -la gpurun_out/
