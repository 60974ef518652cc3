#!/bin/bash
mkdir -p gpurun_out
{
echo "=== slabbench 2048/8"; timeout 300 ./tools/slabbench 2048 8 3
echo "=== ncu xpass split (slab 2048/8)"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:xpass_kernel -c 1 -o gpurun_out/prof5_xsplit -f ./tools/slabbench 2048 8 1 > gpurun_out/ncu5.log 2>&1; tail -3 gpurun_out/ncu5.log
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "=== bench 1024 (default flags)"; timeout 1500 python bench.py
echo "=== bench reference arm"; timeout 1500 python bench.py --impl reference --steps 1 --warmup 1
nproc; free -g | head -2
} > gpurun_out/run5.log 2>&1
grep -v "^$" gpurun_out/run5.log | tail -40
