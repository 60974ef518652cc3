#!/bin/bash
mkdir -p gpurun_out
{
echo "=== passbench"; timeout 600 ./tools/passbench
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
echo "=== bench 1024"; timeout 900 python bench.py --grid 1024 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e
} > gpurun_out/run3.log 2>&1
grep -v "^$" gpurun_out/run3.log | tail -40
