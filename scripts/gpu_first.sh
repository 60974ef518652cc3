#!/bin/bash
# first GPU call: smoke, small + full-size bench, parity tests; logs under gpurun_out/
mkdir -p gpurun_out
{
nvidia-smi; free -g; nproc
echo "=== smoke"; timeout 600 python __graft_entry__.py --smoke
echo "=== bench 256"; timeout 600 python bench.py --grid 256 --steps 2 --warmup 1 --no-cpu-baseline
echo "=== bench 512"; timeout 600 python bench.py --grid 512 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e
echo "=== bench 1024"; timeout 900 python bench.py --grid 1024 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e
echo "=== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40
} > gpurun_out/first.log 2>&1
tail -60 gpurun_out/first.log
