#!/bin/bash
mkdir -p gpurun_out
{
echo "=== slabbench 2048/8"; timeout 300 ./tools/slabbench 2048 8 3
echo "=== slabbench 1024/1"; timeout 300 ./tools/slabbench 1024 1 3
echo "=== pytest gpu, split variant"; PINB200_LIB=$PWD/pinocchio_b200/libpinb200_split.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fft_forward or second_derivatives or fmax_and_displacements or golden" 2>&1 | tail -5
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "=== bench 1024"; timeout 900 python bench.py --grid 1024 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e
} > gpurun_out/run4.log 2>&1
grep -v "^$" gpurun_out/run4.log | tail -40
