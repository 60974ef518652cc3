#!/bin/bash
# r02 call 18: (a) quick parity of the TMA tile loader (default on) and of the split variant; (b) the 1024^3 kernels
# through the product launchers with and without TMA (tools/slabbench 1024 1), then the 2048^3 slab the same way.
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_zgpu_8_split_variant.py -m gpu -q -x -k "not 256 and not large_grid and not 512 and not 1024" --durations=3 > $O/r02_pytest_call18.log 2>&1
echo "pytest rc=$?" >> $O/r02_pytest_call18.log; grep -E "passed|failed|FAILED|ERROR|rc=" $O/r02_pytest_call18.log | tail -5
{
echo "== PINB200_TMA=1 slabbench 1024 1"; PINB200_TMA=1 timeout 200 ./tools/slabbench 1024 1 3
echo "== PINB200_TMA=0 slabbench 1024 1"; PINB200_TMA=0 timeout 200 ./tools/slabbench 1024 1 3
echo "== PINB200_TMA=1 slabbench 2048 8"; PINB200_TMA=1 timeout 200 ./tools/slabbench 2048 8 3
echo "== PINB200_TMA=0 slabbench 2048 8"; PINB200_TMA=0 timeout 200 ./tools/slabbench 2048 8 3
} 2>&1 | tee $O/r02_slabbench_tma.txt
