#!/bin/bash
# 2-GPU call: single-rank regression + 2-rank parity + 2-rank bench
mkdir -p gpurun_out
{
nvidia-smi -L
echo "=== single-GPU pytest"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
echo "=== 2-rank parity 64"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 64
echo "rc=$?"
echo "=== 2-rank parity 128"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/multi_gpu_check.py 128
echo "rc=$?"
echo "=== bench 1 GPU 1024"; timeout 900 python bench.py --grid 1024 --steps 2 --warmup 2 --no-cpu-baseline
echo "=== bench 2 GPU 1024"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu-baseline
} > gpurun_out/multi1.log 2>&1
grep -v "^$" gpurun_out/multi1.log | tail -60
