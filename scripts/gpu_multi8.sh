#!/bin/bash
# 8-GPU call: parity on 8 and 4 ranks, then the scaling points the driver will run
mkdir -p gpurun_out
{
nvidia-smi -L | head -8; free -g | head -2
echo "=== 8-rank parity 256"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tests/multi_gpu_check.py 256
echo "rc=$?"
echo "=== 4-rank parity 128"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 tests/multi_gpu_check.py 128
echo "rc=$?"
echo "=== bench 8 GPU 2048"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 2 --warmup 2 --no-cpu-baseline
echo "=== bench 4 GPU 1024"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e
} > gpurun_out/multi8.log 2>&1
grep -v "^$" gpurun_out/multi8.log | grep -v "OMP_NUM_THREADS\|\*\*\*\*" | tail -50
