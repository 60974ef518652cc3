#!/bin/bash
mkdir -p gpurun_out
{
echo "=== 8-rank parity 128"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tests/multi_gpu_check.py 128
echo "rc=$?"
echo "=== bench 8 GPU 2048"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 2 --warmup 3
echo "rc=$?"
} > gpurun_out/multi8b.log 2>&1
grep -v "^$" gpurun_out/multi8b.log | grep -v "OMP_NUM_THREADS\|\*\*\*\*" | tail -40
