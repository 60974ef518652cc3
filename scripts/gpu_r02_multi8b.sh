#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
{
echo "=== bench 8 GPU 2048 (pipelined, priority streams, one stream per destination)"; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > $O/r02_bench_8gpu_b.json 2> $O/r02_bench_8gpu_b.err; echo "rc=$?"; tail -c 2500 $O/r02_bench_8gpu_b.json; tail -3 $O/r02_bench_8gpu_b.err
echo "=== bench 4 of the 8 GPUs, 1024 (pipelined)"; CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02_bench_4gpu.json 2> $O/r02_bench_4gpu.err; echo "rc=$?"; tail -c 1500 $O/r02_bench_4gpu.json
echo "=== bench 4 GPUs, 1024 (peer stores)"; PINB200_PEER_STORES=1 CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02_bench_4gpu_peerstores.json 2> $O/r02_bench_4gpu_peerstores.err; echo "rc=$?"; tail -c 1500 $O/r02_bench_4gpu_peerstores.json
} > $O/r02_multi8b.log 2>&1
grep -v "^$" $O/r02_multi8b.log | tail -20 | cut -c1-1200
