#!/bin/bash
# r02, 2 GPUs: the pipelined sweep (local x pass + copy-engine transposes under the collapse pass) -- parity against the
# oracle at 64^3 / 128^3, then the 1024^3 bench line with the parity check against the single-GPU fixture, then the
# r01 schedule (peer stores) for comparison
mkdir -p gpurun_out; O=gpurun_out
{
nvidia-smi -L
echo "=== 2-rank parity 64"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 64; echo "rc=$?"
echo "=== 2-rank parity 128"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/multi_gpu_check.py 128; echo "rc=$?"
echo "=== bench 2 GPU 1024 (pipelined)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $O/r02_bench_2gpu.json 2> $O/r02_bench_2gpu.err; echo "rc=$?"; tail -c 2500 $O/r02_bench_2gpu.json
echo "=== bench 2 GPU 1024 (peer stores, r01 schedule)"; PINB200_PEER_STORES=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02_bench_2gpu_peerstores.json 2> $O/r02_bench_2gpu_peerstores.err; echo "rc=$?"; tail -c 1500 $O/r02_bench_2gpu_peerstores.json
} > $O/r02_multi2.log 2>&1
grep -v "^$" $O/r02_multi2.log | tail -70
