#!/bin/bash
# r02, 8 GPUs, 2048^3: pipelined sweep (default from 4 ranks) against the r01 schedule (peer stores); each line carries
# checks.subrun_1024 = parity of the same 8 ranks on the 1024^3 box with the committed single-GPU fixture
mkdir -p gpurun_out; O=gpurun_out
{
echo "=== bench 8 GPU 2048 (pipelined)"; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > $O/r02_bench_8gpu.json 2> $O/r02_bench_8gpu.err; echo "rc=$?"; tail -c 3000 $O/r02_bench_8gpu.json; tail -3 $O/r02_bench_8gpu.err
echo "=== bench 8 GPU 2048 (peer stores, r01 schedule)"; PINB200_PEER_STORES=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > $O/r02_bench_8gpu_peerstores.json 2> $O/r02_bench_8gpu_peerstores.err; echo "rc=$?"; tail -c 2000 $O/r02_bench_8gpu_peerstores.json
} > $O/r02_multi8.log 2>&1
grep -v "^$" $O/r02_multi8.log | tail -30 | cut -c1-1500
