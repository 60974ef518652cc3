#!/bin/bash
# r02, 8 GPUs: the 2048^3 bench line with the unsplit all-local x pass (XCfg LOCAL) and the TMA tile loader
mkdir -p gpurun_out; O=gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > $O/r02_bench_8gpu_xlocal.json 2> $O/r02_bench_8gpu_xlocal.err; echo "rc=$?"
tail -c 4000 $O/r02_bench_8gpu_xlocal.json; tail -3 $O/r02_bench_8gpu_xlocal.err
