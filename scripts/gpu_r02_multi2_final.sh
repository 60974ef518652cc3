#!/bin/bash
# r02, 2 GPUs, last multi-GPU check of the round: peer-store passes with one tile per block (after the 8-GPU run that
# showed walking blocks stall behind their NVLink stores), local passes walking.  Parity of every field against the
# oracle at 64^3, then the 1024^3 bench line.
mkdir -p gpurun_out; O=gpurun_out
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 64 > $O/r02_multi2_final_parity.log 2>&1; echo "parity rc=$?"
grep -E "rel err|ok|GREEN|RED|mismatch" $O/r02_multi2_final_parity.log | tail -14
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/r02_bench_2gpu_final.json 2> $O/r02_bench_2gpu_final.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_2gpu_final.json').read().strip().splitlines()[-1]); r=d['roofline']
print(json.dumps({'value':d['value'],'ms_per_step':d['ms_per_step'],'ms_per_launch':r['ms_per_launch'],'per_radius_ms':r['per_radius_ms'],'lpt':r['lpt_stage_ms'],'checks':d['checks']}))
P
