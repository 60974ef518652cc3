#!/bin/bash
# last GPU seconds of round 2 (2 x B200): the e2e path of bench.py at N > 1 (hand-off begin / end under the displacement
# stage, displacement fields in the Hessian buffers) on a 256^3 box -- it had only run at N = 1 since those changes
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --grid 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_last_e2e_2gpu.json 2> gpurun_out/r02_last_e2e_2gpu.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_last_e2e_2gpu.json').read().strip().splitlines()[-1])
print(json.dumps({'value':d['value'],'ms_per_step':d['ms_per_step'],'e2e':d['e2e'],'checks':d['checks']})[:1500])
P
tail -3 gpurun_out/r02_last_e2e_2gpu.err | cut -c1-300
