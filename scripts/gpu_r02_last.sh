#!/bin/bash
# last GPU call of round 2 (1 x B200, ~1 min): the final tree's bench.py runs and prints its line (new key lpt_breakdown_ms)
mkdir -p gpurun_out
timeout 110 python bench.py --no-e2e --no-cpu-baseline --no-handoff --no-scaledep --steps 2 --warmup 1 > gpurun_out/r02_last_bench.json 2> gpurun_out/r02_last_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_last_bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print(json.dumps({'value':d['value'],'ms_per_step':d['ms_per_step'],'ms_per_launch':r['ms_per_launch'],'lpt':r['lpt_stage_ms'],'lpt_breakdown_ms':r['lpt_breakdown_ms'],'traffic':r['traffic'],'traffic_source':r['traffic_source'],'issue':r['issue'].get('zpass_collapse_kernel')}))
P
tail -2 gpurun_out/r02_last_bench.err
