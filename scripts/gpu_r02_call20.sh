#!/bin/bash
# r02 call 20: y pass with the ky^q table (slabbench 1024 1), quick parity with it, e2e with the hand-off sort under /
# before the displacement stage
mkdir -p gpurun_out; O=gpurun_out
timeout 200 ./tools/slabbench 1024 1 3 2>&1 | tee $O/r02_slabbench_kpow.txt
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_zgpu_8_split_variant.py tests/test_zgpu_10_scaledep_gm.py -m gpu -q -x -k "not 256 and not large_grid and not 512 and not 1024 and not 128" --durations=3 > $O/r02_pytest_call20.log 2>&1
echo "pytest rc=$?" >> $O/r02_pytest_call20.log; grep -E "passed|failed|FAILED|ERROR|rc=" $O/r02_pytest_call20.log | tail -5
for ser in 0 1; do
  PINB200_HANDOFF_SERIAL=$ser timeout 400 python bench.py --no-cpu-baseline --no-handoff --no-scaledep --steps 5 --warmup 3 > $O/r02_bench_call20_ser$ser.json 2> $O/r02_bench_call20_ser$ser.err
  echo "bench ser=$ser rc=$?"
  python - $ser <<'P'
import json,sys
d=json.loads(open(f'gpurun_out/r02_bench_call20_ser{sys.argv[1]}.json').read().strip().splitlines()[-1])
e=d['e2e']; r=d['roofline']
print(json.dumps({'value':d['value'],'ms_per_step':d['ms_per_step'],'ms_per_launch':r['ms_per_launch'],'per_radius_ms':r['per_radius_ms'],'lpt':r['lpt_stage_ms'],'e2e':{k:e.get(k) for k in ('value','ms_per_step','select_sort_ms_device','phases_ms_rank0','handoff_ok')}}))
P
done
