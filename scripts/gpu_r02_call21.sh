#!/bin/bash
# r02 call 21: persistent strided passes (blocks walk the tiles, next tile's load under this tile's stores) A/B with the
# TMA loader, in one build; quick parity; the engine's own per-kernel times
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_zgpu_8_split_variant.py -m gpu -q -x -k "not 256 and not large_grid and not 512 and not 1024 and not 128" > $O/r02_pytest_call21.log 2>&1
echo "pytest rc=$?" >> $O/r02_pytest_call21.log; grep -E "passed|failed|FAILED|ERROR|rc=" $O/r02_pytest_call21.log | tail -5
{
for pers in 1 0; do for tma in 1 0; do
echo "== PINB200_PERSISTENT=$pers PINB200_TMA=$tma slabbench 1024 1"; PINB200_PERSISTENT=$pers PINB200_TMA=$tma timeout 200 ./tools/slabbench 1024 1 3 | grep -E "xpass|ypass"
done; done
for pers in 1 0; do
echo "== PINB200_PERSISTENT=$pers PINB200_TMA=1 slabbench 2048 8"; PINB200_PERSISTENT=$pers timeout 200 ./tools/slabbench 2048 8 3 | grep -E "xpass|ypass|staged"
done
} 2>&1 | tee $O/r02_slabbench_persistent.txt
for pers in 1 0; do
PINB200_PERSISTENT=$pers timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-handoff --no-scaledep --steps 5 --warmup 3 > $O/r02_bench_call21_p$pers.json 2> $O/r02_bench_call21_p$pers.err
echo "bench persistent=$pers rc=$?"
python - $pers <<'P'
import json,sys
d=json.loads(open(f'gpurun_out/r02_bench_call21_p{sys.argv[1]}.json').read().strip().splitlines()[-1]); r=d['roofline']
print(json.dumps({'value':d['value'],'ms_per_step':d['ms_per_step'],'ms_per_launch':r['ms_per_launch'],'per_radius_ms':r['per_radius_ms'],'lpt':r['lpt_stage_ms'],'checks':d['checks']}))
P
done
