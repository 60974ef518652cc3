#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zgpu_1_scaledep.py tests/test_zgpu_3_fragment_handoff.py -m gpu -q -k "not 256 and not large_grid and not 128" --durations=3 > $O/r02_pytest_call16.log 2>&1
echo "pytest rc=$?" >> $O/r02_pytest_call16.log; grep -E "passed|failed|FAILED|ERROR|rc=" $O/r02_pytest_call16.log | tail -5
timeout 600 python bench.py --no-cpu-baseline --no-handoff > $O/r02_bench_call16.json 2> $O/r02_bench_call16.err
echo "bench rc=$?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_call16.json').read().strip().splitlines()[-1])
print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','e2e','scaledep')}, indent=1))
P
tail -3 $O/r02_bench_call16.err
