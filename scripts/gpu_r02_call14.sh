#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_zgpu_3_fragment_handoff.py tests/test_zgpu_1_scaledep.py tests/test_zgpu_9_example_as_shipped.py tests/test_zgpu_6_build_variants.py tests/test_zgpu_5_collapse_tables.py -m gpu -q --durations=5 > $O/r02_pytest_call14.log 2>&1
echo "pytest rc=$?" >> $O/r02_pytest_call14.log; grep -E "passed|failed|FAILED|ERROR|rc=" $O/r02_pytest_call14.log | tail -8
timeout 600 python bench.py --no-cpu-baseline --no-handoff > $O/r02_bench_call14.json 2> $O/r02_bench_call14.err
echo "bench rc=$?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_call14.json').read().strip().splitlines()[-1])
print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','e2e','scaledep')}, indent=1))
P
tail -3 $O/r02_bench_call14.err
