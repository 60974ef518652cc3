#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_zgpu_7_device_writers.py -m gpu -q -x --durations=5 > $O/r02_pytest_call7.log 2>&1
echo "pytest rc=$?" >> $O/r02_pytest_call7.log; tail -15 $O/r02_pytest_call7.log
timeout 600 python bench.py --no-cpu-baseline --no-handoff > $O/r02_bench_call7.json 2> $O/r02_bench_call7.err
echo "bench rc=$?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_call7.json').read().strip().splitlines()[-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','checks')}, indent=1))
P
tail -3 $O/r02_bench_call7.err
