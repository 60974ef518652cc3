#!/bin/bash
# r02: the whole -m gpu suite as the driver runs it (no -x), then the bench line without the CPU leg
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA --durations=20 > $O/r02_pytest_gpu_full.log 2>&1
echo "pytest rc=$?" >> $O/r02_pytest_gpu_full.log; grep -E "passed|failed|FAILED|ERROR|rc=" $O/r02_pytest_gpu_full.log | tail -15
timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
timeout 700 python bench.py --no-cpu-baseline > $O/r02_bench_call8.json 2> $O/r02_bench_call8.err
echo "bench rc=$?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_call8.json').read().strip().splitlines()[-1])
print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','e2e','scaledep','fragment_handoff','checks')}, indent=1))
P
tail -3 $O/r02_bench_call8.err
