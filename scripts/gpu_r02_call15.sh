#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zgpu_1_scaledep.py tests/test_zgpu_3_fragment_handoff.py tests/test_zgpu_4_dropin_catalogues.py tests/test_zgpu_6_build_variants.py tests/test_zgpu_7_device_writers.py -m gpu -q -k "not 256 and not large_grid" --durations=5 > $O/r02_pytest_call15.log 2>&1
echo "pytest rc=$?" >> $O/r02_pytest_call15.log; grep -E "passed|failed|FAILED|ERROR|rc=" $O/r02_pytest_call15.log | tail -8
timeout 600 python bench.py --no-cpu-baseline --no-handoff --no-scaledep > $O/r02_bench_call15.json 2> $O/r02_bench_call15.err
echo "bench rc=$?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_call15.json').read().strip().splitlines()[-1])
print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','e2e')}, indent=1))
P
tail -3 $O/r02_bench_call15.err
