#!/bin/bash
# r02: new device sort + compact hand-off + new epilogue math on hardware: the affected GPU tests, then the bench line
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_zgpu_3_fragment_handoff.py tests/test_gpu_parity.py tests/test_zgpu_4_dropin_catalogues.py -m gpu -q -x --durations=8 > $O/r02_pytest_call6.log 2>&1
echo "pytest rc=$?" >> $O/r02_pytest_call6.log; tail -22 $O/r02_pytest_call6.log
timeout 900 python bench.py --write-parity-fixture $O/b200_1gpu_fmax_1024.json > $O/r02_bench_call6.json 2> $O/r02_bench_call6.err
echo "bench rc=$?"; tail -c 3500 $O/r02_bench_call6.json; tail -5 $O/r02_bench_call6.err
