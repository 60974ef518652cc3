#!/bin/bash
# First GPU call of round 2 (1 x B200, about 25 min of box time): everything written after round 1's GPU
# budget was spent, in order of risk, each step under its own timeout so that a hang cannot eat the call.
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash scripts/gpu_round2_first.sh'
# Outputs: gpurun_out/r02_*.{log,csv,json}; copy the summaries to profiles/r02_* afterwards.
mkdir -p gpurun_out
O=gpurun_out

# 1. the whole -m gpu suite (what the driver runs at round end), verbose, no -x: every failure is listed
timeout 900 python -m pytest tests -m gpu -q -rA --durations=15 > $O/r02_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r02_pytest_gpu.log
tail -25 $O/r02_pytest_gpu.log

# 2. smoke() exactly as the driver calls it
timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' > $O/r02_smoke.log 2>&1
echo "smoke rc=$?" >> $O/r02_smoke.log
tail -2 $O/r02_smoke.log

# 3. the bench line with the final kernels (includes the isolated probes: fragmentation hand-off,
#    linked drop-in, collapse tables; e2e with the bare PCIe rates; CPU reference arm on this box)
timeout 900 python bench.py > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err
echo "bench rc=$?" >> $O/r02_bench_1gpu.err
tail -c 1500 $O/r02_bench_1gpu.json

# 4. launch list of the same command with the final kernels (shares of the step, not absolute times)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_1024.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-handoff > $O/r02_launches.log 2>&1
python tools/launch_summary.py $O/r02_launches_1024.csv > $O/r02_launches_1024_summary.csv 2>/dev/null
head -12 $O/r02_launches_1024_summary.csv

# 5. full-set captures of the three kernels of a radius, sources imported (one launch each)
for K in zpass_collapse xpass_kernel ypass_kernel; do
  timeout 420 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o $O/r02_ncu_$K \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-handoff > $O/r02_ncu_$K.log 2>&1
  timeout 120 python tools/ncu_summary.py $O/r02_ncu_$K.ncu-rep $O/r02_ncu_$K.csv > /dev/null 2>&1
done
ls -la $O | tail -20
