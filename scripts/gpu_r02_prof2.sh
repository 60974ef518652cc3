#!/bin/bash
# r02 call: kernel variants under tools/passbench (old = before the staged twiddles / Fmax row / x-pass factor table)
mkdir -p gpurun_out; O=gpurun_out
for v in old base; do echo "== passbench_$v x z"; timeout 120 ./tools/passbench_$v x z; done 2>&1 | grep -v "job \|timeline" | tee $O/r02_passbench2.txt
for v in nofms noepi; do echo "== passbench_$v z"; timeout 120 ./tools/passbench_$v z; done 2>&1 | grep -v "job \|timeline" | tee -a $O/r02_passbench2.txt
timeout 600 ncu --section SourceCounters --section InstructionStats --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy \
  --import-source on --clock-control none -k regex:zck -s 0 -c 1 -f -o $O/r02_zck_src ./tools/passbench_base z > $O/r02_zck_src.log 2>&1
tail -3 $O/r02_zck_src.log
ncu -i $O/r02_zck_src.ncu-rep --page source --csv > $O/r02_zck_source.csv 2>/dev/null
ncu -i $O/r02_zck_src.ncu-rep --page raw --csv > $O/r02_zck_raw.csv 2>/dev/null
ls -la $O
