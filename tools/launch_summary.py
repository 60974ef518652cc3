"""Per-kernel shares of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`).
    python tools/launch_summary.py gpurun_out/launches.csv > profiles/rNN_launches_summary.csv
Kernel names are cut at the template / argument list; times under ncu are cold-cache and serialised, so the
SHARES are what is compared with the live CUDA-event timers of bench.py, not the absolute values."""
import csv
import re
import sys
from collections import defaultdict

rows = [l for l in open(sys.argv[1], newline="") if l.startswith('"')]
tot, cnt = defaultdict(float), defaultdict(int)
for r in csv.DictReader(rows):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"^void\s+", "", r["Kernel Name"])
    name = re.split(r"[<(]", name)[0].split("::")[-1]
    ns = float(r["Metric Value"].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r["Metric Unit"], 1.0)
    tot[name] += ns
    cnt[name] += 1
allns = sum(tot.values()) or 1.0
print("kernel,launches,total_ms,share_of_listed")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"{k},{cnt[k]},{tot[k] / 1e6:.2f},{tot[k] / allns:.3f}")
