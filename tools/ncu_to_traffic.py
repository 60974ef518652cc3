"""ncu counter pass (CSV, `ncu --metrics ... --csv --log-file x.csv`) -> the per-kernel JSON that bench.py reads for
`roofline.traffic` and for its instruction-issue model (profiles/rNN_traffic.json).

    python tools/ncu_to_traffic.py gpurun_out/r02_ncu_counters.csv 1024 1 profiles/r02_traffic.json

Per kernel (first captured launch of each name): DRAM bytes read / written, warp instructions, FP64-pipe warp
instructions, duration under ncu, and -- for the collapse z pass -- both instruction counts per cell (x 32 lanes).
"""
import csv
import json
import re
import sys


def main():
    src, grid, ngpu, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[hi + 1:]:
        if len(r) <= col["Metric Value"]:
            continue
        name = re.sub(r"^void\s+", "", r[col["Kernel Name"]])
        short = re.split(r"[<(]", name)[0].split("::")[-1]
        key = (short, r[col["ID"]])
        per.setdefault(key, {"kernel": name})[r[col["Metric Name"]]] = float(r[col["Metric Value"]].replace(",", ""))
    res = {"_comment": "per-kernel counters of one launch from an ncu counter pass of the final build; read by bench.py "
                       "(roofline.traffic, roofline.issue) when grid and GPU count match", "_source": src}
    seen = set()
    cells = float(grid) ** 3 / ngpu
    for (short, _), m in per.items():
        if short in seen or "dram__bytes_read.sum" not in m:
            continue
        seen.add(short)
        e = {"grid": grid, "n_gpus": ngpu, "kernel": m["kernel"], "read_gb": m["dram__bytes_read.sum"] / 1e9,
             "write_gb": m["dram__bytes_write.sum"] / 1e9, "ncu_ms": m.get("gpu__time_duration.sum", 0.0) / 1e6,
             "warp_inst": m.get("smsp__inst_executed.sum"), "fp64_warp_inst": m.get("sm__inst_executed_pipe_fp64.sum"),
             "issue_active_pct": m.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
             "fp64_pipe_pct": m.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), "source": src}
        if e["warp_inst"] and e["fp64_warp_inst"]:
            e["inst_per_cell"] = e["warp_inst"] * 32.0 / cells
            e["fp64_inst_per_cell"] = e["fp64_warp_inst"] * 32.0 / cells
        res[short] = e
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk != "source"}
                      for k, v in res.items() if not k.startswith("_")}, indent=1))


if __name__ == "__main__":
    main()
