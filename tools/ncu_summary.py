"""Condense an .ncu-rep (ncu --set full) into the few numbers DESIGN.md and bench.py quote.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/out.csv
One row per metric of the LAST profiled launch; stall reasons as 'warps stalled per issue'."""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_local_ld.sum",
    "smsp__sass_inst_executed_op_local_st.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld_lookup_miss.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
]


def main(rep, out):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True)
    rows = list(csv.reader(r.stdout.splitlines()))
    hdr, units, last = rows[0], rows[1], rows[-1]
    col = {h: i for i, h in enumerate(hdr)}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        w.writerow(["Kernel Name", "", last[col["Kernel Name"]]])
        for m in WANT:
            if m in col:
                w.writerow([m, units[col[m]], last[col[m]]])
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                v = float(last[col[h]] or 0)
                if v >= 0.2:
                    w.writerow([h, "warps", f"{v:.3f}"])
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
