"""Attribute an ncu source-page CSV (SASS level: `ncu -i rep --page source --csv`) to source lines.

    python tools/ncu_source_lines.py <source.csv> <binary-or-.o> <kernel-symbol-substring> [--top N] [--by file|line|op]

The CSV carries per-instruction executed counts and warp-stall samples but no line numbers; the same kernel
is disassembled here with `nvdisasm -g` (needs -lineinfo) and matched by instruction offset.
"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
from pathlib import Path


def disasm_lines(binary: str, pattern: str):
    """offset -> (file, line, opcode) for the kernel whose section name contains `pattern`"""
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", str(Path(binary).resolve())], cwd=td, check=True, capture_output=True)
        out = {}
        for cubin in Path(td).glob("*.cubin"):
            sass = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout
            in_sec, cur = False, ("?", 0)
            for ln in sass.splitlines():
                if ln.startswith("\t.section"):
                    in_sec = ".text." in ln and pattern in ln
                    continue
                if not in_sec:
                    continue
                m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
                if m:
                    cur = (Path(m.group(1)).name, int(m.group(2)))
                    continue
                m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
                if m:
                    out[int(m.group(1), 16)] = (cur[0], cur[1], m.group(2))
            if out:
                return out
    return {}


def main():
    src, binary, pattern = sys.argv[1:4]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    by = sys.argv[sys.argv.index("--by") + 1] if "--by" in sys.argv else "line"
    lines = disasm_lines(binary, pattern)
    rows = list(csv.reader(open(src)))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    base = None
    agg = collections.defaultdict(lambda: collections.Counter())
    tot = collections.Counter()
    for r in rows[2:]:
        if len(r) < len(hdr) - 5:
            continue
        addr = int(r[col["Address"]], 16)
        if base is None:
            base = addr
        off = addr - base
        f, l, op = lines.get(off, ("?", 0, ""))
        # the opcode always comes from the profile itself (the binary given for the line mapping may be a later build)
        toks = [t for t in r[col["Source"]].split() if not t.startswith("@")]
        op = toks[0] if toks else op
        ex = int(r[col["Instructions Executed"]] or 0)
        smp = int(r[col["# Samples"]] or 0)
        key = {"file": f, "line": f"{f}:{l}", "op": op.split(".")[0]}[by]
        agg[key]["inst"] += ex
        agg[key]["samples"] += smp
        isfp64 = op.split(".")[0] in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
        agg[key]["fp64"] += ex if isfp64 else 0
        tot["inst"] += ex
        tot["samples"] += smp
        tot["fp64"] += ex if isfp64 else 0
        for h in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_barrier", "stall_mio", "stall_not_selected",
                  "stall_dispatch", "stall_branch_resolving", "stall_no_inst", "stall_lg"):
            v = int(r[col[h]] or 0)
            agg[key][h] += v
            tot[h] += v
    print(f"total warp instructions {tot['inst']:,}  fp64 {tot['fp64']:,}  samples {tot['samples']:,}")
    print("stalls: " + "  ".join(f"{k[6:]}={v / max(1, tot['samples']):.3f}" for k, v in tot.items() if k.startswith("stall_")))
    print(f"{'key':34s} {'inst%':>6s} {'fp64%':>6s} {'smp%':>6s}  top stalls")
    for k, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = sorted(((h[6:], v) for h, v in c.items() if h.startswith("stall_")), key=lambda x: -x[1])[:3]
        print(f"{k:34s} {100 * c['inst'] / tot['inst']:6.2f} {100 * c['fp64'] / max(1, tot['fp64']):6.2f} {100 * c['samples'] / tot['samples']:6.2f}  "
              + " ".join(f"{h}={v / max(1, c['samples']):.2f}" for h, v in st))


if __name__ == "__main__":
    main()
