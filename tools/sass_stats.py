"""Static SASS statistics of one kernel, attributed to source lines (needs -lineinfo).

    python tools/sass_stats.py pinocchio_b200/_build/k_zpass.cu.o zpass_collapse_kernelILi512ELi1ELi6 [--lines]

Counts instructions per source file and per opcode class (FP64 pipe, MUFU, integer/move, memory,
control).  Straight-line code such as the collapse epilogue executes each instruction once per
cell, so the static count of collapse.cuh + fastmath.cuh is the per-cell instruction budget the
kernel is bound by.  Used to compare variants without a GPU.
"""
import collections
import re
import subprocess
import sys
import tempfile
from pathlib import Path


def classify(op: str) -> str:
    base = op.split(".")[0]
    if base in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"):
        return "fp64"
    if base == "MUFU":
        return "mufu"
    if base in ("LDS", "STS", "LDG", "STG", "LDL", "STL", "LDC", "LDGSTS", "LDSM", "ATOMS", "ATOMG", "RED", "LDGDEPBAR", "DEPBAR"):
        return "mem"
    if base in ("BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "BAR", "WARPSYNC", "NOP", "YIELD", "BREAK", "BMOV"):
        return "ctrl"
    if base in ("F2F", "I2F", "F2I", "FRND", "I2FP", "F2FP", "FSEL", "FADD", "FMUL", "FFMA", "FSETP", "FMNMX", "FCHK"):
        return "fp32/cvt"
    return "int/mov"


def main():
    obj, pattern = sys.argv[1], sys.argv[2]
    show_lines = "--lines" in sys.argv
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", str(Path(obj).resolve())], cwd=td, check=True, capture_output=True)
        cubin = next(Path(td).glob("*.cubin"))
        sass = subprocess.run(["nvdisasm", "-g", str(cubin)], capture_output=True, text=True).stdout
    in_sec = False
    cur = ("?", 0)
    per_file = collections.Counter()
    per_file_class = collections.defaultdict(collections.Counter)
    per_line = collections.defaultdict(collections.Counter)
    ops = collections.defaultdict(collections.Counter)
    total = 0
    for ln in sass.splitlines():
        if ln.startswith("\t.section"):
            in_sec = ".text." in ln and pattern in ln
            continue
        if not in_sec:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            # the innermost frame of an inlined chain is listed first: keep the first of a run
            if "inlined at" in m.group(3) or True:
                cur = (Path(m.group(1)).name, int(m.group(2)))
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
        if not m:
            continue
        op = m.group(1)
        cl = classify(op)
        total += 1
        per_file[cur[0]] += 1
        per_file_class[cur[0]][cl] += 1
        per_line[cur][cl] += 1
        ops[cur[0]][op.split(".")[0]] += 1
    print(f"kernel pattern {pattern}: {total} SASS instructions")
    classes = ["fp64", "mufu", "int/mov", "fp32/cvt", "mem", "ctrl"]
    print(f"{'file':22s} {'total':>6s} " + " ".join(f"{c:>8s}" for c in classes))
    for f, n in per_file.most_common():
        print(f"{f:22s} {n:6d} " + " ".join(f"{per_file_class[f][c]:8d}" for c in classes))
    for f in ("collapse.cuh", "fastmath.cuh"):
        if f in ops:
            print(f"\n{f} opcodes: " + ", ".join(f"{o} {n}" for o, n in ops[f].most_common(14)))
    if show_lines:
        print()
        for (f, l), c in sorted(per_line.items()):
            if f in ("collapse.cuh", "fastmath.cuh", "spline_pack.h"):
                print(f"{f}:{l:<4d} " + " ".join(f"{k}={v}" for k, v in c.items()))


if __name__ == "__main__":
    main()
