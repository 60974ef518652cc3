"""In-tree build of libpinb200.so (sm_100a only) with nvcc.

``python -m pinocchio_b200.build`` compiles pinocchio_b200/csrc/*.cu in parallel and links
pinocchio_b200/libpinb200.so.  The .so is git-ignored but travels to the GPU box with the
snapshot.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = HERE / "_build"
LIB = HERE / "libpinb200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]
SOURCES = ["engine.cu", "k_strided.cu", "k_zpass.cu", "k_misc.cu"]


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                    + [HERE.parent / "include" / "pinb200.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src: str) -> tuple[str, str]:
    obj = OBJ / (src + ".o")
    cmd = [NVCC, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return src, r.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile (if sources changed) and return the path of libpinb200.so."""
    OBJ.mkdir(exist_ok=True)
    stamp = OBJ / "digest.txt"
    dig = _digest()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == dig:
        return LIB
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        logs = list(ex.map(_compile, SOURCES))
    (OBJ / "ptxas.log").write_text("\n".join(f"==== {s}\n{l}" for s, l in logs))
    if verbose:
        for s, l in logs:
            print(f"==== {s}\n{l}")
    cmd = [NVCC, "-shared", "-o", str(LIB), *[str(OBJ / (s + ".o")) for s in SOURCES],
           "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(dig)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
