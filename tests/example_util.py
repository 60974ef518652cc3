"""BASELINE.json configs[0] / the build of configs[4]: example/parameter_file AS SHIPPED (GridSize 128, Box 500 Mpc/h,
CAMB power-spectrum tables with massive-neutrino scale-dependent growth, Hubble table, radiation) with the OPTIONS of
src/Makefile:46-76 (TWO_LPT THREE_LPT ELL_CLASSIC SCALE_DEPENDENT READ_PK_TABLE RECOMPUTE_DISPLACEMENTS
READ_HUBBLE_TABLE): oracle/_ref/pinocchio_{ref,emu,b200}_ex.x.  Fixtures: tests/golden/example/ (verbatim copies of the
shipped parameter_file, outputs, CAMBFiles/ and pinocchio.example.FmaxPDF.out; columns of the shipped catalogues)."""
import os
import re
import shutil
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
EXAMPLE = ROOT / "tests" / "golden" / "example"
REF_EX = ROOT / "oracle" / "_ref" / "pinocchio_ref_ex.x"
EMU_EX = ROOT / "oracle" / "_ref" / "pinocchio_emu_ex.x"
B200_EX = ROOT / "oracle" / "_ref" / "pinocchio_b200_ex.x"


def run_example(exe: Path, workdir: Path, grid: int | None = None, threads: int = 8, timeout: int = 1800) -> str:
    """run `exe parameter_file` on the shipped example (optionally on a smaller grid / box with the same cell size)"""
    workdir.mkdir(parents=True, exist_ok=True)
    text = (EXAMPLE / "parameter_file").read_text()
    if grid is not None:
        text = re.sub(r"(?m)^GridSize\s+\S+", f"GridSize               {grid}", text)
        text = re.sub(r"(?m)^BoxSize\s+\S+", f"BoxSize                {500.0 * grid / 128.0}", text)
    (workdir / "parameter_file").write_text(text)
    shutil.copy(EXAMPLE / "outputs", workdir / "outputs")
    if not (workdir / "CAMBFiles").exists():
        shutil.copytree(EXAMPLE / "CAMBFiles", workdir / "CAMBFiles")
    r = subprocess.run([str(exe), "parameter_file"], cwd=workdir, capture_output=True, text=True, timeout=timeout,
                       env=dict(os.environ, OMP_NUM_THREADS=str(threads)))
    (workdir / "log.txt").write_text(r.stdout + r.stderr)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    return r.stdout
