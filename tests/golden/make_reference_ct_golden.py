"""Generates tests/golden/reference_ct_32.npz: outputs of the WHOLE reference program compiled with
-DTABULATED_CT (ELL_CLASSIC: oracle/_ref/pinocchio_ref_tab.x) and with -DELL_SNG -DTABULATED_CT
(oracle/_ref/pinocchio_ref_sng.x; oracle/Makefile), and with -DMOD_GRAV_FR -DFR0=1.e-5 on top of that
(pinocchio_ref_fr.x: Hu-Sawicki f(R), scale-dependent growth, ten radii), on the HMF_Validation parameter
file scaled to a 32^3 box of 32 Mpc/h (nine smoothing radii, 1 Mpc/h cells as in the 128^3 run).

Stored per variant (`tab_*`, `sng_*`, `fr_*`): every 97th point of the nine collapse-time tables the program
writes to pinocchio.test.CTtable.out (+ the 40-byte header), the FmaxPDF file, the four halo
catalogues and the z = 0 mass function as text.  The ELL_SNG run integrates 2.25 million ODE systems
on one core: about seven minutes.

    make -C oracle all && python tests/golden/make_reference_ct_golden.py [--reuse-tab DIR] [--reuse-sng DIR]
"""
import argparse
import os
import re
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
REF = ROOT / "oracle" / "_ref"
NPOINTS, STRIDE = 100 * 50 * 50, 97
OUTPUT_FILES = ["pinocchio.test.FmaxPDF.out", "pinocchio.0.0000.test.catalog.out", "pinocchio.0.5000.test.catalog.out",
                "pinocchio.1.0000.test.catalog.out", "pinocchio.2.0000.test.catalog.out", "pinocchio.0.0000.test.mf.out"]


def parameter_file_32() -> str:
    text = (HERE / "hmf_validation" / "parameter_file").read_text()
    text = re.sub(r"(?m)^BoxSize\s+\S+", "BoxSize                32", text)
    text = re.sub(r"(?m)^GridSize\s+\S+", "GridSize               32", text)
    text = re.sub(r"(?m)^MaxMemPerParticle\s+\S+", "MaxMemPerParticle      400", text)
    return text + "\nCTtableFile none\n"


def run(exe: Path, workdir: Path):
    workdir.mkdir(parents=True, exist_ok=True)
    (workdir / "parameter_file").write_text(parameter_file_32())
    (workdir / "outputs").write_bytes((HERE / "hmf_validation" / "outputs").read_bytes())
    r = subprocess.run([str(exe), "parameter_file"], cwd=workdir, capture_output=True, text=True,
                       env=dict(os.environ, OMP_NUM_THREADS="8"))
    (workdir / "log.txt").write_text(r.stdout + r.stderr)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]


def read_cttable(path: Path):
    raw = path.read_bytes()
    header = np.frombuffer(raw[:40], dtype=np.uint8)
    off, tabs = 40, []
    while off < len(raw):
        off += 4                                             # int ismooth
        tabs.append(np.frombuffer(raw[off:off + 8 * NPOINTS], dtype=np.float64))
        off += 8 * NPOINTS
    return header, np.array(tabs)


def collect(tag: str, d: Path, out: dict):
    header, tabs = read_cttable(d / "pinocchio.test.CTtable.out")
    idx = np.arange(0, NPOINTS, STRIDE)
    out[f"{tag}_header"] = header
    out[f"{tag}_table_idx"] = idx
    out[f"{tag}_table"] = tabs[:, idx]
    out[f"{tag}_nonzero_per_radius"] = (tabs != 0).sum(axis=1)
    for name in OUTPUT_FILES:
        out[f"{tag}_file_{name}"] = np.frombuffer((d / name).read_bytes(), dtype=np.uint8)
    log = (d / "log.txt").read_text()
    out[f"{tag}_sigma"] = np.array([float(x) for x in re.findall(r"computed sigma:\s+([0-9.]+)", log)])
    rv = re.findall(r"\d+\)\s+Radius=\s*([0-9.]+), Variance=\s*([0-9.]+)", log)
    out[f"{tag}_radius"] = np.array([float(r) for r, _ in rv])
    out[f"{tag}_variance"] = np.array([float(v) for _, v in rv])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reuse-tab")
    ap.add_argument("--reuse-sng")
    ap.add_argument("--reuse-fr")
    a = ap.parse_args()
    out = {}
    for tag, reuse in (("tab", a.reuse_tab), ("sng", a.reuse_sng), ("fr", a.reuse_fr)):
        d = Path(reuse) if reuse else Path(tempfile.mkdtemp(prefix=f"pinref_{tag}_"))
        if not reuse:
            run(REF / f"pinocchio_ref_{tag}.x", d)
        collect(tag, d, out)
    np.savez_compressed(HERE / "reference_ct_32.npz", **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    sys.exit(main())
