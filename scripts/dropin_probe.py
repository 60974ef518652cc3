"""Runs the linked drop-in (oracle/_ref/pinocchio_b200.x = the unchanged reference program + shim +
libpinb200.so) and the reference program itself (oracle/_ref/pinocchio_ref.x) on the HMF_Validation
parameter file (128^3) and prints one JSON line with their own timers and a catalogue comparison.
Run by bench.py in a fresh process after its measurements; usable by hand on a GPU box."""
import json
import os
import re
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden" / "hmf_validation"
REF = ROOT / "oracle" / "_ref"


def run(exe, d, threads):
    d.mkdir(parents=True, exist_ok=True)
    for n in ("parameter_file", "outputs"):
        (d / n).write_bytes((GOLDEN / n).read_bytes())
    r = subprocess.run([str(exe), "parameter_file"], cwd=d, capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, OMP_NUM_THREADS=str(threads)))
    return r.returncode, r.stdout + r.stderr


def timers(log):
    out = {}
    for key, name in (("Total", "total_s"), ("fmax", "fmax_s"), ("Fragmentation", "fragmentation_s"), ("Initialization", "init_s")):
        m = re.search(rf"(?m)^{key}:\s+([0-9.]+)", log)
        if m:
            out[name] = float(m.group(1))
    return out


def main():
    b200, ref = REF / "pinocchio_b200.x", REF / "pinocchio_ref.x"
    if not (b200.exists() and ref.exists()):
        print(json.dumps({"error": "oracle/_ref/pinocchio_{b200,ref}.x not built"}))
        return
    cores = os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as td:
        da, db = Path(td) / "b200", Path(td) / "ref"
        rc_a, log_a = run(b200, da, cores)
        if rc_a != 0:
            print(json.dumps({"error": "pinocchio_b200.x failed", "log_tail": log_a[-400:]}))
            return
        rc_b, log_b = run(ref, db, cores)
        if rc_b != 0:
            print(json.dumps({"error": "pinocchio_ref.x failed", "log_tail": log_b[-400:]}))
            return
        ca = np.loadtxt(da / "pinocchio.0.0000.test.catalog.out")
        cb = np.loadtxt(db / "pinocchio.0.0000.test.catalog.out")
        mb = dict(zip(cb[:, 0].astype(np.int64).tolist(), cb[:, 11].astype(np.int64).tolist()))
        same = sum(1 for i, n in zip(ca[:, 0].astype(np.int64).tolist(), ca[:, 11].astype(np.int64).tolist()) if mb.get(i) == n)
        pa = np.loadtxt(da / "pinocchio.test.FmaxPDF.out")[:, 2]
        pb = np.loadtxt(db / "pinocchio.test.FmaxPDF.out")[:, 2]
        print(json.dumps({"workload": "HMF_Validation/parameter_file, 128^3, one task, whole program (init + fmax + fragmentation)",
                          "host_threads": cores, "dropin": timers(log_a), "reference": timers(log_b),
                          "halos_z0": {"dropin": int(len(ca)), "reference": int(len(cb)), "same_id_and_particle_count": int(same)},
                          "fmaxpdf_max_bin_diff": int(np.abs(pa - pb).max())}))


if __name__ == "__main__":
    main()
