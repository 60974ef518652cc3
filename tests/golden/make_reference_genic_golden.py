"""Generates tests/golden/reference_genic_32.npz: kdensity from the REFERENCE's own GenIC_large
(src/GenIC.c compiled verbatim into oracle/_ref, GSL's ranlxd1/mt19937 restated in
oracle/ref_gsl_rng.c, P(k) from the lattice table of the EH fit) for two seeded 32^3 boxes.

    make -C oracle && python tests/golden/make_reference_genic_golden.py
"""
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
from oracle.reference_runner import ReferenceRun
from pinocchio_b200.cosmology import Cosmology, pk_lattice_table
N, seed, fixed, paired, out = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
cosmo = Cosmology(pk_norm_override=2.03146e7)
box = N / 0.7
x = np.linspace(-2, 0, 8)
run = ReferenceRun(N, box, [0.0], np.ones(4), x, x, threads=2)
np.save(out, run.genic(seed, pk_lattice_table(cosmo, N, box), fixed, paired))
"""


def one(N, seed, fixed, paired):
    with tempfile.TemporaryDirectory() as td:
        out = Path(td) / "kd.npy"
        subprocess.run([sys.executable, "-c", SCRIPT, str(ROOT), str(N), str(seed), str(fixed), str(paired), str(out)], check=True)
        return np.load(out)


def main():
    N = 32
    out = Path(__file__).resolve().parent / "reference_genic_32.npz"
    np.savez_compressed(out, N=N, box=N / 0.7, pk_norm=2.03146e7,
                        kd_486604=one(N, 486604, 0, 0), kd_12345_fixed_paired=one(N, 12345, 1, 1))
    print(out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
