"""Generates tests/golden/reference_scaledep_32.npz: the displacement fields of the REFERENCE's
own compute_fmax() (oracle/_ref, see oracle/Makefile) with a k-dependent growth rate, i.e. the
-DSCALE_DEPENDENT behaviour of compute_derivative (src/fmax-pfft.c:340-364).  The k loop is the
reference's; GrowingMode*(z, k) is the table stand-in of oracle/ref_harness.c (src/cosmo.c needs
GSL's integrators and cannot be compiled here).  Same input field as reference_fmax_32.npz.

    make -C oracle && OMP_NUM_THREADS=8 python tests/golden/make_reference_scaledep_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import pinocchio_oracle as po  # noqa: E402
from oracle.reference_runner import ReferenceRun  # noqa: E402

HERE = Path(__file__).resolve().parent
# grid-unit k runs from 2 pi/32 = 0.196 to pi sqrt(3) = 5.44: bins -1.0, -0.8, ... 0.8 in log10 k
# put modes below kmin, inside every bin and (k > 6.3 never) none above kmax; the shipped
# LOGKMIN = -3, DELTALOGK = 0.5 (src/def_splines.h:41-42) is covered by the emulator test.
LOGKMIN, DLOGK, NK = -0.5, 0.12, 10


def tables():
    rng = np.random.default_rng(11)
    base = np.log10(np.array([0.61, 0.16, 0.05, 0.11]))[:, None]
    return base + 0.3 * np.cumsum(rng.uniform(-0.2, 0.2, (4, NK)), axis=1)


def main():
    gold = dict(np.load(HERE / "reference_fmax_32.npz"))
    N = int(gold["N"])
    run = ReferenceRun(N, float(gold["box"]), list(gold["radii"]), gold["growth"], gold["invgrow_x"], gold["invgrow_y"],
                       threads=8)
    tab = tables()
    run.set_growth_tables(tab, LOGKMIN, DLOGK)
    run.set_kdensity(gold["kdensity"])
    run.compute_fmax()
    prod = run.products(po.PRODUCT_DTYPE_3LPT)
    # the collapse sweep and the LPT sources run with ScaleDep.order = 0: identical to the plain run
    plain = gold["products"].view(po.PRODUCT_DTYPE_3LPT)
    assert np.array_equal(prod["Fmax"], plain["Fmax"]) and np.array_equal(prod["Rmax"], plain["Rmax"])
    assert not np.array_equal(prod["Vel"], plain["Vel"])
    out = HERE / "reference_scaledep_32.npz"
    np.savez_compressed(out, N=N, log10_growth=tab, logkmin=LOGKMIN, dlogk=DLOGK,
                        **{n: prod[n].copy() for n in ("Vel", "Vel_2LPT", "Vel_3LPT_1", "Vel_3LPT_2")})
    print(out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
