"""Pins the oracle (and the host cosmology tables) to the reference's own shipped fixtures.

Fixtures under tests/golden/hmf_validation/ are verbatim copies / excerpts of the reference's
HMF_Validation/ directory (128^3, seed 486604, one MPI rank; see tests/golden/README.md).
"""
from pathlib import Path

import numpy as np
import pytest

from oracle import pinocchio_oracle as po
from pinocchio_b200.cosmology import Cosmology, set_smoothing

GOLDEN = Path(__file__).resolve().parent / "golden" / "hmf_validation"
LOGGED_SIGMA = [0.2032, 0.3258, 0.5051, 0.7505, 1.0850, 1.5527, 2.1897, 2.6563, 2.7733]   # log_RUN.txt:135-335
LOGGED_RADII = [20.635922, 13.996056, 9.026099, 5.465945, 3.058354, 1.548258, 0.689079, 0.258729, 0.0]
LOGGED_VAR = [0.078961, 0.157548, 0.314349, 0.627210, 1.251448, 2.496966, 4.982102, 9.940601, 10.477455]


def test_gsl_rng_known_answers():
    # GSL rng/test.c: mt19937 seed 4357 -> 1000th output; ranlxd1 seed 1 -> 10000th output
    assert po.mt19937_outputs(4357, 1000)[999] == 1186927261
    r = po.RanLxd1([1])
    for _ in range(9999):
        r.get()
    assert int(r.get()[0]) == 1998227290


def test_ranlxd1_signed_seed_quirk():
    # seeds >= 2^31 behave as |int32(seed)|
    a = po.RanLxd1(np.array([0xFFFFFFFF], dtype=np.uint32))       # int32 = -1 -> |.| = 1
    b = po.RanLxd1(np.array([1], dtype=np.uint32))
    assert np.array_equal(a.xdbl, b.xdbl)


def test_spiral_map():
    # first ring of the square spiral around the origin (src/GenIC.c:840-855)
    assert int(po.get_map(0, 0)) == 1
    ords = sorted(int(po.get_map(x, y)) for x in (-1, 0, 1) for y in (-1, 0, 1))
    assert ords == list(range(1, 10))


@pytest.fixture(scope="module")
def cosmo():
    return Cosmology()


def test_cosmology_tables(cosmo):
    # log_RUN.txt:57 "Normalization constant for the power spectrum: 2.03146e+07"
    assert abs(cosmo.PkNorm / 2.03146e7 - 1) < 5e-6
    lad = set_smoothing(cosmo, 128 / 0.7 / 128)
    assert lad.Nsmooth == 9
    assert np.abs(lad.Radius - LOGGED_RADII).max() < 1e-6          # log_RUN.txt:62-70
    assert np.abs(lad.Variance - LOGGED_VAR).max() < 1e-6
    # growth factors at a = 1 (pinocchio.test.cosmology.out, row a=1, columns 7-10)
    assert abs(cosmo.GrowingMode(0.0) - 1) < 1e-12
    assert abs(cosmo.GrowingMode_2LPT(0.0) - 0.432825) < 1e-6
    assert abs(cosmo.GrowingMode_3LPT_1(0.0) + 0.113453) < 1e-6
    assert abs(cosmo.GrowingMode_3LPT_2(0.0) - 0.121639) < 1e-6
    rows = np.loadtxt(GOLDEN / "cosmology.excerpt.out")
    idx = np.searchsorted(cosmo.a_knots, rows[:, 0] * (1 - 1e-5))
    assert np.abs(rows[:, 6] / cosmo.grow1[idx] - 1).max() < 2e-5
    assert np.abs(rows[:, 7] / cosmo.grow2[idx] - 1).max() < 5e-5


def test_hmf_validation_sigma_and_fmaxpdf():
    cosmo = Cosmology(pk_norm_override=2.03146e7)
    N, box = 128, 128 / 0.7
    kd = po.genic(N, box, 486604, cosmo.PowerSpectrum)
    res = po.compute_fmax(kd, LOGGED_RADII, box / N, cosmo.InverseGrowingMode, lpt_order=0)
    assert np.abs(np.sqrt(res["TrueVariance"]) - LOGGED_SIGMA).max() < 6e-5
    pdf = po.fmax_pdf(res["Fmax"]).astype(np.int64)
    gold = np.loadtxt(GOLDEN / "pinocchio.test.FmaxPDF.out")[:, 2].astype(np.int64)
    d = np.abs(pdf - gold)
    assert d.max() <= 6 and d.sum() <= 100, (d.max(), d.sum())
    assert abs(int(pdf[10:].sum()) - 1230386) <= 3            # log_RUN.txt:407


# BASELINE.json configs[0]: example/ as shipped.  The shipped example/log is an EH run (Omega0 = .25,
# h = .7, sigma8 = .8; 128^3, BoxSize 500 Mpc/h, 4 MPI tasks, seed 486604): its seven "computed sigma"
# values (log:161-311) and its collapsed-particle count (log:383,480-481) pin GenIC + the radius sweep
# at a second cell size (3.9 Mpc/h).  (example/pinocchio.example.FmaxPDF.out belongs to another run --
# it holds 741 412 collapsed particles against the log's 687 249 -- and is not used.)
EXAMPLE_LOG_SIGMA = [0.2761, 0.3919, 0.5555, 0.7871, 1.1067, 1.4135, 1.5929]
EXAMPLE_LOG_RADII = [20.635922, 13.996056, 9.026099, 5.465945, 3.058354, 1.548258, 0.0]      # log:55-61
EXAMPLE_LOG_COLLAPSED = 687249


def test_example_log_sigma_and_collapsed_count():
    cosmo = Cosmology(pk_norm_override=2.03146e7)           # log:50
    N, box = 128, 500.0 / 0.7                               # log:25-26
    lad = set_smoothing(cosmo, box / N)
    assert lad.Nsmooth == 7
    assert np.abs(lad.Radius - EXAMPLE_LOG_RADII).max() < 5e-5
    assert abs(lad.Variance[-1] - 3.913232) < 2e-5          # log:61
    kd = po.genic(N, box, 486604, cosmo.PowerSpectrum)
    res = po.compute_fmax(kd, EXAMPLE_LOG_RADII, box / N, cosmo.InverseGrowingMode, lpt_order=0)
    assert np.abs(np.sqrt(res["TrueVariance"]) - EXAMPLE_LOG_SIGMA).max() < 6e-5
    pdf = po.fmax_pdf(res["Fmax"]).astype(np.int64)
    assert abs(int(pdf[10:].sum()) - EXAMPLE_LOG_COLLAPSED) <= 5
    assert abs(int(pdf[:10].sum()) - 1409903) <= 5          # log:481
