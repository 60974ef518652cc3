"""BASELINE.json configs[0] on the GPU: example/ as shipped (128^3, BoxSize 500 Mpc/h, seed 486604,
EH power spectrum): the seven "computed sigma" values of example/log:161-311 and its
collapsed-particle count (log:383).  Needs a B200: -m gpu."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
EXAMPLE_LOG_SIGMA = [0.2761, 0.3919, 0.5555, 0.7871, 1.1067, 1.4135, 1.5929]
EXAMPLE_LOG_RADII = [20.635922, 13.996056, 9.026099, 5.465945, 3.058354, 1.548258, 0.0]


def test_example_config_sigma_and_collapsed_count():
    from pinocchio_b200.cosmology import Cosmology, SmoothingLadder
    from pinocchio_b200.engine import Pinocchio, RunConfig
    cosmo = Cosmology(pk_norm_override=2.03146e7)
    N, box = 128, 500.0 / 0.7
    cfg = RunConfig(GridSize=N, BoxSize_htrue=box, lpt_order=3)
    p = Pinocchio(cfg, cosmo, smoothing=SmoothingLadder(np.array(EXAMPLE_LOG_RADII), np.zeros(7)))
    p.GenIC_large()
    p.compute_fmax()
    assert np.abs(np.sqrt(p.TrueVariance) - EXAMPLE_LOG_SIGMA).max() < 6e-5
    pdf = p.Fmax_PDF().astype(np.int64)
    assert pdf.sum() == N ** 3
    assert abs(int(pdf[10:].sum()) - 687249) <= 5
    # the four output redshifts of the example (outputs: 2, 1, 0.5, 0): RECOMPUTE-style re-entry keeps working
    v0 = p.field("Vel", 0).astype(np.float64)
    p.compute_displacements(0, 0, 2.0)
    ratio = cosmo.GrowingMode(2.0) / cosmo.GrowingMode(0.0)
    assert np.abs(p.field("Vel", 0).astype(np.float64) - ratio * v0).max() <= 2e-7 * np.abs(v0).max()
    p.close()


def test_genic_against_reference_code_golden():
    """CUDA GenIC against kdensity from the REFERENCE'S OWN GenIC_large (src/GenIC.c compiled verbatim
    into oracle/_ref; fixture tests/golden/reference_genic_32.npz, make_reference_genic_golden.py)"""
    from pathlib import Path
    from pinocchio_b200.cosmology import Cosmology, SmoothingLadder
    from pinocchio_b200.engine import Pinocchio, RunConfig
    g = dict(np.load(Path(__file__).resolve().parent / "golden" / "reference_genic_32.npz"))
    N = int(g["N"])
    cosmo = Cosmology(pk_norm_override=float(g["pk_norm"]))
    for key, seed, fixed, paired in (("kd_486604", 486604, 0, 0), ("kd_12345_fixed_paired", 12345, 1, 1)):
        cfg = RunConfig(GridSize=N, BoxSize_htrue=float(g["box"]), RandomSeed=seed, FixedIC=fixed, PairedIC=paired)
        p = Pinocchio(cfg, cosmo, smoothing=SmoothingLadder(np.array([0.0]), np.zeros(1)))
        p.GenIC_large()
        kd = p.read_kdensity()
        ref = g[key]
        assert np.array_equal(kd != 0, ref != 0)
        assert np.abs(kd - ref).max() <= 1e-13 * np.abs(ref).max()
        p.close()
