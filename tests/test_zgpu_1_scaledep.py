"""-DSCALE_DEPENDENT displacements on the GPU (pinb200_displacements_scaledep): growth_rate(|k|)
evaluated per mode in the x-pass loader (src/fmax-pfft.c:340-364, src/cosmo.c:1728-1819).
Needs a B200: -m gpu.  (File name sorts after the round-1 parity tests on purpose.)
"""
from pathlib import Path

import numpy as np
import pytest

from oracle import pinocchio_oracle as po

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def cosmo():
    from pinocchio_b200.cosmology import Cosmology
    return Cosmology(pk_norm_override=2.03146e7)


def make(N, cosmo, radii, box):
    from pinocchio_b200.cosmology import SmoothingLadder
    from pinocchio_b200.engine import Pinocchio, RunConfig
    cfg = RunConfig(GridSize=N, BoxSize_htrue=box, lpt_order=3)
    return Pinocchio(cfg, cosmo, smoothing=SmoothingLadder(np.array(radii, dtype=np.float64), np.zeros(len(radii))))


def test_scale_dependent_against_reference_golden(cosmo):
    """the reference's own k loop with a k-dependent growth rate (reference_scaledep_32.npz)"""
    gold = dict(np.load(GOLDEN / "reference_fmax_32.npz"))
    sd = dict(np.load(GOLDEN / "reference_scaledep_32.npz"))
    N = int(gold["N"])
    p = make(N, cosmo, list(gold["radii"]), float(gold["box"]))
    p.set_scale_dependent_growth(lambda z: sd["log10_growth"], float(sd["logkmin"]), float(sd["dlogk"]))
    p.write_kdensity(gold["kdensity"])
    p.compute_fmax()
    for name in ("Vel", "Vel_2LPT", "Vel_3LPT_1", "Vel_3LPT_2"):
        for a in range(3):
            r = sd[name][:, a].reshape(N, N, N).astype(np.float64)
            assert np.abs(p.field(name, a).astype(np.float64) - r).max() <= 1e-6 * np.abs(r).max(), (name, a)
    # switching back gives the scale-independent fields again (same k-vectors, compute_sources = 0)
    p.set_scale_dependent_growth(None)
    p.compute_displacements(0, 0, 0.0)
    gp = gold["products"].view(po.PRODUCT_DTYPE_3LPT)
    for name in ("Vel", "Vel_3LPT_1"):
        r = gp[name][:, 1].reshape(N, N, N).astype(np.float64)
        assert np.abs(p.field(name, 1).astype(np.float64) - r).max() <= 1e-6 * np.abs(r).max(), name
    p.close()


@pytest.mark.parametrize("N,logkmin,dlogk", [(64, -3.0, 0.5), (128, -0.8, 0.15)])
def test_scale_dependent_against_oracle(N, logkmin, dlogk, cosmo):
    """shipped LOGKMIN/DELTALOGK (src/def_splines.h:41-42) and a ladder whose bins straddle the
    grid's k range (modes below kmin, in every bin); re-entry at a second segment redshift"""
    rng = np.random.default_rng(11)
    base = np.log10(np.array([0.61, 0.16, 0.05, 0.11]))[:, None]
    tabs = {0.0: base + 0.3 * np.cumsum(rng.uniform(-0.2, 0.2, (4, 10)), axis=1),
            1.0: base - 0.3 + 0.2 * np.cumsum(rng.uniform(-0.2, 0.2, (4, 10)), axis=1)}
    radii = [2.0, 0.0]
    p = make(N, cosmo, radii, N / 0.7)
    p.set_scale_dependent_growth(lambda z: tabs[z], logkmin, dlogk)
    p.GenIC_large()
    kd = p.read_kdensity()
    p.compute_fmax()
    h = po.second_derivatives(kd, 0.0, 1.0 / 0.7)
    kv = po.lpt_kvectors(h)
    for z, sources in ((0.0, None), (1.0, 0)):
        if sources is not None:
            p.compute_displacements(sources, 0, z)      # RECOMPUTE_DISPLACEMENTS re-entry, src/fragment.c:409
        g = [po.growth_rate_of_k(N, o, tabs[z], logkmin, dlogk) for o in (1, 2, 3, 4)]
        for name, kvec, growth in (("Vel", kd, g[0]), ("Vel_2LPT", kv[0], g[1]), ("Vel_3LPT_1", kv[1], g[2]),
                                   ("Vel_3LPT_2", kv[2], g[3])):
            ref = po.first_derivatives(kvec, growth)
            for a in range(3):
                assert np.abs(p.field(name, a).astype(np.float64) - ref[a]).max() <= 1e-6 * np.abs(ref[a]).max(), (z, name, a)
    p.close()


def test_recompute_sd_special_mode_3(cosmo):
    """compute_displacements(1, 1, z) (src/pinocchio.c:186, src/fmax.c:301-319): displacements without
    an Fmax sweep -- the R = 0 second derivatives are recomputed first; Fmax/Rmax stay zero."""
    N = 64
    radii = [2.0, 0.0]
    p = make(N, cosmo, radii, N / 0.7)
    p.GenIC_large()
    p.compute_fmax()
    want = {n: [p.field(n, a) for a in range(3)] for n in ("Vel", "Vel_2LPT", "Vel_3LPT_1", "Vel_3LPT_2")}
    p.close()
    q = make(N, cosmo, radii, N / 0.7)
    q.GenIC_large()
    q.compute_displacements(1, 1, 0.0)
    for n, fields in want.items():
        for a in range(3):
            assert np.array_equal(q.field(n, a), fields[a]), (n, a)
    prod = q.products()
    assert not prod["Fmax"].any() and not prod["Rmax"].any()
    assert np.array_equal(prod["Vel_2LPT"][:, 2].reshape(N, N, N), want["Vel_2LPT"][2])
    q.close()
