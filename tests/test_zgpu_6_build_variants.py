"""Compile-time variants of the reference that change what the path delivers (src/Makefile:46-68):
builds without -DTHREE_LPT / -DTWO_LPT (lpt_order 2 and 1: 32- and 20-byte product_data) and
-DDOUBLE_PRECISION_PRODUCTS (PRODFLOAT = double, src/pinocchio.h:225-231: the packer's 8-byte path,
112-byte records, members merged one by one because the 4 bytes after Rmax are padding; the library
keeps its products as float SoA, so the doubles it delivers are the floats widened).
Needs a B200: -m gpu.  Written after round 1's GPU budget was spent; dry-run on the emulated ABI by
tests/test_gpu_tests_dry_run.py, the linked programs by tests/test_dropin_emulated.py.
(File name sorts after the other parity tests.)
"""
import os
import re
import subprocess

import numpy as np
import pytest

from test_reference_full import GOLDEN, REF_X, load_catalog, match_fraction

pytestmark = pytest.mark.gpu


def test_double_records_are_the_float_records_widened():
    from pinocchio_b200.cosmology import Cosmology, SmoothingLadder
    from pinocchio_b200.engine import PRODUCT_DTYPE_3LPT, PRODUCT_DTYPE_3LPT_DOUBLE, Pinocchio, RunConfig
    N = 64
    p = Pinocchio(RunConfig(GridSize=N, BoxSize_htrue=N / 0.7, lpt_order=3), Cosmology(pk_norm_override=2.03146e7),
                  smoothing=SmoothingLadder(np.array([6.0, 2.5, 0.0]), np.zeros(3)))
    p.GenIC_large()
    p.compute_fmax()
    a = p.products(dtype=PRODUCT_DTYPE_3LPT)
    b = p.products(dtype=PRODUCT_DTYPE_3LPT_DOUBLE)
    assert a.dtype.itemsize == 56 and b.dtype.itemsize == 112
    assert np.array_equal(a["Rmax"], b["Rmax"])
    for name in ("Fmax", "Vel", "Vel_2LPT", "Vel_3LPT_1", "Vel_3LPT_2"):
        assert np.array_equal(a[name].astype(np.float64), b[name]), name
    # the padding after Rmax is not a member: the caller's bytes there survive the download
    raw = np.full(1000 * 112, 0xAB, dtype=np.uint8)
    out = raw.view(PRODUCT_DTYPE_3LPT_DOUBLE)
    f = PRODUCT_DTYPE_3LPT_DOUBLE.fields
    from pinocchio_b200.engine import ProductLayout
    import ctypes
    lay = ProductLayout(112, 8, f["Rmax"][1], f["Fmax"][1], f["Vel"][1], f["Vel_2LPT"][1], f["Vel_3LPT_1"][1], f["Vel_3LPT_2"][1])
    p._ck(p.lib.pinb200_download_products(p.h, out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(lay), 4321, 1000))
    assert np.array_equal(out, b[4321:5321])
    assert (raw.reshape(1000, 112)[:, 4:8] == 0xAB).all()
    p.close()


@pytest.mark.parametrize("tag", ["dp", "lpt2", "zel", "rad"])
def test_linked_dropin_build_variants(tag, tmp_path):
    bx, rx = REF_X.parent / f"pinocchio_b200_{tag}.x", REF_X.parent / f"pinocchio_ref_{tag}.x"
    if not (bx.exists() and rx.exists()):
        pytest.skip(f"{tag} variants not built (make -C oracle all)")

    def run(exe, d, threads):
        d.mkdir(parents=True, exist_ok=True)
        text = (GOLDEN / "parameter_file").read_text()
        text = re.sub(r"(?m)^BoxSize\s+\S+", "BoxSize                64", text)
        text = re.sub(r"(?m)^GridSize\s+\S+", "GridSize               64", text)
        text = re.sub(r"(?m)^MaxMemPerParticle\s+\S+", "MaxMemPerParticle      400", text)
        (d / "parameter_file").write_text(text)
        (d / "outputs").write_bytes((GOLDEN / "outputs").read_bytes())
        r = subprocess.run([str(exe), "parameter_file"], cwd=d, capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, OMP_NUM_THREADS=str(threads)))
        assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
        return r.stdout

    da, db = tmp_path / f"b200_{tag}", tmp_path / f"ref_{tag}"
    log_b = run(bx, da, 8)
    run(rx, db, 16)
    assert "B200 path" in log_b and (tag != "dp" or "Products in double precision" in log_b)
    pa = np.loadtxt(da / "pinocchio.test.FmaxPDF.out")[:, 2].astype(np.int64)
    pb = np.loadtxt(db / "pinocchio.test.FmaxPDF.out")[:, 2].astype(np.int64)
    assert pa.sum() == pb.sum() == 64 ** 3 and np.abs(pa - pb).sum() <= 60
    for z in ("0.0000", "2.0000"):
        ia, na, _ = load_catalog(da / f"pinocchio.{z}.test.catalog.out")
        ib, nb, _ = load_catalog(db / f"pinocchio.{z}.test.catalog.out")
        assert abs(len(ia) - len(ib)) <= max(2, 0.005 * len(ib))
        assert match_fraction(ia, na, ib, nb) > 0.99


@pytest.mark.parametrize("order", [1, 2])
def test_lower_lpt_orders(order):
    """builds without -DTHREE_LPT (order 2) or without -DTWO_LPT (order 1): same Fmax / Rmax as the full
    build, displacement fields up to that order against the oracle, the higher ones not delivered"""
    from oracle import pinocchio_oracle as po
    from pinocchio_b200.cosmology import Cosmology, SmoothingLadder
    from pinocchio_b200.engine import Pinocchio, PinocchioError, RunConfig
    N, radii = 64, [6.0, 2.5, 0.0]
    cosmo = Cosmology(pk_norm_override=2.03146e7)
    lad = SmoothingLadder(np.array(radii), np.zeros(3))
    full = Pinocchio(RunConfig(GridSize=N, BoxSize_htrue=N / 0.7, lpt_order=3), cosmo, smoothing=lad)
    full.GenIC_large()
    full.compute_fmax()
    kd = full.read_kdensity()
    F3, R3, V3 = full.field("Fmax"), full.field("Rmax"), [full.field("Vel", a) for a in range(3)]
    full.close()
    p = Pinocchio(RunConfig(GridSize=N, BoxSize_htrue=N / 0.7, lpt_order=order), cosmo, smoothing=lad)
    p.GenIC_large()
    p.compute_fmax()
    assert np.array_equal(p.field("Fmax"), F3) and np.array_equal(p.field("Rmax"), R3)
    for a in range(3):
        assert np.array_equal(p.field("Vel", a), V3[a])
    ref = po.compute_fmax(kd, radii, 1.0 / 0.7, cosmo.InverseGrowingMode, growth=tuple(p.growth_rates(0.0)), lpt_order=order)
    names = ["Vel"] + (["Vel_2LPT"] if order >= 2 else [])
    for name in names:
        for a in range(3):
            scale = np.abs(ref[name][a]).max()
            assert np.abs(p.field(name, a).astype(np.float64) - ref[name][a]).max() <= 1e-6 * scale, (name, a)
    for name in (["Vel_2LPT"] if order < 2 else []) + ["Vel_3LPT_1", "Vel_3LPT_2"]:
        with pytest.raises(PinocchioError):
            p.field(name, 0)
    prod = p.products()                      # 56-byte layout asked for: the absent members stay zero
    assert np.array_equal(prod["Fmax"].reshape(N, N, N), F3)
    assert not prod["Vel_3LPT_1"].any() and not prod["Vel_3LPT_2"].any()
    assert prod["Vel_2LPT"].any() == (order >= 2)
    p.close()


def test_seed_plane_given_by_the_caller():
    """pinb200_set_seed_plane (the MimicOldSeed hand-over of the shim): the spiral table given explicitly
    reproduces the default delta_k bit for bit; another table gives another realisation; wrong sizes are refused"""
    from oracle import pinocchio_oracle as po
    from pinocchio_b200.cosmology import Cosmology, SmoothingLadder
    from pinocchio_b200.engine import Pinocchio, PinocchioError, RunConfig
    N = 64
    cosmo = Cosmology(pk_norm_override=2.03146e7)

    def make():
        return Pinocchio(RunConfig(GridSize=N, BoxSize_htrue=N / 0.7, lpt_order=3), cosmo,
                         smoothing=SmoothingLadder(np.array([2.5, 0.0]), np.zeros(2)))
    p = make()
    p.GenIC_large()
    kd0 = p.read_kdensity()
    p.close()
    table = np.asarray(po.seed_table(N, 486604), dtype=np.uint32).reshape(N, N)
    p = make()
    with pytest.raises(PinocchioError):
        p.set_seed_plane(table[:-1])
    p.set_seed_plane(table)
    p.GenIC_large()
    assert np.array_equal(p.read_kdensity(), kd0)
    p.set_seed_plane(po.seed_table_old(N, 486604))                 # the MimicOldSeed plane: another realisation
    p.GenIC_large()
    assert not np.array_equal(p.read_kdensity(), kd0)
    p.close()
