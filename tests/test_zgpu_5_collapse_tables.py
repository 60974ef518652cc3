"""-DTABULATED_CT / -DELL_SNG on the B200, through the C ABI: the table kernels
(ell_classic or one rkf45 ellipsoid integration per table point), the spline records, the per-cell look-up
as a stand-alone kernel and as the epilogue of the collapse z pass, and the linked drop-in programs
oracle/_ref/pinocchio_b200_{tab,sng,fr}.x against the outputs of the reference program compiled with the
same flags (tests/golden/reference_ct_32.npz).  The CPU-side counterpart, on the same kernel source under
the emulator, is tests/test_collapse_tables.py.  Needs a B200: -m gpu.
"""
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import pinocchio_oracle as po

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden" / "reference_ct_32.npz"
REF = ROOT / "oracle" / "_ref"
N = 32
ND, NXY = po.CT_NBINS_D, po.CT_NBINS_XY
NPOINTS = ND * NXY * NXY


def relerr(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-3)


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def pin():
    """32^3 box of 32 Mpc/h: 1 Mpc/h cells, the nine-radius ladder of the golden run"""
    from pinocchio_b200.cosmology import Cosmology
    from pinocchio_b200.engine import Pinocchio, RunConfig
    c = Cosmology()
    p = Pinocchio(RunConfig(GridSize=N, BoxSize_htrue=N / 0.7), c)
    assert p.Smoothing.Nsmooth == 9
    yield p
    p.close()


def test_delta_vector_host_entry():
    from pinocchio_b200.engine import Pinocchio
    assert np.array_equal(Pinocchio.ct_delta_vector(), po.ct_delta_vector())


def test_ell_sng_tables(pin, gold):
    """the batch ODE kernel: every 97th point of a table against the reference program's, and against the
    NumPy restatement of GSL's rkf45 on identical inputs"""
    from pinocchio_b200.engine import CT_SNG
    c = pin.cosmo
    pin.initialize_collapse_times(CT_SNG)
    import os
    if not os.environ.get("PINB200_LIB"):                  # the emulated ABI (dry runs of this file on a CPU) keeps no timers
        assert pin.timers().coll > 0
    idx = gold["sng_table_idx"]
    D_in = c.GrowingMode(1.0 / 1.0e-5 - 1.0)
    for ism in (0, 3, 8):
        t = pin.collapse_table(ism).ravel()
        # the host cosmology of pinocchio_b200.cosmology agrees with the program's to ~5e-6: a point that collapses
        # just before a = 5 may fall on the other side
        assert abs(int((t != 0).sum()) - int(gold["sng_nonzero_per_radius"][ism])) <= 5
        both = (t[idx] != 0) & (gold["sng_table"][ism] != 0)
        assert (both != (gold["sng_table"][ism] != 0)).sum() <= 2
        assert relerr(t[idx], gold["sng_table"][ism])[both].max() < 2e-5
        l1, l2, l3 = [a.ravel() for a in po.ct_table_lambdas(np.sqrt(pin.Smoothing.Variance[ism]))]
        for k in range(7, idx.size, 301):
            a = po.ell_sng(l1[idx[k]], l2[idx[k]], l3[idx[k]], D_in, c.p.Omega0, c.p.OmegaLambda, c.OmegaRad)
            F = 1.0 / a if a > 0 else 0.0
            assert abs(F - t[idx[k]]) <= 1e-7 * max(F, 1e-3), (ism, idx[k])   # FMA contraction: 5e-10 on the CPU


def test_classic_tables_lookup_and_fmax(pin):
    """ELL_CLASSIC tables against the oracle's; the per-cell look-up against the oracle's interpolation of the
    SAME (downloaded) table; compute_fmax with tables against the oracle's radius loop"""
    from pinocchio_b200.engine import CT_CLASSIC
    c = pin.cosmo
    pin.initialize_collapse_times(CT_CLASSIC)
    dv = po.ct_delta_vector()
    tables = [pin.collapse_table(i) for i in range(9)]
    ref5 = po.ct_table_classic(np.sqrt(pin.Smoothing.Variance[5]), c.InverseGrowingMode)
    e = relerr(tables[5], ref5)
    assert np.array_equal(tables[5] == 0, ref5 == 0)
    assert (e > 1e-9).sum() < 100 and (e > 1e-6).sum() < 10          # ell_classic near den = 0

    rng = np.random.default_rng(3)
    h = rng.normal(0, 1.2, (6, 40000))
    ampl = float(np.sqrt(pin.Smoothing.Variance[6]))
    h[:, :8] = 0.0
    h[:3, :8] = rng.normal(0, 1, 8)                          # isotropic tensors: q == 0, the diagonal branch
    F = pin.inverse_collapse_time(h, ismooth=6)
    Fr = po.inverse_collapse_time_tab(h, tables[6], dv, ampl)
    assert np.array_equal(F == -10.0, Fr == -10.0)
    assert np.abs(F - Fr).max() <= 1e-9 * max(1.0, np.abs(Fr).max())

    pin.GenIC_large()
    kd = pin.read_kdensity()
    pin.compute_fmax(displacements=False)
    Fmax, Rmax = po.init_products((N, N, N))
    Fs = []
    for ism in range(9):
        hh = po.second_derivatives(kd, pin.Smoothing.Radius[ism], pin.CellSize)
        Fnew = po.inverse_collapse_time_tab(hh, tables[ism], dv, float(np.sqrt(pin.Smoothing.Variance[ism])))
        Fs.append(Fnew)
        po.update_fmax(Fmax, Rmax, Fnew, ism)           # in place
    got_F, got_R = pin.field("Fmax"), pin.field("Rmax")
    dF = np.abs(got_F.astype(np.float64) - Fmax.astype(np.float64))
    assert (dF <= 1e-6 * np.maximum(1.0, np.abs(Fmax))).all()
    top2 = np.sort(np.stack(Fs), axis=0)[-2:]
    ties = np.abs(top2[1] - top2[0]) <= 1e-6 * np.maximum(1.0, np.abs(top2[1]))
    assert not ((got_R != Rmax) & ~ties).any()
    assert pin.Fmax_PDF().sum() == N ** 3

    # tables supplied by the caller (a CTtableFile): same result
    before = got_F.copy()
    pin.initialize_collapse_times(CT_CLASSIC, tables=np.stack(tables))
    assert np.array_equal(pin.collapse_table(4), tables[4])
    pin.compute_fmax(displacements=False)
    assert np.array_equal(pin.field("Fmax"), before)
    # back to the direct evaluation: the table is an approximation of it
    pin.initialize_collapse_times(None)
    pin.compute_fmax(displacements=False)
    direct = pin.field("Fmax").astype(np.float64)
    coll = direct > 1.0
    assert np.median(np.abs(before[coll] - direct[coll]) / direct[coll]) < 5e-3


def test_error_paths(pin):
    from pinocchio_b200.engine import PinocchioError
    pin.initialize_collapse_times(None)
    with pytest.raises(PinocchioError, match="collapse tables not set"):
        pin._ct_shape = (NXY, NXY, ND)
        pin.collapse_table(0)
    with pytest.raises(PinocchioError, match="model must be"):
        pin.initialize_collapse_times(2)
    with pytest.raises(PinocchioError, match="strictly increasing"):
        pin.initialize_collapse_times(1, delta_vector=np.zeros(ND))


def _run32(exe: Path, workdir: Path) -> str:
    import os
    import subprocess
    workdir.mkdir(parents=True, exist_ok=True)
    text = (ROOT / "tests" / "golden" / "hmf_validation" / "parameter_file").read_text()
    text = re.sub(r"(?m)^BoxSize\s+\S+", "BoxSize                32", text)
    text = re.sub(r"(?m)^GridSize\s+\S+", "GridSize               32", text)
    text = re.sub(r"(?m)^MaxMemPerParticle\s+\S+", "MaxMemPerParticle      400", text)
    (workdir / "parameter_file").write_text(text + "\nCTtableFile none\n")
    (workdir / "outputs").write_bytes((ROOT / "tests" / "golden" / "hmf_validation" / "outputs").read_bytes())
    r = subprocess.run([str(exe), "parameter_file"], cwd=workdir, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, OMP_NUM_THREADS="8"))
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    return r.stdout


def _catalog(raw: bytes):
    rows = [ln.split() for ln in raw.decode().splitlines() if ln.strip() and not ln.lstrip().startswith("#")]
    a = np.array(rows, dtype=np.float64)
    return a[:, 0].astype(np.int64), a[:, 11].astype(np.int64)


@pytest.mark.parametrize("tag", ["tab", "sng", "fr"])
def test_linked_dropin_program(tag, gold, tmp_path):
    """reference host code + shim + libpinb200.so with -DTABULATED_CT (and -DELL_SNG) against the reference
    program's own outputs for the same parameter file"""
    exe = REF / f"pinocchio_b200_{tag}.x"
    if not exe.exists() or f"{tag}_header" not in gold.files:
        pytest.skip(f"{exe.name} not built (make -C oracle all) or no golden outputs for it")
    log = _run32(exe, tmp_path)
    assert "B200 path" in log and "Collapse times computed for interpolation" in log and "Pinocchio done!" in log
    sig = np.array([float(x) for x in re.findall(r"computed sigma:\s+([0-9.]+)", log)])
    ns = gold[f"{tag}_sigma"].size                     # 9 radii; 10 with the f(R) growth
    assert sig.size == ns and np.abs(sig - gold[f"{tag}_sigma"]).max() <= 1e-4
    raw = (tmp_path / "pinocchio.test.CTtable.out").read_bytes()
    assert raw[:40] == gold[f"{tag}_header"].tobytes() and len(raw) == 40 + ns * (4 + 8 * NPOINTS)
    tabs = np.array([np.frombuffer(raw[40 + i * (4 + 8 * NPOINTS) + 4:40 + (i + 1) * (4 + 8 * NPOINTS)], dtype=np.float64)
                     for i in range(ns)])
    assert np.array_equal((tabs != 0).sum(axis=1), gold[f"{tag}_nonzero_per_radius"])
    e = relerr(tabs[:, gold[f"{tag}_table_idx"]], gold[f"{tag}_table"])
    if tag == "sng":
        assert e.max() < 1e-7
    elif tag == "fr":
        # kinks of the force modification: see tests/test_collapse_tables.py
        assert (e > 1e-6).mean() < 2e-3 and e.max() < 5e-2
    else:
        assert (e > 1e-9).sum() <= 30 and e.max() < 1e-3
    pdf = np.loadtxt(tmp_path / "pinocchio.test.FmaxPDF.out")[:, 2].astype(np.int64)
    gpdf = np.array([ln.split()[2] for ln in gold[f"{tag}_file_pinocchio.test.FmaxPDF.out"].tobytes().decode().splitlines()
                     if ln.strip() and not ln.startswith("#")], dtype=np.float64).astype(np.int64)
    assert pdf.sum() == gpdf.sum() == N ** 3
    assert np.abs(pdf - gpdf).max() <= 3 and np.abs(pdf - gpdf).sum() <= 20
    for z in ("0.0000", "2.0000"):
        name = f"pinocchio.{z}.test.catalog.out"
        ia, na = _catalog((tmp_path / name).read_bytes())
        ib, nb = _catalog(gold[f"{tag}_file_{name}"].tobytes())
        da = dict(zip(ia.tolist(), na.tolist()))
        same = sum(1 for i, n in zip(ib.tolist(), nb.tolist()) if da.get(i) == n)
        assert abs(len(ia) - len(ib)) <= 3 and same >= 0.97 * len(ib)
