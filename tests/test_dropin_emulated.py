"""The linked drop-in, end to end WITHOUT a GPU: oracle/_ref/pinocchio_emu.x is the unchanged
reference program with fmax.c, fmax-pfft.c, collapse_times.c, LPT.c and GenIC.c replaced by
shim/fmax_b200.c, linked against the TEST-ONLY emulated ABI (tests/host/emu_abi.cpp: the C ABI of
include/pinb200.h on host arrays, running the very kernel bodies of kernels.cuh under the pthread
block emulator with the engine's schedule).  It is compared with oracle/_ref/pinocchio_ref.x, the
same program with the reference's own five files, on a 32^3 version of HMF_Validation/parameter_file.

What this pins on a CPU-only box: the shim's run-time logic (order of calls, P(k) lattice table,
smoothing ladder, inverse-growth spline knots taken from the reference's gsl_spline, growth factors,
units, the products[] download into the reference's arena) and the hand-over to the unchanged
fragmentation -- BASELINE.json's "halo catalogue and mass function produced by the unchanged
fragmentation stage must match the reference".  tests/test_zgpu_4_dropin_catalogues.py repeats it at
128^3 with the real libpinb200.so on the B200.
"""
import re

import numpy as np
import pytest

from test_reference_full import GOLDEN, REF_X, load_catalog, match_fraction

EMU_X = REF_X.parent / "pinocchio_emu.x"
pytestmark = pytest.mark.skipif(not (REF_X.exists() and EMU_X.exists()),
                                reason="oracle/_ref/pinocchio_{ref,emu}.x not built (make -C oracle all)")
N = 32
# "": the default build (-DTWO_LPT -DTHREE_LPT -DELL_CLASSIC -DNORADIATION).
# "_sd": the same + -DSCALE_DEPENDENT -DRECOMPUTE_DISPLACEMENTS -DPLC -- growth rates per k bin handed to
# pinb200_displacements_scaledep, one inverse-growth spline per smoothing radius, displacements
# recomputed for each of the four redshift segments (compute_displacements(0,0,z), src/fragment.c:409)
# and the *_prev members of the 104-byte product_data, which belong to the fragmentation, preserved
# by the member-wise download (product_merge.h).
VARIANTS = ["", "_sd"]


def run32(exe, workdir, variant=""):
    import os
    import subprocess
    workdir.mkdir(parents=True, exist_ok=True)
    text = (GOLDEN / "parameter_file").read_text()
    text = re.sub(r"(?m)^BoxSize\s+\S+", f"BoxSize                {N}", text)
    text = re.sub(r"(?m)^GridSize\s+\S+", f"GridSize               {N}", text)
    if variant in ("_sd", "_dp"):                       # 104- / 112-byte records (+ fields kept for the re-entry)
        text = re.sub(r"(?m)^MaxMemPerParticle\s+\S+", "MaxMemPerParticle      400", text)
    (workdir / "parameter_file").write_text(text)
    (workdir / "outputs").write_bytes((GOLDEN / "outputs").read_bytes())
    r = subprocess.run([str(exe), "parameter_file"], cwd=workdir, capture_output=True, text=True, timeout=1500,
                       env=dict(os.environ, OMP_NUM_THREADS="4"))
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    return r.stdout


@pytest.fixture(scope="module", params=VARIANTS)
def runs(request, tmp_path_factory):
    v = request.param
    emu, ref = REF_X.parent / f"pinocchio_emu{v}.x", REF_X.parent / f"pinocchio_ref{v}.x"
    if not (emu.exists() and ref.exists()):
        pytest.skip(f"{emu.name} / {ref.name} not built")
    a = tmp_path_factory.mktemp("emu" + v)
    b = tmp_path_factory.mktemp("ref" + v)
    return a, run32(emu, a, v), b, run32(ref, b, v), v


def test_emulated_dropin_log(runs):
    a, log_a, b, log_b, v = runs
    assert "B200 path" in log_a and "Pinocchio done!" in log_a and "B200 path" not in log_b
    # products[] filled from the device-side selection + sort (shim: download_products_compact) unless the build's
    # records carry members of other stages (RECOMPUTE_DISPLACEMENTS: the _sd variant)
    assert ("compact hand-off" in log_a) == (v != "_sd")
    sig = lambda log: [float(x) for x in re.findall(r"computed sigma:\s+([0-9.]+)", log)]
    assert len(sig(log_a)) == 9 and sig(log_a) == sig(log_b)
    ncoll = lambda log: int(re.search(r"Number of collapsed particles to z=0: (\d+)", log).group(1))
    assert ncoll(log_a) == ncoll(log_b) > 0.4 * N ** 3
    # -DRECOMPUTE_DISPLACEMENTS: both programs recompute the displacements for segments 2..4
    nre = lambda log: len(re.findall(r"Computing displacements for redshift", log))
    assert nre(log_a) == nre(log_b)


def test_emulated_dropin_fmaxpdf_identical(runs):
    a, _, b, _, _ = runs
    pa = np.loadtxt(a / "pinocchio.test.FmaxPDF.out")[:, 2]
    pb = np.loadtxt(b / "pinocchio.test.FmaxPDF.out")[:, 2]
    assert pa.sum() == N ** 3
    assert np.abs(pa - pb).max() <= 1          # float-rounding flips of Fmax at a bin edge at most


def test_emulated_dropin_past_light_cone(runs):
    """the _sd variant is also built with -DPLC (BASELINE.json configs[4]: scale-dependent growth with
    past light-cone output): the light-cone catalogue and n(z) of the two programs are the same bytes"""
    a, _, b, _, _ = runs
    plc = list(b.glob("*.plc.out"))
    if not plc:
        pytest.skip("variant built without -DPLC")
    for f in plc + list(b.glob("*.nz.out")):
        assert (a / f.name).read_bytes() == f.read_bytes(), f.name
    assert plc[0].stat().st_size > 1000


@pytest.mark.parametrize("z", ["0.0000", "0.5000", "1.0000", "2.0000"])
def test_emulated_dropin_catalogues_match(runs, z):
    a, _, b, _, _ = runs
    ia, na, ca = load_catalog(a / f"pinocchio.{z}.test.catalog.out")
    ib, nb, cb = load_catalog(b / f"pinocchio.{z}.test.catalog.out")
    assert len(ib) > 20
    assert len(ia) == len(ib)
    assert match_fraction(ia, na, ib, nb) >= 0.99
    if np.array_equal(ia, ib):
        # same halos in the same order: positions (Mpc/h) and velocities (km/s) to the printed digits
        assert np.abs(ca[:, 2:8] - cb[:, 2:8]).max() <= 0.011
        assert np.abs(ca[:, 8:11] - cb[:, 8:11]).max() <= 0.2


def test_emulated_dropin_mass_function_matches(runs):
    a, _, b, _, _ = runs
    ma = np.loadtxt(a / "pinocchio.0.0000.test.mf.out")
    mb = np.loadtxt(b / "pinocchio.0.0000.test.mf.out")
    assert np.array_equal(ma[:, 4], mb[:, 4])          # halos per mass bin


# ---- -DSNAPSHOT builds: product_data with zacc / group_ID; special mode 3 ---------------------------
SNAP_REF, SNAP_EMU = REF_X.parent / "pinocchio_ref_snap.x", REF_X.parent / "pinocchio_emu_snap.x"
needs_snap = pytest.mark.skipif(not (SNAP_REF.exists() and SNAP_EMU.exists()), reason="SNAPSHOT variants not built")


def run32_args(exe, workdir, extra_param_lines=(), args=()):
    import os
    import subprocess
    workdir.mkdir(parents=True, exist_ok=True)
    text = (GOLDEN / "parameter_file").read_text()
    text = re.sub(r"(?m)^BoxSize\s+\S+", f"BoxSize                {N}", text)
    text = re.sub(r"(?m)^GridSize\s+\S+", f"GridSize               {N}", text)
    (workdir / "parameter_file").write_text(text + "".join(l + "\n" for l in extra_param_lines))
    (workdir / "outputs").write_bytes((GOLDEN / "outputs").read_bytes())
    r = subprocess.run([str(exe), "parameter_file", *args], cwd=workdir, capture_output=True, text=True, timeout=1500,
                       env=dict(os.environ, OMP_NUM_THREADS="4"))
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    return r.stdout


def differing_bytes(a, b):
    x, y = np.frombuffer(a.read_bytes(), dtype=np.uint8), np.frombuffer(b.read_bytes(), dtype=np.uint8)
    assert x.size == y.size
    return int((x != y).sum())


@needs_snap
def test_emulated_dropin_special_mode_3_lpt_snapshot(tmp_path):
    """`pinocchio.x parameter_file 3` (src/pinocchio.c:170-200): compute_displacements(1, 1, z) -- the shim's
    recompute_sd path, pinb200_second_derivatives(pinb, 0, NULL) -- then write_LPT_snapshot: positions and
    velocities of every particle from the 3LPT displacements, compared as bytes"""
    a, b = tmp_path / "emu", tmp_path / "ref"
    log = run32_args(SNAP_EMU, a, args=("3",))
    run32_args(SNAP_REF, b, args=("3",))
    assert "only produce a GADGET snapshot" in log
    fa = next(a.glob("*.LPT_snapshot.out"))
    fb = b / fa.name
    assert fa.stat().st_size > 12 * 4 * N ** 3 // 2
    # float products differ by last-bit rounding flips in ~1e-4 of the values at most
    assert differing_bytes(fa, fb) <= 2e-4 * fa.stat().st_size


@needs_snap
def test_emulated_dropin_snapshot_build_and_timeless_snapshot(tmp_path):
    """-DSNAPSHOT run with WriteTimelessSnapshot: the 64-byte records carry zacc / group_ID, which belong to
    the fragmentation (member-wise download); FMAX, ZEL, 2LPT, 31PT, 32PT, ZACC, GRUP blocks of the timeless
    snapshot (src/write_snapshot.c:207-345) and all catalogues against the reference's"""
    a, b = tmp_path / "emu", tmp_path / "ref"
    run32_args(SNAP_EMU, a, extra_param_lines=("WriteTimelessSnapshot",))
    run32_args(SNAP_REF, b, extra_param_lines=("WriteTimelessSnapshot",))
    for z in ("0.0000", "2.0000"):
        assert (a / f"pinocchio.{z}.test.catalog.out").read_bytes() == (b / f"pinocchio.{z}.test.catalog.out").read_bytes()
    ts = a / "pinocchio.test.t_snapshot.out"
    assert ts.stat().st_size > 14 * 4 * N ** 3
    assert differing_bytes(ts, b / ts.name) <= 2e-4 * ts.stat().st_size


def test_emulated_dropin_dump_products_file_boundary(tmp_path):
    """DumpProducts / ReadProductsFromDumps (src/fmax.c:372-506, src/pinocchio.c:220-234) across the two
    programs: the drop-in dumps products[], the unchanged reference resumes from the dump and fragments;
    and the other way round.  Catalogues must equal those of the straight runs."""
    straight = tmp_path / "straight"
    run32_args(REF_X, straight)
    want = (straight / "pinocchio.0.0000.test.catalog.out").read_bytes()
    for writer, reader in ((EMU_X, REF_X), (REF_X, EMU_X)):
        d = tmp_path / f"{writer.stem}_to_{reader.stem}"
        run32_args(writer, d, extra_param_lines=("DumpProducts",))
        dumps = d / "DumpProducts"
        assert (dumps / "Task.0").stat().st_size == 56 * N ** 3 and (dumps / "summary").exists()
        for f in d.glob("pinocchio.*"):
            f.unlink()
        log = run32_args(reader, d, extra_param_lines=("ReadProductsFromDumps",))
        assert "B200 path" not in log or reader == EMU_X
        assert (d / "pinocchio.0.0000.test.catalog.out").read_bytes() == want


# ---- -DDOUBLE_PRECISION_PRODUCTS: PRODFLOAT = double (src/pinocchio.h:225-231) ---------------------
DP_REF, DP_EMU = REF_X.parent / "pinocchio_ref_dp.x", REF_X.parent / "pinocchio_emu_dp.x"


@pytest.mark.skipif(not (DP_REF.exists() and DP_EMU.exists()), reason="DOUBLE_PRECISION_PRODUCTS variants not built")
def test_emulated_dropin_double_precision_products(tmp_path):
    """112-byte records of doubles (prodfloat_bytes = 8 in the packer).  The library keeps its products as
    float SoA, so the doubles it delivers carry float precision (6e-8 relative, inside the 1e-6 contract)
    while the reference's carry the full double: catalogues, mass functions and the FmaxPDF still come out
    byte for byte, the merger histories to the printed digits but for rounding flips of the last one."""
    a, b = tmp_path / "emu", tmp_path / "ref"
    log = run32(DP_EMU, a, "_dp")
    run32(DP_REF, b, "_dp")
    assert "B200 path" in log
    for z in ("0.0000", "0.5000", "1.0000", "2.0000"):
        for kind in ("catalog", "mf"):
            name = f"pinocchio.{z}.test.{kind}.out"
            assert (a / name).read_bytes() == (b / name).read_bytes(), name
    assert (a / "pinocchio.test.FmaxPDF.out").read_bytes() == (b / "pinocchio.test.FmaxPDF.out").read_bytes()
    def branches(path):       # the nine-column branch lines (tree headers and counts have fewer)
        rows = [l.split() for l in path.read_text().splitlines() if not l.startswith("#")]
        return np.array([r for r in rows if len(r) == 9], dtype=np.float64)

    ha, hb = branches(a / "pinocchio.test.histories.out"), branches(b / "pinocchio.test.histories.out")
    assert ha.shape[0] > 100
    assert ha.shape == hb.shape
    assert np.array_equal(ha[:, :6], hb[:, :6])                 # tree structure: identical
    assert np.abs(ha[:, 6:] - hb[:, 6:]).max() <= 1.5e-4        # redshifts printed with four decimals
    assert (ha != hb).sum() <= 5


# ---- builds without -DTHREE_LPT / -DTWO_LPT (src/Makefile:46-47): lpt_order 2 and 1; and the shipped
#      default without -DNORADIATION (src/Makefile:85: radiation in the host cosmology) -------------
@pytest.mark.parametrize("tag,record_bytes", [("lpt2", 32), ("zel", 20), ("rad", 56)])
def test_emulated_dropin_lower_lpt_orders_and_radiation(tag, record_bytes, tmp_path):
    """lpt2 / zel: product_data shrinks to {Rmax, Fmax, Vel[, Vel_2LPT]}; the shim passes lpt_order 2 / 1 and the
    layout of the smaller record.  rad: growth tables and inverse-growth spline of a cosmology with radiation.  Every output file byte for byte, DumpProducts records of the right size and equal but for float rounding flips."""
    ref, emu = REF_X.parent / f"pinocchio_ref_{tag}.x", REF_X.parent / f"pinocchio_emu_{tag}.x"
    if not (ref.exists() and emu.exists()):
        pytest.skip(f"{ref.name} / {emu.name} not built")
    a, b = tmp_path / "emu", tmp_path / "ref"
    log = run32_args(emu, a, extra_param_lines=("DumpProducts",))
    run32_args(ref, b, extra_param_lines=("DumpProducts",))
    assert "B200 path" in log
    names = sorted(f.name for f in b.glob("pinocchio.*"))
    assert len(names) >= 11
    for name in names:
        assert (a / name).read_bytes() == (b / name).read_bytes(), name
    assert (a / "DumpProducts" / "Task.0").stat().st_size == record_bytes * N ** 3
    # raw float products: last-bit rounding flips in ~1e-4 of the values at most
    assert differing_bytes(a / "DumpProducts" / "Task.0", b / "DumpProducts" / "Task.0") <= 2e-4 * record_bytes * N ** 3


# ---- parameter-file options that reach the path (src/ReadParamfile.c:207-255) ---------------------
@pytest.mark.parametrize("lines", [("FixedIC", "PairedIC"), ("MimicOldSeed",)], ids=["fixed_paired", "mimic_old_seed"])
def test_emulated_dropin_parameter_file_options(lines, tmp_path):
    """FixedIC / PairedIC (amplitudes fixed to the spectrum, phases shifted by pi, src/GenIC.c:370-376) travel
    in pinb200_desc; MimicOldSeed (internal.mimic_original_seedtable) replaces the spiral seed plane by the
    N-GenIC table of src/GenIC.c:493-537, which the shim builds with the host's own generator and hands
    over through pinb200_set_seed_plane.  Each option is a different realisation, reproduced byte for byte."""
    a, b, c = tmp_path / "emu", tmp_path / "ref", tmp_path / "ref_plain"
    run32_args(EMU_X, a, extra_param_lines=lines)
    run32_args(REF_X, b, extra_param_lines=lines)
    run32_args(REF_X, c)
    names = sorted(f.name for f in b.glob("pinocchio.*"))
    assert len(names) >= 11
    for name in names:
        assert (a / name).read_bytes() == (b / name).read_bytes(), name
    assert (b / "pinocchio.0.0000.test.catalog.out").read_bytes() != (c / "pinocchio.0.0000.test.catalog.out").read_bytes()


def test_emulated_dropin_ignores_use_transposed_fft(tmp_path):
    """UseTransposedFFT only chooses PFFT's internal k-space layout (src/fmax-pfft.c:92,159-181,271); the
    real-space results are the same, so the drop-in accepts the option and delivers the same run.  (The
    one-task PFFT stand-in under the reference program does not provide transposed layouts, hence the
    comparison with the drop-in's own run without the option.)"""
    a, b = tmp_path / "transposed", tmp_path / "plain"
    run32_args(EMU_X, a, extra_param_lines=("UseTransposedFFT",))
    run32_args(EMU_X, b)
    for name in sorted(f.name for f in b.glob("pinocchio.*")):
        assert (a / name).read_bytes() == (b / name).read_bytes(), name


@pytest.mark.parametrize("mimic", [False, True], ids=["spiral", "mimic_old_seed"])
def test_oracle_seed_planes_against_the_reference_dump(mimic, tmp_path):
    """`DumpSeedPlane 1` makes the reference program write SEEDTABLE (src/GenIC.c:152,1144-1200): pins the
    oracle's spiral table (seed_table) and its N-GenIC table (seed_table_old) entry by entry"""
    from oracle import pinocchio_oracle as po
    lines = ("DumpSeedPlane 1",) + (("MimicOldSeed",) if mimic else ())
    run32_args(REF_X, tmp_path, extra_param_lines=lines)
    a = np.loadtxt(tmp_path / f"seed_plane.Ng_{N}_Nt_1.dat", dtype=np.int64)
    assert a.shape == (N * N, 4) and np.array_equal(a[:, 2], a[:, 1] * N + a[:, 0])
    table = np.zeros((N, N), dtype=np.uint32)
    table[a[:, 1], a[:, 0]] = a[:, 3]
    want = po.seed_table_old(N, 486604) if mimic else po.seed_table(N, 486604)
    assert np.array_equal(table, want)


@needs_snap
def test_emulated_dropin_special_mode_2_density_snapshot(tmp_path):
    """`pinocchio.x parameter_file 2` (src/pinocchio.c:136-168): write_in_cvector(kdensity) -> reverse_transform ->
    write_from_rvector -> write_density.  delta_k lives on the device: the shim fetches it into the host
    kdensity[] when write_in_cvector is handed that array, and reverse_transform is pinb200_fft_c2r."""
    a, b = tmp_path / "emu", tmp_path / "ref"
    log = run32_args(SNAP_EMU, a, args=("2",))
    run32_args(SNAP_REF, b, args=("2",))
    assert "only writes the linear density field" in log
    fa, fb = a / "pinocchio.test.density0.out", b / "pinocchio.test.density0.out"
    assert fa.stat().st_size == fb.stat().st_size > 4 * N ** 3
    da = np.frombuffer(fa.read_bytes()[-(4 * N ** 3 + 4):-4], dtype=np.float32)
    db = np.frombuffer(fb.read_bytes()[-(4 * N ** 3 + 4):-4], dtype=np.float32)
    assert np.abs(db).max() > 0 and np.abs(da.astype(np.float64) - db).max() <= 1e-6 * np.abs(db).max()
    assert differing_bytes(fa, fb) <= 2e-4 * fa.stat().st_size


def test_emulated_dropin_refuses_an_unsupported_grid_loudly(tmp_path):
    """error convention of the boundary: a non-zero return of the C ABI reaches the caller, which
    prints and aborts (src/pinocchio.c:229-230,259-263); the library accepts powers of two only (the
    emulated ABI 32 and 64), the reference any FFTW size"""
    import os
    import subprocess
    text = (GOLDEN / "parameter_file").read_text()
    text = re.sub(r"(?m)^BoxSize\s+\S+", "BoxSize                48", text)
    text = re.sub(r"(?m)^GridSize\s+\S+", "GridSize               48", text)
    (tmp_path / "parameter_file").write_text(text)
    (tmp_path / "outputs").write_bytes((GOLDEN / "outputs").read_bytes())
    r = subprocess.run([str(EMU_X), "parameter_file"], cwd=tmp_path, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert r.returncode != 0
    assert "pinb200_create" in r.stdout + r.stderr and "aborting" in r.stdout + r.stderr
    assert not list(tmp_path.glob("pinocchio.*.catalog.out"))
