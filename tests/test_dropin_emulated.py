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
fragmentation stage must match the reference".  tests/test_zgpu_dropin_catalogues.py repeats it at
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
# "_sd": the same + -DSCALE_DEPENDENT -DRECOMPUTE_DISPLACEMENTS -- growth rates per k bin handed to
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
    if variant == "_sd":                                # 104-byte records + fields kept for the re-entry
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
    return a, run32(emu, a, v), b, run32(ref, b, v)


def test_emulated_dropin_log(runs):
    a, log_a, b, log_b = runs
    assert "B200 path" in log_a and "Pinocchio done!" in log_a and "B200 path" not in log_b
    sig = lambda log: [float(x) for x in re.findall(r"computed sigma:\s+([0-9.]+)", log)]
    assert len(sig(log_a)) == 9 and sig(log_a) == sig(log_b)
    ncoll = lambda log: int(re.search(r"Number of collapsed particles to z=0: (\d+)", log).group(1))
    assert ncoll(log_a) == ncoll(log_b) > 0.4 * N ** 3
    # -DRECOMPUTE_DISPLACEMENTS: both programs recompute the displacements for segments 2..4
    nre = lambda log: len(re.findall(r"Computing displacements for redshift", log))
    assert nre(log_a) == nre(log_b)


def test_emulated_dropin_fmaxpdf_identical(runs):
    a, _, b, _ = runs
    pa = np.loadtxt(a / "pinocchio.test.FmaxPDF.out")[:, 2]
    pb = np.loadtxt(b / "pinocchio.test.FmaxPDF.out")[:, 2]
    assert pa.sum() == N ** 3
    assert np.abs(pa - pb).max() <= 1          # float-rounding flips of Fmax at a bin edge at most


@pytest.mark.parametrize("z", ["0.0000", "0.5000", "1.0000", "2.0000"])
def test_emulated_dropin_catalogues_match(runs, z):
    a, _, b, _ = runs
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
    a, _, b, _ = runs
    ma = np.loadtxt(a / "pinocchio.0.0000.test.mf.out")
    mb = np.loadtxt(b / "pinocchio.0.0000.test.mf.out")
    assert np.array_equal(ma[:, 4], mb[:, 4])          # halos per mass bin
