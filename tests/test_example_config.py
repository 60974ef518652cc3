"""example/parameter_file as shipped (see tests/example_util.py), without a GPU:
  * the reference program compiled verbatim over the one-task MPI / mini-GSL stand-ins (now with GSL's bicubic
    gsl_spline2d for READ_PK_TABLE, oracle/ref_full/mini_gsl.c) reproduces the SHIPPED outputs of that run:
    example/pinocchio.example.FmaxPDF.out bin for bin (741 412 collapsed particles) and the z = 0 / z = 2 catalogues;
  * the drop-in over the emulated ABI (reference host code + shim + the kernel bodies on the CPU) against that
    reference program on the same parameter file at 64^3 (the emulated ABI's largest grid): every output file.
The B200 run of the same parameter file at the shipped 128^3 is tests/test_zgpu_9_example_as_shipped.py."""
import re

import numpy as np
import pytest

from example_util import EMU_EX, EXAMPLE, REF_EX, run_example
from test_reference_full import load_catalog, match_fraction

pytestmark = pytest.mark.skipif(not REF_EX.exists(), reason="oracle/_ref/pinocchio_ref_ex.x not built (make -C oracle all)")


@pytest.fixture(scope="module")
def refrun(tmp_path_factory):
    d = tmp_path_factory.mktemp("example_ref")
    return d, run_example(REF_EX, d)


def test_reference_reproduces_the_shipped_example(refrun):
    d, log = refrun
    assert "Pinocchio done!" in log
    assert int(re.search(r"Number of collapsed particles to z=0: (\d+)", log).group(1)) == 741412
    pdf = np.loadtxt(d / "pinocchio.example.FmaxPDF.out")[:, 2].astype(np.int64)
    gold = np.loadtxt(EXAMPLE / "pinocchio.example.FmaxPDF.out")[:, 2].astype(np.int64)
    assert pdf.sum() == gold.sum() == 128 ** 3
    assert np.abs(pdf - gold).max() <= 2            # identical bin for bin in this container
    mf = np.loadtxt(d / "pinocchio.0.0000.example.mf.out")
    mfg = np.load(EXAMPLE / "mf_0.0000.npz")
    assert np.abs(mf[:, 4] - mfg["nhalos"]).max() <= 2 and np.allclose(mf[:, 5], mfg["nm_analytic"], rtol=1e-3)


@pytest.mark.parametrize("z", ["0.0000", "2.0000"])
def test_reference_catalogues_against_shipped(refrun, z):
    d, _ = refrun
    ids, npart, _ = load_catalog(d / f"pinocchio.{z}.example.catalog.out")
    g = np.load(EXAMPLE / f"catalog_{z}_id_npart.npz")
    assert abs(len(ids) - len(g["id"])) <= 0.002 * len(g["id"])
    assert match_fraction(ids, npart, g["id"], g["npart"]) > 0.995


@pytest.mark.skipif(not EMU_EX.exists(), reason="oracle/_ref/pinocchio_emu_ex.x not built")
def test_emulated_dropin_on_the_example_configuration(tmp_path):
    """CAMB tables, scale-dependent growth per k bin and per radius, Hubble table, radiation, displacements
    recomputed per redshift segment: the shim's marshalling of all of it, at 64^3 (box 250 Mpc/h, same cells)"""
    a, b = tmp_path / "emu", tmp_path / "ref"
    log_a = run_example(EMU_EX, a, grid=64, threads=4)
    log_b = run_example(REF_EX, b, grid=64, threads=4)
    assert "B200 path" in log_a and "Pinocchio done!" in log_a and "B200 path" not in log_b
    sig = lambda log: [float(x) for x in re.findall(r"computed sigma:\s+([0-9.]+)", log)]
    assert len(sig(log_a)) >= 6 and sig(log_a) == sig(log_b)
    names = ["pinocchio.example.FmaxPDF.out", "pinocchio.0.0000.example.mf.out"] + \
            [f"pinocchio.{z}.example.catalog.out" for z in ("0.0000", "0.5000", "1.0000", "2.0000")]
    for name in names:
        assert (a / name).read_bytes() == (b / name).read_bytes(), name
