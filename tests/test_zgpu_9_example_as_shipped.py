"""BASELINE.json configs[0] on the B200: example/parameter_file AS SHIPPED (128^3, CAMB tables, scale-dependent growth
of a massive-neutrino cosmology, Hubble table, radiation, displacements recomputed per redshift segment; the OPTIONS
of src/Makefile:46-76) through the linked drop-in oracle/_ref/pinocchio_b200_ex.x -- the unchanged reference host
code with the five files of the hot path replaced by shim/fmax_b200.c + libpinb200.so -- against
  (a) the outputs the reference SHIPS for this run (tests/golden/example/: FmaxPDF with 741 412 collapsed particles,
      z = 0 and z = 2 catalogues, z = 0 mass function), and
  (b) the reference program itself run here on the same files (oracle/_ref/pinocchio_ref_ex.x).
Needs a B200: -m gpu."""
import re

import numpy as np
import pytest

from example_util import B200_EX, EXAMPLE, REF_EX, run_example
from test_reference_full import load_catalog, match_fraction

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (REF_EX.exists() and B200_EX.exists()),
                                                  reason="oracle/_ref/pinocchio_{ref,b200}_ex.x not built (make -C oracle all)")]


@pytest.fixture(scope="module")
def runs(tmp_path_factory):
    da, db = tmp_path_factory.mktemp("example_b200"), tmp_path_factory.mktemp("example_ref")
    return da, run_example(B200_EX, da, threads=16), db, run_example(REF_EX, db, threads=16)


def test_example_log_and_fmaxpdf(runs):
    da, log_a, db, log_b = runs
    assert "B200 path" in log_a and "Pinocchio done!" in log_a
    sig = lambda log: np.array([float(x) for x in re.findall(r"computed sigma:\s+([0-9.]+)", log)])
    assert sig(log_a).size == sig(log_b).size >= 6 and np.abs(sig(log_a) - sig(log_b)).max() <= 1e-4
    ncoll = int(re.search(r"Number of collapsed particles to z=0: (\d+)", log_a).group(1))
    assert abs(ncoll - 741412) <= 5                                      # the shipped FmaxPDF's count
    pdf = np.loadtxt(da / "pinocchio.example.FmaxPDF.out")[:, 2].astype(np.int64)
    shipped = np.loadtxt(EXAMPLE / "pinocchio.example.FmaxPDF.out")[:, 2].astype(np.int64)
    assert pdf.sum() == shipped.sum() == 128 ** 3
    assert np.abs(pdf - shipped).max() <= 6 and np.abs(pdf - shipped).sum() <= 250


@pytest.mark.parametrize("z", ["0.0000", "0.5000", "1.0000", "2.0000"])
def test_example_catalogues(runs, z):
    da, _, db, _ = runs
    ia, na, _ = load_catalog(da / f"pinocchio.{z}.example.catalog.out")
    ib, nb, _ = load_catalog(db / f"pinocchio.{z}.example.catalog.out")
    assert abs(len(ia) - len(ib)) <= max(2, 0.005 * len(ib))
    assert match_fraction(ia, na, ib, nb) > 0.99                           # against the reference program run here
    if z in ("0.0000", "2.0000"):                                         # against the shipped catalogues
        g = np.load(EXAMPLE / f"catalog_{z}_id_npart.npz")
        assert match_fraction(ia, na, g["id"], g["npart"]) > 0.99


def test_example_mass_function(runs):
    da, _, db, _ = runs
    a = np.loadtxt(da / "pinocchio.0.0000.example.mf.out")
    b = np.loadtxt(db / "pinocchio.0.0000.example.mf.out")
    g = np.load(EXAMPLE / "mf_0.0000.npz")
    assert np.abs(a[:, 4] - b[:, 4]).max() <= 3 and np.abs(a[:, 4] - g["nhalos"]).max() <= 3
    assert np.allclose(a[:, 5], b[:, 5], rtol=1e-12)
