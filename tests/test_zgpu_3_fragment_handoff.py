"""Device-side hand-off to the fragmentation: pinb200_collapsed_cells and
pinb200_download_products_sorted -- the cells distribute() keeps (Fmax >= F_last,
src/distribute.c:58-175,547-600) in the order sort_and_organize gives frag[] (descending Fmax,
src/fragment.c:484-520).  Needs a B200: -m gpu.  (File name sorts after the round-1 parity tests.)
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HMF_RADII = [20.635922, 13.996056, 9.026099, 5.465945, 3.058354, 1.548258, 0.689079, 0.258729, 0.0]


@pytest.fixture(scope="module")
def cosmo():
    from pinocchio_b200.cosmology import Cosmology
    return Cosmology(pk_norm_override=2.03146e7)


def make(N, cosmo):
    from pinocchio_b200.cosmology import SmoothingLadder
    from pinocchio_b200.engine import Pinocchio, RunConfig
    cfg = RunConfig(GridSize=N, BoxSize_htrue=N / 0.7, lpt_order=3)
    return Pinocchio(cfg, cosmo, smoothing=SmoothingLadder(np.array(HMF_RADII), np.zeros(len(HMF_RADII))))


@pytest.mark.parametrize("N", [64, 128])
def test_collapsed_cells_order_and_records(N, cosmo):
    p = make(N, cosmo)
    p.GenIC_large()
    p.compute_fmax()
    F = p.field("Fmax").ravel()
    prod = p.products()
    for Flast in (1.0, 2.0):
        idx = p.collapsed_cells(Flast)
        sel = np.flatnonzero(F >= np.float32(Flast))
        want = sel[np.argsort(-F[sel].astype(np.float64), kind="stable")]     # ties: ascending cell index
        assert idx.size == want.size
        assert np.array_equal(idx, want.astype(np.uint32))
        frag = p.sorted_products()
        assert np.array_equal(frag, prod[idx])
        part = p.sorted_products(first=100, n=1000)
        assert np.array_equal(part, prod[idx[100:1100]])
    # Flast = 1: exactly the cells the FmaxPDF counts as collapsed (src/fmax.c:533-536)
    assert p.collapsed_cells(1.0).size == int(p.Fmax_PDF()[10:].sum())
    p.close()


def test_collapsed_cells_large_grid_properties(cosmo):
    """512^3 (more than one pass-to-pass buffer of tiles): permutation, order, count"""
    N = 512
    p = make(N, cosmo)
    p.GenIC_large()
    p.compute_fmax(displacements=False)
    F = p.field("Fmax").ravel()
    idx = p.collapsed_cells(1.0)
    assert idx.size == int((F >= 1.0).sum()) == int(p.Fmax_PDF()[10:].sum())
    Fs = F[idx]
    assert (np.diff(Fs) <= 0).all()                       # descending
    same = np.diff(Fs) == 0
    assert (np.diff(idx.astype(np.int64))[same] > 0).all()    # ties in ascending cell index
    assert np.unique(idx).size == idx.size
    p.close()


def test_handoff_started_under_the_displacement_stage(cosmo):
    """pinb200_handoff_begin / _end: the selection, the sort and the two downloads run on side streams while
    pinb200_displacements computes the velocity fields; same list and same records as the synchronous calls."""
    import ctypes
    N = 128
    p = make(N, cosmo)
    p.GenIC_large()
    p.compute_fmax(displacements=False)
    F = p.field("Fmax").ravel()
    want = p.collapsed_cells(1.0).copy()
    fm = np.zeros(N ** 3, dtype=np.float32)
    idx = np.zeros(want.size, dtype=np.uint32)
    n = ctypes.c_size_t(0)
    p._ck(p.lib.pinb200_handoff_begin(p.h, 1.0, fm.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                      idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint)), idx.size, ctypes.byref(n)))
    assert n.value == want.size
    p.compute_displacements(1, 0, 0.0)                  # the stage the hand-off hides under
    p._ck(p.lib.pinb200_handoff_end(p.h))
    assert np.array_equal(fm, F) and np.array_equal(idx, want)
    p._sorted_n = int(n.value)
    prod = p.products()
    assert np.array_equal(p.sorted_products(), prod[idx])
    assert prod["Vel_3LPT_2"].any()
    # a second begin without end in between, and a sweep that starts while a list is pending, are both safe
    p._ck(p.lib.pinb200_handoff_begin(p.h, 2.0, None, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint)), idx.size, ctypes.byref(n)))
    p._ck(p.lib.pinb200_handoff_begin(p.h, 1.0, None, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint)), idx.size, ctypes.byref(n)))
    p.compute_fmax(displacements=False)
    assert np.array_equal(p.collapsed_cells(1.0), want)
    p.close()
