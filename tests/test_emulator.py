"""CPU-only checks of the CUDA kernel bodies through the pthread block emulator.

The templates of pinocchio_b200/csrc/kernels.cuh are compiled for the host with g++ and run
block by block (tests/host/emu.cpp); results are compared with the oracle.  This verifies the
index math, FFT plans, the fused k-space factors, the c2r/r2c glue, the collapse arithmetic
and the GenIC RANLUX chain without a GPU.  The `-m gpu` tests repeat the comparisons on the
real kernels through the C ABI.
"""
import ctypes

import numpy as np
import pytest

from emu_util import (PF, PI32, PU32, empty_field, from_pitched_c, load_emulator, pitch, ptr, ptr_array,
                      real_view, to_pitched_c, twiddles)
from oracle import pinocchio_oracle as po
from pinocchio_b200.cosmology import Cosmology, pk_lattice_table


@pytest.fixture(scope="module")
def lib():
    return load_emulator()


@pytest.fixture(scope="module")
def cosmo():
    return Cosmology()


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("L", [16, 32, 64, 128, 256, 512, 1024, 2048])
def test_strided_tile_fft(lib, L):
    rng = np.random.default_rng(L)
    TK = lib.emu_strided_tk(L)
    x = rng.standard_normal((L, TK)) + 1j * rng.standard_normal((L, TK))
    tw = twiddles(L)
    for d in (+1, -1):
        out = np.zeros_like(x)
        assert lib.emu_strided_tile_fft(L, d, ptr(x), ptr(out), ptr(tw)) == 0
        ref = np.fft.fft(x, axis=0) if d < 0 else np.fft.ifft(x, axis=0) * L
        assert rel(out, ref) < 2e-15


@pytest.mark.parametrize("M", [16, 32, 64, 128, 256, 512, 1024])
def test_contiguous_line_fft(lib, M):
    rng = np.random.default_rng(M)
    x = rng.standard_normal(M) + 1j * rng.standard_normal(M)
    tw = twiddles(2 * M)
    for d in (+1, -1):
        out = np.zeros_like(x)
        assert lib.emu_zline_fft(M, d, ptr(x), ptr(out), ptr(tw)) == 0
        ref = np.fft.fft(x) if d < 0 else np.fft.ifft(x) * M
        assert rel(out, ref) < 2e-15


def _c2r(lib, N, ck):
    """plain c2r through x, y, z passes (in place), returns real [N,N,N]"""
    tw = twiddles(N)
    f = to_pitched_c(ck)
    norm = 1.0 / N ** 3
    assert lib.emu_xpass(N, +1, ptr(f), ptr(f), None, None, 1, 1, None, ctypes.c_double(norm), 0, 0, ptr(tw)) == 0
    jobs = np.array([0, 0, 0], dtype=np.int32)
    assert lib.emu_ypass(N, +1, ptr_array([f], 3), ptr_array([f], 6), ptr(jobs, PI32), 1, 1, ptr(tw)) == 0
    kz = np.zeros(6, dtype=np.int32)
    assert lib.emu_zpass_out(N, 1, ptr_array([f], 6), ptr(kz, PI32), 1, None, 0, ptr_array([f], 6), None, None,
                             None, None, ptr(tw)) == 0
    return real_view(f).copy()


def _r2c(lib, N, r):
    tw = twiddles(N)
    f = empty_field(N)
    real_view(f)[...] = r
    assert lib.emu_zpass_r2c(N, ptr(f), ptr(f), ptr(tw)) == 0
    jobs = np.array([0, 0, 0], dtype=np.int32)
    assert lib.emu_ypass(N, -1, ptr_array([f], 3), ptr_array([f], 6), ptr(jobs, PI32), 1, 1, ptr(tw)) == 0
    assert lib.emu_xpass(N, -1, ptr(f), ptr(f), None, None, 1, 1, None, ctypes.c_double(1.0), 0, 0, ptr(tw)) == 0
    return from_pitched_c(f)


@pytest.mark.parametrize("N", [32, 64])
def test_fft3d_roundtrip_and_nonhermitian(lib, N):
    rng = np.random.default_rng(7)
    r = rng.standard_normal((N, N, N))
    ck = _r2c(lib, N, r)
    assert rel(ck, po.forward_transform(r)) < 5e-15
    # arbitrary (non-Hermitian) half-complex input: FFTW/numpy literal c2r semantics (App. A.5)
    c = rng.standard_normal((N, N, N // 2 + 1)) + 1j * rng.standard_normal((N, N, N // 2 + 1))
    out = _c2r(lib, N, c)
    assert rel(out, po.reverse_transform(c)) < 5e-15


def _hessian_collapse(lib, N, kd, radius, cell, spline_packed, ismooth, Fmax, Rmax, store_h):
    tw = twiddles(N)
    M = N // 2
    knorm = 2 * np.pi / N
    rs = radius / cell
    gauss = np.exp(-0.5 * (knorm * np.arange(M + 1)) ** 2 * rs * rs)
    src = to_pitched_c(kd)
    A = [empty_field(N) for _ in range(3)]
    B = [empty_field(N) for _ in range(6)]
    norm = 1.0 / N ** 3
    assert lib.emu_xpass(N, +1, ptr(src), ptr(A[0]), ptr(A[1]), ptr(A[2]), 7, 0, ptr(gauss), ctypes.c_double(norm),
                         1, 0, ptr(tw)) == 0
    jobs = np.array([2, 0, 0, 0, 2, 1, 0, 0, 2, 1, 1, 3, 1, 0, 4, 0, 1, 5], dtype=np.int32)
    assert lib.emu_ypass(N, +1, ptr_array(A, 3), ptr_array(B, 6), ptr(jobs, PI32), 6, 0, ptr(tw)) == 0
    kz = np.array([0, 0, 2, 0, 1, 1], dtype=np.int32)
    sums = np.zeros(2)
    dc = np.array([kd[0, 0, 0].real * norm])
    hd = ptr_array(B, 6) if store_h else None
    nspl = spline_packed.shape[1]
    assert lib.emu_zpass_collapse(N, ptr_array(B, 6), ptr(kz, PI32), 0, ptr(dc), ptr(spline_packed), nspl, ismooth,
                                  ptr(Fmax, PF), ptr(Rmax, PI32), ptr(sums), hd, ptr(tw)) == 0
    return sums, B


def test_genic_hessian_collapse_lpt(lib, cosmo):
    """End-to-end at 32^3 in the emulator: GenIC -> 3 radii -> Fmax/Rmax -> sources ->
    contraction -> displacements, each stage against the oracle."""
    N = 32
    box = 64.0 / 0.7
    cell = box / N
    seed = 486604
    # ---- GenIC
    seeds = po.seed_table(N, seed)
    pk = pk_lattice_table(cosmo, N, box)
    kdp = np.zeros((N, N, pitch(N)), dtype=np.complex128)
    assert lib.emu_genic(N, ptr(np.ascontiguousarray(seeds), PU32), ptr(pk), ctypes.c_double(box), 0, 0, ptr(kdp)) == 0
    kd = from_pitched_c(kdp)
    kd_ref = po.genic(N, box, seed, cosmo.PowerSpectrum)
    assert rel(kd, kd_ref) < 1e-13
    assert np.count_nonzero(kd) == np.count_nonzero(kd_ref)

    # ---- radii loop
    radii = [6.0, 2.5, 0.0]
    spl = cosmo.sp_invgrow.packed()
    Fmax = np.zeros((N, N, N), dtype=np.float32)
    Rmax = np.zeros((N, N, N), dtype=np.int32)
    Fo, Ro = po.init_products((N, N, N))
    B = None
    for ism, R in enumerate(radii):
        sums, B = _hessian_collapse(lib, N, kd_ref, R, cell, spl, ism, Fmax, Rmax, store_h=(ism == len(radii) - 1))
        h = po.second_derivatives(kd_ref, R, cell)
        Fnew = po.inverse_collapse_time(h, cosmo.InverseGrowingMode)
        po.update_fmax(Fo, Ro, Fnew, ism)
        tv, av = po.true_variance(h)
        assert abs(sums[1] / N ** 3 - tv) < 1e-12 * tv
        # float Fmax: allow last-bit differences from FFT rounding; Rmax may differ only on near-ties
        unstable = po.ill_conditioned_mask(h, cosmo.InverseGrowingMode) if ism == 0 else \
            unstable | po.ill_conditioned_mask(h, cosmo.InverseGrowingMode)
        ok = ~unstable
        assert np.abs(Fmax.astype(np.float64) - Fo)[ok].max() < 1e-5
        mism = (Rmax != Ro) & ok
        assert mism.mean() < 1e-3 and unstable.mean() < 1e-3
    hess = [real_view(b).copy() for b in B]
    for k in range(6):
        assert rel(hess[k], h[k]) < 1e-13

    # ---- LPT sources
    P2 = 2 * pitch(N)
    S = [np.zeros((N, N, P2)) for _ in range(3)]
    Hp = [np.ascontiguousarray(b.view(np.float64).reshape(N, N, P2)) for b in B]
    assert lib.emu_sources(N, ptr_array(Hp, 6), ptr(S[0]), ptr(S[1]), ptr(S[2]), 3) == 0
    s2, s31, s32 = po.lpt_sources(h)
    assert rel(S[0][:, :, :N], s2) < 1e-13
    assert rel(S[1][:, :, :N], s31) < 1e-13
    assert rel(S[2][:, :, :N], s32) < 1e-13

    # ---- kvector_2LPT and the contraction into source_3LPT_2 (three groups)
    k2 = _r2c(lib, N, S[0][:, :, :N])
    k2_ref, k31_ref, k32_ref = po.lpt_kvectors(h)
    assert rel(k2, k2_ref) < 1e-12
    tw = twiddles(N)
    norm = 1.0 / N ** 3
    src = to_pitched_c(k2)
    dc = np.array([k2[0, 0, 0].real * norm])
    acc = S[2]
    groups = [(2, [(0, 0, 0)], [0], [0]),
              (1, [(0, 1, 0), (0, 0, 1)], [0, 1], [3, 4]),
              (0, [(0, 2, 0), (0, 1, 1), (0, 0, 2)], [0, 1, 2], [1, 5, 2])]
    for pw, jobs, kzp, slots in groups:
        C = empty_field(N)
        D = [empty_field(N) for _ in range(3)]
        d = [None, None, None]
        d[pw] = C
        assert lib.emu_xpass(N, +1, ptr(src), ptr(d[0]), ptr(d[1]), ptr(d[2]), 1 << pw, 1, None,
                             ctypes.c_double(norm), 1, 0, ptr(tw)) == 0
        ja = np.array(jobs, dtype=np.int32).ravel()
        assert lib.emu_ypass(N, +1, ptr_array([C], 3), ptr_array(D, 6), ptr(ja, PI32), len(jobs), 1, ptr(tw)) == 0
        kz = np.array(kzp + [0] * (6 - len(kzp)), dtype=np.int32)
        w = np.array([2.0 * (1.0 if s <= 2 else 2.0) for s in slots] + [0.0] * (6 - len(slots)))
        hs = [Hp[s] for s in slots]
        assert lib.emu_zpass_out(N, len(jobs), ptr_array(D, 6), ptr(kz, PI32), 1, ptr(dc), 2, None, None,
                                 ptr_array(hs, 6), ptr(w), ptr(acc), ptr(tw)) == 0
    k32 = _r2c(lib, N, acc[:, :, :N])
    assert rel(k32, k32_ref) < 1e-11

    # ---- displacements from kvector_3LPT_2 (has power on the Nyquist planes) and from delta_k
    for kvec, with_nyq, growth in ((k32_ref, 1, 0.1216), (kd_ref, 0, 1.0)):
        src = to_pitched_c(kvec)
        W = [empty_field(N) for _ in range(5)]
        dc = np.array([-kvec[0, 0, 0].imag * norm])
        assert lib.emu_xpass(N, +1, ptr(src), ptr(W[1]), ptr(W[0]), None, 3, with_nyq, None,
                             ctypes.c_double(norm * growth), 1, 1, ptr(tw)) == 0
        ja = np.array([0, 0, 0, 1, 1, 1, 1, 0, 2], dtype=np.int32)
        assert lib.emu_ypass(N, +1, ptr_array(W[:2], 3), ptr_array(W[2:], 6), ptr(ja, PI32), 3, with_nyq, ptr(tw)) == 0
        kz = np.array([0, 0, 1, 0, 0, 0], dtype=np.int32)
        V = [np.zeros((N, N, N), dtype=np.float32) for _ in range(3)]
        assert lib.emu_zpass_out(N, 3, ptr_array(W[2:], 6), ptr(kz, PI32), with_nyq, ptr(dc), 1, None,
                                 ptr_array(V, 6, PF), None, None, None, ptr(tw)) == 0
        ref = po.first_derivatives(kvec, growth)
        for a in range(3):
            scale = np.abs(ref[a]).max()
            assert np.abs(V[a] - ref[a]).max() < 2e-7 * scale   # float32 storage


def test_collapse_cells_branches(lib, cosmo):
    """inverse_collapse_time on synthetic Hessians covering every branch of ell_classic."""
    rng = np.random.default_rng(3)
    n = 20000
    h = rng.standard_normal((6, n)) * np.array([1.5, 1.5, 1.5, 0.7, 0.7, 0.7])[:, None]
    # special cases: zero tensor, diagonal tensors, tiny l1, negative definite, exact sphere
    h[:, 0] = 0.0
    h[:, 1] = [1.0, 1.0, 1.0, 0, 0, 0]
    h[:, 2] = [2.0, 0.5, -0.3, 0, 0, 0]
    h[:, 3] = [-1.0, -2.0, -3.0, 0, 0, 0]
    h = np.ascontiguousarray(h)
    F = np.zeros(n)
    spl = cosmo.sp_invgrow.packed()
    assert lib.emu_collapse_cells(ptr(h), ctypes.c_longlong(n), ptr(spl), spl.shape[1], ptr(F)) == 0
    hl = [h[i] for i in range(6)]
    ref = po.inverse_collapse_time(hl, cosmo.InverseGrowingMode)
    # cells where the reference algorithm itself is ill-conditioned are flagged, not compared
    unstable = po.ill_conditioned_mask(hl, cosmo.InverseGrowingMode)
    assert unstable.mean() < 1e-3
    ok = ~unstable
    assert np.isfinite(F[ok]).all() and np.isfinite(ref[ok]).all()
    assert (np.abs(F[ok] - ref[ok]) <= 1e-6 * np.maximum(1.0, np.abs(ref[ok]))).all()   # the 1e-6 contract
    assert np.median(np.abs(F[ok] - ref[ok])) < 1e-14
    assert (ref == 0).sum() > 100 and (ref > 1).sum() > 100
