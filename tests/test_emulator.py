"""CPU-only checks of the CUDA kernel bodies through the fiber block emulator.

The templates of pinocchio_b200/csrc/kernels.cuh are compiled for the host with g++ and run
block by block (tests/host/emu.cpp); results are compared with the oracle.  This verifies the
index math, FFT plans, the fused k-space factors, the c2r/r2c glue, the collapse arithmetic,
the GenIC RANLUX chain and -- by emulating 2 and 4 ranks in one process -- the peer-memory
scatter addressing of the slab decomposition, all without a GPU.  The `-m gpu` tests repeat
the comparisons on the real kernels through the C ABI.
"""
import ctypes

import numpy as np
import pytest

from emu_cluster import EmuCluster
from emu_util import load_emulator, packed_spline, ptr, twiddles
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


@pytest.mark.parametrize("N,P,split", [(32, 1, False), (64, 1, False), (32, 2, False), (32, 4, False),
                                       (32, 1, True), (64, 2, True), (32, 4, True)])
def test_fft3d_forward_and_nonhermitian_c2r(N, P, split):
    """r2c and c2r through the three passes, on 1, 2 and 4 emulated ranks; split = the
    decimation-in-frequency long-line path (XCfg/YCfg in kernels.cuh)."""
    cl = EmuCluster(N, P, split)
    rng = np.random.default_rng(7)
    r = rng.standard_normal((N, N, N))
    for k in range(P):
        cl.A[0][k].view(np.float64).reshape(cl.lx, N, 2 * cl.Pc)[:, :, :N] = r[k * cl.lx:(k + 1) * cl.lx]
    cl.r2c(cl.A[0], cl.KV[0])
    assert rel(cl.gather_k(cl.KV[0]), po.forward_transform(r)) < 5e-15
    # arbitrary (non-Hermitian) half-complex input: FFTW/numpy literal c2r semantics (App. A.5)
    c = rng.standard_normal((N, N, N // 2 + 1)) + 1j * rng.standard_normal((N, N, N // 2 + 1))
    cl.scatter_k(c, cl.KV[1])
    assert rel(cl.c2r_plain(cl.KV[1]), po.reverse_transform(c)) < 5e-15


@pytest.mark.parametrize("P,split", [(1, False), (2, False), (4, False), (2, True)])
def test_genic_hessian_collapse_lpt(cosmo, P, split):
    """The engine's whole schedule at 32^3 on P emulated ranks: GenIC -> 3 radii -> Fmax/Rmax ->
    sources -> r2c -> contraction -> r2c -> displacements, each stage against the oracle."""
    N = 32
    box = 64.0 / 0.7
    cell = box / N
    seed = 486604
    cl = EmuCluster(N, P, split)
    # ---- GenIC
    cl.genic(po.seed_table(N, seed), pk_lattice_table(cosmo, N, box), box)
    kd = cl.gather_k(cl.kdens)
    kd_ref = po.genic(N, box, seed, cosmo.PowerSpectrum)
    assert rel(kd, kd_ref) < 1e-13
    assert np.count_nonzero(kd) == np.count_nonzero(kd_ref)
    cl.scatter_k(kd_ref, cl.kdens)       # continue from the oracle's field (identical to ~1e-16)

    # ---- radii loop
    radii = [6.0, 2.5, 0.0]
    spl = packed_spline(cl.lib, cosmo.sp_invgrow)
    nspl = cosmo.sp_invgrow.size
    Fo, Ro = po.init_products((N, N, N))
    unstable = np.zeros((N, N, N), dtype=bool)
    for ism, R in enumerate(radii):
        s2 = cl.hessian_collapse(R, cell, spl, nspl, ism, store_h=(ism == len(radii) - 1))
        h = po.second_derivatives(kd_ref, R, cell)
        Fnew = po.inverse_collapse_time(h, cosmo.InverseGrowingMode)
        po.update_fmax(Fo, Ro, Fnew, ism)
        tv, _ = po.true_variance(h)
        assert abs(s2 / N ** 3 - tv) < 1e-12 * tv
        unstable |= po.ill_conditioned_mask(h, cosmo.InverseGrowingMode)
        ok = ~unstable
        Fmax = np.concatenate(cl.Fmax, axis=0)
        Rmax = np.concatenate(cl.Rmax, axis=0)
        assert (np.abs(Fmax.astype(np.float64) - Fo)[ok] <= 1e-6 * np.maximum(1.0, np.abs(Fo[ok]))).all()
        assert ((Rmax != Ro) & ok).mean() < 1e-3 and unstable.mean() < 1e-3
    for k in range(6):
        assert rel(cl.gather_real(cl.B[k]), h[k]) < 1e-13

    # ---- LPT sources, kvector_2LPT, contraction, kvector_3LPT_1/2
    cl.sources()
    s2r, s31r, s32r = po.lpt_sources(h)
    assert rel(cl.gather_real(cl.A[0]), s2r) < 1e-13
    assert rel(cl.gather_real(cl.A[1]), s31r) < 1e-13
    assert rel(cl.gather_real(cl.A[2]), s32r) < 1e-13
    k2_ref, k31_ref, k32_ref = po.lpt_kvectors(h)
    cl.r2c(cl.A[0], cl.KV[0])
    assert rel(cl.gather_k(cl.KV[0]), k2_ref) < 1e-12
    cl.contraction()
    cl.r2c(cl.A[1], cl.KV[1])
    cl.r2c(cl.A[2], cl.KV[2])
    assert rel(cl.gather_k(cl.KV[1]), k31_ref) < 1e-11
    assert rel(cl.gather_k(cl.KV[2]), k32_ref) < 1e-11

    # ---- displacements from kvector_3LPT_2 (power on the Nyquist planes) and from delta_k
    for kvec, ref_k, with_nyq, growth in ((cl.KV[2], k32_ref, 1, 0.1216), (cl.kdens, kd_ref, 0, 1.0)):
        V = cl.displacement(kvec, growth, with_nyq)
        ref = po.first_derivatives(ref_k, growth)
        for a in range(3):
            assert np.abs(V[a] - ref[a]).max() < 2e-7 * np.abs(ref[a]).max()   # float32 storage


def scaledep_tables(nk=10, seed=11):
    """[4][NkBINS] log10 growth tables with a pronounced k dependence (what InterpolateGrowth's
    k-bin splines return at one redshift, src/cosmo.c:1728-1757)."""
    rng = np.random.default_rng(seed)
    base = np.log10(np.array([0.61, 0.16, 0.05, 0.11]))[:, None]
    return base + 0.3 * np.cumsum(rng.uniform(-0.2, 0.2, (4, nk)), axis=1)


@pytest.mark.parametrize("P,split", [(2, False), (4, False), (2, True)])
def test_staged_transpose_equals_peer_stores(P, split):
    """Multi-GPU sweep, two ways of doing the transpose behind the inverse x pass: peer stores from the kernel
    (r01) and own planes in place + local staging + strided block copies (r02, what the copy engines do under the
    collapse pass).  The R-layout fields must come out bit-identical, pads and Nyquist columns aside -- except where
    the scattering kernel splits its lines (the split build here, N = 2048 in the product): the all-local pass of the
    staged transpose transforms the whole line in one tile (XCfg LOCAL), another factorisation of the same FFT, so
    the two agree to rounding."""
    N = 32
    cl = EmuCluster(N, P, split=split)
    rng = np.random.default_rng(17)
    glob = rng.standard_normal((N, N, N // 2 + 1)) + 1j * rng.standard_normal((N, N, N // 2 + 1))
    cl.scatter_k(glob, cl.kdens)
    M = N // 2
    gauss = np.exp(-0.5 * (2 * np.pi / N * np.arange(M + 1)) ** 2 * 1.7 ** 2)
    cl.xpass_inv(cl.kdens, {0: cl.A[0], 1: cl.A[1], 2: cl.A[2]}, 7, 0, gauss, cl.norm, 1, 0)
    want = [[a.copy() for a in cl.A[pw]] for pw in range(3)]
    A2 = [[cl.rfield() for _ in range(P)] for _ in range(3)]
    S = [[cl.kfield() for _ in range(P)] for _ in range(3)]
    cl.xpass_inv_staged(cl.kdens, S, A2, 7, 0, gauss, cl.norm, 1, 0)
    for pw in range(3):
        for r in range(P):
            got, ref = A2[pw][r][:, :, :M], want[pw][r][:, :, :M]
            if split:
                assert np.abs(got - ref).max() <= 1e-14 * np.abs(ref).max(), (pw, r)
                assert not np.array_equal(got, ref) or P == 1   # the unsplit kernel really ran
            else:
                assert np.array_equal(got, ref), (pw, r)


def test_interpolate_growth_clamps_and_knots():
    """Oracle restatement of InterpolateGrowth: knots are reproduced, clamped outside [kmin, kmax]."""
    tab = scaledep_tables()[0]
    kn = 10.0 ** (-3.0 + 0.5 * np.arange(10))
    v = po.interpolate_growth(kn * (1 + 1e-13), tab)     # just inside each bin
    assert np.abs(v[:-1] - tab[:-1]).max() < 1e-10
    assert po.interpolate_growth(np.array([1e-5]), tab)[0] == tab[0]
    assert po.interpolate_growth(np.array([1e3]), tab)[0] == tab[-1]
    mid = po.interpolate_growth(np.array([10.0 ** -1.75]), tab)[0]      # halfway between bins 2 and 3
    assert abs(mid - 0.5 * (tab[2] + tab[3])) < 1e-13


@pytest.mark.parametrize("P,split", [(1, False), (2, False), (1, True)])
def test_scale_dependent_displacements(P, split):
    """-DSCALE_DEPENDENT: growth_rate(|k|) applied per mode in the x-pass loader
    (src/fmax-pfft.c:340-364) for ScaleDep.order 1..4, k = 0 mode left unscaled."""
    N = 32
    cl = EmuCluster(N, P, split)
    rng = np.random.default_rng(21)
    tabs = scaledep_tables()
    # tables whose bins straddle the grid's k range (2 pi/N .. pi sqrt 3): LOGKMIN -1.5, DELTALOGK 0.25,
    # so that clamping below kmin, interpolation and clamping above kmax all occur
    for logkmin, dlogk in ((-3.0, 0.5), (-0.6, 0.1)):
        for order in (1, 2, 3, 4):
            c = rng.standard_normal((N, N, N // 2 + 1)) + 1j * rng.standard_normal((N, N, N // 2 + 1))
            if order == 1:
                c[N // 2, :, :] = 0; c[:, N // 2, :] = 0; c[:, :, N // 2] = 0     # delta_k has no Nyquist planes
            cl.scatter_k(c, cl.KV[0])
            sign = -1.0 if order == 3 else 1.0
            V = cl.displacement(cl.KV[0], 1.0, 0 if order == 1 else 1, gk=(tabs[order - 1], logkmin, dlogk, sign))
            g = po.growth_rate_of_k(N, order, tabs, logkmin, dlogk)
            assert g.std() > 1e-3 * abs(g.mean())          # the test tables really depend on k
            ref = po.first_derivatives(c, g)
            for a in range(3):
                assert np.abs(V[a] - ref[a]).max() < 2e-7 * np.abs(ref[a]).max()


def test_collapse_cells_branches(lib, cosmo):
    """inverse_collapse_time on synthetic Hessians covering the branches of ell_classic."""
    rng = np.random.default_rng(3)
    n = 20000
    h = rng.standard_normal((6, n)) * np.array([1.5, 1.5, 1.5, 0.7, 0.7, 0.7])[:, None]
    # special cases: zero tensor, sphere (already diagonal, q == 0), diagonal tensors
    h[:, 0] = 0.0
    h[:, 1] = [1.0, 1.0, 1.0, 0, 0, 0]
    h[:, 2] = [2.0, 0.5, -0.3, 0, 0, 0]
    h[:, 3] = [-1.0, -2.0, -3.0, 0, 0, 0]
    h = np.ascontiguousarray(h)
    F = np.zeros(n)
    spl = packed_spline(lib, cosmo.sp_invgrow)
    assert lib.emu_collapse_cells(ptr(h), ctypes.c_longlong(n), ptr(spl), cosmo.sp_invgrow.size, ptr(F)) == 0
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


def test_fastmath(lib):
    """fastmath.cuh (constant-bank elementary functions of the collapse epilogue) against libm."""
    rng = np.random.default_rng(5)

    def run(which, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        assert lib.emu_fastmath(which, ptr(x), ctypes.c_longlong(x.size), ptr(y)) == 0
        return y

    x = np.concatenate([rng.uniform(-1, 1, 200000), [1.0, -1.0, 0.0, 0.5, -0.5, 1 - 1e-16, -1 + 1e-16, 1e-300]])
    assert np.abs(run(0, x) - np.arccos(x)).max() < 1e-15
    with np.errstate(invalid="ignore"):
        assert np.isnan(run(0, np.array([1.0000001, -1.5, np.nan]))).all()
    x = np.concatenate([10.0 ** rng.uniform(-300, 300, 200000), rng.uniform(0.5, 2.0, 100000), [1.0, 2.0, 0.1, 1e-310]])
    ref = np.log10(x)
    assert (np.abs(run(1, x) - ref) <= 4e-16 * np.maximum(1.0, np.abs(ref))).all()
    x = np.concatenate([-(10.0 ** rng.uniform(-8, 2.84, 200000)), [0.0, -1e-300, -699.9, -800.0, -1e10]])
    ref = np.exp(x)
    assert (np.abs(run(2, x) - ref) <= 5e-16 * ref + 1e-300).all()
    x = np.concatenate([rng.uniform(-20, 20, 200000), [0.0, 1.0, -1.0, 0.30102999566, 299.9, -299.9]])
    ref = 10.0 ** x
    assert (np.abs(run(3, x) - ref) <= 3e-15 * ref).all()


def test_fastmath_div_sqrt_newton_schedules(lib):
    """fm_rcp / fm_div / fm_sqrt / 1/sqrt (fastmath.cuh): MUFU-seed + Newton sequences without range
    checks.  On the host the seed is the exact value cut to 20 mantissa bits (no better than the
    hardware's), so this checks that the schedules converge to ~1 ulp from such a seed, over the
    magnitudes the collapse epilogue produces (|den| >= 1e-20 => coefficients up to 1e60)."""
    rng = np.random.default_rng(9)

    def run(which, x, n=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        n = x.size if n is None else n
        y = np.zeros(n)
        assert lib.emu_fastmath(which, ptr(x), ctypes.c_longlong(n), ptr(y)) == 0
        return y

    mag = 10.0 ** rng.uniform(-150, 150, 300000)
    x = np.concatenate([mag * rng.choice([-1.0, 1.0], mag.size), rng.uniform(0.5, 2.5, 100000), [1.0, -1.0, 3.0, 1e-20, 1e60]])
    assert (np.abs(run(4, x) * x - 1.0) <= 3e-16).all()
    a = np.concatenate([rng.standard_normal(x.size - 3) * 10.0 ** rng.uniform(-100, 100, x.size - 3), [0.0, 1.0, -7.0]])
    q = run(5, np.stack([a, x], axis=1).ravel(), n=x.size)
    ref = a / x
    assert (np.abs(q - ref) <= 2.5e-16 * np.abs(ref)).all()
    assert (q == ref).mean() > 0.95                    # the remainder correction makes most quotients exact
    assert q[-3] == 0.0                                # zero numerator: exact zero, no special casing needed
    xp = np.concatenate([10.0 ** rng.uniform(-250, 250, 300000), rng.uniform(0, 0.25, 100000), [0.25, 1.0, 2.0, 4.0, 1e-300]])
    s = run(6, xp)
    ref = np.sqrt(xp)
    assert (np.abs(s - ref) <= 1.2e-16 * ref).all() and (s == ref).mean() > 0.95
    assert run(6, np.array([0.0]))[0] == 0.0           # acos(+-1) needs sqrt(0) = 0
    with np.errstate(invalid="ignore"):
        assert np.isnan(run(6, np.array([-1.0, -1e-30, np.nan]))).all()      # acos(|x| > 1) must stay NaN
    rs = run(7, xp)
    assert (np.abs(rs * ref - 1.0) <= 4e-16).all()


def test_cbrt_pair_and_cos_thirds(lib):
    """fm_cbrt_pair (cube root and its reciprocal from one Newton chain, seed cut to 20 bits on the host) and the
    Estrin form of cos_thirds against libm."""
    rng = np.random.default_rng(21)

    def run(which, x, nout=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(x.size if nout is None else nout)
        assert lib.emu_fastmath(which, ptr(x), ctypes.c_longlong(x.size), ptr(y)) == 0
        return y

    x = np.concatenate([10.0 ** rng.uniform(-300, 300, 300000), rng.uniform(0.5, 9.0, 100000), [1.0, 8.0, 27.0, 1e-20, 1e60, 2.0 ** -1022]])
    ref = np.cbrt(x)
    assert (np.abs(run(8, x) - ref) <= 5e-16 * ref).all()                     # 2 ulp
    assert (np.abs(run(9, x) * ref - 1.0) <= 4e-16).all()
    with np.errstate(invalid="ignore"):
        assert np.isnan(run(8, np.array([-1.0, 0.0, np.nan, np.inf]))).all()
    t = np.concatenate([rng.uniform(0, np.pi, 200000), [0.0, np.pi, np.pi / 2]])
    c = run(10, t, 3 * t.size).reshape(-1, 3)
    for j in range(3):
        assert np.abs(c[:, j] - np.cos((t + 2 * np.pi * j) / 3)).max() < 1e-15


@pytest.mark.parametrize("n,frac_ties", [(1, 0.0), (5000, 0.0), (16384, 0.5), (16385, 0.0), (100003, 0.3)])
def test_collapsed_cells_filter_and_sort(lib, n, frac_ties):
    """pinb200_collapsed_cells' kernels (sort_cells.cuh) under the block emulator: cells with
    Fmax >= F_last in order of descending Fmax, ties in ascending cell index (the selection of
    src/distribute.c and the order of sort_and_organize, src/fragment.c:484-520).  Sizes around
    the 16384-element tile; values as the sweep leaves them (-10, 0, 1 + z)."""
    rng = np.random.default_rng(n)
    F = (1.0 + rng.exponential(1.5, n)).astype(np.float32)
    kind = rng.uniform(size=n)
    F[kind < 0.25] = 0.0            # never collapses
    F[kind < 0.05] = -10.0          # the -10 of src/collapse_times.c:734-736
    if frac_ties:
        t = rng.uniform(size=n) < frac_ties
        F[t] = np.round(F[t] * 4) / 4          # many exactly equal keys
    F[rng.integers(0, n, 3)] = np.float32(np.nan)
    for Flast in (1.0, 1.75):
        out = np.full(n, 0xFFFFFFFF, dtype=np.uint32)
        lib.emu_collapsed_cells.restype = ctypes.c_longlong
        m = lib.emu_collapsed_cells(ptr(F, ctypes.POINTER(ctypes.c_float)), ctypes.c_longlong(n), ctypes.c_float(Flast),
                                    ptr(out, ctypes.POINTER(ctypes.c_uint)))
        with np.errstate(invalid="ignore"):
            sel = np.flatnonzero(F >= np.float32(Flast))
        want = sel[np.argsort(-F[sel].astype(np.float64), kind="stable")]
        assert m == want.size
        assert np.array_equal(out[:m], want.astype(np.uint32))
