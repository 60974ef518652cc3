"""CPU oracle for PINOCCHIO's collapse-time hot path.  TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the reference algorithm (pigimonaco/Pinocchio V5.1) for
GenIC -> per-radius Hessian -> ellipsoidal collapse -> Fmax/Rmax -> 2LPT/3LPT displacements.
It is the checker for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``pinocchio_b200/`` does, and the product path has no CPU fallback.

Parity pin: ``tests/test_oracle_golden.py`` checks this module against the reference's own
shipped run ``HMF_Validation/`` (sigma(R) per smoothing radius to the 4 printed digits, the
210-bin FmaxPDF histogram, the collapsed-particle count; fixtures copied to tests/golden/),
plus the GSL rng known answers (mt19937 seed 4357, ranlxd1 seed 1).

Every function cites the reference file:line it follows (paths relative to the reference root).
Third-party arithmetic that is not in the reference tree (GSL 2.7.1 rng + cspline, FFTW/PFFT)
is restated from its published algorithm.
"""
from __future__ import annotations

import math

import numpy as np

PI = 3.14159265358979323846      # src/pinocchio.h:56
SMALL = 1.0e-20                  # src/collapse_times.c:38
NBINS_PDF = 210                  # NBINS, src/pinocchio.h:65

# slot order of second_derivatives[0][ider-1]: xx,yy,zz,xy,xz,yz (src/fmax.c:235-239)
HESSIAN_PAIRS = ((1, 1), (2, 2), (3, 3), (1, 2), (1, 3), (2, 3))


# ======================================================================================
# GSL random number generators
# ======================================================================================
def mt19937_outputs(seed: int, count: int) -> np.ndarray:
    """First ``count`` 32-bit outputs of gsl_rng_mt19937 seeded with ``seed``.

    gsl mt19937 == reference MT19937 with init_genrand seeding (seed 0 -> 4357).
    """
    if seed == 0:
        seed = 4357
    bg = np.random.MT19937()
    bg._legacy_seeding(int(seed) & 0xFFFFFFFF)
    return bg.random_raw(count).astype(np.uint32)


class RanLxd1:
    """gsl_rng_ranlxd1 (Luescher RANLUX, double precision, luxury pr = 202), vectorised
    over independent generators (one per (kx,ky) column).  All generators advance in lock
    step, so the integer pointers ir/jr/ir_old are scalars.

    Seeding reproduces GSL's signed-int quirk: ``int i = seed & 0xFFFFFFFF`` followed by
    ``i % 2 ; i /= 2`` yields the binary digits of |(int32)seed|.
    """

    ONE_BIT = 1.0 / 281474976710656.0   # 2^-48
    NEXT = tuple((i + 1) % 12 for i in range(12))

    def __init__(self, seeds, pr: int = 202):
        seeds = np.atleast_1d(np.asarray(seeds, dtype=np.uint32)).astype(np.int64)
        seeds = np.where(seeds == 0, 1, seeds)
        i32 = np.where(seeds >= 2 ** 31, seeds - 2 ** 32, seeds)      # signed 32-bit view
        mag = np.abs(i32)                                              # C: digits of |i|
        n = seeds.size
        xbit = np.zeros((31, n), dtype=np.int64)
        for k in range(31):
            xbit[k] = mag % 2
            mag //= 2
        ibit, jbit = 0, 18
        self.xdbl = np.zeros((12, n))
        for k in range(12):
            x = np.zeros(n)
            for _ in range(48):
                y = (xbit[ibit] + 1) % 2
                x = x + x + y
                xbit[ibit] = (xbit[ibit] + xbit[jbit]) % 2
                ibit = (ibit + 1) % 31
                jbit = (jbit + 1) % 31
            self.xdbl[k] = self.ONE_BIT * x
        self.carry = np.zeros(n)
        self.ir, self.jr, self.ir_old = 11, 7, 0
        self.pr = pr

    def _step(self):
        ir, jr = self.ir, self.jr
        y = self.xdbl[jr] - self.xdbl[ir] - self.carry
        neg = y < 0
        self.carry = np.where(neg, self.ONE_BIT, 0.0)
        self.xdbl[ir] = np.where(neg, y + 1.0, y)
        self.ir = self.NEXT[ir]
        self.jr = self.NEXT[jr]

    def _increment_state(self):
        k = 0
        while self.ir > 0:
            self._step()
            k += 1
        while k < self.pr:
            self._step()
            k += 1
        self.ir_old = self.ir

    def get_double(self) -> np.ndarray:
        self.ir = self.NEXT[self.ir]
        if self.ir == self.ir_old:
            self._increment_state()
        return self.xdbl[self.ir].copy()

    def get(self) -> np.ndarray:
        """gsl_rng_get: (unsigned long)(get_double * 2^32)."""
        return (self.get_double() * 4294967296.0).astype(np.uint64)


# ======================================================================================
# Seed plane (src/GenIC.c:482-990, App. A.2)
# ======================================================================================
def get_map(px, py):
    """Ordinal of a point along the square spiral, src/GenIC.c:840-855."""
    px = np.asarray(px, dtype=np.int64)
    py = np.asarray(py, dtype=np.int64)
    l = 2 * np.maximum(np.abs(px), np.abs(py))
    c = (py > px).astype(np.int64) + ((px > 0) & (px == py)).astype(np.int64)
    d = np.where(c != 0, l * 3 + px + py, l - px - py)
    return (l - 1) * (l - 1) + d


def seed_table(nmesh: int, random_seed: int) -> np.ndarray:
    """SEEDTABLE[j*Nmesh + i] for the whole plane (src/GenIC.c:629-647,875-990).

    Plane coordinate c maps to spiral coordinate c (c < N/2) or c - N (transpose_subregion,
    src/GenIC.c:1017-1041); the seed is the get_map()-th output (1-based) of
    mt19937(RandomSeed).  Returned with shape [j, i] (y slow, x fast) like the reference.
    """
    n2 = nmesh // 2
    c = np.arange(nmesh, dtype=np.int64)
    s = np.where(c >= n2, c - nmesh, c)
    sx = s[None, :]          # i (x) fast
    sy = s[:, None]          # j (y) slow
    m = get_map(sx, sy)
    out = mt19937_outputs(random_seed, int(m.max()))
    return out[m - 1].astype(np.uint32)


def seed_table_old(nmesh: int, random_seed: int) -> np.ndarray:
    """The `MimicOldSeed` seed plane (internal.mimic_original_seedtable): OLDSEEDTABLE of
    src/GenIC.c:493-537, which copy_seeds_subregion (:990-1012) hands over column for column.

    N/2 square rings grown inwards from the four corners (0,0), (N,0), (0,N), (N,N) in that order;
    ring r of a corner = r cells of the column at distance r, then r + 1 cells of the row at distance
    r (corner cell last); each cell (unsigned)(0x7fffffff * gsl_rng_uniform) of ONE ranlxd1 stream
    seeded with RandomSeed.  Returned with shape [j, i] like seed_table()."""
    n = nmesh
    rng = RanLxd1([random_seed])
    t = np.zeros((n, n), dtype=np.uint32)

    def draw():
        return np.uint32(int(float(0x7FFFFFFF) * float(rng.get_double()[0])))

    for r in range(n // 2):
        for flip_row, flip_col in ((0, 0), (0, 1), (1, 0), (1, 1)):
            col = n - 1 - r if flip_col else r
            row = n - 1 - r if flip_row else r
            for m in range(r):
                t[n - 1 - m if flip_row else m, col] = draw()
            for m in range(r + 1):
                t[row, n - 1 - m if flip_col else m] = draw()
    return t


# ======================================================================================
# GenIC (src/GenIC.c:73-460)
# ======================================================================================
def genic(nmesh: int, box: float, random_seed: int, power_spectrum,
          fixed_ic: bool = False, paired_ic: bool = False, seeds: np.ndarray | None = None):
    """kdensity[x][y][kz] (complex128, shape [N,N,N/2+1]) as left by GenIC_large, including
    the final N^3 normalisation (:430-445).  ``box`` in true Mpc; ``power_spectrum(k)`` is
    the reference's PowerSpectrum (src/cosmo.c:953-1007).  Even N only."""
    assert nmesh % 2 == 0
    N, N2 = nmesh, nmesh // 2
    if seeds is None:
        seeds = seed_table(N, random_seed)          # [j, i]
    fac = (1.0 / box) ** 1.5
    kd = np.zeros((N, N, N2 + 1), dtype=np.complex128)

    # one generator per (ii, jj) column, flattened as ii*N + jj
    ii = np.repeat(np.arange(N), N)
    jj = np.tile(np.arange(N), N)
    col_seed = seeds[jj, ii]
    rng = RanLxd1(col_seed)

    def kcomp(idx):
        return np.where(idx < N2, idx, -(N - idx)) * 2 * PI / box

    kx = kcomp(ii)
    ky = kcomp(jj)
    kmag2_ij = kx * kx + ky * ky
    col_ok = (ii != N2) & (jj != N2)                 # :193, :212

    # mirrored generator for the k=0 plane (:289-368)
    mirror = (ii > N2) | ((ii == 0) & (jj > N2))
    jjj = np.where(mirror, (N - jj) % N, jj)
    iii = np.where(mirror & (ii > N2), N - ii, ii)
    k0 = RanLxd1(seeds[jjj, iii])
    ph0 = k0.get_double() * 2 * PI
    am0 = k0.get_double()
    if np.any(am0 == 0):
        raise NotImplementedError("zero amplitude draw (p = 2^-48) not vectorised")

    for kk in range(N2):
        phase = rng.get_double() * 2 * PI
        ampl = rng.get_double()
        if np.any(ampl == 0):
            raise NotImplementedError("zero amplitude draw (p = 2^-48) not vectorised")
        kz = kk * 2 * PI / box
        kmag = np.sqrt(kmag2_ij + kz * kz)
        ok = col_ok.copy()
        if kk == 0:
            ok &= ~((ii == 0) & (jj == 0))
        ok &= ~(kmag * box / (2 * PI) > 1.0 * N / 2)          # NYQUIST = 1 (:280)
        sign = np.ones(ii.size)
        if kk == 0:
            ok &= ~((ii == 0) & (jj == N2))
            ok &= ~(ii == N2)
            phase = np.where(mirror, ph0, phase)
            ampl = np.where(mirror, am0, ampl)
            sign = np.where(mirror, -1.0, 1.0)
        if paired_ic:
            phase = phase + PI
        with np.errstate(divide="ignore", invalid="ignore"):
            p_of_k = power_spectrum(np.where(kmag > 0, kmag, 1.0))
        if not fixed_ic:
            p_of_k = p_of_k * (-np.log(ampl))
        delta = fac * np.sqrt(p_of_k)
        val = delta * np.cos(phase) + 1j * sign * delta * np.sin(phase)
        val = np.where(ok, val, 0.0)
        kd[:, :, kk] = val.reshape(N, N)
    kd *= float(N) ** 3
    return kd


# ======================================================================================
# k-space derivative + c2r (src/fmax-pfft.c:255-456, 203-228; App. A.3, A.5)
# ======================================================================================
def _kgrid(N):
    n = np.arange(N)
    n = np.where(n > N // 2, n - N, n)            # index N/2 stays +N/2 (:312-313)
    k = 2.0 * PI / N * n
    kx = k[:, None, None]
    ky = k[None, :, None]
    kz = k[None, None, : N // 2 + 1]
    return kx, ky, kz


def derivative_kspace(ck: np.ndarray, d1: int, d2: int, rsmooth: float, growth: float = 1.0):
    """The k-space multiply of compute_derivative (before the c2r).  ``ck`` is [N,N,N/2+1]
    complex; returns a new array.  d1,d2 in {-1,0,1,2,3} as in the reference."""
    N = ck.shape[0]
    kx, ky, kz = _kgrid(N)
    k2 = (kx * kx + ky * ky) + kz * kz
    comp = (np.ones_like(k2), kx + 0 * k2, ky + 0 * k2, kz + 0 * k2)
    with np.errstate(divide="ignore", invalid="ignore"):
        smoothing = np.exp(-0.5 * k2 * rsmooth * rsmooth)
        if d1 == -1 and d2 == -1:
            green = np.ones_like(k2)
        elif d1 == 0 and d2 == 0:
            green = -comp[d1] * comp[d2] / k2
        else:
            green = comp[d1] * comp[d2] / k2
        factor = green * smoothing * growth
    out = np.where(k2 != 0.0, ck * np.where(k2 != 0.0, factor, 1.0), ck)
    swap = (d1 == 0 and d2 > 0) or (d1 > 0 and d2 == 0)
    if swap:
        out = -out.imag + 1j * out.real         # (re,im) -> (-im,re), all modes (:389-394)
    return out


def interpolate_growth(k, tab, logkmin=-3.0, dlogk=0.5):
    """InterpolateGrowth (src/cosmo.c:1728-1757, -DSCALE_DEPENDENT) at a fixed redshift:
    ``tab[j]`` = my_spline_eval(SPLINE[pointer + j], -log10(1+z)), j = 0..NkBINS-1; linear
    interpolation in log10 k between the k bins, the first / last bin below kmin / above kmax
    (kmin, kmax as src/cosmo.c:169-170; LOGKMIN, DELTALOGK src/def_splines.h:41-42)."""
    tab = np.asarray(tab, dtype=np.float64)
    nk = tab.size
    k = np.asarray(k, dtype=np.float64)
    kmin = 10.0 ** logkmin
    kmax = 10.0 ** (logkmin + (nk - 1) * dlogk)
    with np.errstate(divide="ignore", invalid="ignore"):
        dk = (np.log10(np.where(k > 0, k, 1.0)) - logkmin) / dlogk
    kk = np.clip(dk.astype(np.int64), 0, nk - 1)      # (int)dk; only used inside [kmin, kmax]
    frac = dk - kk
    hi = tab[np.minimum(kk + 1, nk - 1)]              # k == kmax: weight 0 on the entry past the end
    mid = frac * hi + (1 - frac) * tab[kk]
    return np.where(k < kmin, tab[0], np.where(k > kmax, tab[nk - 1], mid))


def growth_rate_of_k(N, order, log10_growth, logkmin=-3.0, dlogk=0.5):
    """growth_rate of src/fmax-pfft.c:340-364 on the whole k grid for ScaleDep.order = 1..4:
    GrowingMode*(z, k_module) = +-10^InterpolateGrowth (src/cosmo.c:1786-1819; the minus sign
    is GrowingMode_3LPT_1's, :1810).  k_module is in GRID units, as the reference passes it.
    ``log10_growth`` is the [4][NkBINS] table at the segment redshift."""
    kx, ky, kz = _kgrid(N)
    kmod = np.sqrt((kx * kx + ky * ky) + kz * kz)
    g = 10.0 ** interpolate_growth(kmod, np.asarray(log10_growth)[order - 1], logkmin, dlogk)
    return -g if order == 3 else g


def reverse_transform(ck: np.ndarray) -> np.ndarray:
    """c2r with the 1/N^3 normalisation (src/fmax-pfft.c:203-228).  numpy.fft.irfftn has
    FFTW's literal half-complex semantics (App. A.5) and includes 1/N^3."""
    N = ck.shape[0]
    return np.fft.irfftn(ck, s=(N, N, N), axes=(0, 1, 2))


def forward_transform(r: np.ndarray) -> np.ndarray:
    """r2c, unnormalised (src/fmax-pfft.c:191-200)."""
    return np.fft.rfftn(r, axes=(0, 1, 2))


def compute_derivative(ck, d1, d2, rsmooth, growth=1.0):
    return reverse_transform(derivative_kspace(ck, d1, d2, rsmooth, growth))


def second_derivatives(kdensity, radius, cell_size):
    """compute_second_derivatives (src/fmax.c:225-258): six real fields, slot order
    xx,yy,zz,xy,xz,yz; ScaleDep.order = 0 (growth 1)."""
    rs = radius / cell_size
    return [compute_derivative(kdensity, a, b, rs) for (a, b) in HESSIAN_PAIRS]


# ======================================================================================
# Ellipsoidal collapse (src/collapse_times.c:114-221, 404-427, 679-776, 1354-1362)
# ======================================================================================
def ell_classic(l1, l2, l3):
    """Vectorised ell_classic (src/collapse_times.c:114-221).  NaNs propagate as in C."""
    l1 = np.asarray(l1, dtype=np.float64)
    with np.errstate(all="ignore"):
        dele = l1 + l2 + l3
        det = l1 * l2 * l3
        den = det / 126.0 + 5.0 * l1 * dele * (dele - l1) / 84.0

        # --- |den| < SMALL branch (:135-155)
        dis = 7.0 * l1 * (l1 + 6.0 * dele)
        e2 = (7.0 * l1 - np.sqrt(dis)) / (3.0 * l1 * (l1 - dele))
        e2 = np.where(e2 < 0.0, -0.1, e2)
        e2 = np.where(dis < 0.0, -0.1, e2)
        e1 = np.where(l1 > 0.0, 1.0 / l1, -0.1)
        ell_small_den = np.where(np.abs(dele - l1) < SMALL, e1, e2)

        # --- 3rd order (:156-212)
        rden = 1.0 / den
        a1 = 3.0 * l1 * (dele - l1) / 14.0 * rden
        a1_2 = a1 * a1
        a2 = l1 * rden
        a3 = -1.0 * rden
        q = (a1_2 - 3.0 * a2) / 9.0
        r = (2.0 * a1_2 * a1 - 9.0 * a1 * a2 + 27.0 * a3) / 54.0
        r_2_q_3 = r * r - q * q * q

        fabs_r = np.abs(r)
        sq = np.power(np.sqrt(r_2_q_3) + fabs_r, 0.333333333333333)
        ea = -fabs_r / r * (sq + q / sq) - a1 / 3.0
        ea = np.where(ea < 0.0, -0.1, ea)

        sq2 = 2 * np.sqrt(q)
        inv_3 = 1.0 / 3
        t = np.arccos(2 * r / q / sq2)
        s1 = -sq2 * np.cos(t * inv_3) - a1 * inv_3
        s2 = -sq2 * np.cos((t + 2.0 * PI) * inv_3) - a1 * inv_3
        s3 = -sq2 * np.cos((t + 4.0 * PI) * inv_3) - a1 * inv_3
        s1 = np.where(s1 < 0.0, 1.0e10, s1)
        s2 = np.where(s2 < 0.0, 1.0e10, s2)
        s3 = np.where(s3 < 0.0, 1.0e10, s3)
        eb = np.where(s1 < s2, s1, s2)
        eb = np.where(s3 < eb, s3, eb)
        eb = np.where(eb == 1.0e10, -0.1, eb)

        ell3 = np.where(r_2_q_3 > 0, ea, eb)
        ell = np.where(np.abs(den) < SMALL, ell_small_den, ell3)
        ell = np.where(np.abs(l1) < SMALL, -0.1, ell)

        # spherical-collapse correction (:215-218)
        inv_del = 1.0 / dele
        corr = -0.364 * inv_del * np.exp(-6.5 * (l1 - l2) * inv_del - 2.8 * (l2 - l3) * inv_del)
        ell = np.where((dele > 0.0) & (ell > 0.0), ell + corr, ell)
    return ell


def eigenvalues(h):
    """Eigenvalues of the Hessian as in inverse_collapse_time (src/collapse_times.c:679-749).
    ``h`` = sequence of six arrays (xx,yy,zz,xy,xz,yz).  Returns (x1,x2,x3,bad) with x1>=x2>=x3
    (``ord``, :1354-1362) and ``bad`` flagging the -10 early return (:734-736)."""
    d0, d1, d2, d3, d4, d5 = [np.asarray(a, dtype=np.float64) for a in h]
    with np.errstate(all="ignore"):
        mu1 = d0 + d1 + d2
        mu1_2 = mu1 * mu1
        mu2 = 0.5 * mu1_2
        mu2 = mu2 - 0.5 * ((d0 * d0 + d1 * d1) + d2 * d2)
        a0, a1_, a2_ = d3 * d3, d4 * d4, d5 * d5
        mu2 = mu2 - ((a0 + a1_) + a2_)
        mu3 = d0 * d1 * d2 + 2.0 * d3 * d4 * d5 - d0 * a2_ - d1 * a1_ - d2 * a0
        q = (mu1_2 - 3.0 * mu2) / 9.0
        r = -(2.0 * mu1_2 * mu1 - 9.0 * mu1 * mu2 + 27.0 * mu3) / 54.0
        bad = (q != 0.0) & ((q * q * q < r * r) | (q < 0.0))
        sq = 2 * np.sqrt(q)
        t = np.arccos(2 * r / q / sq)
        inv_3 = 1.0 / 3.0
        x1 = -sq * np.cos(t * inv_3) + mu1 * inv_3
        x2 = -sq * np.cos((t + 2.0 * PI) * inv_3) + mu1 * inv_3
        x3 = -sq * np.cos((t + 4.0 * PI) * inv_3) + mu1 * inv_3
        diag = q == 0.0
        x1 = np.where(diag, d0, x1)
        x2 = np.where(diag, d1, x2)
        x3 = np.where(diag, d2, x3)
        # ord(): C macros max/min (NaN-propagation as `a>b?a:b`)
        hi = np.where(np.where(x1 > x2, x1, x2) > x3, np.where(x1 > x2, x1, x2), x3)
        lo = np.where(np.where(x1 < x2, x1, x2) < x3, np.where(x1 < x2, x1, x2), x3)
        mid = x1 + x2 + x3 - lo - hi
    return hi, mid, lo, bad


def inverse_collapse_time(h, inverse_growing_mode):
    """F = 1 + z_collapse per cell (src/collapse_times.c:679-776 with ELL_CLASSIC, :404-415).
    ``inverse_growing_mode(b_c)`` is InverseGrowingMode (src/cosmo.c:1822-1832)."""
    x1, x2, x3, bad = eigenvalues(h)
    bc = ell_classic(x1, x2, x3)
    with np.errstate(all="ignore"):
        pos = bc > 0.0
        F = np.where(pos, 1.0 + inverse_growing_mode(np.where(pos, bc, 1.0)), 0.0)
    F = np.where(bad, -10.0, F)
    return F


# ======================================================================================
# TABULATED_CT and ELL_SNG (src/collapse_times.c:239-400, 780-1346)
# ======================================================================================
CT_NBINS_XY, CT_NBINS_D = 50, 100                      # src/collapse_times.c:781-782
CT_SQUEEZE, CT_EXPO, CT_RANGE_D, CT_RANGE_X, CT_DELTA0 = 1.2, 1.75, 7.0, 3.5, -1.0   # :783-787


def ct_delta_vector(nd: int = CT_NBINS_D) -> np.ndarray:
    """delta_vector of initialize_collapse_times (src/collapse_times.c:836-877): bins growing like
    |delta - CT_DELTA0|^(CT_EXPO-1) away from CT_DELTA0, never smaller than CT_SQUEEZE * ref_interval."""
    deltaf = (CT_SQUEEZE / CT_EXPO) ** (1.0 / (CT_EXPO - 1.0))
    ref = (((CT_RANGE_D - CT_DELTA0) ** (2.0 - CT_EXPO) + (CT_RANGE_D + CT_DELTA0) ** (2.0 - CT_EXPO)
            - 2.0 * deltaf ** (2.0 - CT_EXPO)) / CT_EXPO / (2.0 - CT_EXPO) + 2.0 * deltaf / CT_SQUEEZE) / (nd - 2.0)
    dv = np.zeros(nd)
    d = -CT_RANGE_D
    for i in range(nd):
        dv[i] = d
        interval = CT_EXPO * ref * abs(d - CT_DELTA0) ** (CT_EXPO - 1.0)
        if interval / ref < CT_SQUEEZE:
            interval = ref * CT_SQUEEZE
        d += interval
    return dv


def ct_table_lambdas(ampl: float, dv=None, nxy: int = CT_NBINS_XY):
    """(l1, l2, l3) of every table point, arrays [iy][ix][id] (src/collapse_times.c:966-976)."""
    dv = ct_delta_vector() if dv is None else np.asarray(dv)
    bin_x = CT_RANGE_X / nxy
    y, x, d = np.meshgrid(np.arange(nxy) * bin_x, np.arange(nxy) * bin_x, dv, indexing="ij")
    return (d + 2.0 * x + y) / 3.0 * ampl, (d - x + y) / 3.0 * ampl, (d - x - 2.0 * y) / 3.0 * ampl


def ct_table_classic(ampl: float, inverse_growing_mode, dv=None, nxy: int = CT_NBINS_XY) -> np.ndarray:
    """CT_table of one radius with ELL_CLASSIC: ell() = 1 + InverseGrowingMode(b_c) or 0 (:404-415, 977)."""
    l1, l2, l3 = ct_table_lambdas(ampl, dv, nxy)
    bc = ell_classic(l1, l2, l3)
    with np.errstate(all="ignore"):
        pos = bc > 0.0
        return np.where(pos, 1.0 + inverse_growing_mode(np.where(pos, bc, 1.0)), 0.0)


H_OVER_C = 100.0 / 299792.458       # H_over_c = 100 / SPEEDOFLIGHT (src/cosmo.c:109, src/pinocchio.h:63)


def force_modification(size, a, delta, omega0, omega_lambda, fr0):
    """ForceModification of Hu-Sawicki f(R) gravity (-DMOD_GRAV_FR, src/collapse_times.c:294-311)."""
    ff = 4.0 * omega_lambda / omega0
    with np.errstate(all="ignore"):
        thickness = fr0 / omega0 / (H_OVER_C * size) ** 2 * a ** 7 * np.float64(1.0 + delta) ** (-1.0 / 3.0) * \
            (((1.0 + ff) / (1.0 + ff * a ** 3)) ** 2 - ((1.0 + ff) / (1.0 + delta + ff * a ** 3)) ** 2)
    f3 = thickness * (3.0 + thickness * (-3.0 + thickness))
    if f3 < 0.0:
        f3 = 0.0
    return f3 / 3.0 if f3 < 1.0 else 1.0 / 3.0


def sng_system(t, y, omega0, omega_lambda, omega_rad=0.0, omega_k=None, fr0=0.0, fr_size=1.0):
    """r.h.s. of the nine eigenvalue equations of Nadkarni-Ghosh & Singhal (src/collapse_times.c:239-290);
    OmegaMatter / OmegaLambda of src/cosmo.c:1675-1718 for a cosmological constant; fr0 > 0: the gravity term
    of the velocity equations times 1 + ForceModification(fr_size, a, delta) (-DMOD_GRAV_FR, :276-277)."""
    if omega_k is None:
        omega_k = 1.0 - omega0 - omega_lambda - omega_rad
    z = 1.0 / t - 1.0
    e2 = omega_rad * (1 + z) ** 4 + omega0 * (1 + z) ** 3 + omega_k * (1 + z) ** 2 + omega_lambda
    e2 /= omega_rad + omega0 + omega_k + omega_lambda
    om, ol = omega0 * (1 + z) ** 3 / e2, omega_lambda / e2
    delta = y[6] + y[7] + y[8]
    grav = 1.0 + force_modification(fr_size, t, delta, omega0, omega_lambda, fr0) if fr0 else 1.0
    f = [0.0] * 9
    for i in range(3):
        s = 0.0
        for j in range(3):
            if i == j or y[i] == y[j]:
                continue
            s += (y[j + 6] - y[i + 6]) * ((1 - y[i]) ** 2 * (1 + y[i + 3]) - (1 - y[j]) ** 2 * (1 + y[j + 3])) / \
                 ((1 - y[i]) ** 2 - (1 - y[j]) ** 2)
        f[i] = y[i + 3] * (y[i] - 1.0) / t
        f[i + 3] = 0.5 * (y[i + 3] * (om - 2.0 * ol - 2.0) - 3.0 * om * y[i + 6] * grav - 2.0 * y[i + 3] ** 2) / t
        f[i + 6] = ((5.0 / 6.0 + y[i + 6]) * ((3.0 + y[3] + y[4] + y[5]) - (1.0 + delta) / (2.5 + delta) * (y[3] + y[4] + y[5]))
                    - (2.5 + delta) * (1.0 + y[i + 3]) + s) / t
    return np.array(f)


_RKF45_A = (0.25, 0.375, 12.0 / 13.0, 1.0, 0.5)
_RKF45_B = ((0.25,), (3 / 32, 9 / 32), (1932 / 2197, -7200 / 2197, 7296 / 2197), (8341 / 4104, -32832 / 4104, 29440 / 4104, -845 / 4104),
            (-6080 / 20520, 41040 / 20520, -28352 / 20520, 9295 / 20520, -5643 / 20520))
_RKF45_C = (902880 / 7618050, 0.0, 3953664 / 7618050, 3855735 / 7618050, -1371249 / 7618050, 277020 / 7618050)
_RKF45_E = (1 / 360, 0.0, -128 / 4275, -2197 / 75240, 1 / 50, 2 / 55)


def ell_sng(l1, l2, l3, D_in, omega0, omega_lambda, omega_rad=0.0, fr0=0.0, fr_size=1.0):
    """ell_sng (src/collapse_times.c:315-400) for one point: gsl_odeiv2 rkf45 (GSL 2.7 rkf45.c) driven by
    evolve_apply with control_standard_new(1e-6, 1e-6, 1, 1) (cstd.c: shrink by 0.9 r^-1/5 >= 0.2 above 1.1,
    grow by 0.9 r^-1/6 <= 5 below 0.5) from a = 1e-5 to 5; collapse when lambda_a1 >= 0.99999, the epoch
    interpolated linearly from the INITIAL point as the reference does (olda / oldlam are never advanced)."""
    amin, amax = 1.0e-5, 5.0
    rhs = lambda t, y: sng_system(t, y, omega0, omega_lambda, omega_rad, None, fr0, fr_size)
    y = np.array([l1 * D_in, l2 * D_in, l3 * D_in, l1 * D_in / (l1 * D_in - 1.0), l2 * D_in / (l2 * D_in - 1.0),
                  l3 * D_in / (l3 * D_in - 1.0), l1 * D_in, l2 * D_in, l3 * D_in])
    t, hh, olda, oldlam = amin, 1.0e-6, amin, l1 * D_in
    with np.errstate(all="ignore"):
        k1 = rhs(t, y)
        while t < amax:
            h0, final = hh, False
            if h0 > amax - t:
                h0, final = amax - t, True
            while True:
                k = [k1]
                for st in range(5):
                    k.append(rhs(t + _RKF45_A[st] * h0, y + h0 * sum(b * kk for b, kk in zip(_RKF45_B[st], k))))
                ynew = y + h0 * sum(c * kk for c, kk in zip(_RKF45_C, k))
                yerr = h0 * sum(e * kk for e, kk in zip(_RKF45_E, k))
                dnew = rhs(t + h0, ynew)
                d0 = 1.0e-6 * (np.abs(ynew) + np.abs(h0 * dnew)) + 1.0e-6
                r = np.abs(yerr) / np.abs(d0)
                rmax = max(2.2250738585072014e-308, np.nanmax(r) if np.any(~np.isnan(r)) else 0.0)
                if rmax > 1.1:
                    hnew = h0 * max(0.2, 0.9 / rmax ** 0.2)
                    if not abs(hnew) < abs(h0) or t + hnew == t:
                        return -1.0
                    h0, final = hnew, False
                    continue
                break
            hn = h0 * min(5.0, max(1.0, 0.9 / rmax ** (1.0 / 6.0))) if rmax < 0.5 else h0
            y, k1 = ynew, dnew
            t = amax if final else t + h0
            if not final:
                hh = hn
            if y[0] >= 0.99999:
                return olda + (1.0 - oldlam) * (t - olda) / (y[0] - oldlam)
    return 0.0


def _natural_spline_c(x, y):
    """second-derivative coefficients c of gsl_interp_cspline for many columns at once: y [..., n]"""
    n = x.size
    h = np.diff(x)
    dy = np.diff(y, axis=-1)
    diag = 2.0 * (h[:-1] + h[1:])
    off = h[1:-1]
    rhs = 3.0 * (dy[..., 1:] / h[1:] - dy[..., :-1] / h[:-1])
    m = n - 2
    cp = np.zeros(m)
    dp = np.zeros(rhs.shape)
    cp[0] = off[0] / diag[0]
    dp[..., 0] = rhs[..., 0] / diag[0]
    for i in range(1, m):
        den = diag[i] - off[i - 1] * cp[i - 1]
        if i < m - 1:
            cp[i] = off[i] / den
        dp[..., i] = (rhs[..., i] - off[i - 1] * dp[..., i - 1]) / den
    c = np.zeros(y.shape)
    c[..., m] = dp[..., m - 1]
    for i in range(m - 2, -1, -1):
        c[..., i + 1] = dp[..., i] - cp[i] * c[..., i + 2]
    return c


def interpolate_collapse_time(table, dv, ampl, l1, l2, l3):
    """interpolate_collapse_time, BILINEAR_SPLINE (src/collapse_times.c:1132-1147, 1211-1221): four natural
    cubic splines in delta = (l1+l2+l3)/ampl (my_spline_eval: linear extrapolation outside the knots,
    src/cosmo.c:2016-2027) blended bilinearly in x = (l1-l2)/ampl, y = (l2-l3)/ampl.  table [iy][ix][id]."""
    table, dv = np.asarray(table), np.asarray(dv)
    nxy, nd = table.shape[0], dv.size
    bin_x = CT_RANGE_X / nxy
    c = _natural_spline_c(dv, table)
    d = (l1 + l2 + l3) / ampl
    x = (l1 - l2) / ampl
    y = (l2 - l3) / ampl
    ix = np.clip((x / bin_x).astype(np.int64), 0, nxy - 2)
    iy = np.clip((y / bin_x).astype(np.int64), 0, nxy - 2)
    dx, dy = x / bin_x - ix, y / bin_x - iy
    i = np.clip(np.searchsorted(dv, d, side="right") - 1, 0, nd - 2)
    hh = dv[i + 1] - dv[i]
    t = d - dv[i]

    def column(jx, jy):
        y0, y1, c0, c1 = table[jy, jx, i], table[jy, jx, i + 1], c[jy, jx, i], c[jy, jx, i + 1]
        b = (y1 - y0) / hh - hh * (c1 + 2.0 * c0) / 3.0
        val = y0 + t * (b + t * (c0 + t * (c1 - c0) / (3.0 * hh)))
        lo = table[jy, jx, 0] + (d - dv[0]) * (table[jy, jx, 1] - table[jy, jx, 0]) / (dv[1] - dv[0])
        hi = table[jy, jx, -1] + (d - dv[-1]) * (table[jy, jx, -1] - table[jy, jx, -2]) / (dv[-1] - dv[-2])
        return np.where(d < dv[0], lo, np.where(d > dv[-1], hi, val))

    return ((1.0 - dx) * (1.0 - dy) * column(ix, iy) + dx * (1.0 - dy) * column(ix + 1, iy)
            + (1.0 - dx) * dy * column(ix, iy + 1) + dx * dy * column(ix + 1, iy + 1))


def inverse_collapse_time_tab(h, table, dv, ampl):
    """inverse_collapse_time with -DTABULATED_CT (src/collapse_times.c:679-776): eigenvalues as before, F from
    the table of this smoothing radius."""
    x1, x2, x3, bad = eigenvalues(h)
    with np.errstate(all="ignore"):
        F = interpolate_collapse_time(table, dv, ampl, np.where(bad, 0.0, x1), np.where(bad, 0.0, x2), np.where(bad, 0.0, x3))
    return np.where(bad, -10.0, F)


def ill_conditioned_mask(h, inverse_growing_mode, eps=1e-13, ntrial=2, tol=1e-7, seed=0):
    """Cells whose F is numerically ill-conditioned in the REFERENCE algorithm itself.

    ell_classic normalises the cubic by its leading coefficient `den` (src/collapse_times.c:133,
    160-169); when den is close to zero the coefficients reach 1e9..1e26 and r*r - q*q*q cancels
    catastrophically, so F jumps by O(1) under 1e-15 relative changes of the Hessian (i.e. under
    a different compiler, FMA contraction or libm).  For such cells "the reference's value" is
    not defined to 1e-6; parity tests flag them (F changes by more than `tol` under `eps`
    relative perturbations) and count them instead of comparing them."""
    rng = np.random.default_rng(seed)
    F0 = inverse_collapse_time(h, inverse_growing_mode)
    mask = np.zeros(F0.shape, dtype=bool)
    # FFT round-off is absolute (~1e-16 of the field's largest value), not relative to each
    # entry, and most of the jumps are branch switches of ell_classic: a one-sided random
    # trial crosses a nearby switch only half of the time, so every trial is applied with both
    # signs (antithetic pair) and carries an absolute as well as a relative part.
    scale = max(float(np.abs(a).max()) for a in h)
    for _ in range(ntrial):
        noise = [eps * (a * rng.standard_normal(a.shape) + scale * rng.standard_normal(a.shape)) for a in h]
        for sgn in (1.0, -1.0):
            Fp = inverse_collapse_time([a + sgn * d for a, d in zip(h, noise)], inverse_growing_mode)
            with np.errstate(invalid="ignore"):
                mask |= ~(np.abs(Fp - F0) <= tol * np.maximum(1.0, np.abs(F0)))
    return mask


def init_products(shape):
    """ismooth == 0 initialisation (src/collapse_times.c:461-492)."""
    return np.full(shape, -10.0, dtype=np.float32), np.full(shape, -1, dtype=np.int32)


def update_fmax(Fmax, Rmax, Fnew, ismooth):
    """Running max (src/collapse_times.c:587-590): float Fmax promoted to double for the
    compare; store (float)Fnew."""
    upd = Fmax.astype(np.float64) < Fnew
    Fmax[upd] = Fnew[upd].astype(np.float32)
    Rmax[upd] = ismooth
    return upd


def true_variance(h):
    """Sum(delta^2)/N^3 with delta = trace of the Hessian (src/collapse_times.c:554-556,662)."""
    delta = h[0] + h[1] + h[2]
    return float(np.sum(delta * delta) / delta.size), float(np.sum(delta) / delta.size)


def fmax_pdf(Fmax):
    """Fmax_PDF histogram (src/fmax.c:509-550)."""
    xF = (Fmax.astype(np.float64) * 10.0).astype(np.int64)      # (int)(float*10.) truncation
    xF = np.clip(xF, 0, NBINS_PDF - 1)
    return np.bincount(xF.ravel(), minlength=NBINS_PDF).astype(np.uint64)


# ======================================================================================
# LPT sources and displacements (src/LPT.c:32-235, src/fmax.c:193-222, 292-367)
# ======================================================================================
def lpt_sources(h):
    """source_2LPT, source_3LPT_1 and the first part of source_3LPT_2 (src/LPT.c:64-93)."""
    s0, s1, s2, s3, s4, s5 = h
    src2 = s0 * s1 + s0 * s2 + s1 * s2 - s3 * s3 - s4 * s4 - s5 * s5
    src31 = 3.0 * (s0 * (s1 * s2 - s5 * s5) - s3 * (s3 * s2 - s4 * s5) + s4 * (s3 * s5 - s4 * s1))
    src32 = 2.0 * (s0 + s1 + s2) * src2
    return src2, src31, src32


def lpt_kvectors(h):
    """kvector_2LPT, kvector_3LPT_1, kvector_3LPT_2 (src/LPT.c:98-172).  ``h`` are the six
    Hessians of the R = 0 radius; Rsmooth = 0, ScaleDep.order = 0 throughout."""
    src2, src31, src32 = lpt_sources(h)
    k2 = forward_transform(src2)
    src32 = src32.copy()
    for idx, (a, b) in enumerate(HESSIAN_PAIRS):
        ider = idx + 1
        rv = compute_derivative(k2, a, b, 0.0)
        src32 -= 2.0 * (1.0 if ider <= 3 else 2.0) * rv * h[idx]
    k31 = forward_transform(src31)
    k32 = forward_transform(src32)
    return k2, k31, k32


def first_derivatives(kvec, growth):
    """compute_first_derivatives (src/fmax.c:193-222): three real fields, Rsmooth = 0, then
    cast to PRODFLOAT=float by write_from_rvector_to_products (src/fmax-pfft.c:563-631).
    ``growth``: a scalar, or the per-mode array of growth_rate_of_k (-DSCALE_DEPENDENT)."""
    return [compute_derivative(kvec, ia, 0, 0.0, growth).astype(np.float32) for ia in (1, 2, 3)]


# ======================================================================================
# Driver: compute_fmax + compute_displacements (src/fmax.c:36-190, 292-367)
# ======================================================================================
def compute_fmax(kdensity, radii, cell_size, inverse_growing_mode, growth=None,
                 lpt_order=3, keep=False):
    """Returns dict with Fmax (f32), Rmax (i32), TrueVariance[], Vel, Vel_2LPT, Vel_3LPT_1,
    Vel_3LPT_2 (each [3] list of f32 fields).  ``growth`` = (D, D2, D31, D32) at the segment
    redshift (growth_rate of src/fmax-pfft.c:344-364), scalars or per-mode arrays
    (growth_rate_of_k); default all ones."""
    N = kdensity.shape[0]
    if growth is None:
        growth = (1.0, 1.0, 1.0, 1.0)
    Fmax, Rmax = init_products((N, N, N))
    tv = []
    extra = {}
    h = None
    for ismooth, R in enumerate(radii):
        h = second_derivatives(kdensity, R, cell_size)
        Fnew = inverse_collapse_time(h, inverse_growing_mode)
        update_fmax(Fmax, Rmax, Fnew, ismooth)
        tv.append(true_variance(h)[0])
        if keep:
            extra.setdefault("F", []).append(Fnew)
            extra["unstable"] = extra.get("unstable", False) | ill_conditioned_mask(h, inverse_growing_mode)
    out = {"Fmax": Fmax, "Rmax": Rmax, "TrueVariance": np.array(tv)}
    if lpt_order >= 2:
        if lpt_order >= 3:
            k2, k31, k32 = lpt_kvectors(h)
            out["Vel_3LPT_1"] = first_derivatives(k31, growth[2])
            out["Vel_3LPT_2"] = first_derivatives(k32, growth[3])
        else:
            k2 = forward_transform(lpt_sources(h)[0])
        out["Vel_2LPT"] = first_derivatives(k2, growth[1])
        if keep:
            extra["kvector_2LPT"] = k2
            if lpt_order >= 3:
                extra["kvector_3LPT_1"], extra["kvector_3LPT_2"] = k31, k32
    out["Vel"] = first_derivatives(kdensity, growth[0])
    if keep:
        extra["hessian_R0"] = h
        out.update(extra)
    return out


# product_data layout for TWO_LPT+THREE_LPT, float products, no SNAPSHOT/RECOMPUTE
# (src/pinocchio.h:233-259): sizeof = 56
PRODUCT_DTYPE_3LPT = np.dtype([("Rmax", "<i4"), ("Fmax", "<f4"), ("Vel", "<f4", 3),
                               ("Vel_2LPT", "<f4", 3), ("Vel_3LPT_1", "<f4", 3),
                               ("Vel_3LPT_2", "<f4", 3)])
assert PRODUCT_DTYPE_3LPT.itemsize == 56


def pack_products(res) -> np.ndarray:
    """AoS products[] (index = z + N*(y + N*x), src/pinocchio.h:84-85)."""
    n = res["Fmax"].size
    p = np.zeros(n, dtype=PRODUCT_DTYPE_3LPT)
    p["Rmax"] = res["Rmax"].ravel()
    p["Fmax"] = res["Fmax"].ravel()
    for name in ("Vel", "Vel_2LPT", "Vel_3LPT_1", "Vel_3LPT_2"):
        if name in res:
            for a in range(3):
                p[name][:, a] = res[name][a].ravel()
    return p


# ---------------------------------------------------------------------------------------------------------------
# set_scaledep_GM (src/initialization.c:1533-2026): the variance integrals behind the per-radius inverse-growth
# splines and the k_GM_* wavenumbers
# ---------------------------------------------------------------------------------------------------------------
def window_function(kind: int, kr):
    """WindowFunction, src/cosmo.c:1611-1646: 0 Gaussian, 2 top-hat (kr < 1e-5 -> 1)."""
    kr = np.asarray(kr, dtype=np.float64)
    if kind == 0:
        return np.exp(-kr * kr / 2.0)
    out = np.ones_like(kr)
    m = kr >= 1.0e-5
    x = kr[m]
    out[m] = 3.0 * (np.sin(x) / (x * x) / x - np.cos(x) / (x * x))
    return out


def interpolate_growth_table(tab, logk, logkmin, dlogk):
    """InterpolateGrowth (src/cosmo.c:1728-1755) on a [nkbins][ntimes] table of spline values at the time knots:
    returns [ntimes][len(logk)].  Below kmin the first bin, above kmax the last, linear in log10 k between."""
    tab = np.asarray(tab, dtype=np.float64)
    nk = tab.shape[0]
    logk = np.asarray(logk, dtype=np.float64)
    if nk == 1:
        return np.repeat(tab[0][:, None], logk.size, axis=1)
    dk = (logk - logkmin) / dlogk
    kk = np.clip(dk.astype(np.int64), 0, nk - 2)
    fr = dk - kk
    val = fr[None, :] * tab[kk + 1].T + (1.0 - fr)[None, :] * tab[kk].T
    val = np.where((logk < logkmin)[None, :], tab[0][:, None], val)
    val = np.where((logk > logkmin + (nk - 1) * dlogk)[None, :], tab[nk - 1][:, None], val)
    return val


def scaledep_variances(logk, a_dens, a_disp, log10_growth, fomega, logkmin, dlogk, radius_dens, radius_disp):
    """The three integrals of set_scaledep_GM for every radius and time knot on a FIXED quadrature (nodes logk, the
    node weights folded into a_dens = w P k^3 / 2 pi^2 and a_disp = w P k / 2 pi^2):
    out[0] density, IntegrandForSDDensVariance (src/initialization.c:1439-1447, Gaussian window);
    out[1] displacement, IntegrandForSDDisplVariance (:1449-1457, top-hat);
    out[2] velocity, IntegrandForSDVelVariance (:1489-1498, top-hat, times fomega^2); each the sqrt of the integral
    (the reference's vector[i], :1600, :1747, :1891)."""
    logk = np.asarray(logk, dtype=np.float64)
    k = 10.0 ** logk
    D2 = (10.0 ** interpolate_growth_table(log10_growth, logk, logkmin, dlogk)) ** 2          # [nt][n]
    V2 = D2 * interpolate_growth_table(fomega, logk, logkmin, dlogk) ** 2
    rd, rp = np.asarray(radius_dens, dtype=np.float64), np.asarray(radius_disp, dtype=np.float64)
    Td = np.asarray(a_dens)[None, :] * window_function(0, k[None, :] * rd[:, None]) ** 2      # [ns][n]
    Tt = np.asarray(a_disp)[None, :] * window_function(2, k[None, :] * rp[:, None]) ** 2
    return np.sqrt(np.stack([Td @ D2.T, Tt @ D2.T, Tt @ V2.T]))
