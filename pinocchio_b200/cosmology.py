"""Host-side cosmology tables consumed by the collapse-time hot path.

In the reference these tables are built once on the CPU by ``initialize_cosmology``
(src/cosmo.c:84-437), ``normalize_PowerSpectrum`` (src/cosmo.c:1058-1085),
``initialize_MassVariance`` (src/cosmo.c:1507-1560) and ``set_smoothing``
(src/initialization.c:386-435).  They stay on the host here too (SURVEY.md §2.1 marks
cosmo.c out of scope for the GPU); only their *outputs* -- natural-cubic-spline knots of
the growth factors, sqrt(P(k)) on the integer |n|^2 lattice, and the smoothing-radius
ladder -- cross the C ABI (include/pinb200.h).

Only the Lambda-CDM / Eisenstein-Hu branch with scale-independent growth (NkBINS = 1,
src/def_splines.h:37-38) is implemented; that is what configs 1-4 of BASELINE.json use.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
from scipy.integrate import quad, solve_ivp

NBINS = 210          # src/pinocchio.h:65
NBB = 10             # src/cosmo.c:35
NSIGMA = 6.0         # src/pinocchio.h:68
STEP_VAR = 0.3       # src/pinocchio.h:69
TOLERANCE = 1.0e-4   # src/pinocchio.h:452
NWINT = 1000         # src/pinocchio.h:449
DELTA_C = 1.686


# --------------------------------------------------------------------------------------
# natural cubic spline, GSL `gsl_interp_cspline` semantics
# --------------------------------------------------------------------------------------
class NaturalSpline:
    """Natural cubic spline with the reference's linear extrapolation.

    ``c`` follows gsl/interpolation/cspline.c (tridiagonal system with c[0]=c[n-1]=0);
    ``b`` and ``d`` are the per-interval coefficients GSL derives at evaluation time.
    ``__call__`` mirrors ``my_spline_eval`` (src/cosmo.c:2016-2027): secant extrapolation
    outside the knot range.
    """

    def __init__(self, x, y):
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        n = x.size
        assert n >= 3 and np.all(np.diff(x) > 0), "spline abscissae must increase"
        h = np.diff(x)
        dy = np.diff(y)
        c = np.zeros(n)
        if n > 2:
            # rows i = 0..n-3 of the (n-2)x(n-2) system for c[1..n-2]
            diag = 2.0 * (h[:-1] + h[1:])
            off = h[1:-1].copy()
            rhs = 3.0 * (dy[1:] / h[1:] - dy[:-1] / h[:-1])
            # Thomas algorithm (symmetric tridiagonal)
            m = diag.size
            cp = np.zeros(m)
            dp = np.zeros(m)
            cp[0] = off[0] / diag[0] if m > 1 else 0.0
            dp[0] = rhs[0] / diag[0]
            for i in range(1, m):
                den = diag[i] - off[i - 1] * cp[i - 1]
                if i < m - 1:
                    cp[i] = off[i] / den
                dp[i] = (rhs[i] - off[i - 1] * dp[i - 1]) / den
            sol = np.zeros(m)
            sol[-1] = dp[-1]
            for i in range(m - 2, -1, -1):
                sol[i] = dp[i] - cp[i] * sol[i + 1]
            c[1:-1] = sol
        self.x, self.y, self.c = x, y, c
        self.b = dy / h - h * (c[1:] + 2.0 * c[:-1]) / 3.0
        self.d = (c[1:] - c[:-1]) / (3.0 * h)
        self.size = n

    def __call__(self, xq):
        xq = np.asarray(xq, dtype=np.float64)
        x, y = self.x, self.y
        i = np.clip(np.searchsorted(x, xq, side="right") - 1, 0, self.size - 2)
        dx = xq - x[i]
        val = y[i] + dx * (self.b[i] + dx * (self.c[i] + dx * self.d[i]))
        lo = y[0] + (xq - x[0]) * (y[1] - y[0]) / (x[1] - x[0])
        hi = y[-1] + (xq - x[-1]) * (y[-1] - y[-2]) / (x[-1] - x[-2])
        return np.where(xq < x[0], lo, np.where(xq > x[-1], hi, val))

    def packed(self):
        """(5, n) float64 table [x, y, b, c, d] handed to the device (b, d padded)."""
        n = self.size
        t = np.zeros((5, n))
        t[0], t[1], t[3] = self.x, self.y, self.c
        t[2, : n - 1], t[4, : n - 1] = self.b, self.d
        return np.ascontiguousarray(t)


# --------------------------------------------------------------------------------------
@dataclass
class CosmoParams:
    Omega0: float = 0.25
    OmegaLambda: float = 0.75
    OmegaBaryon: float = 0.044
    Hubble100: float = 0.70
    Sigma8: float = 0.8
    PrimordialIndex: float = 0.96
    noradiation: bool = True          # -DNORADIATION (src/cosmo.c:36-40)


@dataclass
class Cosmology:
    """Tables of ``initialize_cosmology`` for simple-Lambda EH cosmologies."""

    p: CosmoParams = field(default_factory=CosmoParams)
    pk_norm_override: float | None = None     # use a logged PkNorm (printed precision) if given

    def __post_init__(self):
        p = self.p
        self.OmegaRad = (0.0 if p.noradiation else 4.2e-5) / p.Hubble100 ** 2
        self.OmegaK = 1.0 - p.Omega0 - p.OmegaLambda - self.OmegaRad
        self._growth_tables()
        self.PkNorm = 1.0
        if self.pk_norm_override is not None:
            self.PkNorm = float(self.pk_norm_override)
        else:
            # normalize_PowerSpectrum, src/cosmo.c:1058-1085 (top-hat window, R = 8/h true Mpc)
            self.PkNorm = p.Sigma8 ** 2 / self.compute_mass_variance(8.0 / p.Hubble100, window=2)
        self._mass_variance_tables()

    # ---- background ------------------------------------------------------------
    def E2(self, a):
        p = self.p
        return p.Omega0 / a ** 3 + self.OmegaK / a ** 2 + self.OmegaRad / a ** 4 + p.OmegaLambda

    def dlnE2_da(self, a):
        p = self.p
        dE2 = -3.0 * p.Omega0 / a ** 4 - 2.0 * self.OmegaK / a ** 3 - 4.0 * self.OmegaRad / a ** 5
        return dE2 / self.E2(a)

    # ---- growth ODEs, src/cosmo.c:659-703 with ICs :203-217 -------------------------
    def _rhs(self, a, y):
        p = self.p
        E2 = self.E2(a)
        a1 = -(3.0 / a + 0.5 * self.dlnE2_da(a))
        b1 = 1.5 * p.Omega0 / (E2 * a ** 5)
        d = np.empty(9)
        d[0] = 1.0 / a / math.sqrt(E2)
        d[1] = a1 * y[1] + b1 * y[2]
        d[2] = y[1]
        d[3] = a1 * y[3] + b1 * y[4] - b1 * y[2] * y[2]
        d[4] = y[3]
        d[5] = a1 * y[5] + b1 * y[6] - 2.0 * b1 * y[2] ** 3
        d[6] = y[5]
        d[7] = a1 * y[7] + b1 * y[8] - 2.0 * b1 * y[2] * y[4] + 2.0 * b1 * y[2] ** 3
        d[8] = y[7]
        return d

    def _growth_tables(self):
        log_amin = -4.0
        dloga = -log_amin / (NBINS - NBB)
        a_knots = np.array([10.0 ** (log_amin + i * dloga) for i in range(NBINS)])
        for i in range(NBINS):
            if abs(log_amin + i * dloga) < dloga / 10.0:
                a_knots[i] = 1.0
        x1 = 10.0 ** (log_amin - 2.0)
        y0 = np.array([2.0 / 3.0 * x1 ** 1.5, 1.0, x1, -6.0 / 7.0 * x1, -3.0 / 7.0 * x1 * x1,
                       -x1 * x1, -x1 ** 3 / 3.0, 10.0 / 7.0 * x1 * x1, 10.0 / 21.0 * x1 ** 3])
        sol = solve_ivp(self._rhs, (x1, a_knots[-1]), y0, method="DOP853", t_eval=a_knots,
                        rtol=1e-12, atol=1e-30)
        assert sol.success
        y = sol.y
        today = int(np.argmax(a_knots >= 1.0))
        norm = y[2, today]
        self.a_knots = a_knots
        self.today = today
        self.grow1 = y[2] / norm
        self.grow2 = -y[4] / norm ** 2
        self.grow31 = -y[6] / 3.0 / norm ** 3
        self.grow32 = y[8] / 4.0 / norm ** 3
        self.cosmtime_hubble = y[0]
        loga = np.log10(a_knots)
        self.sp_grow1 = NaturalSpline(loga, np.log10(self.grow1))
        self.sp_grow2 = NaturalSpline(loga, np.log10(self.grow2))
        self.sp_grow31 = NaturalSpline(loga, np.log10(self.grow31))
        self.sp_grow32 = NaturalSpline(loga, np.log10(self.grow32))
        # SP_INVGROW: x = log10 D, y = log10 a  (src/cosmo.c:401)
        self.sp_invgrow = NaturalSpline(np.log10(self.grow1), loga)

    # GrowingMode* (src/cosmo.c:1776-1819); k is ignored when NkBINS == 1
    def GrowingMode(self, z):
        return float(10.0 ** self.sp_grow1(-math.log10(1.0 + z)))

    def GrowingMode_2LPT(self, z):
        return float(10.0 ** self.sp_grow2(-math.log10(1.0 + z)))

    def GrowingMode_3LPT_1(self, z):
        return float(-(10.0 ** self.sp_grow31(-math.log10(1.0 + z))))

    def GrowingMode_3LPT_2(self, z):
        return float(10.0 ** self.sp_grow32(-math.log10(1.0 + z)))

    def growth_for_order(self, order: int, z: float) -> float:
        """growth_rate switch of compute_derivative (src/fmax-pfft.c:344-364)."""
        return {0: lambda z: 1.0, 1: self.GrowingMode, 2: self.GrowingMode_2LPT,
                3: self.GrowingMode_3LPT_1, 4: self.GrowingMode_3LPT_2}[order](z)

    def InverseGrowingMode(self, D):
        """src/cosmo.c:1822-1832."""
        return 1.0 / 10.0 ** self.sp_invgrow(np.log10(D)) - 1.0

    # ---- power spectrum: Eisenstein & Hu, src/cosmo.c:1447-1498 --------------------
    def transf_EH(self, fk):
        p = self.p
        fk = np.asarray(fk, dtype=np.float64)
        Teta_27 = 1.0104
        OB = p.OmegaBaryon if p.OmegaBaryon > 1e-6 else 1e-6
        Omegac = p.Omega0 - OB
        Oh2 = p.Omega0 * p.Hubble100 ** 2
        Ob2 = OB * p.Hubble100 ** 2
        b1 = 0.313 * Oh2 ** -0.419 * (1 + 0.607 * Oh2 ** 0.674)
        b2 = 0.238 * Oh2 ** 0.223
        zd = 1291.0 * Oh2 ** 0.251 * (1.0 + b1 * Ob2 ** b2) / (1.0 + 0.659 * Oh2 ** 0.828)
        Rd = 31.5 * Ob2 / (Teta_27 ** 4 * 0.001 * zd)
        zeq = 2.5e4 * Oh2 / Teta_27 ** 4
        Req = 31.5 * Ob2 / (Teta_27 ** 4 * 0.001 * zeq)
        keq = 7.46e-2 * Oh2 / Teta_27 / Teta_27
        s = 1.633 * math.log((math.sqrt(1.0 + Rd) + math.sqrt(Rd + Req)) / (1 + math.sqrt(Req))) / (keq * math.sqrt(Req))
        ks = fk * s
        q = fk * Teta_27 * Teta_27 / Oh2
        alc = ((46.9 * Oh2) ** 0.670 * (1.0 + (32.1 * Oh2) ** -0.532)) ** (-OB / p.Omega0) * \
              ((12.0 * Oh2) ** 0.424 * (1.0 + (45.0 * Oh2) ** -0.582)) ** (-(OB / p.Omega0) ** 3)
        bec = 1.0 / (1.0 + (0.944 / (1.0 + (458.0 * Oh2) ** -0.708)) *
                     ((Omegac / p.Omega0) ** ((0.395 * Oh2) ** -0.0266) - 1.0))

        def T0(q, a, b):
            ll = np.log(math.exp(1.0) + 1.8 * b * q)
            C = 14.2 / a + 386.0 / (1.0 + 69.9 * q ** 1.08)
            return ll / (ll + C * q * q)

        f = 1.0 / (1 + (ks / 5.4) ** 4)
        Tc = f * T0(q, 1.0, bec) + (1.0 - f) * T0(q, alc, bec)
        beb = 0.5 + OB / p.Omega0 + (3.0 - 2.0 * OB / p.Omega0) * math.sqrt((17.2 * Oh2) ** 2 + 1.0)
        bno = 8.41 * Oh2 ** 0.435
        kst = ks / (1.0 + (bno / ks) ** 3) ** 0.3333
        ksi = 1.6 * Ob2 ** 0.52 * Oh2 ** 0.73 * (1.0 + (10.4 * Oh2) ** -0.95)
        y = (1.0 + zeq) / (1 + zd)
        alb = 2.07 * keq * s * (1.0 + Rd) ** -0.75 * (
            y * (-6.0 * math.sqrt(1.0 + y) + (2.0 + 3.0 * y) *
                 math.log((math.sqrt(1.0 + y) + 1.0) / (math.sqrt(1.0 + y) - 1.0))))
        Tb = (T0(q, 1.0, 1.0) / (1.0 + (ks / 5.2) ** 2) +
              alb / (1.0 + (beb / ks) ** 3) * np.exp(-(fk / ksi) ** 1.4)) * np.sin(kst) / kst
        return (OB * Tb + Omegac * Tc) / p.Omega0

    def PowerSpectrum(self, k):
        """P(k) in true Mpc^3 for k in true Mpc^-1 (src/cosmo.c:953-1007, WhichSpectrum=1)."""
        k = np.asarray(k, dtype=np.float64)
        return self.PkNorm * k ** self.p.PrimordialIndex * self.transf_EH(k) ** 2

    # ---- mass variance, src/cosmo.c:1507-1640 -----------------------------------
    @staticmethod
    def _window(kr, window):
        if window == 0:
            return math.exp(-kr * kr / 2.0)
        if window == 2:
            if kr < 1e-5:
                return 1.0
            kr2 = kr * kr
            return 3.0 * (math.sin(kr) / kr2 / kr - math.cos(kr) / kr2)
        raise ValueError(window)

    def compute_mass_variance(self, R, window=0):
        def integrand(logk):
            k = math.exp(logk)
            w = self._window(k * R, window)
            return float(self.PowerSpectrum(k)) * w * w * k ** 3 / (2.0 * math.pi ** 2)

        # gsl_integration_qags(-10, log(500/R), epsabs 0, epsrel TOLERANCE, limit NWINT)
        res, _ = quad(integrand, -10.0, math.log(500.0 / R), epsabs=0.0, epsrel=TOLERANCE, limit=NWINT)
        return res

    def _mass_variance_tables(self):
        rmin, dr = -6.0, 0.04
        rv = rmin + dr * np.arange(NBINS)
        massvar = np.zeros(NBINS)
        for i in range(NBINS - 1, -1, -1):
            massvar[i] = math.log10(self.compute_mass_variance(10.0 ** rv[i], window=0))
            if i < NBINS - 1 and massvar[i] - massvar[i + 1] < 1e-6:
                massvar[i] = massvar[i + 1] + 1e-6
        self.sp_massvar = NaturalSpline(rv, massvar)
        # SP_RADIUS has x = -log10(var) increasing with i (variance decreases with radius)
        self.sp_radius = NaturalSpline(-massvar, rv)

    def MassVariance(self, R):
        return float(10.0 ** self.sp_massvar(math.log10(R)))

    def Radius(self, var):
        return float(10.0 ** self.sp_radius(-math.log10(var)))


# --------------------------------------------------------------------------------------
@dataclass
class SmoothingLadder:
    """``smoothing_data`` (src/pinocchio.h:284-292) as filled by set_smoothing."""
    Radius: np.ndarray
    Variance: np.ndarray

    @property
    def Nsmooth(self):
        return int(self.Radius.size)


def set_smoothing(cosmo: Cosmology, inter_part_dist: float, zlast: float = 0.0) -> SmoothingLadder:
    """src/initialization.c:386-435.  ``inter_part_dist`` is in true Mpc."""
    var_min = (DELTA_C / NSIGMA / cosmo.GrowingMode(zlast)) ** 2
    rmin = inter_part_dist / 6.0
    var_max = cosmo.MassVariance(rmin)
    nsmooth = int((math.log10(var_max) - math.log10(var_min)) / STEP_VAR + 2)
    if nsmooth <= 0:
        nsmooth = 1
    radius = np.zeros(nsmooth)
    variance = np.zeros(nsmooth)
    for i in range(nsmooth - 1):
        variance[i] = 10.0 ** (math.log10(var_min) + STEP_VAR * i)
        radius[i] = cosmo.Radius(variance[i])
    radius[nsmooth - 1] = 0.0
    variance[nsmooth - 1] = var_max
    return SmoothingLadder(radius, variance)


def pk_lattice_table(cosmo: Cosmology, grid: int, box_true_mpc: float) -> np.ndarray:
    """P(k) on the integer lattice: entry m holds PowerSpectrum(2 pi sqrt(m) / Box), m = |n|^2.

    GenIC evaluates ``PowerSpectrum(kmag)`` per mode (src/GenIC.c:283) and discards modes with
    |n| > N/2 (:280); P therefore depends on the integer m <= (N/2)^2 only.  Entry 0 is 0.
    """
    mmax = (grid // 2) ** 2
    m = np.arange(mmax + 1, dtype=np.float64)
    k = 2.0 * math.pi * np.sqrt(m) / box_true_mpc
    t = np.zeros(mmax + 1)
    t[1:] = cosmo.PowerSpectrum(k[1:])
    return t
