"""set_scaledep_GM (reference src/initialization.c:1533-2026) -- CPU side.

1. The oracle's fixed quadrature against QUADPACK's adaptive qags (scipy.integrate.quad IS QUADPACK's qagse, the
   routine gsl_integration_qags ports; GSL itself is absent from this image): converged value to 1e-10, and the
   reference's own request (epsrel = TOLERANCE = 1e-4, limit = NWINT = 1000) met with a wide margin.
2. The kernel bodies of scaledep_gm.cuh under the block emulator, through the C ABI, against the oracle.
3. The reference's OWN set_scaledep_GM against shim/scaledep_gm_b200.c in one process (oracle/sdgm_harness.c, every
   reference translation unit linked) on example/parameter_file as shipped (CAMB tables, massive neutrinos: real
   scale dependence): inverse-growth vectors and every k_GM wavenumber.
The -m gpu twin is tests/test_zgpu_10_scaledep_gm.py.
"""
import ctypes
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

from example_util import EXAMPLE, ROOT, run_example
from oracle import pinocchio_oracle as po
from pinocchio_b200.engine import SdgmDesc, gauss_legendre_nodes

REF = ROOT / "oracle" / "_ref"
PD = ctypes.POINTER(ctypes.c_double)


def synthetic_case(nk=10, nt=24, ns=5, seed=3):
    """a BBKS-shaped spectrum, growth tables with a pronounced k dependence, radii from 0 to 60 Mpc"""
    rng = np.random.default_rng(seed)
    t = np.linspace(-1.4, 0.02, nt)                                       # log10 a
    tilt = 0.08 * np.tanh(np.arange(nk) - 4.0)[:, None]
    lg = (1.0 - tilt) * t[None, :] + 0.01 * rng.standard_normal((nk, 1))
    fo = 0.5 + 0.4 * (1.0 - 10.0 ** t)[None, :] + 0.05 * np.tanh(np.arange(nk) - 5.0)[:, None]
    rd = np.array([20.6, 9.0, 3.1, 0.69, 0.0])[:ns]
    rp = np.linspace(60.0, 0.0, ns)
    return lg, fo, rd, rp


def power(k):
    q = k / 0.2
    T = np.log(1 + 2.34 * q) / (2.34 * q) * (1 + 3.89 * q + (16.1 * q) ** 2 + (5.46 * q) ** 3 + (6.71 * q) ** 4) ** -0.25
    return 2.0e4 * k ** 0.96 * T * T


def quadrature(lo=-4.0, hi=3.14, npanels=512):
    logk, w = gauss_legendre_nodes(lo, hi, npanels, breaks=-3.0 + 0.5 * np.arange(10))
    k = 10.0 ** logk
    return logk, w * power(k) * k ** 3 / (2 * np.pi ** 2), w * power(k) * k / (2 * np.pi ** 2)


def test_fixed_quadrature_against_quadpack():
    from scipy import integrate
    lg, fo, rd, rp = synthetic_case()
    logkmin, dlogk, lo, hi = -3.0, 0.5, -4.0, 3.14
    logk, ad, ap = quadrature(lo, hi)
    out = po.scaledep_variances(logk, ad, ap, lg, fo, logkmin, dlogk, rd, rp)
    worst_conv = worst_req = 0.0
    for q, r, i in [(0, 0, 3), (0, 4, 23), (1, 0, 0), (1, 2, 11), (1, 4, 20), (2, 1, 7), (2, 3, 23)]:
        def f(x, q=q, r=r, i=i):
            k = 10.0 ** x
            D = 10.0 ** po.interpolate_growth_table(lg, np.array([x]), logkmin, dlogk)[i, 0]
            w = po.window_function(0 if q == 0 else 2, np.array([k * (rd[r] if q == 0 else rp[r])]))[0]
            v = power(k) * D * D * w * w * (k ** 3 if q == 0 else k) / (2 * np.pi ** 2)
            return v * po.interpolate_growth_table(fo, np.array([x]), logkmin, dlogk)[i, 0] ** 2 if q == 2 else v
        # break points at the k bins: the integrand has kinks there
        conv, _ = integrate.quad(f, lo, hi, epsabs=0, epsrel=1e-12, limit=4000, points=list(logkmin + dlogk * np.arange(10)))
        req, _ = integrate.quad(f, lo, hi, epsabs=0, epsrel=1e-4, limit=1000)      # the reference's call
        worst_conv = max(worst_conv, abs(out[q, r, i] ** 2 - conv) / conv)
        worst_req = max(worst_req, abs(out[q, r, i] ** 2 - req) / req)
    print(f"fixed quadrature vs QUADPACK: converged {worst_conv:.2e}, the reference's epsrel=1e-4 call {worst_req:.2e}")
    assert worst_conv < 1e-10
    assert worst_req < 1e-4           # TOLERANCE, src/initialization.c:1432


def test_quadrature_converged_in_panels():
    lg, fo, rd, rp = synthetic_case()
    a = po.scaledep_variances(*quadrature(npanels=512), lg, fo, -3.0, 0.5, rd, rp)
    b = po.scaledep_variances(*quadrature(npanels=4096), lg, fo, -3.0, 0.5, rd, rp)
    assert np.abs(a / b - 1).max() < 1e-11


def emu_abi():
    lib = REF / "libpinb_emuabi.so"
    if not lib.exists():
        pytest.skip("oracle/_ref/libpinb_emuabi.so not built")
    L = ctypes.CDLL(str(lib))
    L.pinb200_scaledep_variances.argtypes = [ctypes.POINTER(SdgmDesc), PD]
    L.pinb200_last_error.restype = ctypes.c_char_p
    L.pinb200_last_error.argtypes = [ctypes.c_void_p]
    return L


def call_abi(L, logk, ad, ap, lg, fo, logkmin, dlogk, rd, rp):
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (logk, ad, ap, lg, fo, rd, rp)]
    p = [a.ctypes.data_as(PD) for a in arrs]
    d = SdgmDesc(0, arrs[0].size, p[0], p[1], p[2], lg.shape[0], lg.shape[1], logkmin, dlogk, p[3], p[4], arrs[5].size, p[5], p[6])
    out = np.zeros((3, arrs[5].size, lg.shape[1]))
    rc = L.pinb200_scaledep_variances(ctypes.byref(d), out.ctypes.data_as(PD))
    return rc, out


@pytest.mark.parametrize("nk,ns", [(10, 5), (1, 3), (10, 11)])
def test_kernel_bodies_against_oracle(nk, ns):
    """scaledep_gm.cuh block by block (nk = 1: no -DSCALE_DEPENDENT; ns = 11: two radius chunks of the kernel)"""
    L = emu_abi()
    lg, fo, rd, rp = synthetic_case(nk=nk, ns=5)
    if ns != 5:
        rd, rp = np.linspace(25.0, 0.0, ns), np.linspace(60.0, 0.0, ns)
    logk, ad, ap = quadrature(npanels=64)              # the emulator is slow; the arithmetic is the same
    rc, out = call_abi(L, logk, ad, ap, lg, fo, -3.0, 0.5, rd, rp)
    assert rc == 0
    ref = po.scaledep_variances(logk, ad, ap, lg, fo, -3.0, 0.5, rd, rp)
    assert np.abs(out / ref - 1).max() < 1e-12


def test_abi_rejects_bad_arguments():
    L = emu_abi()
    lg, fo, rd, rp = synthetic_case()
    logk, ad, ap = quadrature(npanels=4)
    rc, _ = call_abi(L, logk, ad, ap, lg, fo, -3.0, 0.0, rd, rp)       # dlogk must be positive
    assert rc != 0 and b"dlogk" in L.pinb200_last_error(None)
    rc, _ = call_abi(L, logk, ad, ap, lg, fo, -3.0, 0.5, np.zeros(65), np.zeros(65))
    assert rc != 0


def test_reference_set_scaledep_gm_against_binding(tmp_path):
    """the reference's function and shim/scaledep_gm_b200.c (over the emulated ABI) in one process, shipped example"""
    exe = REF / "sdgm_emu_ex.x"
    if not exe.exists():
        pytest.skip("oracle/_ref/sdgm_emu_ex.x not built")
    out = run_example(exe, tmp_path, grid=32, threads=4, timeout=600)
    d = json.loads(out.strip().splitlines()[-1])
    print({k: v for k, v in d.items() if k != "k_gm"})
    assert d["nkbins"] == 10 and d["nbins"] == 210 and d["nsmooth"] >= 5
    # oracle/ref_full/mini_gsl.c's qags is accurate to ~1e-7 (its header); GSL's own to its 1e-4 request
    assert d["invgrow_vector_max_rel"] < 1e-6
    assert d["rad_gm_max_abs"] == 0.0
    ks = np.array([[a, b] for _, _, a, b in d["k_gm"]])
    assert len(ks) == 3 * d["nsmooth"]
    assert len(np.unique(ks[:, 0])) >= 4                        # a real bisection happened, not one clamped value
    assert np.array_equal(ks[:, 0], ks[:, 1])                   # every wavenumber bit-identical
