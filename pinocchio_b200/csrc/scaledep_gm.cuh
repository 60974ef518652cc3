// Batched variance integrals of set_scaledep_GM (reference src/initialization.c:1533-2026).
//
// The reference computes, for every smoothing radius R_r, every one of the NBINS time knots t_i and three quantities,
//     I = Int dlogk  P(k) D(t_i, k)^2 [f_Omega(t_i, k)^2] W(k R_r)^2 k^p / (2 pi^2)          (logk = log10 k)
// (density: Gaussian window, p = 3, src/initialization.c:1439-1447; displacement and velocity: top-hat window, p = 1,
// :1449-1457, :1489-1498) with one adaptive gsl_integration_qags call each -- 3 x Nsmooth x 210 serial integrals whose
// integrand walks the splines of the host cosmology.  All of them share the interval and the factor P(k) k^p, the
// window depends on (r, k) only and the growth factors on (t, k) only.  Here the integral is one fixed high-order
// quadrature for all of them: the caller supplies the nodes in log10 k and a_q[j] = weight_j P(k_j) k_j^p / (2 pi^2)
// (P from the UNCHANGED host cosmology, whichever spectrum it uses), the device builds T[q][r][j] = a_q[j] W_q(k_j R_r)^2
// once (sdgm_window_body) and contracts it with D^2 (t_i, k_j) [f^2], evaluated per (i, j) from the k-bin tables of
// InterpolateGrowth (src/cosmo.c:1728-1755) -- one block per time knot, fixed summation order (sdgm_integrate_body).
// Composite Gauss-Legendre with 8 points on 512 panels agrees with QUADPACK's converged value to 1e-12; the
// reference's own answer is only defined to its TOLERANCE of 1e-4 (tests/test_scaledep_gm.py).
#pragma once
#include "fft_core.cuh"

namespace pinb {

struct SdgmParams {
  const double* logk;    // [n] quadrature nodes, log10 k
  const double* a_dens;  // [n] weight * P(k) * k^3 / (2 pi^2)
  const double* a_disp;  // [n] weight * P(k) * k   / (2 pi^2)
  int n;
  const double* lg;      // [nk][nt] log10 GrowingMode at the k bins and time knots (SP_GROW1 + kk)
  const double* fo;      // [nk][nt] fomega there (SP_FOMEGA1 + kk)
  int nk, nt;
  double logkmin, dlogk; // LOGKMIN, DELTALOGK (src/def_splines.h:41-42)
  const double* r_dens;  // [ns] Smoothing.Radius (Gaussian window, WindowFunctionType 0)
  const double* r_disp;  // [ns] Smoothing.Rad_GM (top-hat window, WindowFunctionType 2)
  int ns;
  double* T;             // [2][ns][n] scratch: a_q W_q^2
  double* out;           // [3][ns][nt]: sqrt of the density, displacement, velocity integrals
};

// WindowFunction of src/cosmo.c:1611-1646, squared
PINB_HD double sdgm_window2(int type, double kr) {
  double w;
  if (type == 0) {
    w = exp(-kr * kr / 2.);
  } else if (kr < 1.e-5) {
    w = 1.0;
  } else {
    const double kr2 = kr * kr;
    w = 3. * (sin(kr) / kr2 / kr - cos(kr) / kr2);
  }
  return w * w;
}

// one thread per (q, r, j)
template <class Ctx> PINB_HD void sdgm_window_body(Ctx& ctx, const SdgmParams& p) {
  const long long idx = (long long)ctx.bid() * ctx.nthreads() + ctx.tid();
  const long long per = (long long)p.ns * p.n;
  if (idx >= 2 * per) return;
  const int q = (int)(idx / per), r = (int)((idx % per) / p.n), j = (int)(idx % p.n);
  const double k = pow(10., p.logk[j]);
  const double R = q == 0 ? p.r_dens[r] : p.r_disp[r];
  p.T[idx] = (q == 0 ? p.a_dens[j] : p.a_disp[j]) * sdgm_window2(q == 0 ? 0 : 2, k * R);
}

// InterpolateGrowth (src/cosmo.c:1728-1755) on a [nk][nt] table at time knot i
PINB_HD double sdgm_interp(const SdgmParams& p, const double* tab, int i, double logk) {
  const double logkmax = p.logkmin + (p.nk - 1) * p.dlogk;
  if (logk < p.logkmin || p.nk == 1) return tab[i];
  if (logk > logkmax) return tab[(size_t)(p.nk - 1) * p.nt + i];
  double dk = (logk - p.logkmin) / p.dlogk;
  int kk = (int)dk;
  if (kk > p.nk - 2) kk = p.nk - 2;  // logk == logkmax: the reference multiplies the bin past the end by 0
  dk -= kk;
  return dk * tab[(size_t)(kk + 1) * p.nt + i] + (1 - dk) * tab[(size_t)kk * p.nt + i];
}

// block = time knot i; NT threads stride over the nodes; RC radii at a time keep 3 RC accumulators in registers.
// scratch: NT doubles.
template <int NT, int RC, class Ctx> PINB_HD void sdgm_integrate_body(Ctx& ctx, double* scratch, const SdgmParams& p) {
  const int i = ctx.bid(), tid = ctx.tid();
  const size_t per = (size_t)p.ns * p.n;
  for (int r0 = 0; r0 < p.ns; r0 += RC) {
    double acc[3][RC];
#pragma unroll
    for (int c = 0; c < RC; c++) acc[0][c] = acc[1][c] = acc[2][c] = 0.0;
    for (int j = tid; j < p.n; j += NT) {
      const double lk = p.logk[j];
      const double D = pow(10., sdgm_interp(p, p.lg, i, lk));  // GrowingMode, src/cosmo.c:1789-1795
      const double f = sdgm_interp(p, p.fo, i, lk);            // fomega, :1757-1763
      const double D2 = D * D, V2 = D2 * f * f;
#pragma unroll
      for (int c = 0; c < RC; c++) {
        if (r0 + c < p.ns) {
          const double td = p.T[(size_t)(r0 + c) * p.n + j], tt = p.T[per + (size_t)(r0 + c) * p.n + j];
          acc[0][c] += td * D2;
          acc[1][c] += tt * D2;
          acc[2][c] += tt * V2;
        }
      }
    }
    // fixed-order tree over the block, one accumulator at a time
#pragma unroll
    for (int q = 0; q < 3; q++)
#pragma unroll
      for (int c = 0; c < RC; c++) {
        if (r0 + c >= p.ns) continue;  // uniform over the block
        scratch[tid] = acc[q][c];
        ctx.sync();
        for (int h = NT / 2; h > 0; h >>= 1) {
          if (tid < h) scratch[tid] += scratch[tid + h];
          ctx.sync();
        }
        if (tid == 0) p.out[((size_t)q * p.ns + r0 + c) * p.nt + i] = sqrt(scratch[0]);
        ctx.sync();
      }
  }
}

}  // namespace pinb
