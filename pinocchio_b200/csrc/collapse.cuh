// Per-cell ellipsoidal collapse: Hessian -> eigenvalues -> b_c -> F = 1 + z_c.
// Device restatement of inverse_collapse_time / ell / ell_classic / ord
// (reference src/collapse_times.c:679-776, 404-427, 114-221, 1354-1362) and
// InverseGrowingMode / my_spline_eval (src/cosmo.c:1822-1832, 2016-2027).
// Host/device portable so tests/host can check it against the oracle on a CPU-only box.
#pragma once
#include <math.h>
#include "fft_core.cuh"

namespace pinb {

#define PINB_PI 3.14159265358979323846 /* src/pinocchio.h:56 */
#define PINB_SMALL 1.e-20              /* src/collapse_times.c:38 */

// Natural cubic spline table [x | y | b | c | d], n knots each (b, d valid for n-1 intervals);
// coefficients computed on the host exactly as gsl_interp_cspline does (engine.cu).
struct SplineView {
  const double* x;
  const double* y;
  const double* b;
  const double* c;
  const double* d;
  int n;
};

// my_spline_eval: linear (secant) extrapolation outside the knots, cspline inside.
PINB_HD double spline_eval(const SplineView& s, double xq) {
  const int n = s.n;
  if (xq < s.x[0]) return s.y[0] + (xq - s.x[0]) * (s.y[1] - s.y[0]) / (s.x[1] - s.x[0]);
  if (xq > s.x[n - 1])
    return s.y[n - 1] + (xq - s.x[n - 1]) * (s.y[n - 1] - s.y[n - 2]) / (s.x[n - 1] - s.x[n - 2]);
  // gsl_interp_bsearch: largest i in [0, n-2] with x[i] <= xq
  int lo = 0, hi = n - 1;
  while (hi > lo + 1) {
    const int mid = (hi + lo) >> 1;
    if (s.x[mid] > xq) hi = mid; else lo = mid;
  }
  const double dx = xq - s.x[lo];
  return s.y[lo] + dx * (s.b[lo] + dx * (s.c[lo] + dx * s.d[lo]));
}

// InverseGrowingMode(D): 1/10^spline(log10 D) - 1
PINB_HD double inverse_growing_mode(const SplineView& s, double D) {
  return 1.0 / pow(10.0, spline_eval(s, log10(D))) - 1.0;
}

// ell_classic, src/collapse_times.c:114-221 (branch structure kept verbatim so that NaNs and
// the SMALL tests behave as in the reference)
PINB_HD double ell_classic(double l1, double l2, double l3) {
  double ell;
  const double del = l1 + l2 + l3;
  const double det = l1 * l2 * l3;
  if (fabs(l1) < PINB_SMALL) {
    ell = -0.1;
  } else {
    const double den = det / 126. + 5. * l1 * del * (del - l1) / 84.;
    if (fabs(den) < PINB_SMALL) {
      if (fabs(del - l1) < PINB_SMALL) {
        ell = (l1 > 0.0) ? 1. / l1 : -.1;
      } else {
        const double dis = 7. * l1 * (l1 + 6. * del);
        if (dis < 0.0) {
          ell = -.1;
        } else {
          ell = (7. * l1 - sqrt(dis)) / (3. * l1 * (l1 - del));
          if (ell < 0.) ell = -.1;
        }
      }
    } else {
      const double rden = 1.0 / den;
      const double a1 = 3. * l1 * (del - l1) / 14. * rden;
      const double a1_2 = a1 * a1;
      const double a2 = l1 * rden;
      const double a3 = -1.0 * rden;
      const double q = (a1_2 - 3. * a2) / 9.;
      const double r = (2. * a1_2 * a1 - 9. * a1 * a2 + 27. * a3) / 54.;
      const double r_2_q_3 = r * r - q * q * q;
      if (r_2_q_3 > 0) {
        const double fabs_r = fabs(r);
        const double sq = pow(sqrt(r_2_q_3) + fabs_r, 0.333333333333333);
        ell = -fabs_r / r * (sq + q / sq) - a1 / 3.;
        if (ell < 0.) ell = -.1;
      } else {
        const double sq = 2 * sqrt(q);
        const double inv_3 = 1.0 / 3;
        const double t = acos(2 * r / q / sq);
        double s1 = -sq * cos(t * inv_3) - a1 * inv_3;
        double s2 = -sq * cos((t + 2. * PINB_PI) * inv_3) - a1 * inv_3;
        double s3 = -sq * cos((t + 4. * PINB_PI) * inv_3) - a1 * inv_3;
        if (s1 < 0.) s1 = 1.e10;
        if (s2 < 0.) s2 = 1.e10;
        if (s3 < 0.) s3 = 1.e10;
        ell = (s1 < s2 ? s1 : s2);
        ell = (s3 < ell ? s3 : ell);
        if (ell == 1.e10) ell = -.1;
      }
    }
  }
  if (del > 0. && ell > 0.) {
    const double inv_del = 1.0 / del;
    ell += -.364 * inv_del * exp(-6.5 * (l1 - l2) * inv_del - 2.8 * (l2 - l3) * inv_del);
  }
  return ell;
}

// inverse_collapse_time with ELL_CLASSIC: returns F; also delta = trace (for the variance).
// d = {xx, yy, zz, xy, xz, yz}
PINB_HD double inverse_collapse_time(const double* d, const SplineView& sp) {
  const double mu1 = d[0] + d[1] + d[2];
  const double mu1_2 = mu1 * mu1;
  double mu2 = 0.5 * mu1_2;
  mu2 -= 0.5 * (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  const double add0 = d[3] * d[3], add1 = d[4] * d[4], add2 = d[5] * d[5];
  mu2 -= add0 + add1 + add2;
  const double mu3 = d[0] * d[1] * d[2] + 2. * d[3] * d[4] * d[5] - d[0] * add2 - d[1] * add1 - d[2] * add0;
  const double q = (mu1_2 - 3.0 * mu2) / 9.0;
  double x1, x2, x3;
  if (q == 0.) {
    x1 = d[0]; x2 = d[1]; x3 = d[2];
  } else {
    const double r = -(2. * mu1_2 * mu1 - 9.0 * mu1 * mu2 + 27.0 * mu3) / 54.;
    if (q * q * q < r * r || q < 0.0) return -10.0;
    const double sq = 2 * sqrt(q);
    const double t = acos(2 * r / q / sq);
    const double inv_3 = 1.0 / 3.0;
    x1 = -sq * cos(t * inv_3) + mu1 * inv_3;
    x2 = -sq * cos((t + 2. * PINB_PI) * inv_3) + mu1 * inv_3;
    x3 = -sq * cos((t + 4. * PINB_PI) * inv_3) + mu1 * inv_3;
  }
  // ord(): C macros, NaN behaviour of `a>b?a:b`
  const double m12 = (x1 > x2 ? x1 : x2), n12 = (x1 < x2 ? x1 : x2);
  const double hi = (m12 > x3 ? m12 : x3);
  const double lo = (n12 < x3 ? n12 : x3);
  const double mid = x1 + x2 + x3 - lo - hi;
  const double bc = ell_classic(hi, mid, lo);
  if (bc > 0.0) return 1. + inverse_growing_mode(sp, bc);
  return 0.0;
}

}  // namespace pinb
