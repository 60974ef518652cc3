// Per-cell ellipsoidal collapse: Hessian -> eigenvalues -> b_c -> F = 1 + z_c.
// Device restatement of inverse_collapse_time / ell / ell_classic / ord
// (reference src/collapse_times.c:679-776, 404-427, 114-221, 1354-1362) and
// InverseGrowingMode / my_spline_eval (src/cosmo.c:1822-1832, 2016-2027).
// Host/device portable so tests/host can check it against the oracle on a CPU-only box.
//
// The kernel is bound by FP64 instruction issue, not by HBM (SURVEY.md section 7 "hard parts"),
// so the arithmetic is restructured without changing the algorithm:
//   * cos(t/3), cos((t+2pi)/3), cos((t+4pi)/3) come from ONE sin/cos pair of t/3 in [0, pi/3]
//     (no range reduction needed) and the angle-addition formulas;
//   * pow(x, 0.333333333333333) -> cbrt(x);   1/pow(10, y) -> exp10(-y);
//   * divisions by literal constants -> multiplications; a/b/c -> a/(b*c);
//   * branch-free 8-step lower-bound search of the spline interval.
// Each substitution changes the result by a few ulps (1e-15 relative); the contract is 1e-6.
// Branch structure, SMALL tests and NaN propagation are those of the reference (App. A.6).
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
  // gsl_interp_bsearch: largest i in [0, n-2] with x[i] <= xq  (branch-free lower bound)
  int lo = 0;
#pragma unroll
  for (int step = 512; step >= 1; step >>= 1) {
    const int m = lo + step;
    if (m <= n - 2 && s.x[m] <= xq) lo = m;
  }
  const double dx = xq - s.x[lo];
  return s.y[lo] + dx * (s.b[lo] + dx * (s.c[lo] + dx * s.d[lo]));
}

// InverseGrowingMode(D) = 1/10^spline(log10 D) - 1  (src/cosmo.c:1822-1832)
PINB_HD double inverse_growing_mode(const SplineView& s, double D) {
  return exp10(-spline_eval(s, log10(D))) - 1.0;
}

// sin and cos of a in [0, ~1.1] (a = t/3 with t = acos(.) in [0, pi]): Taylor series in a^2,
// truncation < 1e-18 on the interval, no range reduction.
PINB_HD void sincos_third(double a, double& s, double& c) {
  const double z = a * a;
  double ps = -8.2206352466243297e-18;              // -1/19!
  ps = ps * z + 2.8114572543455208e-15;             //  1/17!
  ps = ps * z - 7.6471637318198165e-13;             // -1/15!
  ps = ps * z + 1.6059043836821613e-10;             //  1/13!
  ps = ps * z - 2.5052108385441719e-08;             // -1/11!
  ps = ps * z + 2.7557319223985893e-06;             //  1/9!
  ps = ps * z - 1.9841269841269841e-04;             // -1/7!
  ps = ps * z + 8.3333333333333332e-03;             //  1/5!
  ps = ps * z - 1.6666666666666666e-01;             // -1/3!
  s = a + a * (z * ps);
  double pc = 4.1103176233121648e-19;               //  1/20!
  pc = pc * z - 1.5619206968586225e-16;             // -1/18!
  pc = pc * z + 4.7794773323873853e-14;             //  1/16!
  pc = pc * z - 1.1470745597729725e-11;             // -1/14!
  pc = pc * z + 2.0876756987868099e-09;             //  1/12!
  pc = pc * z - 2.7557319223985888e-07;             // -1/10!
  pc = pc * z + 2.4801587301587302e-05;             //  1/8!
  pc = pc * z - 1.3888888888888889e-03;             // -1/6!
  pc = pc * z + 4.1666666666666664e-02;             //  1/4!
  pc = pc * z - 0.5;
  c = 1.0 + z * pc;
}

// the three values cos(t/3), cos((t+2pi)/3), cos((t+4pi)/3) for t in [0, pi]
PINB_HD void cos_thirds(double t, double& c0, double& c1, double& c2) {
  double s, c;
  sincos_third(t * (1.0 / 3.0), s, c);
  const double h = 0.86602540378443864676;  // sin(2pi/3)
  c0 = c;
  c1 = -0.5 * c - h * s;
  c2 = -0.5 * c + h * s;
}

// ell_classic, src/collapse_times.c:114-221 (branch structure kept so that NaNs and the SMALL
// tests behave as in the reference)
PINB_HD double ell_classic(double l1, double l2, double l3) {
  double ell;
  const double del = l1 + l2 + l3;
  const double det = l1 * l2 * l3;
  if (fabs(l1) < PINB_SMALL) {
    ell = -0.1;
  } else {
    const double den = det * (1. / 126.) + 5. * l1 * del * (del - l1) * (1. / 84.);
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
      const double a1 = 3. * l1 * (del - l1) * (1. / 14.) * rden;
      const double a1_2 = a1 * a1;
      const double a2 = l1 * rden;
      const double a3 = -1.0 * rden;
      const double q = (a1_2 - 3. * a2) * (1. / 9.);
      const double r = (2. * a1_2 * a1 - 9. * a1 * a2 + 27. * a3) * (1. / 54.);
      const double r_2_q_3 = r * r - q * q * q;
      const double a1_3 = a1 * (1. / 3.);
      if (r_2_q_3 > 0) {
        const double fabs_r = fabs(r);
        const double sq = cbrt(sqrt(r_2_q_3) + fabs_r);  // pow(., 0.333333333333333)
        const double sg = (r > 0.) ? -1.0 : ((r < 0.) ? 1.0 : NAN);  // -fabs(r)/r
        ell = sg * (sq + q / sq) - a1_3;
        if (ell < 0.) ell = -.1;
      } else {
        const double sq = 2 * sqrt(q);
        const double t = acos(2 * r / (q * sq));
        double c0, c1, c2;
        cos_thirds(t, c0, c1, c2);
        double s1 = -sq * c0 - a1_3;
        double s2 = -sq * c1 - a1_3;
        double s3 = -sq * c2 - a1_3;
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
    ell += -.364 * inv_del * exp((-6.5 * (l1 - l2) - 2.8 * (l2 - l3)) * inv_del);
  }
  return ell;
}

// inverse_collapse_time with ELL_CLASSIC: returns F.  d = {xx, yy, zz, xy, xz, yz}
PINB_HD double inverse_collapse_time(const double* d, const SplineView& sp) {
  const double mu1 = d[0] + d[1] + d[2];
  const double mu1_2 = mu1 * mu1;
  double mu2 = 0.5 * mu1_2;
  mu2 -= 0.5 * (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  const double add0 = d[3] * d[3], add1 = d[4] * d[4], add2 = d[5] * d[5];
  mu2 -= add0 + add1 + add2;
  const double mu3 = d[0] * d[1] * d[2] + 2. * d[3] * d[4] * d[5] - d[0] * add2 - d[1] * add1 - d[2] * add0;
  const double q = (mu1_2 - 3.0 * mu2) * (1. / 9.);
  double x1, x2, x3;
  if (q == 0.) {
    x1 = d[0]; x2 = d[1]; x3 = d[2];
  } else {
    const double r = -(2. * mu1_2 * mu1 - 9.0 * mu1 * mu2 + 27.0 * mu3) * (1. / 54.);
    if (q * q * q < r * r || q < 0.0) return -10.0;
    const double sq = 2 * sqrt(q);
    const double t = acos(2 * r / (q * sq));
    const double m3 = mu1 * (1. / 3.);
    double c0, c1, c2;
    cos_thirds(t, c0, c1, c2);
    x1 = -sq * c0 + m3;
    x2 = -sq * c1 + m3;
    x3 = -sq * c2 + m3;
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
