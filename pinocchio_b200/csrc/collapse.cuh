// Per-cell ellipsoidal collapse: Hessian -> eigenvalues -> b_c -> F = 1 + z_c.
// Device restatement of inverse_collapse_time / ell / ell_classic / ord
// (reference src/collapse_times.c:679-776, 404-427, 114-221, 1354-1362) and
// InverseGrowingMode / my_spline_eval (src/cosmo.c:1822-1832, 2016-2027).
// Host/device portable so tests/host can check it against the oracle on a CPU-only box.
//
// The kernel is bound by FP64 instruction issue and latency, not by HBM (ncu r01: FP64 pipe 41 %,
// issue slots 54 %, DRAM 11 %), so the arithmetic is restructured without changing the algorithm:
//   * cos(t/3), cos((t+2pi)/3), cos((t+4pi)/3) come from ONE sin/cos pair of t/3 in [0, pi/3]
//     (no range reduction needed) and the angle-addition formulas;
//   * pow(x, 0.333333333333333) -> cbrt(x);   1/pow(10, y) -> exp10(-y);
//   * divisions by literal constants -> multiplications; a/b/c -> a/(b*c);
//   * the spline interval comes from a uniform look-up table instead of a 10-step search;
//   * divisions and square roots on the common path use the MUFU seeds + Newton steps of
//     fastmath.cuh without CUDA's range checks and called slow paths (fm_rcp, fm_div,
//     fm_sqrt_pair); 2r/(q 2 sqrt q) becomes r (1/sqrt q)^3 from the same Newton chain;
//   * the common path of ell_classic is branch free (both cubic cases are evaluated and
//     selected) so that two cells can be interleaved per thread; the rare guarded cases
//     (|l1| < SMALL, |den| < SMALL) fall back to the reference's nested branches.
// Each substitution changes the result by a few ulps (1e-15 relative); the contract is 1e-6.
// SMALL tests and NaN propagation are those of the reference.
#pragma once
#include <math.h>
#include "fastmath.cuh"
#include "fft_core.cuh"
#include "spline_pack.h"

namespace pinb {

#define PINB_PI 3.14159265358979323846 /* src/pinocchio.h:56 */
#define PINB_SMALL 1.e-20              /* src/collapse_times.c:38 */

// packed table of spline_pack.h (in shared memory on the device)
struct SplineView {
  const double* t;
  int n;
};

// my_spline_eval: linear (secant) extrapolation outside the knots, cspline inside.
PINB_HD double spline_eval(const SplineView& s, double xq) {
  const double* t = s.t;
  if (xq < t[0]) return t[5] + (xq - t[0]) * t[3];
  if (xq > t[1]) return t[6] + (xq - t[1]) * t[4];
  const int n = s.n;
  int j = (int)((xq - t[0]) * t[2]);
  j = j < 0 ? 0 : (j > PINB_SPLINE_NLUT - 1 ? PINB_SPLINE_NLUT - 1 : j);
  const unsigned short* lut = reinterpret_cast<const unsigned short*>(t + PINB_SPLINE_HDR + 5 * n);
  int i = lut[j];
  const double* k = t + PINB_SPLINE_HDR + 5 * i;
  // gsl_interp_bsearch semantics: largest i in [0, n-2] with x[i] <= xq
  while (i < n - 2 && k[5] <= xq) { i++; k += 5; }
  if (i > 0 && xq < k[0]) k -= 5;  // rounding of the bin index at a bin edge
  const double dx = xq - k[0];
  return k[1] + dx * (k[2] + dx * (k[3] + dx * k[4]));
}

// InverseGrowingMode(D) = 1/10^spline(log10 D) - 1  (src/cosmo.c:1822-1832)
PINB_HD double inverse_growing_mode(const SplineView& s, double D) {
  return fm_exp10(-spline_eval(s, fm_log10(D))) - 1.0;
}

// Taylor coefficients of sin and cos in a^2 (descending order), kept in the constant bank on the
// device: a literal double costs two UMOV issue slots per use (19 % of the r01 kernel).
#define PINB_SINCOS_TABLE                                                                      \
  {-8.2206352466243297e-18, 2.8114572543455208e-15, -7.6471637318198165e-13,                   \
   1.6059043836821613e-10, -2.5052108385441719e-08, 2.7557319223985893e-06,                    \
   -1.9841269841269841e-04, 8.3333333333333332e-03, -1.6666666666666666e-01,                   \
   4.1103176233121648e-19, -1.5619206968586225e-16, 4.7794773323873853e-14,                    \
   -1.1470745597729725e-11, 2.0876756987868099e-09, -2.7557319223985888e-07,                   \
   2.4801587301587302e-05, -1.3888888888888889e-03, 4.1666666666666664e-02, -0.5,              \
   0.33333333333333333, 0.86602540378443864676}
#if defined(__CUDACC__)
static __constant__ double kSinCosDev[21] = PINB_SINCOS_TABLE;
#endif
static const double kSinCosHost[21] = PINB_SINCOS_TABLE;
PINB_HD double sc_coef(int i) {
#if defined(__CUDA_ARCH__)
  return kSinCosDev[i];
#else
  return kSinCosHost[i];
#endif
}

// the three values cos(t/3), cos((t+2pi)/3), cos((t+4pi)/3) for t in [0, pi]:
// sin and cos of a = t/3 in [0, ~1.05] by Taylor series in a^2 (truncation < 1e-18), then
// cos(a + 2pi/3) = -c/2 - (sqrt3/2) s,  cos(a + 4pi/3) = -c/2 + (sqrt3/2) s
// (both polynomials by Estrin's scheme: four dependent FMA levels instead of nine / ten)
PINB_HD void cos_thirds(double t, double& c0, double& c1, double& c2) {
  const double a = t * sc_coef(19);
  const double z = a * a;
  const double z2 = z * z, z4 = z2 * z2, z8 = z4 * z4;
  const double q0 = fma_rn(sc_coef(7), z, sc_coef(8)), q1 = fma_rn(sc_coef(5), z, sc_coef(6));
  const double q2 = fma_rn(sc_coef(3), z, sc_coef(4)), q3 = fma_rn(sc_coef(1), z, sc_coef(2));
  const double ps = fma_rn(z8, sc_coef(0), fma_rn(z4, fma_rn(z2, q3, q2), fma_rn(z2, q1, q0)));
  const double s = a + a * (z * ps);
  const double u0 = fma_rn(sc_coef(17), z, sc_coef(18)), u1 = fma_rn(sc_coef(15), z, sc_coef(16));
  const double u2 = fma_rn(sc_coef(13), z, sc_coef(14)), u3 = fma_rn(sc_coef(11), z, sc_coef(12));
  const double u4 = fma_rn(sc_coef(9), z, sc_coef(10));
  const double pc = fma_rn(z8, u4, fma_rn(z4, fma_rn(z2, u3, u2), fma_rn(z2, u1, u0)));
  const double c = 1.0 + z * pc;
  const double hs = sc_coef(20) * s;
  c0 = c;
  c1 = -0.5 * c - hs;
  c2 = -0.5 * c + hs;
}

// ell_classic, src/collapse_times.c:114-221, with the reference's nested branches (used for the
// rare guarded cases and as the plain scalar version)
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

// Same function with a branch-free common path: both cases of the 3rd-order cubic are evaluated
// and selected with the reference's conditions, so that the compiler can interleave independent
// cells.  `guard` lanes (|l1| < SMALL or |den| < SMALL) take the nested version above.
PINB_HD double ell_classic_flat(double l1, double l2, double l3) {
  const double del = l1 + l2 + l3;
  const double det = l1 * l2 * l3;
  const double den = det * mc(MC_1_126) + mc(MC_5_84) * l1 * del * (del - l1);
  if (fabs(l1) < PINB_SMALL || fabs(den) < PINB_SMALL) return ell_classic(l1, l2, l3);
  const double rden = fm_rcp(den);  // |den| >= 1e-20 here
  const double a1 = 3. * l1 * (del - l1) * mc(MC_1_14) * rden;
  const double a1_2 = a1 * a1;
  const double a2 = l1 * rden;
  const double a3 = -1.0 * rden;
  const double q = (a1_2 - 3. * a2) * mc(MC_1_9);
  const double r = (2. * a1_2 * a1 - 9. * a1 * a2 + 27. * a3) * mc(MC_1_54);
  const double r_2_q_3 = r * r - q * q * q;
  const double a1_3 = a1 * mc(MC_1_3);
  // Both cases are evaluated on every lane; the lane that will NOT be selected is fed benign
  // operands (1, 0) instead of a negative radicand or an |argument| > 1: NaN operands send
  // sqrt and division into their slow-path subroutines (r02 profile: 230 instructions per
  // cell in __cuda_sm20_div_rn_f64_full / dsqrt_rn_f64_mediumpath before this guard).
  const bool c1 = r_2_q_3 > 0;
  // case 1 (r^2 - q^3 > 0)
  // cube root and its reciprocal from one Newton chain.  q / sq is formed with the remainder step of fm_div
  // (seeded by sq^-1), i.e. as the quotient by the ROUNDED sq itself: s + q/s is stationary in s where s^2 = q,
  // so an error of sq then cancels to second order, as in the reference's expression (independent roundings
  // of sq and sq^-1 would not: 1e-7 on the cells next to r^2 = q^3)
  const CbrtPair sqa = fm_cbrt_pair(fm_sqrt_pair(c1 ? r_2_q_3 : 1.0).s + fabs(r));
  const double sg = (r > 0.) ? -1.0 : ((r < 0.) ? 1.0 : NAN);
  const double qs0 = q * sqa.rc;
  const double qs = fma_rn(fma_rn(-sqa.c, qs0, q), sqa.rc, qs0);
  double ella = sg * (sqa.c + qs) - a1_3;
  ella = (ella < 0.) ? -.1 : ella;
  // case 2 (r^2 <= q^3, hence q >= 0; a NaN discriminant lands here as in the reference)
  // 2 r / (q * 2 sqrt q) = r * (1/sqrt q)^3: sqrt and its reciprocal come from one Newton chain
  const double q2 = c1 ? 1.0 : q;
  const SqrtPair sp = fm_sqrt_pair(q2);
  const double sqb = 2 * sp.s;
  const double t = fm_acos(c1 ? 0.5 : r * (sp.rs * sp.rs * sp.rs));
  double c0, c1c, c2;
  cos_thirds(t, c0, c1c, c2);
  double s1 = -sqb * c0 - a1_3;
  double s2 = -sqb * c1c - a1_3;
  double s3 = -sqb * c2 - a1_3;
  s1 = (s1 < 0.) ? 1.e10 : s1;
  s2 = (s2 < 0.) ? 1.e10 : s2;
  s3 = (s3 < 0.) ? 1.e10 : s3;
  double ellb = (s1 < s2 ? s1 : s2);
  ellb = (s3 < ellb ? s3 : ellb);
  ellb = (ellb == 1.e10) ? -.1 : ellb;
  double ell = c1 ? ella : ellb;
  // the correction is used for del > 0 only, where its exponent is <= 0 (l1 >= l2 >= l3); the other lanes get a
  // benign 0 instead of a positive argument that would send the warp through libm's exp (r02 ncu: 3 % of the
  // kernel's instructions, del < 0 in half of the cells)
  const bool use_corr = del > 0.;
  const double inv_del = fm_rcp(use_corr ? del : 1.0);
  const double xe = (mc(MC_M65) * (l1 - l2) + mc(MC_M28) * (l2 - l3)) * inv_del;
  const double corr = mc(MC_M0364) * inv_del * fm_exp_neg(use_corr ? xe : 0.0);
  if (use_corr && ell > 0.) ell += corr;
  return ell;
}

// inverse_collapse_time with ELL_CLASSIC: returns F.  d = {xx, yy, zz, xy, xz, yz}
PINB_HD double inverse_collapse_time(const double* d, const SplineView& sp) {
  // mu1, mu2 and q rounded operation by operation (mul_rn / add_rn): `q == 0.` must hold for an
  // isotropic tensor exactly as in the reference's unfused arithmetic (:694-724)
  const double mu1 = add_rn(add_rn(d[0], d[1]), d[2]);
  const double mu1_2 = mul_rn(mu1, mu1);
  double mu2 = 0.5 * mu1_2;
  mu2 = add_rn(mu2, -0.5 * add_rn(add_rn(mul_rn(d[0], d[0]), mul_rn(d[1], d[1])), mul_rn(d[2], d[2])));
  const double add0 = mul_rn(d[3], d[3]), add1 = mul_rn(d[4], d[4]), add2 = mul_rn(d[5], d[5]);
  mu2 = add_rn(mu2, -add_rn(add_rn(add0, add1), add2));
  const double mu3 = d[0] * d[1] * d[2] + 2. * d[3] * d[4] * d[5] - d[0] * add2 - d[1] * add1 - d[2] * add0;
  const double q = add_rn(mu1_2, -mul_rn(3.0, mu2)) * mc(MC_1_9);
  const double r = -(2. * mu1_2 * mu1 - 9.0 * mu1 * mu2 + 27.0 * mu3) * mc(MC_1_54);
  const bool diag = (q == 0.);                                  // already diagonal (:724-728)
  const bool bad = !diag && (q * q * q < r * r || q < 0.0);     // :734-736
  // benign operands on the lanes whose trigonometric solution is not used (see ell_classic_flat)
  const bool unused = diag || bad;
  const double qs = unused ? 1.0 : q;
  const SqrtPair rt = fm_sqrt_pair(qs);
  const double sq = 2 * rt.s;
  const double t = fm_acos(unused ? 0.5 : r * (rt.rs * rt.rs * rt.rs));  // 2 r / (q * 2 sqrt q)
  const double m3 = mu1 * mc(MC_1_3);
  double c0, c1, c2;
  cos_thirds(t, c0, c1, c2);
  const double x1 = diag ? d[0] : -sq * c0 + m3;
  const double x2 = diag ? d[1] : -sq * c1 + m3;
  const double x3 = diag ? d[2] : -sq * c2 + m3;
  // ord(): C macros, NaN behaviour of `a>b?a:b`
  const double m12 = (x1 > x2 ? x1 : x2), n12 = (x1 < x2 ? x1 : x2);
  const double hi = (m12 > x3 ? m12 : x3);
  const double lo = (n12 < x3 ? n12 : x3);
  const double mid = x1 + x2 + x3 - lo - hi;
  const double bc = ell_classic_flat(hi, mid, lo);
  // 92 % of the cells have b_c > 0: evaluated on every lane (benign operand 1 elsewhere) so that
  // the code stays convergent
  const bool pos = bc > 0.0;
  const double igm = inverse_growing_mode(sp, pos ? bc : 1.0);
  const double F = pos ? 1. + igm : 0.0;
  return bad ? -10.0 : F;
}

}  // namespace pinb
