// Tabulated collapse times (TABULATED_CT) and the numerical ellipsoidal collapse of Nadkarni-Ghosh &
// Singhal (ELL_SNG): SURVEY.md section 8 row a19.
// Device restatement of initialize_collapse_times / interpolate_collapse_time (BILINEAR_SPLINE, the
// mode the reference compiles: src/collapse_times.c:39-41, 780-1046, 1132-1222), of sng_system /
// ell_sng (:239-290, 315-400) and of the GSL pieces they run on: gsl_interp_cspline (natural spline,
// symmetric tridiagonal solve) and gsl_odeiv2 rkf45 with the standard step control (GSL 2.7,
// ode-initval2/rkf45.c, cstd.c, evolve.c -- restated from the published algorithm, GSL is not in
// this image).  Host/device portable like collapse.cuh: tests/host runs the same source on the CPU.
//
// The reference fills, per smoothing radius, a table F(delta, x, y) of CT_NBINS_D x CT_NBINS_XY^2
// = 250 000 points with `ell()` -- one 9-variable adaptive ODE integration per point when ELL_SNG is
// chosen (403 of 417 s of the reference's f(R) test) -- then evaluates per cell four cubic splines in
// delta and blends them bilinearly in (x, y).  Here:
//   * ct_build_point: one thread per table point (the batch ODE job);
//   * ct_spline_column: one thread per (x, y) column turns the 100 values into {y, b, c, d} records
//     (32 bytes, one sector) per interval, plus two records that encode my_spline_eval's linear
//     extrapolation below / above the knots (src/cosmo.c:2016-2027) as ordinary cubic records;
//   * ct_interpolate: per cell one interval search (uniform look-up + one step) shared by the four
//     columns, four 32-byte gathers from the L2-resident 8 MB table, four Horner evaluations.
#pragma once
#include "collapse.cuh"

namespace pinb {

#define PINB_CT_MAXD 128   /* upper bound on CT_NBINS_D (the reference compiles 100) */
#define PINB_CT_NLUT 1024  /* uniform bins of the interval look-up */

// the reference's compile-time sampling (src/collapse_times.c:781-787)
#define PINB_CT_NBINS_XY 50
#define PINB_CT_NBINS_D 100
#define PINB_CT_SQUEEZE 1.2
#define PINB_CT_EXPO 1.75
#define PINB_CT_RANGE_D 7.0
#define PINB_CT_RANGE_X 3.5
#define PINB_CT_DELTA0 (-1.0)

// delta_vector of initialize_collapse_times (:836-877), CT_EXPO != 1, != 2 branch generalised
inline void ct_default_delta_vector(double* dv, int nd = PINB_CT_NBINS_D) {
  const double E = PINB_CT_EXPO, SQ = PINB_CT_SQUEEZE, RD = PINB_CT_RANGE_D, D0 = PINB_CT_DELTA0;
  if (E == 1.0) {
    const double interval = 2. * RD / (double)nd;
    for (int id = 0; id < nd; id++) dv[id] = id * interval - RD;
    return;
  }
  const double deltaf = pow(SQ / E, 1. / (E - 1.));
  double ref_interval;
  if (E == 2.0)
    ref_interval = ((log((RD - D0) / deltaf) + log((RD + D0) / deltaf)) / E + 2. * deltaf / SQ) / (nd - 2.0);
  else
    ref_interval = ((pow(RD - D0, 2. - E) + pow(RD + D0, 2. - E) - 2. * pow(deltaf, 2. - E)) / E / (2. - E) + 2. * deltaf / SQ) / (nd - 2.0);
  double del = -RD;
  for (int id = 0; id < nd; id++) {
    dv[id] = del;
    double interval = E * ref_interval * pow(fabs(del - D0), E - 1.0);
    interval = (interval / ref_interval < SQ ? ref_interval * SQ : interval);
    del += interval;
  }
}

// one cubic record: value = y + t (b + t (c + t d)), t = delta - knot; 32 bytes = one L2 sector
struct alignas(32) CTRec {
  double y, b, c, d;
};
PINB_HD CTRec ct_rec(double y, double b, double c, double d) {
  CTRec r;
  r.y = y; r.b = b; r.c = c; r.d = d;
  return r;
}
PINB_HD CTRec ct_ld(const CTRec* p) {
#if defined(__CUDA_ARCH__)
  const double2 lo = __ldg(reinterpret_cast<const double2*>(p)), hi = __ldg(reinterpret_cast<const double2*>(p) + 1);
  return ct_rec(lo.x, lo.y, hi.x, hi.y);
#else
  return *p;
#endif
}

// ---- geometry of one table ---------------------------------------------------------------------
// knots: [0] x0  [1] xlast  [2] 1/h  [3] nd   [4 .. 4+nd) delta knots   then PINB_CT_NLUT bytes
struct CTView {
  const double* knots;  // header + delta knots + interval look-up (global memory, L1 resident)
  const CTRec* coef;    // [nxy*nxy][nd + 2] records {y, b, c, d}
  int nd, nxy;
  double inv_ampl;      // 1 / sqrt(Smoothing.Variance[ismooth])
  double inv_bin_x;     // CT_NBINS_XY / CT_RANGE_X
};
inline size_t ct_knots_doubles(int nd) { return 4 + (size_t)nd + PINB_CT_NLUT / 8; }
inline void ct_pack_knots(const double* dv, int nd, double* out) {
  const double h = (dv[nd - 1] - dv[0]) / PINB_CT_NLUT;
  out[0] = dv[0];
  out[1] = dv[nd - 1];
  out[2] = 1.0 / h;
  out[3] = (double)nd;
  for (int i = 0; i < nd; i++) out[4 + i] = dv[i];
  unsigned char* lut = reinterpret_cast<unsigned char*>(out + 4 + nd);
  int i = 0;
  for (int j = 0; j < PINB_CT_NLUT; j++) {
    const double xl = dv[0] + j * h;
    while (i < nd - 2 && dv[i + 1] <= xl) i++;
    lut[j] = (unsigned char)i;
  }
}

// (l1, l2, l3) of table point i (:968-976): i = id + nd * (ix + nxy * iy)
PINB_HD void ct_point_lambdas(int i, const double* dv, int nd, int nxy, double bin_x, double ampl, double& l1, double& l2, double& l3) {
  const int id = i % nd, ix = (i / nd) % nxy, iy = i / nd / nxy;
  const double x = ix * bin_x, y = iy * bin_x;
  l1 = (dv[id] + 2. * x + y) / 3.0 * ampl;
  l2 = (dv[id] - x + y) / 3.0 * ampl;
  l3 = (dv[id] - x - 2. * y) / 3.0 * ampl;
}

// ---- ELL_SNG -----------------------------------------------------------------------------------
// OmegaMatter(z), OmegaLambda(z) of src/cosmo.c:1675-1718 for a cosmological constant
// (params.simpleLambda): E^2(z) = [OmegaRad (1+z)^4 + Omega0 (1+z)^3 + OmegaK (1+z)^2 + OmegaLambda] / E^2(0)
struct SngCosmo {
  double omega0, omega_lambda, omega_rad, omega_k;
  // MOD_GRAV_FR (Hu-Sawicki f(R), src/collapse_times.c:292-311): fr0 = FR0 (0: standard gravity),
  // h_over_c = 100 / SPEEDOFLIGHT (src/cosmo.c:109), fr_size = the smoothing radius handed to sng_system
  double fr0, h_over_c, fr_size;
};

// ForceModification(size, a, delta), src/collapse_times.c:294-311
PINB_HD double sng_force_modification(const SngCosmo& c, double a, double delta) {
  const double ff = 4. * c.omega_lambda / c.omega0;
  const double a3 = a * a * a;
  const double hs = c.h_over_c * c.fr_size;
  const double u = (1.0 + ff) / (1.0 + ff * a3), w = (1.0 + ff) / (1.0 + delta + ff * a3);
  const double thickness = c.fr0 / c.omega0 / (hs * hs) * (a3 * a3 * a) * pow(1. + delta, -1. / 3.) * (u * u - w * w);
  double F3 = thickness * (3. + thickness * (-3. + thickness));
  if (F3 < 0.) F3 = 0.;
  return (F3 < 1. ? F3 / 3. : 1. / 3);
}

// r.h.s. of the nine eigenvalue equations, src/collapse_times.c:239-290; with MOD_GRAV_FR the gravity
// term of the velocity equations is multiplied by 1 + ForceModification
PINB_HD void sng_rhs(double t, const double* y, double* f, const SngCosmo& c) {
  const double z = 1. / t - 1.;
  const double zp = 1. + z, zp2 = zp * zp;
  const double esq = c.omega_rad * zp2 * zp2 + c.omega0 * zp2 * zp + c.omega_k * zp2 + c.omega_lambda;
  const double esq0 = c.omega_rad + c.omega0 + c.omega_k + c.omega_lambda;
  const double ez2 = esq / esq0;
  const double omegam = c.omega0 * (zp2 * zp) / ez2;
  const double omegal = c.omega_lambda / ez2;
  const double delta = y[6] + y[7] + y[8];
  const double sv = y[3] + y[4] + y[5];
  const double rt = 1.0 / t;
  const double grav = (c.fr0 != 0.0) ? 1. + sng_force_modification(c, t, delta) : 1.0;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    double sum = 0.;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      if (i == j || y[i] == y[j]) continue;
      const double ai = (1. - y[i]) * (1. - y[i]), aj = (1. - y[j]) * (1. - y[j]);
      sum += (y[j + 6] - y[i + 6]) * (ai * (1. + y[i + 3]) - aj * (1. + y[j + 3])) / (ai - aj);
    }
    f[i] = (y[i + 3] * (y[i] - 1.0)) * rt;
    f[i + 3] = (0.5 * (y[i + 3] * (omegam - 2.0 * omegal - 2.0) - 3.0 * omegam * y[i + 6] * grav - 2.0 * y[i + 3] * y[i + 3])) * rt;
    f[i + 6] = ((5. / 6. + y[i + 6]) * ((3. + sv) - (1. + delta) / (2.5 + delta) * sv) - (2.5 + delta) * (1. + y[i + 3]) + sum) * rt;
  }
}

// One attempted Runge-Kutta-Fehlberg 4(5) step (GSL rkf45.c): 5th-order solution, error estimate,
// derivative at the new point.  k1 is passed in (GSL reuses dydt_out of the previous step).
PINB_HD void rkf45_try(double t, double h, const double* y, const double* k1, double* ynew, double* yerr, double* dydt_out, const SngCosmo& c) {
  double k2[9], k3[9], k4[9], k5[9], k6[9], yt[9];
#pragma unroll
  for (int i = 0; i < 9; i++) yt[i] = y[i] + (1.0 / 4.0) * h * k1[i];
  sng_rhs(t + (1.0 / 4.0) * h, yt, k2, c);
#pragma unroll
  for (int i = 0; i < 9; i++) yt[i] = y[i] + h * ((3.0 / 32.0) * k1[i] + (9.0 / 32.0) * k2[i]);
  sng_rhs(t + (3.0 / 8.0) * h, yt, k3, c);
#pragma unroll
  for (int i = 0; i < 9; i++) yt[i] = y[i] + h * ((1932.0 / 2197.0) * k1[i] + (-7200.0 / 2197.0) * k2[i] + (7296.0 / 2197.0) * k3[i]);
  sng_rhs(t + (12.0 / 13.0) * h, yt, k4, c);
#pragma unroll
  for (int i = 0; i < 9; i++)
    yt[i] = y[i] + h * ((8341.0 / 4104.0) * k1[i] + (-32832.0 / 4104.0) * k2[i] + (29440.0 / 4104.0) * k3[i] + (-845.0 / 4104.0) * k4[i]);
  sng_rhs(t + h, yt, k5, c);
#pragma unroll
  for (int i = 0; i < 9; i++)
    yt[i] = y[i] + h * ((-6080.0 / 20520.0) * k1[i] + (41040.0 / 20520.0) * k2[i] + (-28352.0 / 20520.0) * k3[i] + (9295.0 / 20520.0) * k4[i] +
                        (-5643.0 / 20520.0) * k5[i]);
  sng_rhs(t + (1.0 / 2.0) * h, yt, k6, c);
#pragma unroll
  for (int i = 0; i < 9; i++) {
    const double d = (902880.0 / 7618050.0) * k1[i] + (3953664.0 / 7618050.0) * k3[i] + (3855735.0 / 7618050.0) * k4[i] +
                     (-1371249.0 / 7618050.0) * k5[i] + (277020.0 / 7618050.0) * k6[i];
    ynew[i] = y[i] + h * d;
    yerr[i] = h * ((1.0 / 360.0) * k1[i] + (-128.0 / 4275.0) * k3[i] + (-2197.0 / 75240.0) * k4[i] + (1.0 / 50.0) * k5[i] + (2.0 / 55.0) * k6[i]);
  }
  sng_rhs(t + h, ynew, dydt_out, c);
}

// ell_sng, src/collapse_times.c:315-400: integrate from a = 1e-5 until lambda_a1 >= 0.99999 (collapse
// of the first axis) or a = 5; gsl_odeiv2_evolve_apply with control_standard_new(1e-6, 1e-6, 1, 1):
// reject and shrink when the scaled error exceeds 1.1 (factor 0.9 r^-1/5, at least 0.2), grow when it
// is below 0.5 (0.9 r^-1/6, at most 5), no new step size after a step cut short to land on amax.
// The collapse epoch is interpolated linearly between the INITIAL point (olda, oldlam are never
// advanced in the reference's loop, :364-386) and the first point past 0.99999 -- reproduced as is.
// Returns -1 on a failed integration (step size underflow), 0 when the ellipsoid does not collapse.
PINB_HD double ell_sng(double l1, double l2, double l3, double D_in, const SngCosmo& c) {
  const double amin = 1.e-5, amax = 5.0;
  double hh = 1.e-6, mya = amin;
  double y[9] = {l1 * D_in, l2 * D_in, l3 * D_in, l1 * D_in / (l1 * D_in - 1.), l2 * D_in / (l2 * D_in - 1.), l3 * D_in / (l3 * D_in - 1.),
                 l1 * D_in, l2 * D_in, l3 * D_in};
  const double olda = mya, oldlam = y[0];
  double k1[9], ynew[9], yerr[9], dnew[9];
  sng_rhs(mya, y, k1, c);
  for (int nstep = 0; nstep < 1000000 && mya < amax; nstep++) {
    double h0 = hh;
    const double dt = amax - mya;
    bool final_step = false;
    if (h0 > dt) {
      h0 = dt;
      final_step = true;
    }
    bool done = false;
    for (int attempt = 0; attempt < 1000; attempt++) {
      rkf45_try(mya, h0, y, k1, ynew, yerr, dnew, c);
      double rmax = 2.2250738585072014e-308;
#pragma unroll
      for (int i = 0; i < 9; i++) {
        const double D0 = 1.0e-6 * (fabs(ynew[i]) + fabs(h0 * dnew[i])) + 1.0e-6;
        const double r = fabs(yerr[i]) / fabs(D0);
        if (r > rmax) rmax = r;
      }
      if (rmax > 1.1) {
        double r = 0.9 / pow(rmax, 1.0 / 5.0);
        if (r < 0.2) r = 0.2;
        const double hnew = h0 * r;
        // GSL's evolve gives up when the shrunken step no longer advances the time
        if (!(fabs(hnew) < fabs(h0)) || mya + hnew == mya) return -1.0;
        h0 = hnew;
        final_step = false;
        continue;
      }
      double hn = h0;
      if (rmax < 0.5) {
        double r = 0.9 / pow(rmax, 1.0 / 6.0);
        if (r > 5.0) r = 5.0;
        if (r < 1.0) r = 1.0;
        hn = h0 * r;
      }
#pragma unroll
      for (int i = 0; i < 9; i++) {
        y[i] = ynew[i];
        k1[i] = dnew[i];
      }
      mya = final_step ? amax : mya + h0;
      if (!final_step) hh = hn;
      done = true;
      break;
    }
    if (!done) return -1.0;
    if (y[0] >= 0.99999) return olda + (1. - oldlam) * (mya - olda) / (y[0] - oldlam);
  }
  return 0.0;
}

// ---- table points -------------------------------------------------------------------------------
// CT_table[i] = ell(ismooth, l1, l2, l3), src/collapse_times.c:404-427, 977
// model 1: ELL_CLASSIC (1 + InverseGrowingMode(b_c));  model 3: ELL_SNG (1 / a_c)
#define PINB_CT_MODEL_CLASSIC 1
#define PINB_CT_MODEL_SNG 3
#define PINB_CT_MODEL_SNG_FR 4 /* same integration, SngCosmo::fr0 != 0 */
PINB_HD double ct_build_point(int model, int i, const double* dv, int nd, int nxy, double bin_x, double ampl, const SplineView& invgrow,
                              double D_in, const SngCosmo& cosmo) {
  double l1, l2, l3;
  ct_point_lambdas(i, dv, nd, nxy, bin_x, ampl, l1, l2, l3);
  if (model == PINB_CT_MODEL_CLASSIC) {
    const double bc = ell_classic(l1, l2, l3);
    return bc > 0.0 ? 1. + inverse_growing_mode(invgrow, bc) : 0.0;
  }
  const double bc = ell_sng(l1, l2, l3, D_in, cosmo);
  return bc > 0.0 ? 1. / bc : 0.0;
}

// ---- spline records of one (x, y) column ----------------------------------------------------------
// gsl_interp_cspline, natural boundary: c[0] = c[n-1] = 0, interior c from the symmetric tridiagonal
// system (diag 2(h_i + h_i+1), off-diagonal h_i+1, rhs 3(dy_i+1/h_i+1 - dy_i/h_i)) solved by the
// L D L^T recurrences of gsl_linalg_solve_symm_tridiag; b_i, d_i derived per interval as cspline_eval
// does.  Records nd and nd+1 are the secant extrapolations below x[0] and above x[nd-1].
PINB_HD void ct_spline_column(const double* x, const double* yv, int n, CTRec* rec) {
  double cc[PINB_CT_MAXD], gamma[PINB_CT_MAXD], alpha[PINB_CT_MAXD];
  const int m = n - 2;
  cc[0] = 0.0;
  cc[n - 1] = 0.0;
  // forward: alpha, gamma, z (z kept in cc[1..m])
  for (int i = 0; i < m; i++) {
    const double h_i = x[i + 1] - x[i], h_ip1 = x[i + 2] - x[i + 1];
    const double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0, g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
    const double diag = 2.0 * (h_ip1 + h_i);
    const double rhs = 3.0 * ((yv[i + 2] - yv[i + 1]) * g_ip1 - (yv[i + 1] - yv[i]) * g_i);
    if (i == 0) {
      alpha[0] = diag;
      cc[1] = rhs;
    } else {
      const double off_prev = x[i + 1] - x[i];  // offdiag[i-1] = h_(i-1)+1 = h_i
      alpha[i] = diag - off_prev * gamma[i - 1];
      cc[i + 1] = rhs - gamma[i - 1] * cc[i];
    }
    gamma[i] = h_ip1 / alpha[i];
  }
  for (int i = 0; i < m; i++) cc[i + 1] = cc[i + 1] / alpha[i];
  for (int i = m - 2; i >= 0; i--) cc[i + 1] = cc[i + 1] - gamma[i] * cc[i + 2];
  for (int i = 0; i < n - 1; i++) {
    const double dx = x[i + 1] - x[i], dy = yv[i + 1] - yv[i];
    rec[i] = ct_rec(yv[i], dy / dx - dx * (cc[i + 1] + 2.0 * cc[i]) / 3.0, cc[i], (cc[i + 1] - cc[i]) / (3.0 * dx));
  }
  rec[n - 1] = ct_rec(yv[n - 1], 0.0, 0.0, 0.0);  // never selected (bsearch stops at n-2)
  rec[n] = ct_rec(yv[0], (yv[1] - yv[0]) / (x[1] - x[0]), 0.0, 0.0);
  rec[n + 1] = ct_rec(yv[n - 1], (yv[n - 1] - yv[n - 2]) / (x[n - 1] - x[n - 2]), 0.0, 0.0);
}

// ---- per-cell evaluation ---------------------------------------------------------------------------
// interpolate_collapse_time (BILINEAR_SPLINE), src/collapse_times.c:1132-1147, 1211-1221
PINB_HD double ct_interpolate(const CTView& v, double l1, double l2, double l3) {
  const double* kn = v.knots;
  const int nd = v.nd, nxy = v.nxy;
  const double d = (l1 + l2 + l3) * v.inv_ampl;
  const double xs = (l1 - l2) * v.inv_ampl * v.inv_bin_x;
  const double ys = (l2 - l3) * v.inv_ampl * v.inv_bin_x;
  int ix = (int)xs, iy = (int)ys;
  ix = (ix >= nxy - 1) ? nxy - 2 : (ix < 0) ? 0 : ix;
  iy = (iy >= nxy - 1) ? nxy - 2 : (iy < 0) ? 0 : iy;
  const double dx = xs - ix, dy = ys - iy;
  // record and reference abscissa, shared by the four columns
  int e;
  double xr;
  if (d < kn[0]) {
    e = nd;
    xr = kn[0];
  } else if (d > kn[1]) {
    e = nd + 1;
    xr = kn[1];
  } else {
    int j = (int)((d - kn[0]) * kn[2]);
    j = j < 0 ? 0 : (j > PINB_CT_NLUT - 1 ? PINB_CT_NLUT - 1 : j);
    e = reinterpret_cast<const unsigned char*>(kn + 4 + nd)[j];
    // gsl_interp_bsearch semantics: largest i in [0, nd-2] with x[i] <= d
    while (e < nd - 2 && kn[4 + e + 1] <= d) e++;
    if (e > 0 && d < kn[4 + e]) e--;
    xr = kn[4 + e];
  }
  const double t = d - xr;
  const size_t ne = (size_t)nd + 2;
  const CTRec* c00 = v.coef + ((size_t)ix + (size_t)iy * nxy) * ne + e;
  const CTRec r00 = ct_ld(c00), r10 = ct_ld(c00 + ne), r01 = ct_ld(c00 + (size_t)nxy * ne), r11 = ct_ld(c00 + (size_t)nxy * ne + ne);
  const double v00 = r00.y + t * (r00.b + t * (r00.c + t * r00.d));
  const double v10 = r10.y + t * (r10.b + t * (r10.c + t * r10.d));
  const double v01 = r01.y + t * (r01.b + t * (r01.c + t * r01.d));
  const double v11 = r11.y + t * (r11.b + t * (r11.c + t * r11.d));
  return (1. - dx) * (1. - dy) * v00 + dx * (1. - dy) * v10 + (1. - dx) * dy * v01 + dx * dy * v11;
}

// inverse_collapse_time with TABULATED_CT (src/collapse_times.c:679-776): same eigenvalues and ordering
// as the ELL_CLASSIC version of collapse.cuh, F read from the table
// (the eigenvalue block is repeated here rather than shared so that the tuned ELL_CLASSIC kernel's
// code generation is untouched)
PINB_HD void hessian_eigenvalues(const double* d, double& hi, double& mid, double& lo, bool& bad) {
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
  const bool diag = (q == 0.);
  bad = !diag && (q * q * q < r * r || q < 0.0);
  const bool unused = diag || bad;
  const SqrtPair rt = fm_sqrt_pair(unused ? 1.0 : q);
  const double sq = 2 * rt.s;
  const double t = fm_acos(unused ? 0.5 : r * (rt.rs * rt.rs * rt.rs));
  const double m3 = mu1 * mc(MC_1_3);
  double c0, c1, c2;
  cos_thirds(t, c0, c1, c2);
  const double x1 = diag ? d[0] : -sq * c0 + m3;
  const double x2 = diag ? d[1] : -sq * c1 + m3;
  const double x3 = diag ? d[2] : -sq * c2 + m3;
  const double m12 = (x1 > x2 ? x1 : x2), n12 = (x1 < x2 ? x1 : x2);
  hi = (m12 > x3 ? m12 : x3);
  lo = (n12 < x3 ? n12 : x3);
  mid = x1 + x2 + x3 - lo - hi;
}
PINB_HD double inverse_collapse_time_tab(const double* d, const CTView& v) {
  double hi, mid, lo;
  bool bad;
  hessian_eigenvalues(d, hi, mid, lo, bad);
  // benign operands on the lanes that return -10 (NaN eigenvalues would index out of the table)
  const double F = ct_interpolate(v, bad ? 0.0 : hi, bad ? 0.0 : mid, bad ? 0.0 : lo);
  return bad ? -10.0 : F;
}

// ---- kernel bodies (execution-context template as in kernels.cuh) ------------------------------------
struct CTBuildParams {
  int model;            // PINB_CT_MODEL_CLASSIC / PINB_CT_MODEL_SNG
  const double* dv;     // nd delta knots
  int nd, nxy;
  double bin_x, ampl;
  const double* spline; // packed inverse-growth spline of this radius (ELL_CLASSIC), global memory
  int nspl;
  double D_in;          // GrowingMode(1/amin - 1, k(R)) (ELL_SNG)
  SngCosmo cosmo;
  double* table;        // [nxy*nxy][nd] as the reference's CT_table
  int npoints;
};
template <class Ctx> PINB_HD void ct_build_body(Ctx& ctx, const CTBuildParams& p) {
  const int i = ctx.bid() * ctx.nthreads() + ctx.tid();
  if (i >= p.npoints) return;
  const SplineView sp{p.spline, p.nspl};
  p.table[i] = ct_build_point(p.model, i, p.dv, p.nd, p.nxy, p.bin_x, p.ampl, sp, p.D_in, p.cosmo);
}

struct CTSplineParams {
  const double* dv;
  int nd, ncols;
  const double* table;
  CTRec* coef;  // [ncols][nd + 2]
};
template <class Ctx> PINB_HD void ct_spline_body(Ctx& ctx, const CTSplineParams& p) {
  const int col = ctx.bid() * ctx.nthreads() + ctx.tid();
  if (col >= p.ncols) return;
  ct_spline_column(p.dv, p.table + (size_t)col * p.nd, p.nd, p.coef + (size_t)col * (p.nd + 2));
}

}  // namespace pinb
