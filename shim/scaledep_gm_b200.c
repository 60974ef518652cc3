/* Reference-side binding of pinb200_scaledep_variances: set_scaledep_GM with its
 * 3 x Nsmooth x NBINS gsl_integration_qags calls replaced by ONE device call.
 *
 * Replaces: set_scaledep_GM, /root/reference/src/initialization.c:1533-2026 (called at :115).  Compiled against the
 * reference's own pinocchio.h / def_splines.h like shim/fmax_b200.c; everything the function leaves behind is the
 * reference's: Smoothing.Rad_GM, Smoothing.k_GM_dens / _displ / _vel and, with -DELL_CLASSIC, the per-radius
 * inverse-growth splines SPLINE_INVGROW[] that InverseGrowingMode (src/cosmo.c:1828) and, through
 * pinb200_set_invgrow_spline, the collapse kernel read.  A maintainer's one-line change is at the call site
 * (INTEGRATION.md section 7a): `if (set_scaledep_GM_b200()) return 1;`.
 *
 * What runs where.  Host, unchanged reference calls: PowerSpectrum(k) at the quadrature nodes (any WhichSpectrum,
 * WDM cut included), the k-bin growth splines at the time knots (my_spline_eval on SPLINE[SP_GROW1 + kk],
 * SPLINE[SP_FOMEGA1 + kk]: what InterpolateGrowth blends, src/cosmo.c:1728-1755), SizeForMass, and the bisection
 * for the k whose growth history matches each integral (:1608-1681, :1755-1828, :1899-1972 -- three copies of one
 * loop in the reference, one function here, with the reference's two slips kept as they are, see bisect_k).
 * Device: the integrals.  The quadrature is fixed, not adaptive -- composite 8-point Gauss-Legendre on about 512
 * panels of the reference's own interval [-4, nyquist] in log10 k (the reference passes the Nyquist WAVENUMBER as the
 * upper limit of log10 k, :1599; kept), panel edges on the k bins where the growth interpolation has its kinks -- and
 * converged to 1e-11 on these integrands, far inside the reference's TOLERANCE of 1e-4 (tests/test_scaledep_gm.py pins
 * it against QUADPACK).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pinocchio.h"
#include "pinb200.h"

#ifdef SCALE_DEPENDENT
#include "def_splines.h"

#define SDGM_SMALLDIFF ((double)1.e-5) /* SMALLDIFF, src/initialization.c:1431 */
#define SDGM_MAXITER 20                /* MAXITER, :1433 */
#define SDGM_PANELS 512
#define SDGM_ORDER 8

static const double gl8_x[SDGM_ORDER] = {-9.60289856497536176e-01, -7.96666477413626728e-01, -5.25532409916328991e-01,
                                         -1.83434642495649780e-01, 1.83434642495649780e-01,  5.25532409916328991e-01,
                                         7.96666477413626728e-01,  9.60289856497536176e-01};
static const double gl8_w[SDGM_ORDER] = {1.01228536290377064e-01, 2.22381034453374427e-01, 3.13706645877886880e-01,
                                         3.62683783378361657e-01, 3.62683783378361657e-01, 3.13706645877886880e-01,
                                         2.22381034453374427e-01, 1.01228536290377064e-01};

/* growth history the integral is matched against: D(z, k), or D f_Omega for the velocities */
static double gm_of(int with_fomega, double z, double k, double k_fomega) {
  return with_fomega ? GrowingMode(z, k) * fomega(z, k_fomega) : GrowingMode(z, k);
}

/* mean over the knots Z20..Today of vector[i] - history(t_i, k) / history(today, k), divided by NBINS (:1620-1627).
 * k_fo: the wavenumber handed to fomega -- the reference's first guess at k2 uses k1 there (:1916), kept */
static double mean_diff(const double *vector, int Z20, int Today, int with_fomega, double k, double k_fo_hist) {
  const double norm = gm_of(with_fomega, 0.0, k, k);
  double diff = 0.0;
  for (int i = Z20; i <= Today; i++) {
    const double Time = pow(10., SPLINE[SP_TIME]->x[i]);
    diff += vector[i] - gm_of(with_fomega, 1. / Time - 1., k, k_fo_hist) / norm;
  }
  return diff / (double)NBINS;
}

/* the bisector search of the reference (three copies there: density :1608-1681, displacements :1755-1828,
 * velocities :1899-1972).  which = 0 density: when both ends miss with the same sign the reference stores log10(k1)
 * instead of k1 (:1648); kept. */
static double bisect_k(const double *vector, int Z20, int Today, int which, int ismooth, double radius) {
  const int with_fomega = (which == 2);
  double logk1 = LOGKMIN, logk2 = LOGKMIN + (NkBINS - 1) * DELTALOGK;
  const double k1 = pow(10., logk1), k2 = pow(10., logk2);
  double diff1 = mean_diff(vector, Z20, Today, with_fomega, k1, k1);
  double diff2 = mean_diff(vector, Z20, Today, with_fomega, k2, with_fomega ? k1 : k2);
  if (fabs(diff1) < SDGM_SMALLDIFF) return k1;
  if (fabs(diff2) < SDGM_SMALLDIFF) return k2;
  if (diff1 * diff2 > 0) {
    if (!ThisTask)
      printf("WARNING in scale-dependent density growth rate for smoothing radius %d (%f): accuracy not guaranteed [diff1=%g - diff2=%g]\n",
             ismooth, radius, diff1, diff2);
    if (fabs(diff1) < fabs(diff2)) return which == 0 ? logk1 : k1;
    return k2;
  }
  double mindiff = fabs(diff1), kmid = k1;
  mindiff = (fabs(diff2) < mindiff ? fabs(diff2) : mindiff);
  int iter = 0;
  do {
    const double logkmid = 0.5 * (logk1 + logk2);
    kmid = pow(10., logkmid);
    const double diffm = mean_diff(vector, Z20, Today, with_fomega, kmid, kmid);
    mindiff = (fabs(diffm) < mindiff ? fabs(diffm) : mindiff);
    if (diff1 * diffm > 0) {
      logk1 = logkmid;
      diff1 = diffm;
    } else {
      logk2 = logkmid;
      diff2 = diffm;
    }
    ++iter;
  } while (fabs(mindiff) > SDGM_SMALLDIFF && iter <= SDGM_MAXITER);
  return kmid;
}

int set_scaledep_GM_b200(void) {
  const int S = Smoothing.Nsmooth;
  int i, ismooth, Today, Z20;

#ifdef ELL_CLASSIC
  SPLINE_INVGROW = (gsl_spline **)calloc(S, sizeof(gsl_spline *));
  for (i = 0; i < S; i++) SPLINE_INVGROW[i] = gsl_spline_alloc(gsl_interp_cspline, NBINS);
  ACCEL_INVGROW = (gsl_interp_accel **)calloc(S, sizeof(gsl_interp_accel *));
  for (i = 0; i < S; i++) ACCEL_INVGROW[i] = gsl_interp_accel_alloc();
#endif
  Smoothing.Rad_GM = (double *)malloc(S * sizeof(double));
  Smoothing.k_GM_dens = (double *)malloc(S * sizeof(double));
  Smoothing.k_GM_displ = (double *)malloc(S * sizeof(double));
  Smoothing.k_GM_vel = (double *)malloc(S * sizeof(double));

  for (i = 0; i < NBINS; i++)
    if (pow(10., SPLINE[SP_TIME]->x[i]) > 1.0) break;
  Today = i - 1;
  for (i = 0; i < NBINS; i++)
    if (pow(10., SPLINE[SP_TIME]->x[i]) > 1. / 21.) break;
  Z20 = i - 1;

  /* radii: the Gaussian ones of the sweep for the density, a linear ladder down from the largest halo's size for the
   * displacements and velocities (:1734-1738) */
  const double Largest = SizeForMass(pow(10., mf.mmax));
  for (ismooth = 0; ismooth < S; ismooth++)
    Smoothing.Rad_GM[ismooth] = Largest * (S - 1 - ismooth) / (double)(S - 1);

  /* quadrature nodes and the factor of the integrand that is the host cosmology's alone.  InterpolateGrowth is
   * piecewise linear in log10 k, so the integrand has kinks at the k bins: panel edges are put there, every segment
   * between two cuts getting its share of the SDGM_PANELS panels (at least one). */
  const double lo = -4., hi = NYQUIST * PI / params.InterPartDist; /* sic: :1585, :1599 */
  double cuts[NkBINS + 2];
  int ncuts = 0;
  cuts[ncuts++] = lo;
  for (int kk = 0; kk < NkBINS; kk++) {
    const double b = LOGKMIN + kk * DELTALOGK;
    if (b > lo && b < hi) cuts[ncuts++] = b;
  }
  cuts[ncuts++] = hi;
  int npanels = 0;
  for (int c = 0; c + 1 < ncuts; c++) {
    int m = (int)floor(SDGM_PANELS * (cuts[c + 1] - cuts[c]) / (hi - lo) + 0.5);
    npanels += m < 1 ? 1 : m;
  }
  const int n = npanels * SDGM_ORDER;
  double *logk = (double *)malloc(n * sizeof(double)), *a_dens = (double *)malloc(n * sizeof(double)),
         *a_disp = (double *)malloc(n * sizeof(double));
  int j = 0;
  for (int c = 0; c + 1 < ncuts; c++) {
    int m = (int)floor(SDGM_PANELS * (cuts[c + 1] - cuts[c]) / (hi - lo) + 0.5);
    if (m < 1) m = 1;
    const double h = 0.5 * (cuts[c + 1] - cuts[c]) / m;
    for (int pnl = 0; pnl < m; pnl++)
      for (int g = 0; g < SDGM_ORDER; g++, j++) {
        logk[j] = cuts[c] + (2 * pnl + 1) * h + h * gl8_x[g];
        const double k = pow(10., logk[j]);
        const double P = PowerSpectrum(k);
        a_dens[j] = h * gl8_w[g] * P * k * k * k / (2. * PI * PI);
        a_disp[j] = h * gl8_w[g] * P * k / (2. * PI * PI);
      }
  }

  /* what InterpolateGrowth blends: the k-bin splines at the time knots */
  double *lg = (double *)malloc((size_t)NkBINS * NBINS * sizeof(double)), *fo = (double *)malloc((size_t)NkBINS * NBINS * sizeof(double));
  for (int kk = 0; kk < NkBINS; kk++)
    for (i = 0; i < NBINS; i++) {
      const double Time = pow(10., SPLINE[SP_TIME]->x[i]);
      const double arg = -log10(1. + (1. / Time - 1.)); /* as InterpolateGrowth is handed z = 1/Time - 1 */
      lg[(size_t)kk * NBINS + i] = my_spline_eval(SPLINE[SP_GROW1 + kk], arg, ACCEL[SP_GROW1 + kk]);
      fo[(size_t)kk * NBINS + i] = my_spline_eval(SPLINE[SP_FOMEGA1 + kk], arg, ACCEL[SP_FOMEGA1 + kk]);
    }

  pinb200_sdgm_desc d;
  memset(&d, 0, sizeof(d));
  {
    const char *dev = getenv("PINB200_DEVICE");
    d.device = dev ? atoi(dev) : 0;
  }
  d.nnodes = n;
  d.logk = logk;
  d.a_dens = a_dens;
  d.a_disp = a_disp;
  d.nkbins = NkBINS;
  d.ntimes = NBINS;
  d.logkmin = LOGKMIN;
  d.dlogk = DELTALOGK;
  d.log10_growth = lg;
  d.fomega = fo;
  d.nsmooth = S;
  d.radius_dens = Smoothing.Radius;
  d.radius_disp = Smoothing.Rad_GM;
  double *all = (double *)malloc((size_t)3 * S * NBINS * sizeof(double));
  if (pinb200_scaledep_variances(&d, all)) {
    if (!ThisTask) printf("ERROR on task %d: pinb200_scaledep_variances: %s\n", ThisTask, pinb200_last_error(NULL));
    return 1;
  }

  double *vector = (double *)malloc(NBINS * sizeof(double));
  for (int which = 0; which < 3; which++)
    for (ismooth = 0; ismooth < S; ismooth++) {
      const double *v = all + ((size_t)which * S + ismooth) * NBINS;
      const double normv = v[Today];
      for (i = 0; i < NBINS; i++) vector[i] = v[i] / normv;
      const double radius = which == 0 ? Smoothing.Radius[ismooth] : Smoothing.Rad_GM[ismooth];
      const double kbest = bisect_k(vector, Z20, Today, which, ismooth, radius);
      if (which == 0) Smoothing.k_GM_dens[ismooth] = kbest;
      else if (which == 1) Smoothing.k_GM_displ[ismooth] = kbest;
      else Smoothing.k_GM_vel[ismooth] = kbest;
#ifdef ELL_CLASSIC
      if (which == 0) {
        for (i = 0; i < NBINS; i++) vector[i] = log10(vector[i]);
        gsl_spline_init(SPLINE_INVGROW[ismooth], vector, &(SPLINE[SP_TIME]->x[0]), NBINS);
      }
#endif
    }

  free(vector);
  free(all);
  free(lg);
  free(fo);
  free(logk);
  free(a_dens);
  free(a_disp);
  WindowFunctionType = 2; /* as the reference leaves it, :2024 */
  return 0;
}
#endif /* SCALE_DEPENDENT */
