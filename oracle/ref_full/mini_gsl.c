/* TEST INFRASTRUCTURE: the part of GSL 2.7.1 that the reference's host code calls outside the hot
 * path (src/cosmo.c, src/initialization.c, src/fragment.c, src/build_groups.c), restated from the
 * published algorithms because GSL is absent from this image.  With it the WHOLE
 * reference program links and runs on one task (oracle/_ref/pinocchio_ref.x), which is what makes
 * the catalogue comparison of north_star possible here.  Not bit-compatible with GSL where GSL's
 * own result is only defined by a tolerance:
 *   - gsl_interp linear / cspline (natural), accel = plain bisection: same arithmetic as GSL
 *     (tridiagonal system of interpolation/cspline.c, evaluation y + d(b + d(c + d e)));
 *   - gsl_integration_qags: adaptive bisection on the interval of largest error estimate, Gauss-
 *     Legendre 20 against 10 points, WITHOUT the epsilon extrapolation of QUADPACK; the requested
 *     relative tolerance is tightened by 1e-3 so that the result is accurate to ~1e-7 where GSL's
 *     extrapolated one is accurate to better than its 1e-4 request (both differ from the true
 *     integral by less than they differ from each other's request);
 *   - gsl_odeiv2: Runge-Kutta-Fehlberg 4(5) step, standard step-size control (S = 0.9, factors
 *     0.2 .. 5), one accepted step per evolve_apply;
 *   - gsl_root_fsolver_brent: bracketing secant/bisection steps;
 *   - gsl_spline2d bicubic (only -DREAD_PK_TABLE): GSL's construction (derivatives from 1-D natural splines,
 *     bicubic Hermite patch per cell), evaluated in Hermite-basis form.
 * Nothing in the product links this file. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <gsl/gsl_errno.h>
#include <gsl/gsl_integration.h>
#include <gsl/gsl_math.h>
#include <gsl/gsl_odeiv2.h>
#include <gsl/gsl_roots.h>
#include <gsl/gsl_spline.h>
#include <gsl/gsl_spline2d.h>

gsl_error_handler_t* gsl_set_error_handler_off(void) { return NULL; }
const char* gsl_strerror(int e) { (void)e; return "mini-gsl error"; }

/* ---- interpolation ------------------------------------------------------------------------ */
struct gsl_interp_type_s { int kind; };
static const gsl_interp_type t_linear = {0}, t_cspline = {1};
const gsl_interp_type* gsl_interp_linear = &t_linear;
const gsl_interp_type* gsl_interp_cspline = &t_cspline;
struct gsl_interp_s { int kind; size_t n; double xmin, xmax; double* c; };
struct gsl_interp_accel_s { size_t cache; };

gsl_interp_accel* gsl_interp_accel_alloc(void) { return calloc(1, sizeof(gsl_interp_accel)); }
void gsl_interp_accel_free(gsl_interp_accel* a) { free(a); }
int gsl_interp_accel_reset(gsl_interp_accel* a) { if (a) a->cache = 0; return GSL_SUCCESS; }

gsl_interp* gsl_interp_alloc(const gsl_interp_type* T, size_t n) {
  gsl_interp* it = calloc(1, sizeof(gsl_interp));
  it->kind = T->kind;
  it->n = n;
  it->c = calloc(n ? n : 1, sizeof(double));
  return it;
}
void gsl_interp_free(gsl_interp* it) { if (it) { free(it->c); free(it); } }

int gsl_interp_init(gsl_interp* it, const double* x, const double* y, size_t n) {
  it->n = n;
  it->xmin = x[0];
  it->xmax = x[n - 1];
  if (it->kind == 0 || n < 3) { memset(it->c, 0, n * sizeof(double)); return GSL_SUCCESS; }
  /* natural cubic spline: c[0] = c[n-1] = 0, symmetric tridiagonal system for c[1..n-2] */
  const size_t m = n - 2;
  double *diag = malloc(m * sizeof(double)), *off = malloc(m * sizeof(double)), *g = malloc(m * sizeof(double));
  for (size_t i = 0; i < m; i++) {
    const double h_i = x[i + 1] - x[i], h_ip1 = x[i + 2] - x[i + 1];
    off[i] = h_ip1;
    diag[i] = 2.0 * (h_ip1 + h_i);
    g[i] = 3.0 * ((y[i + 2] - y[i + 1]) / h_ip1 - (y[i + 1] - y[i]) / h_i);
  }
  for (size_t i = 1; i < m; i++) {
    const double w = off[i - 1] / diag[i - 1];
    diag[i] -= w * off[i - 1];
    g[i] -= w * g[i - 1];
  }
  it->c[0] = it->c[n - 1] = 0.0;
  it->c[m] = g[m - 1] / diag[m - 1];
  for (size_t i = m - 1; i-- > 0;) it->c[i + 1] = (g[i] - off[i] * it->c[i + 2]) / diag[i];
  free(diag); free(off); free(g);
  return GSL_SUCCESS;
}

/* largest i in [0, n-2] with x[i] <= xq (the last interval is closed on the right) */
static size_t bsearch_interval(const double* x, size_t n, double xq) {
  size_t lo = 0, hi = n - 1;
  while (hi > lo + 1) {
    const size_t mid = (hi + lo) / 2;
    if (x[mid] > xq) hi = mid; else lo = mid;
  }
  return lo;
}

static double interp_eval_common(const gsl_interp* it, const double* x, const double* y, double xq, int deriv) {
  if (xq < it->xmin || xq > it->xmax) return NAN; /* GSL: domain error */
  const size_t i = bsearch_interval(x, it->n, xq);
  const double dx = x[i + 1] - x[i], dy = y[i + 1] - y[i], d = xq - x[i];
  if (it->kind == 0 || it->n < 3) return deriv ? dy / dx : y[i] + d * dy / dx;
  const double ci = it->c[i], cip1 = it->c[i + 1];
  const double b = dy / dx - dx * (cip1 + 2.0 * ci) / 3.0;
  const double e = (cip1 - ci) / (3.0 * dx);
  return deriv ? b + d * (2.0 * ci + 3.0 * e * d) : y[i] + d * (b + d * (ci + d * e));
}
double gsl_interp_eval(const gsl_interp* it, const double* x, const double* y, double xq, gsl_interp_accel* a) {
  (void)a;
  return interp_eval_common(it, x, y, xq, 0);
}

gsl_spline* gsl_spline_alloc(const gsl_interp_type* T, size_t n) {
  gsl_spline* s = calloc(1, sizeof(gsl_spline));
  s->interp = gsl_interp_alloc(T, n);
  s->x = calloc(n ? n : 1, sizeof(double));
  s->y = calloc(n ? n : 1, sizeof(double));
  s->size = n;
  return s;
}
int gsl_spline_init(gsl_spline* s, const double* x, const double* y, size_t n) {
  memcpy(s->x, x, n * sizeof(double));
  memcpy(s->y, y, n * sizeof(double));
  s->size = n;
  return gsl_interp_init(s->interp, s->x, s->y, n);
}
double gsl_spline_eval(const gsl_spline* s, double xq, gsl_interp_accel* a) {
  (void)a;
  return interp_eval_common(s->interp, s->x, s->y, xq, 0);
}
double gsl_spline_eval_deriv(const gsl_spline* s, double xq, gsl_interp_accel* a) {
  (void)a;
  return interp_eval_common(s->interp, s->x, s->y, xq, 1);
}
void gsl_spline_free(gsl_spline* s) { if (s) { gsl_interp_free(s->interp); free(s->x); free(s->y); free(s); } }

/* ---- 2-D splines: only reached with -DREAD_PK_TABLE ------------------------------------------- */
struct gsl_interp2d_type_s { int kind; };
static const gsl_interp2d_type t_bicubic = {0};
const gsl_interp2d_type* gsl_interp2d_bicubic = &t_bicubic;
/* gsl_interp2d_bicubic (GSL 2.7 interpolation/bicubic.c): the partial derivatives z_x, z_y, z_xy at the grid
 * points come from one-dimensional natural cubic splines through the rows and columns (z_xy: through the
 * columns of z_y), and every cell is the bicubic Hermite patch through the four corner values and derivatives.
 * z is stored as GSL does: z[j * nx + i] = z(x_i, y_j). */
struct gsl_spline2d_s { size_t nx, ny; double *x, *y, *z, *zx, *zy, *zxy; };
gsl_spline2d* gsl_spline2d_alloc(const gsl_interp2d_type* T, size_t nx, size_t ny) {
  (void)T;
  gsl_spline2d* s = calloc(1, sizeof(gsl_spline2d));
  s->nx = nx;
  s->ny = ny;
  s->x = calloc(nx, sizeof(double));
  s->y = calloc(ny, sizeof(double));
  s->z = calloc(nx * ny, sizeof(double));
  s->zx = calloc(nx * ny, sizeof(double));
  s->zy = calloc(nx * ny, sizeof(double));
  s->zxy = calloc(nx * ny, sizeof(double));
  return s;
}
int gsl_spline2d_init(gsl_spline2d* s, const double* x, const double* y, const double* z, size_t nx, size_t ny) {
  memcpy(s->x, x, nx * sizeof(double));
  memcpy(s->y, y, ny * sizeof(double));
  memcpy(s->z, z, nx * ny * sizeof(double));
  const size_t nmax = nx > ny ? nx : ny;
  double* col = malloc(nmax * sizeof(double));
  gsl_spline* sx = gsl_spline_alloc(gsl_interp_cspline, nx);
  gsl_spline* sy = gsl_spline_alloc(gsl_interp_cspline, ny);
  for (size_t j = 0; j < ny; j++) { /* z_x: splines in x along every row j */
    for (size_t i = 0; i < nx; i++) col[i] = z[j * nx + i];
    gsl_spline_init(sx, x, col, nx);
    for (size_t i = 0; i < nx; i++) s->zx[j * nx + i] = gsl_spline_eval_deriv(sx, x[i], NULL);
  }
  for (size_t i = 0; i < nx; i++) { /* z_y: splines in y along every column i */
    for (size_t j = 0; j < ny; j++) col[j] = z[j * nx + i];
    gsl_spline_init(sy, y, col, ny);
    for (size_t j = 0; j < ny; j++) s->zy[j * nx + i] = gsl_spline_eval_deriv(sy, y[j], NULL);
  }
  for (size_t j = 0; j < ny; j++) { /* z_xy: splines in x through z_y */
    for (size_t i = 0; i < nx; i++) col[i] = s->zy[j * nx + i];
    gsl_spline_init(sx, x, col, nx);
    for (size_t i = 0; i < nx; i++) s->zxy[j * nx + i] = gsl_spline_eval_deriv(sx, x[i], NULL);
  }
  gsl_spline_free(sx);
  gsl_spline_free(sy);
  free(col);
  return GSL_SUCCESS;
}
double gsl_spline2d_eval(const gsl_spline2d* s, double x, double y, gsl_interp_accel* a, gsl_interp_accel* b) {
  (void)a; (void)b;
  const size_t nx = s->nx;
  if (x < s->x[0] || x > s->x[nx - 1] || y < s->y[0] || y > s->y[s->ny - 1]) return NAN; /* GSL: domain error */
  const size_t xi = bsearch_interval(s->x, nx, x), yi = bsearch_interval(s->y, s->ny, y);
  const double dx = s->x[xi + 1] - s->x[xi], dy = s->y[yi + 1] - s->y[yi];
  const double t = (x - s->x[xi]) / dx, u = (y - s->y[yi]) / dy;
  const size_t i00 = yi * nx + xi, i10 = yi * nx + xi + 1, i01 = (yi + 1) * nx + xi, i11 = (yi + 1) * nx + xi + 1;
  /* corner data in the unit square: f, f_t = z_x dx, f_u = z_y dy, f_tu = z_xy dx dy */
  const double f[4] = {s->z[i00], s->z[i10], s->z[i01], s->z[i11]};
  const double ft[4] = {s->zx[i00] * dx, s->zx[i10] * dx, s->zx[i01] * dx, s->zx[i11] * dx};
  const double fu[4] = {s->zy[i00] * dy, s->zy[i10] * dy, s->zy[i01] * dy, s->zy[i11] * dy};
  const double ftu[4] = {s->zxy[i00] * dx * dy, s->zxy[i10] * dx * dy, s->zxy[i01] * dx * dy, s->zxy[i11] * dx * dy};
  /* cubic Hermite basis in each direction: h00 value at 0, h01 value at 1, h10 slope at 0, h11 slope at 1 */
  const double t2 = t * t, t3 = t2 * t, u2 = u * u, u3 = u2 * u;
  const double ht[4] = {2 * t3 - 3 * t2 + 1, -2 * t3 + 3 * t2, t3 - 2 * t2 + t, t3 - t2};
  const double hu[4] = {2 * u3 - 3 * u2 + 1, -2 * u3 + 3 * u2, u3 - 2 * u2 + u, u3 - u2};
  double zq = 0.0;
  for (int jc = 0; jc < 2; jc++)   /* corner (ic, jc) is element ic + 2 jc of the arrays above */
    for (int ic = 0; ic < 2; ic++) {
      const int c = ic + 2 * jc;
      zq += f[c] * ht[ic] * hu[jc] + ft[c] * ht[2 + ic] * hu[jc] + fu[c] * ht[ic] * hu[2 + jc] + ftu[c] * ht[2 + ic] * hu[2 + jc];
    }
  return zq;
}
void gsl_spline2d_free(gsl_spline2d* s) {
  if (!s) return;
  free(s->x); free(s->y); free(s->z); free(s->zx); free(s->zy); free(s->zxy); free(s);
}

/* ---- quadrature ------------------------------------------------------------------------------ */
struct gsl_integration_workspace_s { size_t limit; double *a, *b, *r, *e; };
gsl_integration_workspace* gsl_integration_workspace_alloc(size_t n) {
  gsl_integration_workspace* w = calloc(1, sizeof(*w));
  w->limit = n;
  w->a = malloc(n * sizeof(double));
  w->b = malloc(n * sizeof(double));
  w->r = malloc(n * sizeof(double));
  w->e = malloc(n * sizeof(double));
  return w;
}
void gsl_integration_workspace_free(gsl_integration_workspace* w) { if (w) { free(w->a); free(w->b); free(w->r); free(w->e); free(w); } }

/* Gauss-Legendre nodes and weights on [-1, 1] by Newton iteration on P_n */
static void gauss_legendre(int n, double* x, double* w) {
  for (int i = 0; i < n; i++) {
    double z = cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 1.0;
    for (int it = 0; it < 100; it++) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 1; j <= n; j++) {
        const double p3 = p2;
        p2 = p1;
        p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
      }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      const double dz = p1 / pp;
      z -= dz;
      if (fabs(dz) < 1e-16) break;
    }
    x[i] = z;
    w[i] = 2.0 / ((1.0 - z * z) * pp * pp);
  }
}
static double x10[10], w10[10], x20[20], w20[20];
static int gl_ready = 0;
static void rule(const gsl_function* f, double a, double b, double* res, double* err) {
  if (!gl_ready) { gauss_legendre(10, x10, w10); gauss_legendre(20, x20, w20); gl_ready = 1; }
  const double c = 0.5 * (a + b), h = 0.5 * (b - a);
  double s10 = 0.0, s20 = 0.0;
  for (int i = 0; i < 10; i++) s10 += w10[i] * GSL_FN_EVAL(f, c + h * x10[i]);
  for (int i = 0; i < 20; i++) s20 += w20[i] * GSL_FN_EVAL(f, c + h * x20[i]);
  *res = h * s20;
  *err = fabs(h * (s20 - s10));
}
int gsl_integration_qags(const gsl_function* f, double a, double b, double epsabs, double epsrel, size_t limit,
                         gsl_integration_workspace* w, double* result, double* abserr) {
  if (limit > w->limit) limit = w->limit;
  const double rel = epsrel * 1e-3; /* see the header */
  size_t n = 1;
  w->a[0] = a;
  w->b[0] = b;
  rule(f, a, b, &w->r[0], &w->e[0]);
  for (;;) {
    double tot = 0.0, err = 0.0;
    size_t worst = 0;
    for (size_t i = 0; i < n; i++) {
      tot += w->r[i];
      err += w->e[i];
      if (w->e[i] > w->e[worst]) worst = i;
    }
    const double tol = fmax(epsabs, rel * fabs(tot));
    if (err <= tol || n >= limit || w->e[worst] <= 1e-15 * fabs(tot)) {
      *result = tot;
      *abserr = err;
      return (err <= tol || w->e[worst] <= 1e-15 * fabs(tot)) ? GSL_SUCCESS : GSL_EMAXITER;
    }
    const double lo = w->a[worst], hi = w->b[worst], mid = 0.5 * (lo + hi);
    w->b[worst] = mid;
    rule(f, lo, mid, &w->r[worst], &w->e[worst]);
    w->a[n] = mid;
    w->b[n] = hi;
    rule(f, mid, hi, &w->r[n], &w->e[n]);
    n++;
  }
}

/* ---- ODE: RKF45 + standard control ----------------------------------------------------------- */
struct gsl_odeiv2_step_type_s { int kind; };
static const gsl_odeiv2_step_type t_rkf45 = {45};
const gsl_odeiv2_step_type* gsl_odeiv2_step_rkf45 = &t_rkf45;
struct gsl_odeiv2_step_s { size_t dim; double* k[6]; double *ytmp, *ynew, *yerr, *dydt_out; };
struct gsl_odeiv2_control_s { double eps_abs, eps_rel, a_y, a_dydt; };
struct gsl_odeiv2_evolve_s { size_t dim; };

gsl_odeiv2_step* gsl_odeiv2_step_alloc(const gsl_odeiv2_step_type* T, size_t dim) {
  (void)T;
  gsl_odeiv2_step* s = calloc(1, sizeof(*s));
  s->dim = dim;
  for (int i = 0; i < 6; i++) s->k[i] = calloc(dim, sizeof(double));
  s->ytmp = calloc(dim, sizeof(double));
  s->ynew = calloc(dim, sizeof(double));
  s->yerr = calloc(dim, sizeof(double));
  s->dydt_out = calloc(dim, sizeof(double));
  return s;
}
void gsl_odeiv2_step_free(gsl_odeiv2_step* s) {
  if (!s) return;
  for (int i = 0; i < 6; i++) free(s->k[i]);
  free(s->ytmp); free(s->ynew); free(s->yerr); free(s->dydt_out); free(s);
}
gsl_odeiv2_control* gsl_odeiv2_control_standard_new(double eps_abs, double eps_rel, double a_y, double a_dydt) {
  gsl_odeiv2_control* c = calloc(1, sizeof(*c));
  c->eps_abs = eps_abs; c->eps_rel = eps_rel; c->a_y = a_y; c->a_dydt = a_dydt;
  return c;
}
void gsl_odeiv2_control_free(gsl_odeiv2_control* c) { free(c); }
gsl_odeiv2_evolve* gsl_odeiv2_evolve_alloc(size_t dim) {
  gsl_odeiv2_evolve* e = calloc(1, sizeof(*e));
  e->dim = dim;
  return e;
}
void gsl_odeiv2_evolve_free(gsl_odeiv2_evolve* e) { free(e); }

static int rkf45_try(gsl_odeiv2_step* s, const gsl_odeiv2_system* sys, double t, double h, const double* y) {
  static const double A[6] = {0.0, 0.25, 0.375, 12.0 / 13.0, 1.0, 0.5};
  static const double B[6][5] = {{0},
                                 {0.25},
                                 {3.0 / 32.0, 9.0 / 32.0},
                                 {1932.0 / 2197.0, -7200.0 / 2197.0, 7296.0 / 2197.0},
                                 {439.0 / 216.0, -8.0, 3680.0 / 513.0, -845.0 / 4104.0},
                                 {-8.0 / 27.0, 2.0, -3544.0 / 2565.0, 1859.0 / 4104.0, -11.0 / 40.0}};
  static const double C5[6] = {16.0 / 135.0, 0.0, 6656.0 / 12825.0, 28561.0 / 56430.0, -9.0 / 50.0, 2.0 / 55.0};
  static const double EC[6] = {1.0 / 360.0, 0.0, -128.0 / 4275.0, -2197.0 / 75240.0, 1.0 / 50.0, 2.0 / 55.0};
  const size_t n = s->dim;
  for (int st = 0; st < 6; st++) {
    for (size_t i = 0; i < n; i++) {
      double acc = 0.0;
      for (int j = 0; j < st; j++) acc += B[st][j] * s->k[j][i];
      s->ytmp[i] = y[i] + h * acc;
    }
    if (sys->function(t + A[st] * h, s->ytmp, s->k[st], sys->params) != GSL_SUCCESS) return GSL_FAILURE;
  }
  for (size_t i = 0; i < n; i++) {
    double a5 = 0.0, ae = 0.0;
    for (int j = 0; j < 6; j++) { a5 += C5[j] * s->k[j][i]; ae += EC[j] * s->k[j][i]; }
    s->ynew[i] = y[i] + h * a5;
    s->yerr[i] = h * ae;
  }
  return sys->function(t + h, s->ynew, s->dydt_out, sys->params);
}

int gsl_odeiv2_evolve_apply(gsl_odeiv2_evolve* e, gsl_odeiv2_control* c, gsl_odeiv2_step* s, const gsl_odeiv2_system* sys, double* t,
                            double t1, double* h, double y[]) {
  (void)e;
  const size_t n = s->dim;
  const double S = 0.9, ord = 5.0;
  double h0 = *h;
  const double dt = t1 - *t;
  int final_step = 0;
  if ((dt > 0 && h0 > dt) || (dt < 0 && h0 < dt)) { h0 = dt; final_step = 1; }
  for (int attempt = 0; attempt < 200; attempt++) {
    if (rkf45_try(s, sys, *t, h0, y) != GSL_SUCCESS) return GSL_FAILURE;
    double rmax = 2.2250738585072014e-308;
    for (size_t i = 0; i < n; i++) {
      const double D0 = c->eps_rel * (c->a_y * fabs(s->ynew[i]) + c->a_dydt * fabs(h0 * s->dydt_out[i])) + c->eps_abs;
      const double r = fabs(s->yerr[i]) / fabs(D0);
      if (r > rmax) rmax = r;
    }
    if (rmax > 1.1) { /* reject: shrink and retry */
      double r = S / pow(rmax, 1.0 / ord);
      if (r < 0.2) r = 0.2;
      h0 *= r;
      final_step = 0;
      continue;
    }
    memcpy(y, s->ynew, n * sizeof(double));
    *t = final_step ? t1 : *t + h0;
    double hn = h0;
    if (rmax < 0.5) {
      double r = S / pow(rmax, 1.0 / (ord + 1.0));
      if (r > 5.0) r = 5.0;
      if (r < 1.0) r = 1.0;
      hn = h0 * r;
    }
    /* no new step size is suggested after a step cut short to land on t1 (as GSL's evolve) */
    if (!final_step) *h = hn;
    return GSL_SUCCESS;
  }
  return GSL_FAILURE;
}

/* ---- root bracketing --------------------------------------------------------------------------- */
struct gsl_root_fsolver_type_s { int kind; };
static const gsl_root_fsolver_type t_brent = {0};
const gsl_root_fsolver_type* gsl_root_fsolver_brent = &t_brent;
struct gsl_root_fsolver_s { gsl_function* f; double lo, hi, flo, fhi, root; int flip; };
gsl_root_fsolver* gsl_root_fsolver_alloc(const gsl_root_fsolver_type* T) { (void)T; return calloc(1, sizeof(gsl_root_fsolver)); }
void gsl_root_fsolver_free(gsl_root_fsolver* s) { free(s); }
int gsl_root_fsolver_set(gsl_root_fsolver* s, gsl_function* f, double lo, double hi) {
  s->f = f;
  s->lo = lo;
  s->hi = hi;
  s->flo = GSL_FN_EVAL(f, lo);
  s->fhi = GSL_FN_EVAL(f, hi);
  s->root = 0.5 * (lo + hi);
  s->flip = 0;
  return ((s->flo < 0) == (s->fhi < 0) && s->flo != 0 && s->fhi != 0) ? GSL_FAILURE : GSL_SUCCESS;
}
int gsl_root_fsolver_iterate(gsl_root_fsolver* s) {
  /* alternate a secant (regula falsi) step with a bisection step: keeps the bracket, converges
     super-linearly on smooth functions and never slower than bisection over two iterations */
  double x;
  if (s->flo == 0.0) { s->root = s->hi = s->lo; return GSL_SUCCESS; }
  if (s->fhi == 0.0) { s->root = s->lo = s->hi; return GSL_SUCCESS; }
  if (s->flip) x = 0.5 * (s->lo + s->hi);
  else {
    x = s->lo - s->flo * (s->hi - s->lo) / (s->fhi - s->flo);
    if (!(x > s->lo && x < s->hi)) x = 0.5 * (s->lo + s->hi);
  }
  s->flip = !s->flip;
  const double fx = GSL_FN_EVAL(s->f, x);
  if ((fx < 0) == (s->flo < 0) && fx != 0.0) { s->lo = x; s->flo = fx; } else { s->hi = x; s->fhi = fx; }
  s->root = x;
  return GSL_SUCCESS;
}
double gsl_root_fsolver_root(const gsl_root_fsolver* s) { return s->root; }
double gsl_root_fsolver_x_lower(const gsl_root_fsolver* s) { return s->lo; }
double gsl_root_fsolver_x_upper(const gsl_root_fsolver* s) { return s->hi; }
int gsl_root_test_interval(double lo, double hi, double epsabs, double epsrel) {
  const double al = fabs(lo), ah = fabs(hi);
  const double mn = ((lo > 0 && hi > 0) || (lo < 0 && hi < 0)) ? (al < ah ? al : ah) : 0.0;
  return fabs(hi - lo) < epsabs + epsrel * mn ? GSL_SUCCESS : GSL_CONTINUE;
}
