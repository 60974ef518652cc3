/* TEST INFRASTRUCTURE (oracle/_ref): glue that lets the reference's OWN collapse-time code run
 * in this container, so that oracle/pinocchio_oracle.py and the CUDA path can be checked
 * against the reference's compiled arithmetic and not only against a restatement.
 *
 * Linked with /root/reference/src/{collapse_times,variables,fmax,fmax-pfft,LPT,GenIC}.c, all
 * compiled verbatim from where they lie (oracle/Makefile; never copied into this repo).  What
 * those translation units need from the rest of PINOCCHIO and from MPI/GSL is provided
 * here, for one rank:
 *   - MPI_Wtime / MPI_Reduce / MPI_Bcast for a single task,
 *   - InverseGrowingMode (src/cosmo.c:1822-1832): 1/10^s(log10 D) - 1 with s the natural cubic
 *     spline of gsl_interp_cspline (GSL 2.7 interpolation/cspline.c) and the linear
 *     extrapolation of my_spline_eval (src/cosmo.c:2016-2027), knots supplied by the caller,
 *   - GrowingMode* returning caller-supplied constants (scale-independent growth),
 *   - traps for what only ELL_SNG / TABULATED_CT call (OmegaMatter, gsl_odeiv2_*),
 *   - the carving of the host arrays that src/allocations.c:335-394 does inside main_memory.
 * With src/fmax.c, src/fmax-pfft.c and src/LPT.c compiled next to them (oracle/Makefile) and the
 * one-task PFFT of oracle/ref_fft.c, the reference's own compute_fmax() runs end to end.
 * Nothing in the product links or loads this file.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <omp.h>

#include "pinocchio.h"
#include <gsl/gsl_odeiv2.h>
#include <gsl/gsl_rng.h>

/* ---- one-task MPI ---------------------------------------------------------------------- */
double MPI_Wtime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
static size_t mpi_size(MPI_Datatype t) {
  switch (t) {
    case MPI_DOUBLE: return 8;
    case MPI_FLOAT: case MPI_INT: case MPI_UNSIGNED: return 4;
    case MPI_UNSIGNED_LONG_LONG: return 8;
    default: return 1;
  }
}
int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) {
  (void)op; (void)root; (void)c;
  memcpy(r, s, n * mpi_size(t));
  return MPI_SUCCESS;
}
int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }
int MPI_Comm_free(MPI_Comm* c) { (void)c; return MPI_SUCCESS; }
int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) { return MPI_Reduce(s, r, n, t, op, 0, c); }
int MPI_Allgather(const void* s, int n, MPI_Datatype t, void* r, int rn, MPI_Datatype rt, MPI_Comm c) {
  (void)rn; (void)rt; (void)c;
  memcpy(r, s, n * mpi_size(t));
  return MPI_SUCCESS;
}
int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c) {
  (void)b; (void)n; (void)t; (void)root; (void)c;
  return MPI_SUCCESS;
}

/* ---- natural cubic spline, GSL cspline semantics ------------------------------------------ */
static int sp_n = 0;
static double *sp_x, *sp_y, *sp_c;

int ref_set_invgrow(int n, const double* x, const double* y) {
  free(sp_x); free(sp_y); free(sp_c);
  sp_n = n;
  sp_x = malloc(n * sizeof(double));
  sp_y = malloc(n * sizeof(double));
  sp_c = calloc(n, sizeof(double));
  memcpy(sp_x, x, n * sizeof(double));
  memcpy(sp_y, y, n * sizeof(double));
  const int m = n - 2; /* interior unknowns c[1..n-2], c[0] = c[n-1] = 0 */
  if (m <= 0) return 0;
  double *diag = malloc(m * sizeof(double)), *off = malloc(m * sizeof(double)), *g = malloc(m * sizeof(double));
  for (int i = 0; i < m; i++) {
    const double h_i = x[i + 1] - x[i], h_ip1 = x[i + 2] - x[i + 1];
    off[i] = h_ip1;
    diag[i] = 2.0 * (h_ip1 + h_i);
    g[i] = 3.0 * ((y[i + 2] - y[i + 1]) / h_ip1 - (y[i + 1] - y[i]) / h_i);
  }
  /* symmetric tridiagonal solve (Thomas algorithm) */
  for (int i = 1; i < m; i++) {
    const double w = off[i - 1] / diag[i - 1];
    diag[i] -= w * off[i - 1];
    g[i] -= w * g[i - 1];
  }
  sp_c[m] = g[m - 1] / diag[m - 1];
  for (int i = m - 2; i >= 0; i--) sp_c[i + 1] = (g[i] - off[i] * sp_c[i + 2]) / diag[i];
  free(diag); free(off); free(g);
  return 0;
}

static double spline_eval(double v) {
  const int n = sp_n;
  if (v < sp_x[0]) return sp_y[0] + (v - sp_x[0]) * (sp_y[1] - sp_y[0]) / (sp_x[1] - sp_x[0]);
  if (v > sp_x[n - 1]) return sp_y[n - 1] + (v - sp_x[n - 1]) * (sp_y[n - 1] - sp_y[n - 2]) / (sp_x[n - 1] - sp_x[n - 2]);
  int lo = 0, hi = n - 1; /* gsl_interp_bsearch: x[lo] <= v < x[lo+1], last interval closed */
  while (hi > lo + 1) {
    const int mid = (hi + lo) / 2;
    if (sp_x[mid] > v) hi = mid; else lo = mid;
  }
  const double dx = sp_x[lo + 1] - sp_x[lo], dy = sp_y[lo + 1] - sp_y[lo], d = v - sp_x[lo];
  const double b = dy / dx - dx * (sp_c[lo + 1] + 2.0 * sp_c[lo]) / 3.0;
  const double dd = (sp_c[lo + 1] - sp_c[lo]) / (3.0 * dx);
  return sp_y[lo] + d * (b + d * (sp_c[lo] + d * dd));
}

double InverseGrowingMode(double D, int ismooth) {
  (void)ismooth;
  return 1. / pow(10., spline_eval(log10(D))) - 1.;
}

/* ---- never reached with -DELL_CLASSIC ----------------------------------------------------- */
static void trap(const char* what) {
  fprintf(stderr, "oracle/_ref: %s reached; only the ELL_CLASSIC path is provided\n", what);
  abort();
}
/* growth factors at the segment redshift, supplied by the caller (scale-independent runs: the
 * reference evaluates one spline of z and ignores k, src/cosmo.c:1728-1819).  Sign conventions
 * are those of the reference's return values (GrowingMode_3LPT_1 is negative, :1810). */
static double ref_growth[4] = {1.0, 1.0, 1.0, 1.0};
/* -DSCALE_DEPENDENT stand-in: src/cosmo.c cannot be compiled here (GSL integrators), so the k
 * interpolation of InterpolateGrowth (src/cosmo.c:1728-1757) is restated on caller-supplied
 * tables tab[o*nk + j] = value of the j-th k-bin spline of order o+1 at the segment redshift.
 * The k loop that calls it -- which k is passed, which modes are scaled -- is the reference's own
 * compute_derivative (src/fmax-pfft.c:306-397). */
static int gk_n = 0;
static double gk_logkmin = -3.0, gk_dlogk = 0.5, *gk_tab = NULL;
int ref_set_growth_tables(int nk, double logkmin, double dlogk, const double* tab) {
  free(gk_tab);
  gk_tab = NULL;
  gk_n = nk;
  if (nk <= 0) return 0;
  gk_logkmin = logkmin;
  gk_dlogk = dlogk;
  gk_tab = malloc((size_t)4 * nk * sizeof(double));
  memcpy(gk_tab, tab, (size_t)4 * nk * sizeof(double));
  return 0;
}
static double interpolate_growth(double k, int o) {
  const double* t = gk_tab + (size_t)o * gk_n;
  const double kmin = pow(10., gk_logkmin), kmax = pow(10., gk_logkmin + (gk_n - 1) * gk_dlogk);
  if (k < kmin) return t[0];
  if (k > kmax) return t[gk_n - 1];
  double dk = (log10(k) - gk_logkmin) / gk_dlogk;
  const int kk = (int)dk;
  dk -= kk;
  return (kk + 1 < gk_n) ? dk * t[kk + 1] + (1 - dk) * t[kk] : t[kk];
}
double GrowingMode(double z, double k) { (void)z; return gk_n ? pow(10., interpolate_growth(k, 0)) : ref_growth[0]; }
double GrowingMode_2LPT(double z, double k) { (void)z; return gk_n ? pow(10., interpolate_growth(k, 1)) : ref_growth[1]; }
double GrowingMode_3LPT_1(double z, double k) { (void)z; return gk_n ? -pow(10., interpolate_growth(k, 2)) : ref_growth[2]; }
double GrowingMode_3LPT_2(double z, double k) { (void)z; return gk_n ? pow(10., interpolate_growth(k, 3)) : ref_growth[3]; }
double OmegaMatter(double z) { (void)z; trap("OmegaMatter"); return 0; }
double OmegaLambda(double z) { (void)z; trap("OmegaLambda"); return 0; }
int jac(double t, const double y[], double* dfdy, double dfdt[], void* p) { (void)t; (void)y; (void)dfdy; (void)dfdt; (void)p; trap("jac"); return 0; }
const gsl_odeiv2_step_type* gsl_odeiv2_step_rkf45 = NULL;
gsl_odeiv2_step* gsl_odeiv2_step_alloc(const gsl_odeiv2_step_type* t, size_t n) { (void)t; (void)n; trap("gsl_odeiv2"); return NULL; }
gsl_odeiv2_control* gsl_odeiv2_control_standard_new(double a, double b, double c, double d) { (void)a; (void)b; (void)c; (void)d; trap("gsl_odeiv2"); return NULL; }
gsl_odeiv2_evolve* gsl_odeiv2_evolve_alloc(size_t n) { (void)n; trap("gsl_odeiv2"); return NULL; }
int gsl_odeiv2_evolve_apply(gsl_odeiv2_evolve* e, gsl_odeiv2_control* c, gsl_odeiv2_step* s, const gsl_odeiv2_system* sys, double* t,
                            double t1, double* h, double y[]) {
  (void)e; (void)c; (void)s; (void)sys; (void)t; (void)t1; (void)h; (void)y;
  trap("gsl_odeiv2");
  return 0;
}
void gsl_odeiv2_evolve_free(gsl_odeiv2_evolve* e) { (void)e; }
void gsl_odeiv2_control_free(gsl_odeiv2_control* c) { (void)c; }
void gsl_odeiv2_step_free(gsl_odeiv2_step* s) { (void)s; }

/* ---- entry points for tests/ and bench.py ------------------------------------------------- */
double inverse_collapse_time(int, double*, double*, double*, double*, int*);
int compute_collapse_times(int);

/* the reference's per-cell solver on n Hessians (h6[c*n + i], c = xx,yy,zz,xy,xz,yz) */
int ref_inverse_collapse_time(long n, const double* h6, double* F, double* lambda) {
  int fails = 0;
  for (long i = 0; i < n; i++) {
    double t[6], l1, l2, l3;
    int fail = 0;
    for (int c = 0; c < 6; c++) t[c] = h6[c * n + i];
    F[i] = inverse_collapse_time(0, t, &l1, &l2, &l3, &fail);
    if (lambda) { lambda[3 * i] = l1; lambda[3 * i + 1] = l2; lambda[3 * i + 2] = l3; }
    fails += fail;
  }
  return fails;
}

/* the reference's compute_collapse_times(ismooth) (src/collapse_times.c:431-673) over n cells:
 * Fmax/Rmax persist between calls exactly as products[] does; returns TrueVariance[ismooth]. */
static long ref_n = 0;
static double ref_tv[64];
static grid_data ref_grid;
static double* ref_sd_ptrs[6];
static double** ref_sd_grids[1] = {ref_sd_ptrs};

int ref_collapse_begin(long n, int nthreads) {
  free(products);
  products = calloc(n, sizeof(product_data));
  if (!products) return 1;
  ref_n = n;
  MyGrids = &ref_grid;
  second_derivatives = ref_sd_grids;
  MyGrids[0].total_local_size = (unsigned int)n;
  MyGrids[0].Ntotal = (unsigned long long)n;
  ThisTask = 0;
  NTasks = 1;
  internal.nthreads_omp = nthreads;
  Smoothing.TrueVariance = ref_tv;
  return 0;
}

int ref_collapse_times(int ismooth, const double* h6, double* true_variance) {
  for (int c = 0; c < 6; c++) second_derivatives[0][c] = (double*)(h6 + (size_t)c * ref_n);
  const int rc = compute_collapse_times(ismooth);
  if (true_variance) *true_variance = Smoothing.TrueVariance[ismooth];
  return rc;
}

int ref_collapse_fetch(float* Fmax, int* Rmax) {
  for (long i = 0; i < ref_n; i++) {
    Fmax[i] = (float)products[i].Fmax;
    Rmax[i] = products[i].Rmax;
  }
  return 0;
}

int ref_sizeof_product(void) { return (int)sizeof(product_data); }

/* ---- the whole Fmax + LPT path: reference compute_fmax() on one task -------------------------- */
int set_one_grid(int);
int compute_fft_plans(void);
int compute_fmax(void);

static void* xalloc(size_t bytes) {
  void* p = calloc(bytes ? bytes : 1, 1);
  if (!p) { fprintf(stderr, "oracle/_ref: out of memory\n"); abort(); }
  return p;
}

/* N^3 box of side `box` (true Mpc); radii/variances of the smoothing ladder; growth[4] =
 * GrowingMode, _2LPT, _3LPT_1, _3LPT_2 at the segment redshift; output files (FmaxPDF) go to the
 * current directory.  Arrays persist until the next ref_setup. */
int ref_setup(int N, double box, int nsmooth, const double* radius, const double* variance, const double* growth, int nthreads) {
  static int done = 0;
  if (done) { fprintf(stderr, "oracle/_ref: ref_setup once per process (the reference's globals are never released)\n"); return 1; }
  done = 1;
  ThisTask = 0;
  NTasks = 1;
  Ngrids = 1;
  FFT_Comm = MPI_COMM_WORLD;
  memset(&params, 0, sizeof(params));
  for (int i = 0; i < 3; i++) params.GridSize[i] = N;
  params.use_transposed_fft = 0;
  strcpy(params.RunFlag, "ref");
  internal.nthreads_omp = nthreads;
  internal.nthreads_fft = nthreads;
  internal.tasks_subdivision_dim = 1;
  internal.dump_kdensity = 0;
  omp_set_num_threads(nthreads);
  for (int i = 0; i < 4; i++) ref_growth[i] = growth[i];

  MyGrids = xalloc(sizeof(grid_data));
  for (int i = 0; i < 3; i++) MyGrids[0].GSglobal[i] = N;
  MyGrids[0].Ntotal = (unsigned long long)N * N * N;
  MyGrids[0].BoxSize = box;
  if (set_one_grid(0)) return 1;

  Smoothing.Nsmooth = nsmooth;
  Smoothing.Radius = xalloc(nsmooth * sizeof(double));
  Smoothing.Variance = xalloc(nsmooth * sizeof(double));
  Smoothing.TrueVariance = xalloc(nsmooth * sizeof(double));
  memcpy(Smoothing.Radius, radius, nsmooth * sizeof(double));
  memcpy(Smoothing.Variance, variance, nsmooth * sizeof(double));
  ScaleDep.nseg = 1;
  ScaleDep.z[0] = 0.0;

  /* same carving as src/allocations.c:335-394 (sources alias the k-vectors, first derivatives
   * and density alias the second derivatives), as separate heap blocks */
  const size_t nr = MyGrids[0].total_local_size, nfft = MyGrids[0].total_local_size_fft;
  products = xalloc(nr * sizeof(product_data));
  kdensity = xalloc(sizeof(double*));
  kdensity[0] = xalloc(nfft * sizeof(double));
  kvector_2LPT = xalloc(nfft * sizeof(double));
  source_2LPT = kvector_2LPT;
  kvector_3LPT_1 = xalloc(nfft * sizeof(double));
  kvector_3LPT_2 = xalloc(nfft * sizeof(double));
  source_3LPT_1 = kvector_3LPT_1;
  source_3LPT_2 = kvector_3LPT_2;
  second_derivatives = xalloc(sizeof(double**));
  second_derivatives[0] = xalloc(6 * sizeof(double*));
  first_derivatives = xalloc(sizeof(double**));
  first_derivatives[0] = xalloc(3 * sizeof(double*));
  density = xalloc(sizeof(double*));
  for (int i = 0; i < 6; i++) second_derivatives[0][i] = xalloc(nr * sizeof(double));
  for (int i = 0; i < 3; i++) first_derivatives[0][i] = second_derivatives[0][i];
  density[0] = second_derivatives[0][0];
  rvector_fft = xalloc(sizeof(double*));
  cvector_fft = xalloc(sizeof(pfft_complex*));
  rvector_fft[0] = pfft_alloc_real(nfft);
  cvector_fft[0] = pfft_alloc_complex(nfft / 2);
  ref_n = (long)nr;
  return 0;
}

/* ---- the reference's own GenIC_large (src/GenIC.c:73-460, seed plane :482-990) ---------------
 * GSL's generators come from oracle/ref_gsl_rng.c; PowerSpectrum (src/cosmo.c, not compilable
 * here) is answered from the caller's table on the integer lattice, pk[m] = P(2 pi sqrt(m) / Box),
 * m = |n|^2 -- the same table the CUDA path is given (pinb200_set_power_table). */
int GenIC_large(int);
static const double* pk_tab = NULL;
static long pk_n = 0;
static double pk_box = 1.0;
double PowerSpectrum(double k) {
  const double x = k * pk_box / (2. * PI);
  const long m = lround(x * x);
  if (m < 0 || m >= pk_n) { fprintf(stderr, "oracle/_ref: PowerSpectrum(%g) outside the lattice table\n", k); abort(); }
  return pk_tab[m];
}
/* after ref_setup: fills kdensity[0] for RandomSeed = seed; returns GenIC_large's status */
int ref_genic(int seed, int fixed_ic, int paired_ic, const double* pk, long npk) {
  pk_tab = pk;
  pk_n = npk;
  pk_box = MyGrids[0].BoxSize;
  params.RandomSeed = seed;
  params.FixedIC = fixed_ic;
  params.PairedIC = paired_ic;
  /* defaults of set_parameters, src/initialization.c:214-219 */
  internal.verbose_level = VDIAG;
  internal.dump_seedplane = 0;
  internal.dump_kdensity = 0;
  internal.large_plane = 1;
  internal.mimic_original_seedtable = 0;
  if (!random_generator) random_generator = gsl_rng_alloc(gsl_rng_ranlxd1); /* src/initialization.c:73 */
  memset(kdensity[0], 0, (size_t)MyGrids[0].total_local_size_fft * sizeof(double));
  return GenIC_large(0);
}
int ref_get_kdensity(double* kd) {
  memcpy(kd, kdensity[0], (size_t)MyGrids[0].total_local_size_fft * sizeof(double));
  return 0;
}
/* known answers of GSL's rng/test.c for the two restated generators */
unsigned long ref_rng_nth(int kind, unsigned long seed, long n) {
  gsl_rng* r = gsl_rng_alloc(kind == 1 ? gsl_rng_ranlxd1 : gsl_rng_mt19937);
  unsigned long v = 0;
  gsl_rng_set(r, seed);
  for (long i = 0; i < n; i++) v = gsl_rng_get(r);
  gsl_rng_free(r);
  return v;
}

/* kd: [N][N][N/2+1] complex128 (the reference's non-transposed k layout) */
int ref_set_kdensity(const double* kd) {
  memcpy(kdensity[0], kd, (size_t)MyGrids[0].total_local_size_fft * sizeof(double));
  return 0;
}

/* returns the reference's compute_fmax() status; seconds = its own cputime.fmax */
int ref_compute_fmax(double* seconds, double* true_variance) {
  memset(&cputime, 0, sizeof(cputime));
  FFT_Comm = MPI_COMM_WORLD;
  if (compute_fft_plans()) return 1;
  const int rc = compute_fmax();
  if (seconds) *seconds = cputime.fmax;
  if (true_variance) memcpy(true_variance, Smoothing.TrueVariance, Smoothing.Nsmooth * sizeof(double));
  return rc;
}

/* raw products[] (56-byte records with TWO_LPT+THREE_LPT) */
int ref_fetch_products(void* out) {
  memcpy(out, products, (size_t)ref_n * sizeof(product_data));
  return 0;
}

int ref_fetch_kvector(int which, double* out) {
  const double* src = which == 0 ? kvector_2LPT : (which == 1 ? kvector_3LPT_1 : kvector_3LPT_2);
  memcpy(out, src, (size_t)MyGrids[0].total_local_size_fft * sizeof(double));
  return 0;
}

/* the reference's DumpProducts file boundary (src/fmax.c:372-506) on the arrays of ref_setup */
int dump_products(void);
int read_dumps(void);
int ref_dump_setup(const char* dumpdir, int seed) {
  snprintf(params.DumpDir, SBLENGTH, "%s", dumpdir);
  params.RandomSeed = seed;
  return 0;
}
int ref_dump_products(void) { return dump_products(); }
int ref_read_dumps(double* true_variance) {
  const int rc = read_dumps();
  if (true_variance) memcpy(true_variance, Smoothing.TrueVariance, Smoothing.Nsmooth * sizeof(double));
  return rc;
}
int ref_set_products(const void* rec, const double* true_variance) {
  memcpy(products, rec, (size_t)ref_n * sizeof(product_data));
  if (true_variance) memcpy(Smoothing.TrueVariance, true_variance, Smoothing.Nsmooth * sizeof(double));
  return 0;
}

int ref_timers(double* t) { /* fmax, deriv, fft, coll, lpt, mem_transf */
  t[0] = cputime.fmax; t[1] = cputime.deriv; t[2] = cputime.fft; t[3] = cputime.coll; t[4] = cputime.lpt; t[5] = cputime.mem_transf;
  return 0;
}
