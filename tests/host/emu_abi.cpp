// TEST ONLY: the C ABI of include/pinb200.h implemented on HOST arrays with the kernel bodies of
// kernels.cuh under the fiber block emulator (emu.cpp), single rank, grids 32 and 64.
//
// Purpose: run the linked drop-in -- the unchanged reference program + shim/fmax_b200.c -- end to
// end on a machine WITHOUT a GPU (oracle/_ref/pinocchio_emu.x, tests/test_dropin_emulated.py), so
// that the shim's run-time logic (order of calls, table marshalling, units, products[] download)
// and the hand-over to the reference's fragmentation are checked here and not only on the B200.
// The schedule below mirrors pinocchio_b200/csrc/engine.cu (pinb200_fmax, pinb200_displacements)
// call by call; the kernels are the very templates the GPU runs.  This library is never part of
// the product: libpinb200.so has no CPU path and pinocchio_b200/ does not know this file exists.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>

#include "../../include/pinb200.h"
#include "../../pinocchio_b200/csrc/product_merge.h"
#include "../../pinocchio_b200/csrc/seed_plane.h"

// entry points of emu.cpp (same shared object)
extern "C" {
int emu_xpass(int N, int dir, int rank, int nranks, const double* src, double** dsts, int dst_klayout, int pmask, int with_nyq,
              const double* gauss, double scalar, int green, int times_i, const double* tw);
int emu_xpass_gk(int N, int dir, int rank, int nranks, const double* src, double** dsts, int dst_klayout, int pmask, int with_nyq,
                 const double* gauss, double scalar, int green, int times_i, const double* tw, const double* gk, int gk_n,
                 double gk_logkmin, double gk_dlogk, double gk_sign);
int emu_ypass(int N, int dir, int rank, int nranks, double** srcs, double** dsts, double** kdsts, int dst_klayout, const int* jobs,
              int njobs, int with_nyq, const double* tw);
int emu_zpass_collapse(int N, int nranks, double** srcs, const int* kzpow, int has_nyq, const double* dc, const double* spline,
                       int nspl, int ismooth, float* fmax, int* rmax, double* sums, double** hdst, const double* tw);
int emu_zpass_out(int N, int nranks, int ncomp, double** srcs, const int* kzpow, int has_nyq, const double* dc, int mode,
                  double** rdst, float** fdst, double** hsrc, const double* weight, double* acc, const double* tw);
int emu_zpass_r2c(int N, int nranks, const double* src, double* dst, const double* tw);
int emu_sources(int N, int nranks, double** h, double* s2, double* s31, double* s32, int lpt_order);
int emu_genic(int N, int rank, int nranks, const unsigned int* seeds, const double* pk, double box, int fixed_ic, int paired_ic,
              double* kd);
int emu_zpass_collapse_tab(int N, int nranks, double** srcs, const int* kzpow, int has_nyq, const double* dc, const double* knots,
                           const double* coef, int nd, int nxy, double ampl, double bin_x, int ismooth, float* fmax, int* rmax,
                           double* sums, double** hdst, const double* tw);
int emu_scaledep_variances(int n, const double* logk, const double* a_dens, const double* a_disp, int nk, int nt, double logkmin,
                           double dlogk, const double* lg, const double* fo, int ns, const double* r_dens, const double* r_disp,
                           double* out);
int emu_ct_delta_vector(double* dv, int nd);
int emu_ct_build(int model, const double* dv, int nd, int nxy, double bin_x, double ampl, const double* spline, int nspl, double D_in,
                 const double* cosmo4, int first, int npoints, double* table);
long long emu_ct_knots_doubles(int nd);
int emu_ct_pack_knots(const double* dv, int nd, double* out);
int emu_ct_spline(const double* dv, int nd, int ncols, const double* table, double* coef);
int emu_ct_cells(int which, const double* in, long long n, const double* knots, const double* coef, int nd, int nxy, double ampl,
                 double bin_x, double* F);
int emu_collapse_cells(const double* h6, long long n, const double* spline, int nspl, double* F);
long long emu_collapsed_cells(const float* fmax, long long n, float f_last, unsigned int* idx_out);
long long emu_spline_table_doubles(int n);
int emu_pack_spline(const double* x, const double* y, int n, double* out);
}

typedef std::complex<double> cplx;

struct pinb200_ctx {
  pinb200_desc d{};
  int N = 0, M = 0, P = 0;
  std::string err;
  std::vector<cplx> tw, kdens, A[3], B[6], D[3], KV[3];
  std::vector<double> pk, radius, spline;
  std::vector<std::vector<double>> spline_r;  // SPLINE_INVGROW[ismooth] of -DSCALE_DEPENDENT builds
  int nspl = 0;
  std::vector<float> fmax, vel[12];
  std::vector<int> rmax;
  std::vector<unsigned int> seed_plane;  // pinb200_set_seed_plane (empty: the spiral table)
  std::vector<unsigned int> sorted_idx;  // pinb200_collapsed_cells keeps the ordered list for the sorted download
  bool sorted_valid = false;
  bool kdens_valid = false, hessian_valid = false, kvec_valid = false;
  // TABULATED_CT: per radius the table and its spline records (32-byte aligned storage), shared knots
  bool ct_on = false;
  int ct_nd = 0, ct_nxy = 0;
  double ct_bin_x = 0.0;
  std::vector<double> ct_ampl, ct_knots;
  std::vector<std::vector<double>> ct_tables;
  std::vector<void*> ct_coef;
  unsigned long long launches = 0;
  size_t field() const { return (size_t)N * N * P; }
  size_t ncells() const { return (size_t)N * N * N; }
};

static std::string g_err;
#define FAIL(msg) do { ctx->err = (msg); return 1; } while (0)
static double* dp(std::vector<cplx>& v) { return reinterpret_cast<double*>(v.data()); }

extern "C" const char* pinb200_last_error(const pinb200_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

extern "C" int pinb200_create(const pinb200_desc* desc, pinb200_ctx** out) {
  if (!desc || !out) { g_err = "null argument"; return 1; }
  if (desc->grid_size != 32 && desc->grid_size != 64) { g_err = "emulated ABI: grid_size must be 32 or 64"; return 1; }
  if (desc->nranks != 1) { g_err = "emulated ABI: one rank only"; return 1; }
  pinb200_ctx* c = new pinb200_ctx;
  c->d = *desc;
  c->N = desc->grid_size;
  c->M = c->N / 2;
  c->P = c->M + 8;
  c->tw.resize(c->N);
  for (int k = 0; k < c->N; k++) c->tw[k] = std::polar(1.0, 2.0 * M_PI * k / c->N);
  // garbage on purpose wherever a kernel must write before anybody reads
  const cplx junk(1e300, 0.0);
  c->kdens.assign(c->field(), cplx(0.0, 0.0));
  for (auto& f : c->A) f.assign(c->field(), junk);
  for (auto& f : c->B) f.assign(c->field(), junk);
  for (auto& f : c->D) f.assign(c->field(), junk);
  for (auto& f : c->KV) f.assign(c->field(), junk);
  *out = c;
  return 0;
}
extern "C" int pinb200_destroy(pinb200_ctx* ctx) { delete ctx; return 0; }
extern "C" int pinb200_ipc_handle(pinb200_ctx* ctx, void*) { FAIL("emulated ABI: one rank only"); }
extern "C" int pinb200_connect(pinb200_ctx* ctx, const void*) { FAIL("emulated ABI: one rank only"); }
extern "C" int pinb200_set_stream(pinb200_ctx*, void*) { return 0; }
extern "C" int pinb200_synchronize(pinb200_ctx*) { return 0; }

extern "C" int pinb200_set_power_table(pinb200_ctx* ctx, const double* pk, size_t n) {
  if (!ctx || !pk) return 1;
  if (n != (size_t)ctx->M * ctx->M + 1) FAIL("power table must have (N/2)^2 + 1 entries");
  ctx->pk.assign(pk, pk + n);
  return 0;
}
extern "C" int pinb200_set_smoothing(pinb200_ctx* ctx, int nsmooth, const double* radius) {
  if (!ctx || !radius || nsmooth < 1 || nsmooth > 64) return 1;
  ctx->radius.assign(radius, radius + nsmooth);
  return 0;
}
extern "C" int pinb200_set_invgrow_spline(pinb200_ctx* ctx, int ismooth, const double* x, const double* y, int n) {
  if (!ctx || !x || !y) return 1;
  if (ctx->nspl && ctx->nspl != n) FAIL("all inverse-growth splines must have the same number of knots");
  ctx->nspl = n;
  std::vector<double>* t = &ctx->spline;
  if (ismooth >= 0) {
    if (ismooth >= 64) FAIL("ismooth out of range");
    if ((int)ctx->spline_r.size() <= ismooth) ctx->spline_r.resize(ismooth + 1);
    t = &ctx->spline_r[ismooth];
  }
  t->assign((size_t)emu_spline_table_doubles(n), 0.0);
  return emu_pack_spline(x, y, n, t->data());
}

extern "C" int pinb200_set_seed_plane(pinb200_ctx* ctx, const unsigned int* seeds, size_t n) {
  if (!ctx || !seeds) return 1;
  if (n != (size_t)ctx->N * ctx->N) FAIL("seed plane must hold GridSize^2 entries");
  ctx->seed_plane.assign(seeds, seeds + n);
  return 0;
}
extern "C" int pinb200_genic(pinb200_ctx* ctx) {
  if (!ctx) return 1;
  if (ctx->pk.empty()) FAIL("power table not set");
  std::vector<unsigned int> seeds = ctx->seed_plane;
  if (seeds.empty()) pinb::build_seed_plane(ctx->N, ctx->d.random_seed, seeds);
  std::fill(ctx->kdens.begin(), ctx->kdens.end(), cplx(0.0, 0.0));
  if (emu_genic(ctx->N, 0, 1, seeds.data(), ctx->pk.data(), ctx->d.box_size, ctx->d.fixed_ic, ctx->d.paired_ic, dp(ctx->kdens)))
    FAIL("genic");
  ctx->kdens_valid = true;
  ctx->launches++;
  return 0;
}
extern "C" int pinb200_upload_kdensity(pinb200_ctx* ctx, const double* kd) {
  if (!ctx || !kd) return 1;
  const int N = ctx->N, M = ctx->M, P = ctx->P;
  std::fill(ctx->kdens.begin(), ctx->kdens.end(), cplx(0.0, 0.0));
  for (size_t r = 0; r < (size_t)N * N; r++) std::memcpy(&ctx->kdens[r * P], kd + 2 * r * (M + 1), sizeof(cplx) * (M + 1));
  ctx->kdens_valid = true;
  return 0;
}
static void download_c(pinb200_ctx* ctx, const std::vector<cplx>& f, double* out) {
  const int N = ctx->N, M = ctx->M, P = ctx->P;
  for (size_t r = 0; r < (size_t)N * N; r++) std::memcpy(out + 2 * r * (M + 1), &f[r * P], sizeof(cplx) * (M + 1));
}
extern "C" int pinb200_download_kdensity(pinb200_ctx* ctx, double* kd) {
  if (!ctx || !kd) return 1;
  if (!ctx->kdens_valid) FAIL("kdensity not resident");
  download_c(ctx, ctx->kdens, kd);
  return 0;
}

// ---- pass helpers: the engine's run_xpass_inv / run_ypass_inv / run_r2c on one rank ------------
static int xpass_inv(pinb200_ctx* ctx, std::vector<cplx>& src, std::vector<cplx>* dst[3], int pmask, const double* gauss, int green,
                     int times_i, double scalar, int with_nyq) {
  double* d[3] = {dst[0] ? dp(*dst[0]) : nullptr, dst[1] ? dp(*dst[1]) : nullptr, dst[2] ? dp(*dst[2]) : nullptr};
  ctx->launches++;
  return emu_xpass(ctx->N, +1, 0, 1, dp(src), d, 0, pmask, with_nyq, gauss, scalar, green, times_i, dp(ctx->tw));
}
static int ypass_inv(pinb200_ctx* ctx, std::vector<cplx>* src[3], std::vector<cplx>* dst[6], const int* jobs, int njobs, int with_nyq) {
  double *s[3], *d[6];
  for (int i = 0; i < 3; i++) s[i] = src[i] ? dp(*src[i]) : nullptr;
  for (int i = 0; i < 6; i++) d[i] = dst[i] ? dp(*dst[i]) : nullptr;
  ctx->launches++;
  return emu_ypass(ctx->N, +1, 0, 1, s, d, nullptr, 0, jobs, njobs, with_nyq, dp(ctx->tw));
}
static int r2c(pinb200_ctx* ctx, std::vector<cplx>& src, std::vector<cplx>& kdst) {
  if (emu_zpass_r2c(ctx->N, 1, dp(src), dp(src), dp(ctx->tw))) return 1;
  double* s[3] = {dp(src), nullptr, nullptr};
  double* k[1] = {dp(kdst)};
  const int job[3] = {0, 0, 0};
  if (emu_ypass(ctx->N, -1, 0, 1, s, nullptr, k, 1, job, 1, 1, dp(ctx->tw))) return 1;
  double* d[3] = {dp(kdst), nullptr, nullptr};
  ctx->launches += 3;
  return emu_xpass(ctx->N, -1, 0, 1, dp(kdst), d, 1, 1, 1, nullptr, 1.0, 0, 0, dp(ctx->tw));
}

static const int kHessJobs[18] = {2, 0, 0, 0, 2, 1, 0, 0, 2, 1, 1, 3, 1, 0, 4, 0, 1, 5};
static const int kHessKzPow[6] = {0, 0, 2, 0, 1, 1};

static int hessian_xy(pinb200_ctx* ctx, double rs, double& dc) {
  const int M = ctx->M;
  const double knorm = 2.0 * M_PI / ctx->N, norm = 1.0 / ((double)ctx->N * ctx->N * ctx->N);
  std::vector<double> gauss(M + 1);
  for (int n = 0; n <= M; n++) { const double k = knorm * n; gauss[n] = exp(-0.5 * (k * k) * rs * rs); }
  dc = norm * ctx->kdens[0].real();
  std::vector<cplx>* xd[3] = {&ctx->A[0], &ctx->A[1], &ctx->A[2]};
  if (xpass_inv(ctx, ctx->kdens, xd, 7, gauss.data(), 1, 0, norm, 0)) return 1;
  std::vector<cplx>* ys[3] = {&ctx->A[0], &ctx->A[1], &ctx->A[2]};
  std::vector<cplx>* yd[6] = {&ctx->B[0], &ctx->B[1], &ctx->B[2], &ctx->B[3], &ctx->B[4], &ctx->B[5]};
  return ypass_inv(ctx, ys, yd, kHessJobs, 6, 0);
}

// ---- TABULATED_CT: the engine's pinb200_set_collapse_tables on host arrays -------------------------
extern "C" int pinb200_scaledep_variances(const pinb200_sdgm_desc* d, double* out) {
  if (!d || !out || !d->logk || !d->a_dens || !d->a_disp || !d->log10_growth || !d->fomega || !d->radius_dens || !d->radius_disp) {
    g_err = "pinb200_scaledep_variances: null argument";
    return 1;
  }
  if (d->nnodes < 1 || d->nkbins < 1 || d->ntimes < 1 || d->nsmooth < 1 || d->nsmooth > 64 || !(d->dlogk > 0.0)) {
    g_err = "pinb200_scaledep_variances: nnodes, nkbins, ntimes >= 1, 1 <= nsmooth <= 64, dlogk > 0 required";
    return 1;
  }
  return emu_scaledep_variances(d->nnodes, d->logk, d->a_dens, d->a_disp, d->nkbins, d->ntimes, d->logkmin, d->dlogk, d->log10_growth,
                                d->fomega, d->nsmooth, d->radius_dens, d->radius_disp, out);
}
extern "C" int pinb200_ct_delta_vector(double* dv, int nd) { return (dv && nd >= 4 && nd <= 128) ? emu_ct_delta_vector(dv, nd) : 1; }
extern "C" int pinb200_set_collapse_tables(pinb200_ctx* ctx, const pinb200_ct_desc* desc, const double* variance, const double* d_in,
                                           const double* tables) {
  if (!ctx) return 1;
  for (void* p : ctx->ct_coef) free(p);
  ctx->ct_coef.clear();
  ctx->ct_tables.clear();
  ctx->ct_on = false;
  if (!desc) return 0;
  if (ctx->radius.empty()) FAIL("smoothing ladder not set (pinb200_set_smoothing)");
  if (!variance) FAIL("variance[] missing");
  if (desc->model != PINB200_CT_CLASSIC && desc->model != PINB200_CT_SNG && desc->model != PINB200_CT_SNG_FR)
    FAIL("model must be PINB200_CT_CLASSIC, PINB200_CT_SNG or PINB200_CT_SNG_FR");
  if (desc->model == PINB200_CT_CLASSIC && !tables && ctx->spline.empty()) FAIL("inverse-growth spline not set (pinb200_set_invgrow_spline)");
  if (desc->model != PINB200_CT_CLASSIC && !tables && !d_in) FAIL("d_in[] missing (ELL_SNG)");
  if (desc->model == PINB200_CT_SNG_FR && !tables && (!(desc->fr0 > 0.0) || !(desc->h_over_c > 0.0) || !desc->fr_size))
    FAIL("fr0, h_over_c and fr_size[] must be set (MOD_GRAV_FR)");
  const int ns = (int)ctx->radius.size(), nd = desc->nbins_d, nxy = desc->nbins_xy;
  const size_t ncols = (size_t)nxy * nxy, npoints = ncols * nd;
  std::vector<double> dv(nd);
  if (desc->delta_vector) dv.assign(desc->delta_vector, desc->delta_vector + nd);
  else emu_ct_delta_vector(dv.data(), nd);
  for (int i = 1; i < nd; i++)
    if (!(dv[i] > dv[i - 1])) FAIL("delta_vector must be strictly increasing");
  ctx->ct_nd = nd;
  ctx->ct_nxy = nxy;
  ctx->ct_bin_x = desc->range_x / (double)nxy;
  ctx->ct_ampl.resize(ns);
  ctx->ct_knots.assign((size_t)emu_ct_knots_doubles(nd), 0.0);
  emu_ct_pack_knots(dv.data(), nd, ctx->ct_knots.data());
  ctx->ct_tables.resize(ns);
  double cosmo4[7] = {desc->omega0, desc->omega_lambda, desc->omega_rad, desc->omega_k, 0.0, 0.0, 0.0};
  for (int is = 0; is < ns; is++) {
    if (desc->model == PINB200_CT_SNG_FR && !tables) {
      cosmo4[4] = desc->fr0;
      cosmo4[5] = desc->h_over_c;
      cosmo4[6] = desc->fr_size[is];
    }
    ctx->ct_ampl[is] = sqrt(variance[is]);
    ctx->ct_tables[is].assign(npoints, 0.0);
    if (tables) {
      memcpy(ctx->ct_tables[is].data(), tables + (size_t)is * npoints, npoints * sizeof(double));
    } else {
      const std::vector<double>& spl = (is < (int)ctx->spline_r.size() && !ctx->spline_r[is].empty()) ? ctx->spline_r[is] : ctx->spline;
      emu_ct_build(desc->model, dv.data(), nd, nxy, ctx->ct_bin_x, ctx->ct_ampl[is], spl.empty() ? nullptr : spl.data(), ctx->nspl,
                   d_in ? d_in[is] : 0.0, cosmo4, 0, (int)npoints, ctx->ct_tables[is].data());
      ctx->launches++;
    }
    void* coef = aligned_alloc(32, ncols * (size_t)(nd + 2) * 32);
    ctx->ct_coef.push_back(coef);
    emu_ct_spline(dv.data(), nd, (int)ncols, ctx->ct_tables[is].data(), static_cast<double*>(coef));
    ctx->launches++;
  }
  ctx->ct_on = true;
  return 0;
}
extern "C" int pinb200_download_collapse_table(pinb200_ctx* ctx, int ismooth, double* table) {
  if (!ctx || !table) return 1;
  if (!ctx->ct_on) FAIL("collapse tables not set (pinb200_set_collapse_tables)");
  if (ismooth < 0 || ismooth >= (int)ctx->ct_tables.size()) FAIL("ismooth out of range");
  memcpy(table, ctx->ct_tables[ismooth].data(), ctx->ct_tables[ismooth].size() * sizeof(double));
  return 0;
}

extern "C" int pinb200_fmax(pinb200_ctx* ctx, double* true_variance) {
  if (!ctx) return 1;
  if (!ctx->kdens_valid) FAIL("kdensity not resident (call pinb200_genic or pinb200_upload_kdensity)");
  if (ctx->radius.empty()) FAIL("smoothing ladder not set (pinb200_set_smoothing)");
  if (!ctx->ct_on && ctx->spline.empty()) FAIL("inverse-growth spline not set (pinb200_set_invgrow_spline)");
  if (ctx->ct_on && ctx->ct_tables.size() != ctx->radius.size()) FAIL("collapse tables were set for another smoothing ladder");
  const int ns = (int)ctx->radius.size();
  const double cell = ctx->d.box_size / ctx->N;
  ctx->fmax.assign(ctx->ncells(), 0.0f);
  ctx->rmax.assign(ctx->ncells(), 0);
  for (auto& v : ctx->vel) v.clear();
  ctx->kvec_valid = false;
  for (int is = 0; is < ns; is++) {
    double dc = 0.0, sums[2] = {0.0, 0.0};
    if (hessian_xy(ctx, ctx->radius[is] / cell, dc)) FAIL("hessian passes");
    double* b[6];
    for (int k = 0; k < 6; k++) b[k] = dp(ctx->B[k]);
    // spline_for() of the engine: the per-radius table when one was given, else the global one
    const std::vector<double>& spl = (is < (int)ctx->spline_r.size() && !ctx->spline_r[is].empty()) ? ctx->spline_r[is] : ctx->spline;
    if (ctx->ct_on) {
      if (emu_zpass_collapse_tab(ctx->N, 1, b, kHessKzPow, 0, &dc, ctx->ct_knots.data(), static_cast<const double*>(ctx->ct_coef[is]),
                                 ctx->ct_nd, ctx->ct_nxy, ctx->ct_ampl[is], ctx->ct_bin_x, is, ctx->fmax.data(), ctx->rmax.data(), sums,
                                 is == ns - 1 ? b : nullptr, dp(ctx->tw)))
        FAIL("collapse pass (tabulated)");
    } else if (emu_zpass_collapse(ctx->N, 1, b, kHessKzPow, 0, &dc, spl.data(), ctx->nspl, is, ctx->fmax.data(), ctx->rmax.data(),
                           sums, is == ns - 1 ? b : nullptr, dp(ctx->tw)))
      FAIL("collapse pass");
    ctx->launches++;
    if (true_variance) true_variance[is] = sums[1] / (double)ctx->ncells();
  }
  ctx->hessian_valid = true;
  return 0;
}

struct GrowthK { const double* tab = nullptr; int n = 0; double logkmin = 0, dlogk = 1, sign = 1; };
static int first_derivs_to_vel(pinb200_ctx* ctx, std::vector<cplx>& kvec, double growth, int first, int with_nyq,
                               const GrowthK* gk = nullptr) {
  const double norm = 1.0 / ((double)ctx->N * ctx->N * ctx->N);
  double dc = -norm * kvec[0].imag();  // run_dc with times_i = 1
  std::vector<cplx>* xd[3] = {&ctx->A[1], &ctx->A[0], nullptr};  // p = 0 -> A1, p = 1 -> A0
  if (gk) {
    double* d[3] = {dp(ctx->A[1]), dp(ctx->A[0]), nullptr};
    ctx->launches++;
    if (emu_xpass_gk(ctx->N, +1, 0, 1, dp(kvec), d, 0, 3, with_nyq, nullptr, norm * growth, 1, 1, dp(ctx->tw), gk->tab, gk->n,
                     gk->logkmin, gk->dlogk, gk->sign))
      return 1;
  } else if (xpass_inv(ctx, kvec, xd, 3, nullptr, 1, 1, norm * growth, with_nyq)) return 1;
  std::vector<cplx>* ys[3] = {&ctx->A[0], &ctx->A[1], nullptr};
  std::vector<cplx>* yd[6] = {&ctx->D[0], &ctx->D[1], &ctx->D[2], nullptr, nullptr, nullptr};
  static const int jobs[9] = {0, 0, 0, 1, 1, 1, 1, 0, 2};
  if (ypass_inv(ctx, ys, yd, jobs, 3, with_nyq)) return 1;
  double* s[6] = {dp(ctx->D[0]), dp(ctx->D[1]), dp(ctx->D[2]), nullptr, nullptr, nullptr};
  float* f[6] = {nullptr};
  for (int a = 0; a < 3; a++) {
    ctx->vel[first + a].assign(ctx->ncells(), 0.0f);
    f[a] = ctx->vel[first + a].data();
  }
  static const int kz[6] = {0, 0, 1, 0, 0, 0};
  ctx->launches++;
  return emu_zpass_out(ctx->N, 1, 3, s, kz, with_nyq, &dc, 1, nullptr, f, nullptr, nullptr, nullptr, dp(ctx->tw));
}

static int displacements_impl(pinb200_ctx* ctx, int compute_sources, const double growth[4], const GrowthK* gk) {
  if (!ctx->kdens_valid) FAIL("kdensity not resident");
  const int N = ctx->N, order = ctx->d.lpt_order;
  const double norm = 1.0 / ((double)N * N * N);
  if (order >= 2 && compute_sources) {
    if (!ctx->hessian_valid) FAIL("second derivatives of the R=0 radius are not in place (call pinb200_fmax first)");
    double* h[6];
    for (int k = 0; k < 6; k++) h[k] = dp(ctx->B[k]);
    if (emu_sources(N, 1, h, dp(ctx->A[0]), dp(ctx->A[1]), dp(ctx->A[2]), order)) FAIL("sources");
    ctx->launches++;
    if (r2c(ctx, ctx->A[0], ctx->KV[0])) FAIL("r2c of source_2LPT");
    if (order >= 3) {
      double dc = norm * ctx->KV[0][0].real();
      struct Grp { int pw, n; int jobs[9]; int kz[6]; int slot[3]; };
      const Grp grp[3] = {{2, 1, {0, 0, 0}, {0, 0, 0, 0, 0, 0}, {0, 0, 0}},
                          {1, 2, {0, 1, 0, 0, 0, 1}, {0, 1, 0, 0, 0, 0}, {3, 4, 0}},
                          {0, 3, {0, 2, 0, 0, 1, 1, 0, 0, 2}, {0, 1, 2, 0, 0, 0}, {1, 5, 2}}};
      for (const Grp& gr : grp) {
        std::vector<cplx>* xd[3] = {nullptr, nullptr, nullptr};
        xd[gr.pw] = &ctx->A[0];
        if (xpass_inv(ctx, ctx->KV[0], xd, 1 << gr.pw, nullptr, 1, 0, norm, 1)) FAIL("contraction x pass");
        std::vector<cplx>* ys[3] = {&ctx->A[0], nullptr, nullptr};
        std::vector<cplx>* yd[6] = {&ctx->D[0], &ctx->D[1], &ctx->D[2], nullptr, nullptr, nullptr};
        if (ypass_inv(ctx, ys, yd, gr.jobs, gr.n, 1)) FAIL("contraction y pass");
        double* s[6] = {dp(ctx->D[0]), dp(ctx->D[1]), dp(ctx->D[2]), nullptr, nullptr, nullptr};
        double* hs[6] = {nullptr};
        double w[6] = {0};
        for (int k = 0; k < gr.n; k++) {
          hs[k] = dp(ctx->B[gr.slot[k]]);
          w[k] = 2.0 * (gr.slot[k] <= 2 ? 1.0 : 2.0);
        }
        ctx->launches++;
        if (emu_zpass_out(N, 1, gr.n, s, gr.kz, 1, &dc, 2, nullptr, nullptr, hs, w, dp(ctx->A[2]), dp(ctx->tw))) FAIL("contraction z pass");
      }
      if (r2c(ctx, ctx->A[1], ctx->KV[1])) FAIL("r2c of source_3LPT_1");
      if (r2c(ctx, ctx->A[2], ctx->KV[2])) FAIL("r2c of source_3LPT_2");
    }
    ctx->hessian_valid = false;
    ctx->kvec_valid = true;
  }
  if (order >= 2 && !ctx->kvec_valid) FAIL("LPT k-vectors are not resident");
  if (order >= 2 && first_derivs_to_vel(ctx, ctx->KV[0], growth[1], 3, 1, gk ? gk + 1 : nullptr)) FAIL("Vel_2LPT");
  if (order >= 3) {
    if (first_derivs_to_vel(ctx, ctx->KV[1], growth[2], 6, 1, gk ? gk + 2 : nullptr)) FAIL("Vel_3LPT_1");
    if (first_derivs_to_vel(ctx, ctx->KV[2], growth[3], 9, 1, gk ? gk + 3 : nullptr)) FAIL("Vel_3LPT_2");
  }
  if (first_derivs_to_vel(ctx, ctx->kdens, growth[0], 0, 0, gk)) FAIL("Vel");
  return 0;
}
extern "C" int pinb200_displacements(pinb200_ctx* ctx, int compute_sources, const double growth[4]) {
  if (!ctx || !growth) return 1;
  return displacements_impl(ctx, compute_sources, growth, nullptr);
}
extern "C" int pinb200_displacements_scaledep(pinb200_ctx* ctx, int compute_sources, int nk, double logkmin, double dlogk,
                                              const double* log10_growth) {
  if (!ctx || !log10_growth || nk < 1 || !(dlogk > 0.0)) return 1;
  GrowthK gk[4];
  for (int o = 0; o < 4; o++) {
    gk[o].tab = log10_growth + (size_t)o * nk;
    gk[o].n = nk;
    gk[o].logkmin = logkmin;
    gk[o].dlogk = dlogk;
    gk[o].sign = (o == 2) ? -1.0 : 1.0;
  }
  static const double ones[4] = {1.0, 1.0, 1.0, 1.0};
  return displacements_impl(ctx, compute_sources, ones, gk);
}

extern "C" int pinb200_fmax_pdf(pinb200_ctx* ctx, unsigned long long* counts) {
  if (!ctx || !counts) return 1;
  if (ctx->fmax.empty()) FAIL("Fmax not computed");
  std::memset(counts, 0, sizeof(unsigned long long) * PINB200_NBINS);
  for (float f : ctx->fmax) {
    int x = (int)(f * 10.);
    if (x < 0) x = 0;
    if (x >= PINB200_NBINS) x = PINB200_NBINS - 1;
    counts[x]++;
  }
  return 0;
}

// the device packer (pack_kernel): record i = cell cell_begin + i, or cell gather[cell_begin + i]
static void pack_records(pinb200_ctx* ctx, const pinb200_product_layout* L, size_t cell_begin, size_t ncells, const unsigned int* gather,
                         unsigned char* out) {
  const int off_vel[4] = {L->off_Vel, L->off_Vel_2LPT, L->off_Vel_3LPT_1, L->off_Vel_3LPT_2};
  for (size_t i = 0; i < ncells; i++) {
    const size_t cell = gather ? (size_t)gather[cell_begin + i] : cell_begin + i;
    unsigned char* rec = out + i * L->stride;
    if (L->off_Rmax >= 0 && !ctx->rmax.empty()) *reinterpret_cast<int*>(rec + L->off_Rmax) = ctx->rmax[cell];
    auto put = [&](int off, float v) {
      if (L->prodfloat_bytes == 4) *reinterpret_cast<float*>(rec + off) = v;
      else *reinterpret_cast<double*>(rec + off) = (double)v;
    };
    if (L->off_Fmax >= 0 && !ctx->fmax.empty()) put(L->off_Fmax, ctx->fmax[cell]);
    for (int v = 0; v < 4; v++)
      if (off_vel[v] >= 0 && !ctx->vel[3 * v].empty())
        for (int a = 0; a < 3; a++) put(off_vel[v] + a * L->prodfloat_bytes, ctx->vel[3 * v + a][cell]);
  }
}
extern "C" int pinb200_download_products(pinb200_ctx* ctx, void* products, const pinb200_product_layout* L, size_t cell_begin,
                                         size_t ncells) {
  if (!ctx || !products || !L) return 1;
  if (ctx->fmax.empty() && ctx->vel[0].empty()) FAIL("products not computed");
  if (cell_begin + ncells > ctx->ncells()) FAIL("cell range outside the local slab");
  if (L->prodfloat_bytes != 4 && L->prodfloat_bytes != 8) FAIL("prodfloat_bytes must be 4 or 8");
  // packed records first (as the device packer writes them), then the engine's copy-or-merge rule
  std::vector<unsigned char> packed(ncells * L->stride, 0);
  pack_records(ctx, L, cell_begin, ncells, nullptr, packed.data());
  const bool has_vel[4] = {!ctx->vel[0].empty(), !ctx->vel[3].empty(), !ctx->vel[6].empty(), !ctx->vel[9].empty()};
  const std::vector<pinb::MemberRange> members = pinb::product_members(*L, !ctx->fmax.empty(), has_vel);
  if (pinb::members_cover_record(members, L->stride)) std::memcpy(products, packed.data(), packed.size());
  else pinb::merge_product_members(static_cast<unsigned char*>(products), packed.data(), L->stride, ncells, members);
  return 0;
}
// the file writers of engine.cu (records / snapshot blocks straight to a descriptor), here from the host arrays
static int emu_write_records(pinb200_ctx* ctx, int fd, const pinb200_product_layout* L, size_t cell_begin, size_t ncells) {
  std::vector<unsigned char> packed(ncells * L->stride, 0);
  pack_records(ctx, L, cell_begin, ncells, nullptr, packed.data());
  size_t done = 0;
  while (done < packed.size()) {
    const ssize_t w = write(fd, packed.data() + done, packed.size() - done);
    if (w <= 0) FAIL("write failed");
    done += (size_t)w;
  }
  return 0;
}
extern "C" int pinb200_write_products(pinb200_ctx* ctx, int fd, const pinb200_product_layout* L, size_t cell_begin, size_t ncells) {
  if (!ctx || !L || fd < 0) return 1;
  if (ctx->fmax.empty() && ctx->vel[0].empty()) FAIL("products not computed");
  if (cell_begin + ncells > ctx->ncells()) FAIL("cell range outside the local slab");
  return emu_write_records(ctx, fd, L, cell_begin, ncells);
}
extern "C" int pinb200_write_block(pinb200_ctx* ctx, int fd, int block, size_t cell_begin, size_t ncells) {
  if (!ctx || fd < 0) return 1;
  if (cell_begin + ncells > ctx->ncells()) FAIL("cell range outside the local slab");
  pinb200_product_layout L{};
  L.prodfloat_bytes = 4;
  L.off_Rmax = L.off_Fmax = L.off_Vel = L.off_Vel_2LPT = L.off_Vel_3LPT_1 = L.off_Vel_3LPT_2 = -1;
  bool have = false;
  switch (block) {
    case PINB200_BLOCK_FMAX: L.stride = 4; L.off_Fmax = 0; have = !ctx->fmax.empty(); break;
    case PINB200_BLOCK_RMAX: L.stride = 4; L.off_Rmax = 0; have = !ctx->rmax.empty(); break;
    case PINB200_BLOCK_ZEL: L.stride = 12; L.off_Vel = 0; have = !ctx->vel[0].empty(); break;
    case PINB200_BLOCK_2LPT: L.stride = 12; L.off_Vel_2LPT = 0; have = !ctx->vel[3].empty(); break;
    case PINB200_BLOCK_3LPT_1: L.stride = 12; L.off_Vel_3LPT_1 = 0; have = !ctx->vel[6].empty(); break;
    case PINB200_BLOCK_3LPT_2: L.stride = 12; L.off_Vel_3LPT_2 = 0; have = !ctx->vel[9].empty(); break;
    default: FAIL("unknown snapshot block");
  }
  if (!have) FAIL("the field of this block is not resident");
  return emu_write_records(ctx, fd, &L, cell_begin, ncells);
}
extern "C" int pinb200_download_field(pinb200_ctx* ctx, int which, void* dst) {
  if (!ctx || !dst || which < 0 || which > 13) return 1;
  if (which == 0 ? ctx->fmax.empty() : which == 1 ? ctx->rmax.empty() : ctx->vel[which - 2].empty())
    FAIL("requested field is not resident");  // as engine.cu: fields beyond lpt_order are never allocated
  if (which == 0) std::memcpy(dst, ctx->fmax.data(), ctx->fmax.size() * 4);
  else if (which == 1) std::memcpy(dst, ctx->rmax.data(), ctx->rmax.size() * 4);
  else std::memcpy(dst, ctx->vel[which - 2].data(), ctx->vel[which - 2].size() * 4);
  return 0;
}
extern "C" int pinb200_get_timers(pinb200_ctx* ctx, pinb200_timers* t) {
  if (!ctx || !t) return 1;
  std::memset(t, 0, sizeof *t);
  t->kernel_launches = ctx->launches;
  return 0;
}
extern "C" int pinb200_download_kvector(pinb200_ctx* ctx, int which, double* kvec) {
  if (!ctx || !kvec || which < 0 || which > 2) return 1;
  if (!ctx->kvec_valid) FAIL("LPT k-vectors are not resident");
  download_c(ctx, ctx->KV[which], kvec);
  return 0;
}
// one 3-D transform of a host work vector, engine.cu's schedule (forward_transform / reverse_transform of the shim:
// special mode 2 of src/pinocchio.c:136-168 writes the linear density field through them)
extern "C" int pinb200_fft_c2r(pinb200_ctx* ctx, const double* cplx_in, double* real_out) {
  if (!ctx || !cplx_in || !real_out) return 1;
  const int N = ctx->N, M = ctx->M, P = ctx->P;
  const double norm = 1.0 / ((double)N * N * N);
  ctx->hessian_valid = false;
  std::fill(ctx->A[0].begin(), ctx->A[0].end(), cplx(0.0, 0.0));
  for (size_t r = 0; r < (size_t)N * N; r++) std::memcpy(&ctx->A[0][r * P], cplx_in + 2 * r * (M + 1), sizeof(cplx) * (M + 1));
  std::vector<cplx>* xdst[3] = {&ctx->A[1], nullptr, nullptr};
  if (xpass_inv(ctx, ctx->A[0], xdst, 1, nullptr, 0, 0, norm, 1)) FAIL("x pass");
  std::vector<cplx>* ysrc[3] = {&ctx->A[1], nullptr, nullptr};
  std::vector<cplx>* ydst[6] = {&ctx->A[2], nullptr, nullptr, nullptr, nullptr, nullptr};
  const int job[3] = {0, 0, 0};
  if (ypass_inv(ctx, ysrc, ydst, job, 1, 1)) FAIL("y pass");
  double* zsrc[1] = {dp(ctx->A[2])};
  const int kz[1] = {0};
  ctx->launches++;
  if (emu_zpass_out(N, 1, 1, zsrc, kz, 1, nullptr, 0, zsrc, nullptr, nullptr, nullptr, nullptr, dp(ctx->tw))) FAIL("z pass");
  for (size_t r = 0; r < (size_t)N * N; r++) std::memcpy(real_out + r * N, dp(ctx->A[2]) + r * 2 * P, sizeof(double) * N);
  return 0;
}
extern "C" int pinb200_fft_r2c(pinb200_ctx* ctx, const double* real_in, double* cplx_out) {
  if (!ctx || !real_in || !cplx_out) return 1;
  const int N = ctx->N, P = ctx->P;
  std::fill(ctx->A[0].begin(), ctx->A[0].end(), cplx(0.0, 0.0));
  for (size_t r = 0; r < (size_t)N * N; r++) std::memcpy(dp(ctx->A[0]) + r * 2 * P, real_in + r * N, sizeof(double) * N);
  if (r2c(ctx, ctx->A[0], ctx->A[1])) FAIL("r2c");
  download_c(ctx, ctx->A[1], cplx_out);
  return 0;
}
// compute_second_derivatives(R): x/y passes, then the plain c2r z pass in place (engine: zpass_out mode 0)
extern "C" int pinb200_second_derivatives(pinb200_ctx* ctx, double radius, double* hessian_out) {
  if (!ctx) return 1;
  if (!ctx->kdens_valid) FAIL("kdensity not resident");
  double dc = 0.0;
  if (hessian_xy(ctx, radius / (ctx->d.box_size / ctx->N), dc)) FAIL("hessian passes");
  double* b[6];
  for (int k = 0; k < 6; k++) b[k] = dp(ctx->B[k]);
  ctx->launches++;
  if (emu_zpass_out(ctx->N, 1, 6, b, kHessKzPow, 0, &dc, 0, b, nullptr, nullptr, nullptr, nullptr, dp(ctx->tw))) FAIL("z pass");
  if (hessian_out) {
    const int N = ctx->N, P2 = 2 * ctx->P;
    for (int k = 0; k < 6; k++)
      for (size_t r = 0; r < (size_t)N * N; r++)
        std::memcpy(hessian_out + ((size_t)k * N * N + r) * N, reinterpret_cast<double*>(ctx->B[k].data()) + r * P2, sizeof(double) * N);
  }
  ctx->hessian_valid = (radius == 0.0);
  return 0;
}
extern "C" int pinb200_collapse_cells(pinb200_ctx* ctx, int ismooth, const double* h6, size_t n, double* F) {
  if (!ctx || !h6 || !F) return 1;
  if (ctx->ct_on) {
    if (ismooth < 0 || ismooth >= (int)ctx->ct_tables.size()) FAIL("ismooth out of range of the collapse tables");
    return emu_ct_cells(1, h6, (long long)n, ctx->ct_knots.data(), static_cast<const double*>(ctx->ct_coef[ismooth]), ctx->ct_nd, ctx->ct_nxy,
                        ctx->ct_ampl[ismooth], ctx->ct_bin_x, F);
  }
  if (ctx->spline.empty()) FAIL("inverse-growth spline not set (pinb200_set_invgrow_spline)");
  const std::vector<double>& spl = (ismooth >= 0 && ismooth < (int)ctx->spline_r.size() && !ctx->spline_r[ismooth].empty()) ? ctx->spline_r[ismooth] : ctx->spline;
  return emu_collapse_cells(h6, (long long)n, spl.data(), ctx->nspl, F);
}
// engine.cu's argument checks and hand-over rules; the passes themselves are the kernel bodies of
// sort_cells.cuh under the block emulator (emu_collapsed_cells in emu.cpp)
extern "C" int pinb200_collapsed_cells(pinb200_ctx* ctx, float f_last, unsigned int* cell_index_out, size_t capacity, size_t* count) {
  if (!ctx || !count) return 1;
  if (ctx->fmax.empty()) FAIL("Fmax not computed");
  if (!(f_last > 0.0f)) FAIL("f_last must be positive (F = 1 + z_collapse; the float keys are ordered by their bit patterns)");
  std::vector<unsigned int> idx(ctx->ncells());
  const long long n = emu_collapsed_cells(ctx->fmax.data(), (long long)ctx->ncells(), f_last, idx.data());
  ctx->launches += 2;
  *count = (size_t)n;
  if (n > 0 && cell_index_out && capacity > 0) {
    idx.resize((size_t)n);
    std::memcpy(cell_index_out, idx.data(), sizeof(unsigned int) * (capacity < (size_t)n ? capacity : (size_t)n));
    ctx->sorted_idx.swap(idx);
    ctx->sorted_valid = true;
    ctx->launches += 12;
  }
  return 0;
}
// begin / end: synchronous here (the emulated ABI has no streams)
extern "C" int pinb200_handoff_begin(pinb200_ctx* ctx, float f_last, float* fmax_out, unsigned int* cell_index_out, size_t capacity, size_t* count) {
  if (!ctx || !count) return 1;
  if (pinb200_collapsed_cells(ctx, f_last, cell_index_out, capacity, count)) return 1;
  if (fmax_out) std::memcpy(fmax_out, ctx->fmax.data(), ctx->fmax.size() * 4);
  return 0;
}
extern "C" int pinb200_handoff_end(pinb200_ctx* ctx) { return ctx ? 0 : 1; }
extern "C" int pinb200_download_products_sorted(pinb200_ctx* ctx, void* products, const pinb200_product_layout* L, size_t first, size_t n) {
  if (!ctx || !products || !L) return 1;
  if (!ctx->sorted_valid) FAIL("no ordered cell list (call pinb200_collapsed_cells with an output array first)");
  if (first + n > ctx->sorted_idx.size()) FAIL("record range outside the ordered cell list");
  if (L->prodfloat_bytes != 4 && L->prodfloat_bytes != 8) FAIL("prodfloat_bytes must be 4 or 8");
  if (n == 0) return 0;
  std::memset(products, 0, n * L->stride);
  pack_records(ctx, L, first, n, ctx->sorted_idx.data(), static_cast<unsigned char*>(products));
  ctx->launches++;
  return 0;
}
