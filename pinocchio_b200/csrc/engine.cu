// pinb200 engine: C ABI (include/pinb200.h) + host orchestration of the sm_100a kernels.
//
// Schedules (reference counterparts in brackets):
//   pinb200_fmax            [compute_fmax radii loop, src/fmax.c:66-150]
//       per radius: x pass (K2 fused, 3 outputs, scattered to the owner ranks) -> barrier ->
//                   y pass (6 outputs) -> z pass + collapse
//   pinb200_displacements   [compute_displacements, src/fmax.c:292-367; compute_LPT_displacements,
//                            src/LPT.c:32-235]
//       sources -> r2c -> 3 groups of {x,y,z-contraction} -> 2 r2c -> 4 x {x,y,z-float}
//
// Memory.  One cudaMalloc'ed, cudaIpc-exported *exchange arena* per rank holds every buffer that
// a peer writes into or reads from: the barrier flags, kdensity (K layout), the three x-pass
// destinations A[0..2] (R layout) and the three LPT k-vectors KV[0..2] (K layout).  Everything
// else (y-pass outputs, products, work fields) comes from the stream-ordered pool.  Nothing
// points into the caller's arena (SURVEY.md 8b "ownership").
#include <cuda_runtime.h>

#include <cmath>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <unistd.h>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pinb200.h"
#include "launch.h"
#include "product_merge.h"
#include "seed_plane.h"

using namespace pinb;

static thread_local std::string g_create_error;
namespace pinb { void set_create_error(const std::string& s) { g_create_error = s; } }

// pinb200_handoff_begin .. _end: buffers and streams of a selection + sort in flight
struct HandoffState {
  bool active = false, sorted = false;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_ready = nullptr;
  unsigned int *tile_counts = nullptr, *bsums = nullptr, *bs2 = nullptr, *range = nullptr, *counts = nullptr;
  unsigned int *key[2] = {nullptr, nullptr}, *idx[2] = {nullptr, nullptr};
  unsigned long long* d_total = nullptr;
  unsigned long long n = 0;
  int cur = 0;
  float count_ms = 0;
};

struct pinb200_ctx {
  pinb200_desc d{};
  Geom g{};
  int P = 1;               // ranks
  int lx_shift = 0, ly_shift = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t copy_stream = nullptr;          // D2H of packed records (stream_records)
  cudaEvent_t ev_stage[4] = {nullptr, nullptr, nullptr, nullptr};  // staging buffer b: [b] packed, [2 + b] copied out
  cudaStream_t xfer[PINB_MAXR] = {nullptr};    // copy-engine transposes of the multi-GPU sweep: one stream per field (three in use)
  cudaEvent_t ev_xfer[8 + PINB_MAXR] = {nullptr};  // [0] x pass done, [1] first barrier passed, [4] all transposes landed, [8 + f] stream f done
  double2* stage_extra[3] = {nullptr, nullptr, nullptr};  // x-pass staging when the arena has no k-vector slots (lpt_order < 3)
  unsigned char* pinned = nullptr;             // two pinned host buffers of the file writers
  size_t pinned_bytes = 0;
  std::string err;
  size_t field_elems = 0;  // double2 elements of one field (lx*N*P == N*ly*P)
  size_t ncells = 0;       // lx*N*N

  // exchange arena
  unsigned char* arena = nullptr;
  size_t arena_bytes = 0;
  unsigned char* peer_arena[PINB_MAXR] = {nullptr};
  bool connected = false;
  size_t off_flags = 0, off_kdens = 0, off_A[3] = {0, 0, 0}, off_KV[3] = {0, 0, 0}, off_A2[3] = {0, 0, 0};
  unsigned long long epoch = 0;
  double barrier_timeout_s = 600.0;
  int* d_error = nullptr;

  // tables
  double2* tw = nullptr;
  double* gauss = nullptr;
  double* dc = nullptr;
  double* growthk = nullptr;  // [4][nk] log10 growth tables of a scale-dependent displacement call
  double* sums = nullptr;  // [2*64]
  unsigned int* seeds = nullptr;
  double* pk = nullptr;
  std::vector<double> radius;
  std::vector<std::vector<double>> spl_host;  // index 0: global spline, 1+i: per radius
  double* spl_dev = nullptr;                  // (1+nsmooth) packed tables (spline_pack.h)
  int nspl = 0;
  bool spl_dirty = true;
  // TABULATED_CT (collapse_table.cuh): per radius the table, its spline records; shared knots
  bool ct_on = false;
  int ct_nd = 0, ct_nxy = 0, ct_ns = 0;
  double ct_bin_x = 0.0;
  std::vector<double> ct_ampl;
  double* ct_tables = nullptr;  // [ns][nxy*nxy][nd]
  CTRec* ct_coef = nullptr;     // [ns][nxy*nxy][nd + 2]
  double* ct_knots = nullptr;   // packed knots + look-up (ct_pack_knots), then the plain nd knots

  // fields
  double2* kdens = nullptr;                       // arena
  double2* A[3] = {nullptr, nullptr, nullptr};    // arena: x-pass destinations; LPT sources
  double2* KV[3] = {nullptr, nullptr, nullptr};   // arena: kvector_2LPT, kvector_3LPT_1, kvector_3LPT_2
  double2* B[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // y-pass outputs; Hessian of the last radius
  double2* D[3] = {nullptr, nullptr, nullptr};    // arena: y-pass outputs of the displacement stage; during the multi-GPU sweep the
                                                  // second set of x-pass destinations (the transposes of radius r+1 land there while
                                                  // the y and z passes of radius r still read A)
  bool kdens_valid = false, hessian_valid = false, kvec_valid = false;
  bool kdens_has_nyq = false;  // uploaded fields may carry power on the kz = N/2 plane; GenIC never fills it (src/GenIC.c:280)
  float* fmax = nullptr;
  int* rmax = nullptr;
  float* vel[12] = {nullptr};
  bool d_in_arena = false;  // D[] are arena slots (never freed) or pool allocations of the displacement stage
  bool vel_in_B = false;  // the displacement fields live in the (dead) Hessian buffers B: two float fields per double2 field
  unsigned int* sorted_idx = nullptr;  // pinb200_collapsed_cells: cell indices in order of descending Fmax
  size_t sorted_n = 0;

  HandoffState ho;
  pinb200_timers tm{};
  unsigned long long launches = 0;
  cudaEvent_t ev[3 * 64 + 8] = {nullptr};
  cudaEvent_t ev_xt[2 * 64] = {nullptr};  // timing of the transposes of radius r: [2r] first barrier passed, [2r + 1] copies done
  int nxt = 0;
  cudaEvent_t ev_mid[64] = {nullptr};  // pipelined multi-GPU sweep: between the y pass of a radius and the x pass of the next
  cudaEvent_t ev_dx[8] = {nullptr};   // x passes of the four first-derivative calls of the displacement stage
  int ndx = 0;
};

static int handoff_end_impl(pinb200_ctx* ctx);
static int release_vel(pinb200_ctx* ctx);

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      char buf__[512];                                                                        \
      snprintf(buf__, sizeof buf__, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      ctx->err = buf__;                                                                       \
      return 1;                                                                               \
    }                                                                                         \
  } while (0)
#define LAUNCH(call) do { CK(call); ctx->launches++; } while (0)
#define FAIL(msg) do { ctx->err = (msg); return 1; } while (0)
#define TRY(x) do { if (x) return 1; } while (0)

template <class T> static int dev_alloc(pinb200_ctx* ctx, T** p, size_t n) {
  if (*p) return 0;
  CK(cudaMallocAsync((void**)p, n * sizeof(T), ctx->stream));
  return 0;
}
template <class T> static int dev_free(pinb200_ctx* ctx, T** p) {
  if (!*p) return 0;
  CK(cudaFreeAsync(*p, ctx->stream));
  *p = nullptr;
  return 0;
}

static int ilog2(int v) { int s = 0; while ((1 << s) < v) s++; return s; }

// The twelve float displacement fields are as large as the six double2 Hessian fields they are computed from, and the
// Hessian is dead by the time they are written: they are placed there (vel[2k], vel[2k+1] in B[k]) instead of in
// another 52 GB of pool memory (1024^3) -- with the hand-off buffers on top of both the stream-ordered pool had to give
// memory back to the driver and re-map it in every step (r02: 0.7 s of host-side stalls per e2e step).
static int release_vel(pinb200_ctx* ctx) {
  if (ctx->vel_in_B) {
    for (auto& v : ctx->vel) v = nullptr;
    ctx->vel_in_B = false;
    return 0;
  }
  for (auto& v : ctx->vel) {
    if (!v) continue;
    CK(cudaFreeAsync(v, ctx->stream));
    v = nullptr;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
extern "C" const char* pinb200_last_error(const pinb200_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" int pinb200_create(const pinb200_desc* desc, pinb200_ctx** out) {
  if (!desc || !out) { g_create_error = "null argument"; return 1; }
  *out = nullptr;
  if (!grid_supported(desc->grid_size)) { g_create_error = "grid_size must be a power of two in [32, 2048]"; return 1; }
  const int P = desc->nranks;
  if (P < 1 || P > PINB_MAXR || (P & (P - 1)) || desc->rank < 0 || desc->rank >= P) {
    g_create_error = "nranks must be 1, 2, 4 or 8 and 0 <= rank < nranks";
    return 1;
  }
  if (desc->grid_size / P < 8) { g_create_error = "slab thinner than 8 planes"; return 1; }
  if (desc->lpt_order < 1 || desc->lpt_order > 3) { g_create_error = "lpt_order must be 1, 2 or 3"; return 1; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no usable CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e);
    return 1;
  }
  if ((e = cudaSetDevice(desc->device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return 1; }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, desc->device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return 1; }
  if (prop.major != 10) {
    g_create_error = "device is not sm_100 (this library ships sm_100a code only)";
    return 1;
  }
  pinb200_ctx* ctx = new pinb200_ctx;
  ctx->d = *desc;
  ctx->P = P;
  Geom& g = ctx->g;
  g.N = desc->grid_size;
  g.M = g.N / 2;
  g.P = g.M + 8;
  g.lx = g.N / P;
  g.ly = g.N / P;
  g.x0 = desc->rank * g.lx;
  g.y0 = desc->rank * g.ly;
  g.knorm = 2. * PINB_PI / (double)g.N;
  ctx->lx_shift = ilog2(g.lx);
  ctx->ly_shift = ilog2(g.ly);
  ctx->field_elems = (size_t)g.lx * g.N * g.P;
  ctx->ncells = (size_t)g.lx * g.N * g.N;
  auto fail = [&](cudaError_t err) {
    g_create_error = cudaGetErrorString(err);
    delete ctx;
    return 1;
  };
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e);
  ctx->own_stream = true;
  cudaMemPool_t pool;
  if ((e = cudaDeviceGetDefaultMemPool(&pool, desc->device)) != cudaSuccess) return fail(e);
  unsigned long long thr = ~0ull;
  cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  // twiddles: N-th roots of unity, computed in long double
  std::vector<double2> tw(g.N);
  for (int k = 0; k < g.N; k++) {
    const long double a = 2.0L * 3.141592653589793238462643383279502884L * (long double)k / (long double)g.N;
    tw[k] = make_double2((double)cosl(a), (double)sinl(a));
  }
  if ((e = cudaMalloc(&ctx->tw, sizeof(double2) * g.N)) != cudaSuccess) return fail(e);
  if ((e = cudaMemcpy(ctx->tw, tw.data(), sizeof(double2) * g.N, cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&ctx->gauss, sizeof(double) * (g.M + 1))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&ctx->dc, sizeof(double))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&ctx->sums, sizeof(double) * 2 * 64)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&ctx->d_error, sizeof(int))) != cudaSuccess) return fail(e);
  if ((e = cudaMemset(ctx->d_error, 0, sizeof(int))) != cudaSuccess) return fail(e);
  // exchange arena: [flags | kdens | A0 A1 A2 | D0 D1 D2 | KV0 KV1 KV2]
  const size_t fb = ctx->field_elems * sizeof(double2);
  size_t off = 4096;
  ctx->off_flags = 0;
  ctx->off_kdens = off; off += fb;
  for (int i = 0; i < 3; i++) { ctx->off_A[i] = off; off += fb; }
  // second set of x-pass destinations (pipelined sweep, from four ranks on): in the arena, doubling as the y-pass
  // outputs of the displacement stage.  With fewer ranks those three fields come from the pool when the displacement
  // stage needs them, and the pool re-uses their memory for the hand-off buffers afterwards.
  const bool arena_d = P >= 4;
  if (arena_d) for (int i = 0; i < 3; i++) { ctx->off_A2[i] = off; off += fb; }
  const int nkv = desc->lpt_order >= 3 ? 3 : (desc->lpt_order == 2 ? 1 : 0);
  for (int i = 0; i < nkv; i++) { ctx->off_KV[i] = off; off += fb; }
  ctx->arena_bytes = off;
  if ((e = cudaMalloc(&ctx->arena, ctx->arena_bytes)) != cudaSuccess) {
    g_create_error = std::string("exchange arena allocation failed: ") + cudaGetErrorString(e);
    delete ctx;
    return 1;
  }
  if ((e = cudaMemset(ctx->arena, 0, 4096)) != cudaSuccess) return fail(e);
  ctx->kdens = reinterpret_cast<double2*>(ctx->arena + ctx->off_kdens);
  for (int i = 0; i < 3; i++) ctx->A[i] = reinterpret_cast<double2*>(ctx->arena + ctx->off_A[i]);
  ctx->d_in_arena = arena_d;
  if (arena_d) for (int i = 0; i < 3; i++) ctx->D[i] = reinterpret_cast<double2*>(ctx->arena + ctx->off_A2[i]);
  for (int i = 0; i < nkv; i++) ctx->KV[i] = reinterpret_cast<double2*>(ctx->arena + ctx->off_KV[i]);
  ctx->peer_arena[desc->rank] = ctx->arena;
  ctx->connected = (P == 1);
  if (const char* t = getenv("PINB200_BARRIER_TIMEOUT_S")) {
    const double v = atof(t);
    if (v > 0.0) ctx->barrier_timeout_s = v;
  }
  for (auto& ev : ctx->ev)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return fail(e);
  for (auto& ev : ctx->ev_dx)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return fail(e);
  for (auto& ev : ctx->ev_mid)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return fail(e);
  for (auto& ev : ctx->ev_xt)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return fail(e);
  *out = ctx;
  return 0;
}

extern "C" int pinb200_ipc_handle(pinb200_ctx* ctx, void* handle64) {
  if (!ctx || !handle64) return 1;
  static_assert(sizeof(cudaIpcMemHandle_t) == PINB200_IPC_HANDLE_BYTES, "ipc handle size");
  CK(cudaSetDevice(ctx->d.device));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, ctx->arena));
  memcpy(handle64, &h, sizeof h);
  return 0;
}

extern "C" int pinb200_connect(pinb200_ctx* ctx, const void* all_handles) {
  if (!ctx || !all_handles) return 1;
  CK(cudaSetDevice(ctx->d.device));
  const unsigned char* hb = static_cast<const unsigned char*>(all_handles);
  for (int r = 0; r < ctx->P; r++) {
    if (r == ctx->d.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, hb + (size_t)r * sizeof h, sizeof h);
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->peer_arena[r] = static_cast<unsigned char*>(p);
  }
  ctx->connected = true;
  return 0;
}

extern "C" int pinb200_destroy(pinb200_ctx* ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->d.device);
  cudaStreamSynchronize(ctx->stream);
  for (int r = 0; r < ctx->P; r++)
    if (r != ctx->d.rank && ctx->peer_arena[r]) cudaIpcCloseMemHandle(ctx->peer_arena[r]);
  auto fr = [&](void* p) { if (p) cudaFree(p); };
  fr(ctx->tw); fr(ctx->gauss); fr(ctx->dc); fr(ctx->growthk); fr(ctx->sums); fr(ctx->seeds); fr(ctx->pk); fr(ctx->spl_dev);
  fr(ctx->d_error); fr(ctx->ct_tables); fr(ctx->ct_coef); fr(ctx->ct_knots);
  for (auto p : ctx->B) fr(p);
  fr(ctx->fmax); fr(ctx->rmax); fr(ctx->sorted_idx);
  if (ctx->vel_in_B) for (auto& v : ctx->vel) v = nullptr;
  for (auto p : ctx->vel) fr(p);
  if (!ctx->d_in_arena) for (auto p : ctx->D) fr(p);
  fr(ctx->arena);
  {  // hand the stream-ordered pool's cached blocks back to the driver (another context may need them)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->d.device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
  }
  for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
  for (auto& ev : ctx->ev_dx) if (ev) cudaEventDestroy(ev);
  for (auto& ev : ctx->ev_mid) if (ev) cudaEventDestroy(ev);
  for (auto& ev : ctx->ev_xt) if (ev) cudaEventDestroy(ev);
  handoff_end_impl(ctx);
  if (ctx->ho.stream) { cudaStreamDestroy(ctx->ho.stream); cudaEventDestroy(ctx->ho.ev0); cudaEventDestroy(ctx->ho.ev1); cudaEventDestroy(ctx->ho.ev_ready); }
  for (auto& ev : ctx->ev_stage) if (ev) cudaEventDestroy(ev);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  for (auto& st : ctx->xfer) if (st) cudaStreamDestroy(st);
  for (auto& ev : ctx->ev_xfer) if (ev) cudaEventDestroy(ev);
  for (auto& q : ctx->stage_extra) if (q) cudaFree(q);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return 0;
}

extern "C" int pinb200_set_stream(pinb200_ctx* ctx, void* s) {
  if (!ctx) return 1;
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream) CK(cudaStreamDestroy(ctx->stream));
  ctx->stream = (cudaStream_t)s;
  ctx->own_stream = false;
  return 0;
}

static int check_peer_error(pinb200_ctx* ctx) {
  int err = 0;
  CK(cudaMemcpy(&err, ctx->d_error, sizeof(int), cudaMemcpyDeviceToHost));
  if (err) FAIL("cross-GPU barrier timed out: a peer rank did not arrive (all ranks must make the same calls)");
  return 0;
}

extern "C" int pinb200_synchronize(pinb200_ctx* ctx) {
  if (!ctx) return 1;
  CK(cudaStreamSynchronize(ctx->stream));
  return check_peer_error(ctx);
}

extern "C" int pinb200_set_power_table(pinb200_ctx* ctx, const double* pk, size_t n) {
  if (!ctx || !pk) return 1;
  const size_t need = (size_t)ctx->g.M * ctx->g.M + 1;
  if (n < need) FAIL("power table too short: need (N/2)^2 + 1 entries");
  if (ctx->pk) { CK(cudaFree(ctx->pk)); ctx->pk = nullptr; }
  CK(cudaMalloc(&ctx->pk, need * sizeof(double)));
  CK(cudaMemcpy(ctx->pk, pk, need * sizeof(double), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int pinb200_set_smoothing(pinb200_ctx* ctx, int nsmooth, const double* radius) {
  if (!ctx || !radius) return 1;
  if (nsmooth < 1 || nsmooth > 63) FAIL("nsmooth out of range [1, 63]");
  ctx->radius.assign(radius, radius + nsmooth);
  ctx->spl_host.resize(1 + (size_t)nsmooth);
  ctx->spl_dirty = true;
  return 0;
}

extern "C" int pinb200_set_invgrow_spline(pinb200_ctx* ctx, int ismooth, const double* x, const double* y, int n) {
  if (!ctx || !x || !y) return 1;
  if (n < 3 || n > 1024) FAIL("spline size out of range [3, 1024]");
  for (int i = 1; i < n; i++)
    if (!(x[i] > x[i - 1])) FAIL("spline abscissae must be strictly increasing");
  if (ctx->nspl && ctx->nspl != n) FAIL("all inverse-growth splines must have the same number of knots");
  const size_t slot = ismooth < 0 ? 0 : (size_t)ismooth + 1;
  if (ctx->spl_host.size() <= slot) ctx->spl_host.resize(slot + 1);
  pack_spline(x, y, n, ctx->spl_host[slot]);
  ctx->nspl = n;
  ctx->spl_dirty = true;
  return 0;
}

static int upload_splines(pinb200_ctx* ctx) {
  if (!ctx->spl_dirty) return 0;
  if (ctx->spl_host.empty() || ctx->nspl == 0) FAIL("inverse-growth spline not set (pinb200_set_invgrow_spline)");
  const size_t per = spline_table_doubles(ctx->nspl), ns = ctx->spl_host.size();
  std::vector<double> all(per * ns, 0.0);
  for (size_t s = 0; s < ns; s++) {
    const std::vector<double>& src = ctx->spl_host[s].empty() ? ctx->spl_host[0] : ctx->spl_host[s];
    if (src.empty()) FAIL("global inverse-growth spline (ismooth < 0) missing");
    memcpy(all.data() + s * per, src.data(), per * sizeof(double));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->spl_dev) { CK(cudaFree(ctx->spl_dev)); ctx->spl_dev = nullptr; }
  CK(cudaMalloc(&ctx->spl_dev, all.size() * sizeof(double)));
  CK(cudaMemcpy(ctx->spl_dev, all.data(), all.size() * sizeof(double), cudaMemcpyHostToDevice));
  ctx->spl_dirty = false;
  return 0;
}

static int ct_release(pinb200_ctx* ctx) {
  ctx->ct_on = false;
  if (ctx->ct_tables || ctx->ct_coef || ctx->ct_knots) CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->ct_tables) { CK(cudaFree(ctx->ct_tables)); ctx->ct_tables = nullptr; }
  if (ctx->ct_coef) { CK(cudaFree(ctx->ct_coef)); ctx->ct_coef = nullptr; }
  if (ctx->ct_knots) { CK(cudaFree(ctx->ct_knots)); ctx->ct_knots = nullptr; }
  return 0;
}

static CTView ct_view(const pinb200_ctx* ctx, int ismooth) {
  CTView v{};
  const size_t ncols = (size_t)ctx->ct_nxy * ctx->ct_nxy;
  v.knots = ctx->ct_knots;
  v.coef = ctx->ct_coef + (size_t)ismooth * ncols * (ctx->ct_nd + 2);
  v.nd = ctx->ct_nd;
  v.nxy = ctx->ct_nxy;
  v.inv_ampl = 1.0 / ctx->ct_ampl[ismooth];
  v.inv_bin_x = 1.0 / ctx->ct_bin_x;
  return v;
}

extern "C" int pinb200_ct_delta_vector(double* delta_vector, int nbins_d) {
  if (!delta_vector || nbins_d < 4 || nbins_d > PINB_CT_MAXD) return 1;
  ct_default_delta_vector(delta_vector, nbins_d);
  return 0;
}

static const double* spline_for(pinb200_ctx* ctx, int ismooth);
extern "C" int pinb200_set_collapse_tables(pinb200_ctx* ctx, const pinb200_ct_desc* desc, const double* variance, const double* d_in,
                                           const double* tables) {
  if (!ctx) return 1;
  CK(cudaSetDevice(ctx->d.device));
  TRY(ct_release(ctx));
  if (!desc) return 0;
  if (ctx->radius.empty()) FAIL("smoothing ladder not set (pinb200_set_smoothing)");
  if (!variance) FAIL("variance[] missing");
  if (desc->model != PINB200_CT_CLASSIC && desc->model != PINB200_CT_SNG && desc->model != PINB200_CT_SNG_FR)
    FAIL("model must be PINB200_CT_CLASSIC, PINB200_CT_SNG or PINB200_CT_SNG_FR");
  if (desc->nbins_d < 4 || desc->nbins_d > PINB_CT_MAXD) FAIL("nbins_d out of range [4, 128]");
  if (desc->nbins_xy < 2 || desc->nbins_xy > 1024) FAIL("nbins_xy out of range [2, 1024]");
  if (!(desc->range_x > 0.0)) FAIL("range_x must be positive");
  const int ns = (int)ctx->radius.size(), nd = desc->nbins_d, nxy = desc->nbins_xy;
  const size_t ncols = (size_t)nxy * nxy, npoints = ncols * nd;
  if (npoints > 0x7fffffffull) FAIL("table too large");
  std::vector<double> dv(nd);
  if (desc->delta_vector) dv.assign(desc->delta_vector, desc->delta_vector + nd);
  else ct_default_delta_vector(dv.data(), nd);
  for (int i = 1; i < nd; i++)
    if (!(dv[i] > dv[i - 1])) FAIL("delta_vector must be strictly increasing");
  // the interval look-up holds one knot boundary per bin at most (one forward step in ct_interpolate
  // is then enough; more are still handled by its loop)
  for (int is = 0; is < ns; is++)
    if (!(variance[is] > 0.0)) FAIL("variance[] must be positive");
  const bool sng = desc->model == PINB200_CT_SNG || desc->model == PINB200_CT_SNG_FR;
  if (sng && !tables) {
    if (!d_in) FAIL("d_in[] missing (ELL_SNG)");
    if (!(desc->omega0 > 0.0)) FAIL("omega0 must be positive (ELL_SNG)");
    if (desc->model == PINB200_CT_SNG_FR && (!(desc->fr0 > 0.0) || !(desc->h_over_c > 0.0) || !desc->fr_size))
      FAIL("fr0, h_over_c and fr_size[] must be set (MOD_GRAV_FR)");
  }
  if (desc->model == PINB200_CT_CLASSIC && !tables) TRY(upload_splines(ctx));
  ctx->ct_nd = nd;
  ctx->ct_nxy = nxy;
  ctx->ct_ns = ns;
  ctx->ct_bin_x = desc->range_x / (double)nxy;
  ctx->ct_ampl.resize(ns);
  for (int is = 0; is < ns; is++) ctx->ct_ampl[is] = sqrt(variance[is]);
  const size_t nk = ct_knots_doubles(nd);
  std::vector<double> knots(nk + nd);
  ct_pack_knots(dv.data(), nd, knots.data());
  memcpy(knots.data() + nk, dv.data(), nd * sizeof(double));
  CK(cudaMalloc(&ctx->ct_knots, knots.size() * sizeof(double)));
  CK(cudaMalloc(&ctx->ct_tables, (size_t)ns * npoints * sizeof(double)));
  CK(cudaMalloc(&ctx->ct_coef, (size_t)ns * ncols * (nd + 2) * sizeof(CTRec)));
  CK(cudaMemcpyAsync(ctx->ct_knots, knots.data(), knots.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  const double* dv_dev = ctx->ct_knots + nk;
  CK(cudaEventRecord(ctx->ev[5], ctx->stream));
  if (tables) CK(cudaMemcpyAsync(ctx->ct_tables, tables, (size_t)ns * npoints * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  for (int is = 0; is < ns; is++) {
    double* tab = ctx->ct_tables + (size_t)is * npoints;
    if (!tables) {
      CTBuildParams b{};
      b.model = desc->model;
      b.dv = dv_dev;
      b.nd = nd;
      b.nxy = nxy;
      b.bin_x = ctx->ct_bin_x;
      b.ampl = ctx->ct_ampl[is];
      if (desc->model == PINB200_CT_CLASSIC) {
        b.spline = spline_for(ctx, is);
        b.nspl = ctx->nspl;
      } else {
        b.D_in = d_in[is];
        b.cosmo = SngCosmo{desc->omega0, desc->omega_lambda, desc->omega_rad, desc->omega_k, 0.0, 0.0, 0.0};
        if (desc->model == PINB200_CT_SNG_FR) {
          b.cosmo.fr0 = desc->fr0;
          b.cosmo.h_over_c = desc->h_over_c;
          b.cosmo.fr_size = desc->fr_size[is];
        }
      }
      b.table = tab;
      b.npoints = (int)npoints;
      LAUNCH(launch_ct_build(b, ctx->stream));
    }
    CTSplineParams sp{dv_dev, nd, (int)ncols, tab, ctx->ct_coef + (size_t)is * ncols * (nd + 2)};
    LAUNCH(launch_ct_spline(sp, ctx->stream));
  }
  CK(cudaEventRecord(ctx->ev[6], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));  // `knots`, `tables` may go out of scope
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]));
  ctx->tm.coll += ms * 1e-3;  // the reference books the table build under cputime.coll (src/fmax.c:102-118)
  ctx->ct_on = true;
  return 0;
}

extern "C" int pinb200_download_collapse_table(pinb200_ctx* ctx, int ismooth, double* table) {
  if (!ctx || !table) return 1;
  if (!ctx->ct_on) FAIL("collapse tables not set (pinb200_set_collapse_tables)");
  if (ismooth < 0 || ismooth >= ctx->ct_ns) FAIL("ismooth out of range");
  CK(cudaSetDevice(ctx->d.device));
  const size_t npoints = (size_t)ctx->ct_nxy * ctx->ct_nxy * ctx->ct_nd;
  CK(cudaMemcpyAsync(table, ctx->ct_tables + (size_t)ismooth * npoints, npoints * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

static const double* spline_for(pinb200_ctx* ctx, int ismooth) {
  size_t slot = (size_t)ismooth + 1;
  if (slot >= ctx->spl_host.size() || ctx->spl_host[slot].empty()) slot = 0;
  return ctx->spl_dev + slot * spline_table_doubles(ctx->nspl);
}

#define NEED_PEERS() do { if (!ctx->connected) FAIL("peer arenas not connected: exchange pinb200_ipc_handle() and call pinb200_connect()"); } while (0)

// cross-GPU barrier on the stream (no-op on one rank)
static int peer_barrier(pinb200_ctx* ctx, cudaStream_t stream = nullptr) {
  if (ctx->P == 1) return 0;
  if (!stream) stream = ctx->stream;
  BarrierParams b{};
  for (int r = 0; r < ctx->P; r++) b.flags[r] = reinterpret_cast<unsigned long long*>(ctx->peer_arena[r] + ctx->off_flags);
  b.rank = ctx->d.rank;
  b.nranks = ctx->P;
  b.epoch = ++ctx->epoch;
  b.error = ctx->d_error;
  b.timeout_ns = (unsigned long long)(ctx->barrier_timeout_s * 1e9);
  LAUNCH(launch_barrier(b, stream));
  return 0;
}

static PeerPtrs peers_of(pinb200_ctx* ctx, size_t arena_off) {
  PeerPtrs p{};
  for (int r = 0; r < ctx->P; r++) p.r[r] = reinterpret_cast<double2*>(ctx->peer_arena[r] + arena_off);
  return p;
}
static size_t arena_off_of(pinb200_ctx* ctx, const void* p) { return (size_t)((const unsigned char*)p - ctx->arena); }

// ------------------------------------------------------------------------------------------
extern "C" int pinb200_set_seed_plane(pinb200_ctx* ctx, const unsigned int* seeds, size_t n) {
  if (!ctx || !seeds) return 1;
  if (n != (size_t)ctx->g.N * ctx->g.N) FAIL("seed plane must hold GridSize^2 entries");
  CK(cudaSetDevice(ctx->d.device));
  if (!ctx->seeds) CK(cudaMalloc(&ctx->seeds, n * sizeof(unsigned int)));
  CK(cudaMemcpy(ctx->seeds, seeds, n * sizeof(unsigned int), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int pinb200_genic(pinb200_ctx* ctx) {
  if (!ctx) return 1;
  if (!ctx->pk) FAIL("power table not set (pinb200_set_power_table)");
  const Geom& g = ctx->g;
  CK(cudaSetDevice(ctx->d.device));
  if (!ctx->seeds) {
    std::vector<unsigned int> seeds;
    build_seed_plane(g.N, ctx->d.random_seed, seeds);
    CK(cudaMalloc(&ctx->seeds, seeds.size() * sizeof(unsigned int)));
    CK(cudaMemcpy(ctx->seeds, seeds.data(), seeds.size() * sizeof(unsigned int), cudaMemcpyHostToDevice));
  }
  CK(cudaEventRecord(ctx->ev[0], ctx->stream));
  CK(cudaMemsetAsync(ctx->kdens, 0, ctx->field_elems * sizeof(double2), ctx->stream));
  GenicParams p{};
  p.seeds = ctx->seeds;
  p.pk = ctx->pk;
  p.kd = ctx->kdens;
  p.box = ctx->d.box_size;
  p.fixed_ic = ctx->d.fixed_ic;
  p.paired_ic = ctx->d.paired_ic;
  p.g = g;
  LAUNCH(launch_genic(p, ctx->stream));
  CK(cudaEventRecord(ctx->ev[1], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
  ctx->tm.dens += ms * 1e-3;
  ctx->kdens_valid = true;
  ctx->kdens_has_nyq = false;
  return 0;
}

// host half-complex K-layout slab [N][ly][M+1] <-> device [N][ly][P]
static int upload_cplx(pinb200_ctx* ctx, const double* host, double2* dev) {
  // rows of M+1 elements on the host, pitch P on the device: one strided copy, no device temporary (r01 staged the
  // slab in another 8.6 GB and re-pitched it with a kernel)
  const Geom& g = ctx->g;
  const size_t rows = (size_t)g.N * g.ly;
  CK(cudaMemsetAsync(dev, 0, ctx->field_elems * sizeof(double2), ctx->stream));
  CK(cudaMemcpy2DAsync(dev, (size_t)g.P * sizeof(double2), host, (size_t)(g.M + 1) * sizeof(double2), (size_t)(g.M + 1) * sizeof(double2),
                       rows, cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}
static int download_cplx(pinb200_ctx* ctx, const double2* dev, double* host) {
  const Geom& g = ctx->g;
  const size_t rows = (size_t)g.N * g.ly;
  double2* tmp = nullptr;
  TRY(dev_alloc(ctx, &tmp, rows * (g.M + 1)));
  LAUNCH(launch_repitch_c(dev, tmp, rows, g.M + 1, g.P, g.M + 1, ctx->stream));
  CK(cudaMemcpyAsync(host, tmp, rows * (g.M + 1) * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  TRY(dev_free(ctx, &tmp));
  return 0;
}
static int download_real(pinb200_ctx* ctx, const double2* dev, double* host) {
  const Geom& g = ctx->g;
  const size_t rows = (size_t)g.lx * g.N;
  double* tmp = nullptr;
  TRY(dev_alloc(ctx, &tmp, rows * g.N));
  LAUNCH(launch_repitch_r(reinterpret_cast<const double*>(dev), tmp, rows, g.N, 2 * g.P, g.N, ctx->stream));
  CK(cudaMemcpyAsync(host, tmp, rows * g.N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  TRY(dev_free(ctx, &tmp));
  return 0;
}

extern "C" int pinb200_upload_kdensity(pinb200_ctx* ctx, const double* kd) {
  if (!ctx || !kd) return 1;
  CK(cudaSetDevice(ctx->d.device));
  TRY(upload_cplx(ctx, kd, ctx->kdens));
  // a field that did not come from pinb200_genic (forward_transform output, white noise) may have power at
  // kz = N/2: the Hessian and Zel'dovich passes then process the Nyquist tile as compute_derivative does
  // for idz = N/2 (src/fmax-pfft.c:306-397).  On one rank the plane is probed (a re-uploaded GenIC field
  // keeps the cheaper schedule); on several ranks the decision must be the same everywhere, so an uploaded
  // field is always treated as carrying it.
  int nyq = 1;
  if (ctx->P == 1) {
    int* flag = nullptr;
    TRY(dev_alloc(ctx, &flag, (size_t)1));
    CK(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    LAUNCH(launch_nyq_probe(ctx->kdens, (size_t)ctx->g.N * ctx->g.ly, ctx->g.P, ctx->g.M, flag, ctx->stream));
    CK(cudaMemcpyAsync(&nyq, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    TRY(dev_free(ctx, &flag));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->kdens_valid = true;
  ctx->kdens_has_nyq = (nyq != 0);
  return 0;
}

extern "C" int pinb200_download_kdensity(pinb200_ctx* ctx, double* kd) {
  if (!ctx || !kd) return 1;
  if (!ctx->kdens_valid) FAIL("kdensity not resident (call pinb200_genic or pinb200_upload_kdensity)");
  CK(cudaSetDevice(ctx->d.device));
  return download_cplx(ctx, ctx->kdens, kd);
}

// ---- pass helpers --------------------------------------------------------------------------
static int ntiles(const pinb200_ctx* ctx, int tk, bool with_nyq) { return ctx->g.M / tk + (with_nyq ? 1 : 0); }

// inverse x pass of the local K-layout field `src`; the output for power p of kx is scattered
// into the arena buffer dst[p] (R layout) of the rank owning each x.  Barriers on both sides:
// before, so that no peer still reads the destinations; after, so that readers see all stores.
struct GrowthK {        // one row of the scale-dependent growth tables (KFactor::gk*), device pointer
  const double* tab = nullptr;
  int n = 0;
  double logkmin = 0, dlogk = 1, sign = 1;
};
static int run_xpass_inv(pinb200_ctx* ctx, const double2* src, double2* const dst[3], int pmask, bool gauss, int green,
                         int times_i, double scalar, bool with_nyq, const GrowthK* gk = nullptr) {
  XPassParams p{};
  p.src = src;
  for (int i = 0; i < 3; i++)
    if (dst[i]) p.dst[i] = peers_of(ctx, arena_off_of(ctx, dst[i]));
  p.dst_klayout = 0;
  p.lx_shift = ctx->lx_shift;
  p.pmask = pmask;
  p.ntiles_z = ntiles(ctx, xpass_tk(ctx->g.N, +1), with_nyq);
  p.kf.gauss = gauss ? ctx->gauss : nullptr;
  p.kf.scalar = scalar;
  p.kf.green = green;
  p.kf.times_i = times_i;
  if (gk && gk->tab) {
    p.kf.gk = gk->tab;
    p.kf.gk_n = gk->n;
    p.kf.gk_logkmin = gk->logkmin;
    p.kf.gk_dlogk = gk->dlogk;
    p.kf.gk_sign = gk->sign;
  }
  p.g = ctx->g;
  p.tw = ctx->tw;
  TRY(peer_barrier(ctx));
  LAUNCH(launch_xpass(ctx->g.N, +1, p, ctx->g.ly, ctx->stream));
  TRY(peer_barrier(ctx));
  return 0;
}

static int run_ypass_inv(pinb200_ctx* ctx, const double2* const src[3], double2* const dst[6], const YJob* jobs, int njobs,
                         bool with_nyq) {
  YPassParams p{};
  for (int i = 0; i < 3; i++) p.src[i] = src[i];
  for (int i = 0; i < 6; i++) p.dst[i] = dst[i];
  for (int i = 0; i < njobs; i++) p.job[i] = jobs[i];
  p.njobs = njobs;
  p.dst_klayout = 0;
  p.ly_shift = ctx->ly_shift;
  p.ntiles_z = ntiles(ctx, ypass_tk(ctx->g.N), with_nyq);
  p.g = ctx->g;
  p.tw = ctx->tw;
  LAUNCH(launch_ypass(ctx->g.N, +1, p, ctx->g.lx, ctx->stream));
  return 0;
}

// forward r2c: real field in `src` (R layout, arena or pool; destroyed) -> half-complex K-layout
// field `kdst` (arena buffer).  z r2c in place, y forward scattered to the owner of each y,
// x forward in place on the K layout.
static int run_r2c(pinb200_ctx* ctx, double2* src, double2* kdst) {
  const Geom& g = ctx->g;
  ZR2CParams z{};
  z.src = src;
  z.dst = src;
  z.g = g;
  z.tw = ctx->tw;
  LAUNCH(launch_zpass_r2c(g.N, z, (size_t)g.lx * g.N, ctx->stream));
  YPassParams y{};
  y.src[0] = src;
  y.kdst = peers_of(ctx, arena_off_of(ctx, kdst));
  y.dst_klayout = 1;
  y.ly_shift = ctx->ly_shift;
  y.job[0] = YJob{0, 0, 0};
  y.njobs = 1;
  y.ntiles_z = ntiles(ctx, ypass_tk(g.N), true);
  y.g = g;
  y.tw = ctx->tw;
  TRY(peer_barrier(ctx));
  LAUNCH(launch_ypass(g.N, -1, y, g.lx, ctx->stream));
  TRY(peer_barrier(ctx));
  XPassParams p{};
  p.src = kdst;
  p.dst[0].r[0] = kdst;
  p.dst_klayout = 1;
  p.lx_shift = ctx->lx_shift;
  p.pmask = 1;
  p.ntiles_z = ntiles(ctx, xpass_tk(g.N, -1), true);
  p.kf.gauss = nullptr;
  p.kf.scalar = 1.0;
  p.kf.green = 0;
  p.kf.times_i = 0;
  p.g = g;
  p.tw = ctx->tw;
  LAUNCH(launch_xpass(g.N, -1, p, g.ly, ctx->stream));
  return 0;
}

// the k = 0 mode lives on rank 0 (x = 0, yl = 0): read it through the peer mapping
static int run_dc(pinb200_ctx* ctx, const double2* arena_field, double scale, int times_i) {
  const double2* on_rank0 = reinterpret_cast<const double2*>(ctx->peer_arena[0] + arena_off_of(ctx, arena_field));
  LAUNCH(launch_dc_scalar(on_rank0, ctx->dc, scale, times_i, ctx->stream));
  return 0;
}

static int ensure_products(pinb200_ctx* ctx) {
  TRY(dev_alloc(ctx, &ctx->fmax, ctx->ncells));
  TRY(dev_alloc(ctx, &ctx->rmax, ctx->ncells));
  return 0;
}

// Hessian passes for one radius: fills B[0..5] (half-complex, after x and y passes).
// slot order xx,yy,zz,xy,xz,yz (src/fmax.c:239)
// ---- multi-GPU sweep: the transposes leave the SMs ----------------------------------------------------------
// r01 fused the all-to-all of every inverse FFT into the x pass as peer stores: at 2048^3 on 8 GPUs that pass
// took 55 ms per radius (32 ms of it for the line FFTs themselves, the rest waiting on NVLink: SM stores to peer
// memory block the issuing warps), strictly before the 26 ms y pass and the 56 ms collapse pass.  Here the x pass
// of radius r+1 writes its three fields LOCALLY, in the K layout it reads (x-major: the part destined to rank d,
// x in [d lx, (d+1) lx), is one contiguous block), into the arena slots of the LPT k-vectors, which are idle
// during the sweep; the copy engines then move every (field, destination) block into the owner's R-layout buffer
// -- lx rows of ly P elements, one strided cudaMemcpy2DAsync each -- on side streams WHILE the SMs run the
// collapse pass of radius r.  Order: y(r) -> x(r+1) [main] ; barrier, 3 x P copies, barrier [side] || z(r) [main];
// y(r+1) waits for the side streams.  The first barrier says "every rank has finished reading A(r)", the second
// "every block of A(r+1) has landed everywhere"; all sweep barriers are issued on one side stream, in order.
// PINB200_PEER_STORES=1 selects the r01 schedule (peer stores, nothing overlapped).
static int xfer_setup(pinb200_ctx* ctx) {
  if (ctx->xfer[0]) return 0;
  // one stream per field.  (r02, 8 x B200, 2048^3: one high-priority stream per DESTINATION -- seven concurrent copies
  // per GPU -- was slower, 1792 ms per step against 1685 ms; the copy engines sustain ~390-440 GB/s per GPU either way)
  for (int k = 0; k < 3; k++) CK(cudaStreamCreateWithFlags(&ctx->xfer[k], cudaStreamNonBlocking));
  for (auto& ev : ctx->ev_xfer) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  return 0;
}
static int xpass_staging(pinb200_ctx* ctx, double2* S[3]) {
  for (int i = 0; i < 3; i++) {
    if (ctx->KV[i]) { S[i] = ctx->KV[i]; continue; }
    if (!ctx->stage_extra[i]) CK(cudaMalloc(&ctx->stage_extra[i], ctx->field_elems * sizeof(double2)));
    S[i] = ctx->stage_extra[i];
  }
  return 0;
}
// inverse x pass of kdens for one radius into local K-layout staging (no peer traffic, no barrier)
static int run_xpass_local(pinb200_ctx* ctx, double2* const S[3], double2* const own[3], double rsmooth) {
  const Geom& g = ctx->g;
  LAUNCH(launch_gauss_table(ctx->gauss, g.M, g.knorm, rsmooth, ctx->stream));
  XPassParams p{};
  p.src = ctx->kdens;
  for (int i = 0; i < 3; i++) {
    p.dst[i].r[0] = S[i];
    p.dst[i].r[1] = own[i];      // this rank's own x planes skip the staging
  }
  p.dst_klayout = 2;
  p.lx_shift = ctx->lx_shift;
  p.pmask = 0x7;
  p.ntiles_z = ntiles(ctx, xpass_tk(g.N, +1, 2), ctx->kdens_has_nyq);
  p.kf.gauss = ctx->gauss;
  p.kf.scalar = 1.0 / ((double)g.N * g.N * g.N);
  p.kf.green = 1;
  p.kf.times_i = 0;
  p.g = g;
  p.tw = ctx->tw;
  LAUNCH(launch_xpass(g.N, +1, p, g.ly, ctx->stream));
  return 0;
}
// S (K layout, local) -> A (R layout) of every rank, by the copy engines; after_event: what the copies must follow
static int transpose_dma(pinb200_ctx* ctx, double2* const S[3], const size_t dst_off[3], cudaEvent_t after_event, cudaEvent_t done_event) {
  const Geom& g = ctx->g;
  const size_t row = (size_t)g.ly * g.P * sizeof(double2);          // one x plane of the local K-layout slab
  const size_t dpitch = (size_t)g.N * g.P * sizeof(double2);        // one x plane of an R-layout slab
  CK(cudaStreamWaitEvent(ctx->xfer[0], after_event, 0));
  TRY(peer_barrier(ctx, ctx->xfer[0]));                             // nobody reads the destination buffers any more
  CK(cudaEventRecord(ctx->ev_xfer[1], ctx->xfer[0]));
  const int slot = ctx->nxt < 64 ? ctx->nxt++ : -1;
  if (slot >= 0) CK(cudaEventRecord(ctx->ev_xt[2 * slot], ctx->xfer[0]));
  // one stream per field, the destinations of a field in a row starting with the neighbour (the ranks do not all hit
  // rank 0 first); the block of this rank itself was written in place by the x pass
  for (int f = 0; f < 3; f++) {
    cudaStream_t st = ctx->xfer[f];
    if (f) CK(cudaStreamWaitEvent(st, ctx->ev_xfer[1], 0));
    for (int k = 0; k + 1 < ctx->P; k++) {
      const int d = (ctx->d.rank + 1 + k) % ctx->P;
      const unsigned char* src = reinterpret_cast<const unsigned char*>(S[f]) + (size_t)d * g.lx * row;
      unsigned char* dst = ctx->peer_arena[d] + dst_off[f] + (size_t)ctx->d.rank * row;
      CK(cudaMemcpy2DAsync(dst, dpitch, src, row, row, (size_t)g.lx, cudaMemcpyDeviceToDevice, st));
    }
    if (f) {
      CK(cudaEventRecord(ctx->ev_xfer[8 + f], st));
      CK(cudaStreamWaitEvent(ctx->xfer[0], ctx->ev_xfer[8 + f], 0));
    }
  }
  if (slot >= 0) CK(cudaEventRecord(ctx->ev_xt[2 * slot + 1], ctx->xfer[0]));
  TRY(peer_barrier(ctx, ctx->xfer[0]));                             // every block has landed on every rank
  CK(cudaEventRecord(done_event, ctx->xfer[0]));
  return 0;
}

// The k = 0 constant of delta_k (the same for every radius: the window is 1 at k = 0) is read from rank 0
// through the peer mapping.  The barrier orders the read after rank 0's GenIC / upload on ITS stream:
// those calls end with a local synchronisation only, so without it a rank > 0 could read a stale value.
static int hessian_dc(pinb200_ctx* ctx) {
  const Geom& g = ctx->g;
  TRY(peer_barrier(ctx));
  TRY(run_dc(ctx, ctx->kdens, 1.0 / ((double)g.N * g.N * g.N), 0));
  return 0;
}
static int hessian_xy(pinb200_ctx* ctx, double rsmooth, cudaEvent_t mid_event = nullptr) {
  const Geom& g = ctx->g;
  const double norm = 1.0 / ((double)g.N * g.N * g.N);
  const bool nyq = ctx->kdens_has_nyq;
  LAUNCH(launch_gauss_table(ctx->gauss, g.M, g.knorm, rsmooth, ctx->stream));
  TRY(run_xpass_inv(ctx, ctx->kdens, ctx->A, 0x7, true, 1, 0, norm, nyq));
  if (mid_event) CK(cudaEventRecord(mid_event, ctx->stream));
  static const YJob jobs[6] = {{2, 0, 0}, {0, 2, 1}, {0, 0, 2}, {1, 1, 3}, {1, 0, 4}, {0, 1, 5}};
  TRY(run_ypass_inv(ctx, ctx->A, ctx->B, jobs, 6, nyq));
  return 0;
}
static const int kHessKzPow[6] = {0, 0, 2, 0, 1, 1};

extern "C" int pinb200_fmax(pinb200_ctx* ctx, double* true_variance) {
  if (!ctx) return 1;
  if (!ctx->kdens_valid) FAIL("kdensity not resident (call pinb200_genic or pinb200_upload_kdensity)");
  if (ctx->radius.empty()) FAIL("smoothing ladder not set (pinb200_set_smoothing)");
  NEED_PEERS();
  CK(cudaSetDevice(ctx->d.device));
  if (ctx->ct_on && ctx->ct_ns != (int)ctx->radius.size()) FAIL("collapse tables were set for another smoothing ladder");
  if (!ctx->ct_on) TRY(upload_splines(ctx));
  const Geom& g = ctx->g;
  const int ns = (int)ctx->radius.size();
  const double cell = ctx->d.box_size / g.N;  // GRID.CellSize, src/fmax-pfft.c:88
  if (!ctx->d_in_arena) for (auto& w : ctx->D) TRY(dev_free(ctx, &w));
  // a new Fmax sweep re-initialises the products (src/collapse_times.c:461-492 zeroes Vel*):
  // displacement fields of an earlier call are released here and read back as zeros
  TRY(release_vel(ctx));
  TRY(handoff_end_impl(ctx));
  TRY(dev_free(ctx, &ctx->sorted_idx));
  ctx->sorted_n = 0;
  ctx->kvec_valid = false;
  TRY(ensure_products(ctx));
  for (int i = 0; i < 6; i++) TRY(dev_alloc(ctx, &ctx->B[i], ctx->field_elems));
  CK(cudaMemsetAsync(ctx->sums, 0, sizeof(double) * 2 * 64, ctx->stream));
  CK(cudaEventRecord(ctx->ev[0], ctx->stream));
  TRY(hessian_dc(ctx));
  // Worth it when the transposes fit under the collapse pass: from four ranks on.  On two ranks every GPU ships half
  // of its x-pass output (13 GB per radius at 1024^3, 15-20 ms of NVLink) against a 24 ms collapse pass whose HBM
  // traffic it also disturbs: measured 589 ms per step against 548 ms with peer stores (r02, 2 x B200).
  const char* ps_env = getenv("PINB200_PEER_STORES");
  const bool pipelined = ps_env ? (ctx->P > 1 && !atoi(ps_env)) : (ctx->P >= 4);
  double2* S[3] = {nullptr, nullptr, nullptr};
  static const YJob hess_jobs[6] = {{2, 0, 0}, {0, 2, 1}, {0, 0, 2}, {1, 1, 3}, {1, 0, 4}, {0, 1, 5}};
  ctx->nxt = 0;
  // x-pass destinations are double-buffered in the pipelined sweep: radius r is read from X[r & 1] while the
  // transposes of radius r + 1 fill X[(r + 1) & 1] (the arena slots of the displacement stage's y-pass outputs)
  double2* const* X[2] = {ctx->A, ctx->D};
  const size_t* Xoff[2] = {ctx->off_A, ctx->off_A2};
  if (pipelined) {
    TRY(xfer_setup(ctx));
    TRY(xpass_staging(ctx, S));
    // prologue: x pass and transposes of the first radius (nothing to hide them behind)
    TRY(run_xpass_local(ctx, S, X[0], ctx->radius[0] / cell));
    CK(cudaEventRecord(ctx->ev_xfer[0], ctx->stream));
    TRY(transpose_dma(ctx, S, Xoff[0], ctx->ev_xfer[0], ctx->ev_xfer[4]));
  }
  for (int is = 0; is < ns; is++) {
    const double rs = ctx->radius[is] / cell;  // Rsmooth in grid units, src/fmax.c:233
    if (!pipelined) {
      TRY(hessian_xy(ctx, rs, ctx->ev[8 + 3 * is]));
    } else {
      // X[is & 1] complete on this rank, and the staging free again
      CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_xfer[4 + (is & 1)], 0));
      CK(cudaEventRecord(ctx->ev[8 + 3 * is], ctx->stream));
      if (is + 1 < ns) {
        // x pass of the NEXT radius first: its transposes then have the y pass AND the collapse pass of this
        // radius to hide behind (83 ms against ~60 ms of copies at 2048^3 on 8 GPUs)
        TRY(run_xpass_local(ctx, S, X[(is + 1) & 1], ctx->radius[is + 1] / cell));
        CK(cudaEventRecord(ctx->ev_xfer[0], ctx->stream));
        TRY(transpose_dma(ctx, S, Xoff[(is + 1) & 1], ctx->ev_xfer[0], ctx->ev_xfer[4 + ((is + 1) & 1)]));
      }
      CK(cudaEventRecord(ctx->ev_mid[is], ctx->stream));
      TRY(run_ypass_inv(ctx, X[is & 1], ctx->B, hess_jobs, 6, ctx->kdens_has_nyq));
    }
    CK(cudaEventRecord(ctx->ev[8 + 3 * is + 1], ctx->stream));
    CollapseParams c{};
    for (int k = 0; k < 6; k++) {
      c.zs.src[k] = ctx->B[k];
      c.zs.kzpow[k] = kHessKzPow[k];
      c.hdst[k] = (is == ns - 1) ? ctx->B[k] : nullptr;  // keep the R=0 Hessian for the LPT sources
    }
    c.zs.ncomp = 6;
    c.zs.has_nyq = ctx->kdens_has_nyq ? 1 : 0;
    c.zs.dc_add = ctx->dc;
    c.g = g;
    c.tw = ctx->tw;
    if (ctx->ct_on) {
      c.ct = ct_view(ctx, is);  // TABULATED_CT: F from the table of this radius
    } else {
      c.spline = spline_for(ctx, is);
      c.nspl = ctx->nspl;
      c.spl_doubles = (int)spline_table_doubles(ctx->nspl);
    }
    c.ismooth = is;
    c.Fmax = ctx->fmax;
    c.Rmax = ctx->rmax;
    c.sums = ctx->sums + 2 * is;
    LAUNCH(launch_zpass_collapse(g.N, c, (size_t)g.lx * g.N, ctx->stream));
    CK(cudaEventRecord(ctx->ev[8 + 3 * is + 2], ctx->stream));
  }
  CK(cudaEventRecord(ctx->ev[1], ctx->stream));
  // the Hessian kept for the LPT sources is that of the LAST radius: it is the unsmoothed one only when the
  // ladder ends with R = 0 as set_smoothing guarantees (src/initialization.c:386-435); any other ladder
  // must go through pinb200_second_derivatives(ctx, 0, NULL) before pinb200_displacements(compute_sources = 1)
  ctx->hessian_valid = (ctx->radius.back() == 0.0);
  std::vector<double> sums(2 * 64);
  CK(cudaMemcpyAsync(sums.data(), ctx->sums, sizeof(double) * 2 * 64, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  TRY(check_peer_error(ctx));
  const double ntot = (double)g.N * g.N * g.N;
  // local share of Sum(delta^2)/Ntotal: the caller adds the ranks up (MPI_Reduce, src/collapse_times.c:656-662)
  if (true_variance)
    for (int is = 0; is < ns; is++) true_variance[is] = sums[2 * is + 1] / ntot;
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
  ctx->tm.fmax += ms * 1e-3;
  if (pipelined) {
    for (auto& st : ctx->xfer) if (st) CK(cudaStreamSynchronize(st));
    for (int i = 0; i < ctx->nxt; i++) {
      float t = 0;
      CK(cudaEventElapsedTime(&t, ctx->ev_xt[2 * i], ctx->ev_xt[2 * i + 1]));
      ctx->tm.xfer += t * 1e-3;
    }
  }
  for (int is = 0; is < ns; is++) {
    float x = 0, y = 0, z = 0;
    CK(cudaEventElapsedTime(&x, is == 0 ? ctx->ev[0] : ctx->ev[8 + 3 * is - 1], ctx->ev[8 + 3 * is]));
    CK(cudaEventElapsedTime(&y, ctx->ev[8 + 3 * is], ctx->ev[8 + 3 * is + 1]));
    CK(cudaEventElapsedTime(&z, ctx->ev[8 + 3 * is + 1], ctx->ev[8 + 3 * is + 2]));
    if (pipelined) {
      // x = exposed wait for the transposes (radius 0: the whole prologue) + the local x pass of the NEXT radius,
      // which runs just before this radius' y pass
      float xl = 0;
      CK(cudaEventElapsedTime(&xl, ctx->ev[8 + 3 * is], ctx->ev_mid[is]));
      x += xl;
      y -= xl;
    }
    ctx->tm.hess_x += x * 1e-3;
    ctx->tm.hess_y += y * 1e-3;
    ctx->tm.hess_z += z * 1e-3;
    ctx->tm.deriv += (x + y) * 1e-3;
    ctx->tm.coll += z * 1e-3;
    ctx->tm.per_radius[is] = (x + y + z) * 1e-3;
  }
  return 0;
}

// first derivatives of a K-layout k-vector -> three float fields (compute_first_derivatives,
// src/fmax.c:193-222).  Uses A[0], A[1] as x-pass destinations and D[0..2] as y-pass outputs.
static int first_derivs_to_vel(pinb200_ctx* ctx, const double2* kvec, double growth, float* const out[3], bool with_nyq,
                               const GrowthK* gk = nullptr) {
  const Geom& g = ctx->g;
  const double norm = 1.0 / ((double)g.N * g.N * g.N);
  TRY(run_dc(ctx, kvec, norm, 1));
  double2* xdst[3] = {ctx->A[1], ctx->A[0], nullptr};  // p=0 -> A1, p=1 -> A0
  // Rsmooth = 0 (src/fmax.c:200 with R = 0): window = 1
  const int slot = ctx->ndx < 4 ? ctx->ndx++ : -1;
  if (slot >= 0) CK(cudaEventRecord(ctx->ev_dx[2 * slot], ctx->stream));
  TRY(run_xpass_inv(ctx, kvec, xdst, 0x3, false, 1, 1, norm * growth, with_nyq, gk));
  if (slot >= 0) CK(cudaEventRecord(ctx->ev_dx[2 * slot + 1], ctx->stream));
  const double2* ysrc[3] = {ctx->A[0], ctx->A[1], nullptr};
  double2* ydst[6] = {ctx->D[0], ctx->D[1], ctx->D[2], nullptr, nullptr, nullptr};
  static const YJob jobs[3] = {{0, 0, 0}, {1, 1, 1}, {1, 0, 2}};
  TRY(run_ypass_inv(ctx, ysrc, ydst, jobs, 3, with_nyq));
  ZOutParams z{};
  for (int k = 0; k < 3; k++) {
    z.zs.src[k] = ydst[k];
    z.fdst[k] = out[k];
  }
  z.zs.kzpow[0] = 0; z.zs.kzpow[1] = 0; z.zs.kzpow[2] = 1;
  z.zs.ncomp = 3;
  z.zs.has_nyq = with_nyq ? 1 : 0;
  z.zs.dc_add = ctx->dc;
  z.g = g;
  z.tw = ctx->tw;
  z.mode = 1;
  LAUNCH(launch_zpass_out(g.N, z, (size_t)g.lx * g.N, ctx->stream));
  return 0;
}

// growth: scale-independent rates (gk == nullptr) or all ones with gk[0..3] the per-order tables
static int displacements_impl(pinb200_ctx* ctx, int compute_sources, const double growth[4], const GrowthK* gk) {
  if (!ctx->kdens_valid) FAIL("kdensity not resident");
  NEED_PEERS();
  CK(cudaSetDevice(ctx->d.device));
  const Geom& g = ctx->g;
  const int order = ctx->d.lpt_order;
  const size_t nrows = (size_t)g.lx * g.N;
  const double norm = 1.0 / ((double)g.N * g.N * g.N);
  CK(cudaEventRecord(ctx->ev[2], ctx->stream));
  if (!ctx->d_in_arena) for (int i = 0; i < 3; i++) TRY(dev_alloc(ctx, &ctx->D[i], ctx->field_elems));
  if (order >= 2 && compute_sources) {
    if (!ctx->hessian_valid)
      FAIL("second derivatives of the R=0 radius are not in place (call pinb200_fmax with a ladder ending in R = 0, or pinb200_second_derivatives(ctx, 0, NULL), first)");
    // ---- sources (src/LPT.c:64-93): A0 = S2, A1 = S31, A2 = S32 (real, R layout)
    SourcesParams sp{};
    for (int k = 0; k < 6; k++) sp.h[k] = reinterpret_cast<const double*>(ctx->B[k]);
    sp.s2 = reinterpret_cast<double*>(ctx->A[0]);
    sp.s31 = reinterpret_cast<double*>(ctx->A[1]);
    sp.s32 = reinterpret_cast<double*>(ctx->A[2]);
    sp.nrows = nrows;
    sp.N = g.N;
    sp.pitch = 2 * g.P;
    sp.lpt_order = order;
    LAUNCH(launch_sources(sp, ctx->stream));
    TRY(run_r2c(ctx, ctx->A[0], ctx->KV[0]));  // kvector_2LPT; A0 is free afterwards
    if (order >= 3) {
      // ---- second derivatives of phi_2 contracted with the Hessian (src/LPT.c:116-141), in
      //      three groups (by power of kx); x-pass destination A0, y-pass outputs D0..D2
      TRY(peer_barrier(ctx));  // kvector_2LPT complete on rank 0 before its k = 0 mode is read
      TRY(run_dc(ctx, ctx->KV[0], norm, 0));
      struct Grp { int pw; int n; YJob jobs[3]; int kz[3]; int slot[3]; };
      // slots: 0 xx,1 yy,2 zz,3 xy,4 xz,5 yz ; weight 2 (diagonal) or 4 (off-diagonal)
      const Grp grp[3] = {{2, 1, {{0, 0, 0}}, {0, 0, 0}, {0, 0, 0}},
                          {1, 2, {{0, 1, 0}, {0, 0, 1}}, {0, 1, 0}, {3, 4, 0}},
                          {0, 3, {{0, 2, 0}, {0, 1, 1}, {0, 0, 2}}, {0, 1, 2}, {1, 5, 2}}};
      for (const Grp& gr : grp) {
        double2* xdst[3] = {nullptr, nullptr, nullptr};
        xdst[gr.pw] = ctx->A[0];
        TRY(run_xpass_inv(ctx, ctx->KV[0], xdst, 1 << gr.pw, false, 1, 0, norm, true));
        const double2* ysrc[3] = {ctx->A[0], nullptr, nullptr};
        double2* ydst[6] = {ctx->D[0], ctx->D[1], ctx->D[2], nullptr, nullptr, nullptr};
        TRY(run_ypass_inv(ctx, ysrc, ydst, gr.jobs, gr.n, true));
        ZOutParams z{};
        for (int k = 0; k < gr.n; k++) {
          z.zs.src[k] = ydst[k];
          z.zs.kzpow[k] = gr.kz[k];
          z.hsrc[k] = reinterpret_cast<const double*>(ctx->B[gr.slot[k]]);
          z.weight[k] = 2.0 * (gr.slot[k] <= 2 ? 1.0 : 2.0);
        }
        z.zs.ncomp = gr.n;
        z.zs.has_nyq = 1;
        z.zs.dc_add = ctx->dc;
        z.g = g;
        z.tw = ctx->tw;
        z.mode = 2;
        z.acc = reinterpret_cast<double*>(ctx->A[2]);
        LAUNCH(launch_zpass_out(g.N, z, nrows, ctx->stream));
      }
      TRY(run_r2c(ctx, ctx->A[1], ctx->KV[1]));  // kvector_3LPT_1
      TRY(run_r2c(ctx, ctx->A[2], ctx->KV[2]));  // kvector_3LPT_2
    }
    // the Hessian fields are dead now: their memory takes the displacement fields below
    ctx->hessian_valid = false;
    ctx->kvec_valid = true;
  }
  if (order >= 2 && !ctx->kvec_valid) FAIL("LPT k-vectors are not resident (compute_sources = 0 needs an earlier call with compute_sources = 1)");
  CK(cudaEventRecord(ctx->ev[3], ctx->stream));
  const int nvel = order == 1 ? 3 : (order == 2 ? 6 : 12);
  if (!ctx->vel[0] && ctx->B[0] && ctx->B[5] && !ctx->hessian_valid) {
    for (int i = 0; i < nvel; i++) ctx->vel[i] = reinterpret_cast<float*>(ctx->B[i / 2]) + (size_t)(i % 2) * ctx->ncells;
    ctx->vel_in_B = true;
  }
  for (int i = 0; i < nvel; i++) TRY(dev_alloc(ctx, &ctx->vel[i], ctx->ncells));
  ctx->ndx = 0;
  TRY(peer_barrier(ctx));  // all k-vectors complete everywhere before rank 0's k = 0 modes are read
  if (order >= 2) TRY(first_derivs_to_vel(ctx, ctx->KV[0], growth[1], ctx->vel + 3, true, gk ? gk + 1 : nullptr));   // ScaleDep.order = 2
  if (order >= 3) {
    TRY(first_derivs_to_vel(ctx, ctx->KV[1], growth[2], ctx->vel + 6, true, gk ? gk + 2 : nullptr));                 // order 3
    TRY(first_derivs_to_vel(ctx, ctx->KV[2], growth[3], ctx->vel + 9, true, gk ? gk + 3 : nullptr));                 // order 4
  }
  TRY(first_derivs_to_vel(ctx, ctx->kdens, growth[0], ctx->vel + 0, ctx->kdens_has_nyq, gk));    // order 1, src/fmax.c:342-346
  CK(cudaEventRecord(ctx->ev[4], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  TRY(check_peer_error(ctx));
  float a = 0, b = 0;
  CK(cudaEventElapsedTime(&a, ctx->ev[2], ctx->ev[3]));
  CK(cudaEventElapsedTime(&b, ctx->ev[3], ctx->ev[4]));
  ctx->tm.lpt += (a + b) * 1e-3;
  ctx->tm.disp_sources += a * 1e-3;
  ctx->tm.disp_vel += b * 1e-3;
  for (int i = 0; i < ctx->ndx; i++) {
    float x = 0;
    CK(cudaEventElapsedTime(&x, ctx->ev_dx[2 * i], ctx->ev_dx[2 * i + 1]));
    ctx->tm.disp_x += x * 1e-3;
  }
  return 0;
}

extern "C" int pinb200_displacements(pinb200_ctx* ctx, int compute_sources, const double growth[4]) {
  if (!ctx || !growth) return 1;
  return displacements_impl(ctx, compute_sources, growth, nullptr);
}

extern "C" int pinb200_displacements_scaledep(pinb200_ctx* ctx, int compute_sources, int nk, double logkmin, double dlogk,
                                              const double* log10_growth) {
  if (!ctx || !log10_growth) return 1;
  if (nk < 1 || nk > 4096) FAIL("scale-dependent growth: nk out of range");
  if (!(dlogk > 0.0)) FAIL("scale-dependent growth: dlogk must be positive");
  CK(cudaSetDevice(ctx->d.device));
  TRY(dev_free(ctx, &ctx->growthk));
  TRY(dev_alloc(ctx, &ctx->growthk, (size_t)4 * nk));
  // pageable source: the copy is staged before the call returns, so the caller's array may go away
  CK(cudaMemcpyAsync(ctx->growthk, log10_growth, sizeof(double) * 4 * nk, cudaMemcpyHostToDevice, ctx->stream));
  GrowthK gk[4];
  for (int o = 0; o < 4; o++) {
    gk[o].tab = ctx->growthk + (size_t)o * nk;
    gk[o].n = nk;
    gk[o].logkmin = logkmin;
    gk[o].dlogk = dlogk;
    gk[o].sign = (o == 2) ? -1.0 : 1.0;  // GrowingMode_3LPT_1 returns the negative, src/cosmo.c:1810
  }
  static const double ones[4] = {1.0, 1.0, 1.0, 1.0};
  return displacements_impl(ctx, compute_sources, ones, gk);
}

// Cells with Fmax >= f_last in order of descending Fmax (ties: ascending cell index): the selection
// of src/distribute.c:58-175,547-600 and the ordering of sort_and_organize (src/fragment.c:484-520),
// done on the device (k_sort.cu): stable compaction of (key, index) pairs, then an LSD radix sort over the
// bits in which the keys differ.  One host synchronisation (the count of selected cells sizes the buffers).
//
// pinb200_handoff_begin / _end split it so that it can run UNDER the displacement stage: Fmax and Rmax are final
// when pinb200_fmax returns, the twelve displacement fields only 250 ms later (1024^3).  begin counts (one
// synchronisation), then puts the compaction, the sort and the copy of the index list on a side stream and the copy
// of the Fmax field on the copy stream, and returns; the caller runs pinb200_displacements; end waits.  With pinned
// host arrays the two downloads (6.8 GB at 1024^3) and the 94 ms of sorting cost no wall-clock time.
static int handoff_end_impl(pinb200_ctx* ctx) {
  HandoffState& h = ctx->ho;
  if (!h.active) return 0;
  CK(cudaStreamSynchronize(h.stream));
  if (ctx->copy_stream) CK(cudaStreamSynchronize(ctx->copy_stream));
  if (h.sorted) {
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h.ev0, h.ev1));
    ctx->tm.sort_ms = ms + h.count_ms;  // selection + sort on the device (without the downloads)
    TRY(dev_free(ctx, &ctx->sorted_idx));
    ctx->sorted_idx = h.idx[h.cur];      // the ordered indices stay resident for pinb200_download_products_sorted
    ctx->sorted_n = (size_t)h.n;
    h.idx[h.cur] = nullptr;
  }
  for (int b = 0; b < 2; b++) {
    TRY(dev_free(ctx, &h.key[b]));
    TRY(dev_free(ctx, &h.idx[b]));
  }
  TRY(dev_free(ctx, &h.counts));
  TRY(dev_free(ctx, &h.tile_counts));
  TRY(dev_free(ctx, &h.bsums));
  TRY(dev_free(ctx, &h.bs2));
  TRY(dev_free(ctx, &h.range));
  TRY(dev_free(ctx, &h.d_total));
  h.active = false;
  h.sorted = false;
  return 0;
}

extern "C" int pinb200_handoff_begin(pinb200_ctx* ctx, float f_last, float* fmax_out, unsigned int* cell_index_out, size_t capacity,
                                     size_t* count) {
  if (!ctx || !count) return 1;
  if (!ctx->fmax) FAIL("Fmax not computed");
  if (!(f_last > 0.0f)) FAIL("f_last must be positive (F = 1 + z_collapse; the float keys are ordered by their bit patterns)");
  if (ctx->ncells > 0xffffffffull) FAIL("more than 2^32 local cells");
  CK(cudaSetDevice(ctx->d.device));
  TRY(handoff_end_impl(ctx));
  HandoffState& h = ctx->ho;
  // (the y-pass scratch of an earlier displacement stage is dead: its memory serves the buffers below)
  if (!ctx->d_in_arena && ctx->kvec_valid && ctx->vel[0]) for (auto& w : ctx->D) TRY(dev_free(ctx, &w));
  if (!h.stream) {
    CK(cudaStreamCreateWithFlags(&h.stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&h.ev0));
    CK(cudaEventCreate(&h.ev1));
    CK(cudaEventCreateWithFlags(&h.ev_ready, cudaEventDisableTiming));
  }
  if (!ctx->copy_stream) {
    CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (auto& e : ctx->ev_stage) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  h.active = true;
  // the Fmax field itself, on the copy stream (nothing below depends on it)
  if (fmax_out) CK(cudaMemcpyAsync(fmax_out, ctx->fmax, ctx->ncells * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_stream));
  CK(cudaEventRecord(ctx->ev[5], ctx->stream));
  const unsigned long long nc = ctx->ncells;
  const size_t ntsel = cell_sort_ntiles_select(nc);
  TRY(dev_alloc(ctx, &h.tile_counts, ntsel));
  TRY(dev_alloc(ctx, &h.bsums, cell_sort_scan_blocks(ntsel) + 1));
  TRY(dev_alloc(ctx, &h.range, (size_t)2));
  TRY(dev_alloc(ctx, &h.d_total, (size_t)1));
  const unsigned int range_init[2] = {0xffffffffu, 0u};
  CK(cudaMemcpyAsync(h.range, range_init, sizeof range_init, cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(launch_select_count(ctx->fmax, nc, f_last, h.tile_counts, h.range, ctx->stream));
  LAUNCH(launch_scan_u32(h.tile_counts, ntsel, h.bsums, h.d_total, ctx->stream));
  ctx->launches += 2;
  unsigned long long n = 0;
  unsigned int hrange[2] = {0, 0};
  CK(cudaMemcpyAsync(&n, h.d_total, sizeof n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(hrange, h.range, sizeof hrange, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaEventRecord(ctx->ev[6], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  float cms = 0;
  CK(cudaEventElapsedTime(&cms, ctx->ev[5], ctx->ev[6]));
  h.count_ms = cms;
  h.n = n;
  *count = (size_t)n;
  if (n == 0 || !cell_index_out || capacity == 0) return 0;
  for (int b = 0; b < 2; b++) {
    TRY(dev_alloc(ctx, &h.key[b], (size_t)n));
    TRY(dev_alloc(ctx, &h.idx[b], (size_t)n));
  }
  // digits: the keys are offsets from the smallest one, so only the bits of (kmax - kmin) take part
  int sig = 0;
  while (sig < 32 && ((unsigned long long)(hrange[1] - hrange[0]) >> sig) != 0) sig++;
  int maxbits = 0;
  while ((1 << maxbits) < cell_sort_max_bins()) maxbits++;
  const int npass = (sig + maxbits - 1) / maxbits;
  const size_t ntr = cell_sort_ntiles_radix(n);
  if (npass > 0) {
    TRY(dev_alloc(ctx, &h.counts, (size_t)cell_sort_max_bins() * ntr));
    TRY(dev_alloc(ctx, &h.bs2, cell_sort_scan_blocks((unsigned long long)cell_sort_max_bins() * ntr) + 1));
  }
  // the buffers were allocated in the order of the main stream: the side stream starts behind that point
  CK(cudaEventRecord(h.ev_ready, ctx->stream));
  CK(cudaStreamWaitEvent(h.stream, h.ev_ready, 0));
  cudaStream_t st = h.stream;
  CK(cudaEventRecord(h.ev0, st));
  LAUNCH(launch_select_write(ctx->fmax, nc, f_last, h.tile_counts, h.range, h.key[0], h.idx[0], st));
  h.cur = 0;
  int shift = 0;
  for (int pass = 0; pass < npass; pass++) {
    const int bits = (sig - shift + (npass - pass) - 1) / (npass - pass);  // spread the significant bits evenly
    LAUNCH(launch_radix_hist(h.key[h.cur], n, shift, bits, h.counts, st));
    LAUNCH(launch_scan_u32(h.counts, ((unsigned long long)1 << bits) * ntr, h.bs2, nullptr, st));
    LAUNCH(launch_radix_scatter(h.key[h.cur], h.idx[h.cur], h.key[h.cur ^ 1], h.idx[h.cur ^ 1], n, shift, bits, h.counts, st));
    ctx->launches += 2;
    shift += bits;
    h.cur ^= 1;
  }
  CK(cudaEventRecord(h.ev1, st));
  const size_t ncopy = capacity < (size_t)n ? capacity : (size_t)n;
  CK(cudaMemcpyAsync(cell_index_out, h.idx[h.cur], ncopy * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  h.sorted = true;
  return 0;
}

extern "C" int pinb200_handoff_end(pinb200_ctx* ctx) {
  if (!ctx) return 1;
  CK(cudaSetDevice(ctx->d.device));
  return handoff_end_impl(ctx);
}

extern "C" int pinb200_collapsed_cells(pinb200_ctx* ctx, float f_last, unsigned int* cell_index_out, size_t capacity,
                                       size_t* count) {
  if (!ctx || !count) return 1;
  TRY(pinb200_handoff_begin(ctx, f_last, nullptr, cell_index_out, capacity, count));
  return handoff_end_impl(ctx);
}

extern "C" int pinb200_fmax_pdf(pinb200_ctx* ctx, unsigned long long* counts) {
  if (!ctx || !counts) return 1;
  if (!ctx->fmax) FAIL("Fmax not computed");
  CK(cudaSetDevice(ctx->d.device));
  unsigned long long* d = nullptr;
  TRY(dev_alloc(ctx, &d, (size_t)PINB200_NBINS));
  CK(cudaMemsetAsync(d, 0, sizeof(unsigned long long) * PINB200_NBINS, ctx->stream));
  LAUNCH(launch_fmax_pdf(ctx->fmax, ctx->ncells, d, ctx->stream));
  CK(cudaMemcpyAsync(counts, d, sizeof(unsigned long long) * PINB200_NBINS, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  TRY(dev_free(ctx, &d));
  return 0;
}

// Records [first, first + n) of a record stream -- cells in index order (gather == nullptr) or in the order of an
// index list -- packed on the device and copied to the host in chunks through TWO device staging buffers: the
// pack kernel of chunk k+1 runs on the compute stream while the copy engine moves chunk k on a second stream,
// so the PCIe link never waits for a kernel and the device needs 2 x 16 Mi records of staging, not n records
// (r01 staged a whole call at once: 60 GB for the AoS of a 1024^3 slab).  sink != nullptr: records holding
// members that are not ours are staged on the host and merged member by member instead (product_merge.h).
// fd >= 0: the records go to that file descriptor instead of `host`, through two pinned host buffers: the copy
// engine fills one while write(2) drains the other (snapshot blocks and product dumps written from the device SoA).
static int write_all(pinb200_ctx* ctx, int fd, const unsigned char* p, size_t bytes) {
  while (bytes > 0) {
    const ssize_t w = write(fd, p, bytes);
    if (w < 0) {
      if (errno == EINTR) continue;
      ctx->err = std::string("write failed: ") + strerror(errno);
      return 1;
    }
    p += w;
    bytes -= (size_t)w;
  }
  return 0;
}

static int stream_records(pinb200_ctx* ctx, void* host, const pinb200_product_layout* L, const unsigned int* gather, size_t first, size_t n,
                          int fd = -1) {
  if (n == 0) return 0;
  const size_t max_chunk = fd >= 0 ? ((size_t)1 << 21) : ((size_t)1 << 24);
  const size_t chunk = n < max_chunk ? n : max_chunk;
  PackParams p{};
  p.fmax = ctx->fmax;
  p.rmax = ctx->rmax;
  for (int i = 0; i < 12; i++) p.vel[i] = ctx->vel[i];
  p.stride = L->stride;
  p.prodfloat_bytes = L->prodfloat_bytes;
  p.off_rmax = L->off_Rmax;
  p.off_fmax = L->off_Fmax;
  p.off_vel[0] = L->off_Vel;
  p.off_vel[1] = L->off_Vel_2LPT;
  p.off_vel[2] = L->off_Vel_3LPT_1;
  p.off_vel[3] = L->off_Vel_3LPT_2;
  p.gather = gather;
  const bool has_vel[4] = {ctx->vel[0] != nullptr, ctx->vel[3] != nullptr, ctx->vel[6] != nullptr, ctx->vel[9] != nullptr};
  const std::vector<MemberRange> members = product_members(*L, ctx->fmax != nullptr, has_vel);
  const bool direct = fd >= 0 || members_cover_record(members, L->stride);  // a file gets zero bytes for foreign members
  if (!ctx->copy_stream) {
    CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (auto& e : ctx->ev_stage) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  unsigned char* d[2] = {nullptr, nullptr};
  const int nbuf = n > chunk ? 2 : 1;
  for (int b = 0; b < nbuf; b++) TRY(dev_alloc(ctx, &d[b], chunk * L->stride));
  std::vector<unsigned char> stage;
  if (!direct) stage.resize(chunk * L->stride);
  unsigned char* out = static_cast<unsigned char*>(host);
  const bool zero_fill = !members_cover_record(members, L->stride);
  unsigned char* pinned[2] = {nullptr, nullptr};
  size_t pending_bytes[2] = {0, 0};
  if (fd >= 0) {
    if (ctx->pinned_bytes < 2 * chunk * L->stride) {
      if (ctx->pinned) CK(cudaFreeHost(ctx->pinned));
      ctx->pinned = nullptr;
      CK(cudaMallocHost((void**)&ctx->pinned, 2 * chunk * L->stride));
      ctx->pinned_bytes = 2 * chunk * L->stride;
    }
    pinned[0] = ctx->pinned;
    pinned[1] = ctx->pinned + chunk * L->stride;
  }
  CK(cudaEventRecord(ctx->ev[5], ctx->stream));
  size_t k = 0;
  for (size_t off = 0; off < n; off += chunk, k++) {
    const int b = (int)(k & 1);
    const size_t m = n - off < chunk ? n - off : chunk;
    if (k >= 2) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_stage[2 + b], 0));  // the copy of chunk k-2 has left this buffer
    if (zero_fill) CK(cudaMemsetAsync(d[b], 0, m * L->stride, ctx->stream));
    p.out = d[b];
    p.cell_begin = first + off;
    p.ncells = m;
    LAUNCH(launch_pack_products(p, ctx->stream));
    CK(cudaEventRecord(ctx->ev_stage[b], ctx->stream));
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_stage[b], 0));
    if (fd >= 0) {
      // pinned[b] was written out before chunk k-1's copy was issued (see below), so it is free
      CK(cudaMemcpyAsync(pinned[b], d[b], m * L->stride, cudaMemcpyDeviceToHost, ctx->copy_stream));
      CK(cudaEventRecord(ctx->ev_stage[2 + b], ctx->copy_stream));
      pending_bytes[b] = m * L->stride;
      // while chunk k crosses PCIe, chunk k-1 goes to the file
      if (k >= 1) {
        CK(cudaEventSynchronize(ctx->ev_stage[2 + (b ^ 1)]));
        TRY(write_all(ctx, fd, pinned[b ^ 1], pending_bytes[b ^ 1]));
        pending_bytes[b ^ 1] = 0;
      }
      continue;
    }
    if (direct) {
      CK(cudaMemcpyAsync(out + off * L->stride, d[b], m * L->stride, cudaMemcpyDeviceToHost, ctx->copy_stream));
    } else {
      CK(cudaMemcpyAsync(stage.data(), d[b], m * L->stride, cudaMemcpyDeviceToHost, ctx->copy_stream));
      CK(cudaStreamSynchronize(ctx->copy_stream));
      merge_product_members(out + off * L->stride, stage.data(), L->stride, m, members);
    }
    CK(cudaEventRecord(ctx->ev_stage[2 + b], ctx->copy_stream));
  }
  CK(cudaStreamSynchronize(ctx->copy_stream));
  if (fd >= 0)
    for (int b = 0; b < 2; b++)
      if (pending_bytes[(k + b) & 1]) TRY(write_all(ctx, fd, pinned[(k + b) & 1], pending_bytes[(k + b) & 1]));  // the last chunk
  CK(cudaEventRecord(ctx->ev[6], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]));
  ctx->tm.mem_transf += ms * 1e-3;
  for (int b = 0; b < nbuf; b++) TRY(dev_free(ctx, &d[b]));
  return 0;
}

extern "C" int pinb200_write_products(pinb200_ctx* ctx, int fd, const pinb200_product_layout* L, size_t cell_begin, size_t ncells) {
  if (!ctx || !L || fd < 0) return 1;
  if (!ctx->fmax && !ctx->vel[0]) FAIL("products not computed");
  if (cell_begin + ncells > ctx->ncells) FAIL("cell range outside the local slab");
  if (L->prodfloat_bytes != 4 && L->prodfloat_bytes != 8) FAIL("prodfloat_bytes must be 4 or 8");
  CK(cudaSetDevice(ctx->d.device));
  return stream_records(ctx, nullptr, L, nullptr, cell_begin, ncells, fd);
}

extern "C" int pinb200_write_block(pinb200_ctx* ctx, int fd, int block, size_t cell_begin, size_t ncells) {
  if (!ctx || fd < 0) return 1;
  if (cell_begin + ncells > ctx->ncells) FAIL("cell range outside the local slab");
  // a block is a record stream with one member: 4 bytes (FMAX, RMAX) or three floats (AuxStruct of src/write_snapshot.c)
  pinb200_product_layout L{};
  L.prodfloat_bytes = 4;
  L.off_Rmax = L.off_Fmax = L.off_Vel = L.off_Vel_2LPT = L.off_Vel_3LPT_1 = L.off_Vel_3LPT_2 = -1;
  const void* need = nullptr;
  switch (block) {
    case PINB200_BLOCK_FMAX: L.stride = 4; L.off_Fmax = 0; need = ctx->fmax; break;
    case PINB200_BLOCK_RMAX: L.stride = 4; L.off_Rmax = 0; need = ctx->rmax; break;
    case PINB200_BLOCK_ZEL: L.stride = 12; L.off_Vel = 0; need = ctx->vel[0]; break;
    case PINB200_BLOCK_2LPT: L.stride = 12; L.off_Vel_2LPT = 0; need = ctx->vel[3]; break;
    case PINB200_BLOCK_3LPT_1: L.stride = 12; L.off_Vel_3LPT_1 = 0; need = ctx->vel[6]; break;
    case PINB200_BLOCK_3LPT_2: L.stride = 12; L.off_Vel_3LPT_2 = 0; need = ctx->vel[9]; break;
    default: FAIL("unknown snapshot block");
  }
  if (!need) FAIL("the field of this block is not resident");
  CK(cudaSetDevice(ctx->d.device));
  return stream_records(ctx, nullptr, &L, nullptr, cell_begin, ncells, fd);
}

extern "C" int pinb200_download_products(pinb200_ctx* ctx, void* products, const pinb200_product_layout* L, size_t cell_begin,
                                         size_t ncells) {
  if (!ctx || !products || !L) return 1;
  // special mode 3 (displacements without an Fmax sweep, src/pinocchio.c:170-200) leaves Fmax/Rmax zero
  if (!ctx->fmax && !ctx->vel[0]) FAIL("products not computed");
  if (cell_begin + ncells > ctx->ncells) FAIL("cell range outside the local slab");
  if (L->prodfloat_bytes != 4 && L->prodfloat_bytes != 8) FAIL("prodfloat_bytes must be 4 or 8");
  CK(cudaSetDevice(ctx->d.device));
  // Records whose every byte is a member we deliver (the default 56-byte product_data) are copied
  // straight over the caller's; records with foreign members (*_prev of RECOMPUTE_DISPLACEMENTS,
  // zacc/group_ID of SNAPSHOT) are staged and merged member by member (product_merge.h).
  return stream_records(ctx, products, L, nullptr, cell_begin, ncells);
}

// frag[first .. first+n) as sort_and_organize leaves it (src/fragment.c:484-520): the records of the
// collapsed cells in order of descending Fmax, gathered on the device through the index list of the
// last pinb200_collapsed_cells call.
extern "C" int pinb200_download_products_sorted(pinb200_ctx* ctx, void* products, const pinb200_product_layout* L, size_t first,
                                                size_t n) {
  if (!ctx || !products || !L) return 1;
  if (!ctx->sorted_idx) FAIL("no ordered cell list (call pinb200_collapsed_cells with an output array first)");
  if (first + n > ctx->sorted_n) FAIL("record range outside the ordered cell list");
  if (L->prodfloat_bytes != 4 && L->prodfloat_bytes != 8) FAIL("prodfloat_bytes must be 4 or 8");
  CK(cudaSetDevice(ctx->d.device));
  return stream_records(ctx, products, L, ctx->sorted_idx, first, n);
}

extern "C" int pinb200_download_field(pinb200_ctx* ctx, int which, void* dst) {
  if (!ctx || !dst) return 1;
  const void* src = nullptr;
  if (which == 0) src = ctx->fmax;
  else if (which == 1) src = ctx->rmax;
  else if (which >= 2 && which < 14) src = ctx->vel[which - 2];
  if (!src) FAIL("requested field is not resident");
  CK(cudaSetDevice(ctx->d.device));
  CK(cudaMemcpyAsync(dst, src, ctx->ncells * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int pinb200_get_timers(pinb200_ctx* ctx, pinb200_timers* t) {
  if (!ctx || !t) return 1;
  ctx->tm.kernel_launches = ctx->launches;
  *t = ctx->tm;
  return 0;
}

// ---- finer-grained entry points -------------------------------------------------------------
extern "C" int pinb200_fft_c2r(pinb200_ctx* ctx, const double* cplx_in, double* real_out) {
  if (!ctx || !cplx_in || !real_out) return 1;
  NEED_PEERS();  // collective on several ranks: every rank passes its own slab
  CK(cudaSetDevice(ctx->d.device));
  const Geom& g = ctx->g;
  const double norm = 1.0 / ((double)g.N * g.N * g.N);
  // K-layout input in A0, x pass into A1 (R layout), y pass into A2 (the long-line y pass is
  // not usable in place, see YCfg), z pass in place
  ctx->hessian_valid = false;
  TRY(upload_cplx(ctx, cplx_in, ctx->A[0]));
  double2* xdst[3] = {ctx->A[1], nullptr, nullptr};
  TRY(run_xpass_inv(ctx, ctx->A[0], xdst, 1, false, 0, 0, norm, true));
  const double2* ysrc[3] = {ctx->A[1], nullptr, nullptr};
  double2* ydst[6] = {ctx->A[2], nullptr, nullptr, nullptr, nullptr, nullptr};
  YJob job{0, 0, 0};
  TRY(run_ypass_inv(ctx, ysrc, ydst, &job, 1, true));
  ZOutParams z{};
  z.zs.src[0] = ctx->A[2];
  z.zs.kzpow[0] = 0;
  z.zs.ncomp = 1;
  z.zs.has_nyq = 1;
  z.zs.dc_add = nullptr;
  z.g = g;
  z.tw = ctx->tw;
  z.mode = 0;
  z.rdst[0] = ctx->A[2];
  LAUNCH(launch_zpass_out(g.N, z, (size_t)g.lx * g.N, ctx->stream));
  TRY(download_real(ctx, ctx->A[2], real_out));
  return 0;
}

extern "C" int pinb200_fft_r2c(pinb200_ctx* ctx, const double* real_in, double* cplx_out) {
  if (!ctx || !real_in || !cplx_out) return 1;
  NEED_PEERS();
  CK(cudaSetDevice(ctx->d.device));
  const Geom& g = ctx->g;
  const size_t rows = (size_t)g.lx * g.N;
  double* tmp = nullptr;
  TRY(dev_alloc(ctx, &tmp, rows * g.N));
  CK(cudaMemcpyAsync(tmp, real_in, rows * g.N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(ctx->A[0], 0, ctx->field_elems * sizeof(double2), ctx->stream));
  LAUNCH(launch_repitch_r(tmp, reinterpret_cast<double*>(ctx->A[0]), rows, g.N, g.N, 2 * g.P, ctx->stream));
  TRY(run_r2c(ctx, ctx->A[0], ctx->A[1]));
  TRY(download_cplx(ctx, ctx->A[1], cplx_out));
  TRY(dev_free(ctx, &tmp));
  return 0;
}

extern "C" int pinb200_second_derivatives(pinb200_ctx* ctx, double radius, double* hessian_out) {
  if (!ctx) return 1;
  if (!ctx->kdens_valid) FAIL("kdensity not resident");
  NEED_PEERS();
  CK(cudaSetDevice(ctx->d.device));
  const Geom& g = ctx->g;
  if (ctx->vel_in_B) TRY(release_vel(ctx));  // the displacement fields live in B: they are gone with this call
  for (int i = 0; i < 6; i++) TRY(dev_alloc(ctx, &ctx->B[i], ctx->field_elems));
  const double cell = ctx->d.box_size / g.N;
  TRY(hessian_dc(ctx));
  TRY(hessian_xy(ctx, radius / cell));
  ZOutParams z{};
  for (int k = 0; k < 6; k++) {
    z.zs.src[k] = ctx->B[k];
    z.zs.kzpow[k] = kHessKzPow[k];
    z.rdst[k] = ctx->B[k];
  }
  z.zs.ncomp = 6;
  z.zs.has_nyq = ctx->kdens_has_nyq ? 1 : 0;
  z.zs.dc_add = ctx->dc;
  z.g = g;
  z.tw = ctx->tw;
  z.mode = 0;
  LAUNCH(launch_zpass_out(g.N, z, (size_t)g.lx * g.N, ctx->stream));
  if (hessian_out)
    for (int k = 0; k < 6; k++) TRY(download_real(ctx, ctx->B[k], hessian_out + (size_t)k * ctx->ncells));
  // the R = 0 Hessian is what the LPT sources need (recompute_sd of compute_displacements,
  // src/fmax.c:301-319): it stays resident for pinb200_displacements(compute_sources = 1)
  ctx->hessian_valid = (radius == 0.0);
  CK(cudaStreamSynchronize(ctx->stream));
  return check_peer_error(ctx);
}

extern "C" int pinb200_collapse_cells(pinb200_ctx* ctx, int ismooth, const double* hessian6, size_t ncells, double* F_out) {
  if (!ctx || !hessian6 || !F_out) return 1;
  CK(cudaSetDevice(ctx->d.device));
  if (ctx->ct_on && (ismooth < 0 || ismooth >= ctx->ct_ns)) FAIL("ismooth out of range of the collapse tables");
  if (!ctx->ct_on) TRY(upload_splines(ctx));
  if (ncells == 0) return 0;
  double *dh = nullptr, *dF = nullptr;
  TRY(dev_alloc(ctx, &dh, 6 * ncells));
  TRY(dev_alloc(ctx, &dF, ncells));
  CK(cudaMemcpyAsync(dh, hessian6, 6 * ncells * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (ctx->ct_on) LAUNCH(launch_collapse_cells_tab(dh, ncells, ct_view(ctx, ismooth), dF, ctx->stream));
  else LAUNCH(launch_collapse_cells(dh, ncells, spline_for(ctx, ismooth), ctx->nspl, dF, ctx->stream));
  CK(cudaMemcpyAsync(F_out, dF, ncells * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  TRY(dev_free(ctx, &dh));
  TRY(dev_free(ctx, &dF));
  return 0;
}

extern "C" int pinb200_download_kvector(pinb200_ctx* ctx, int which, double* kvec) {
  if (!ctx || !kvec) return 1;
  if (which < 0 || which > 2 || !ctx->KV[which]) FAIL("which must name a k-vector of the configured LPT order");
  if (!ctx->kvec_valid) FAIL("LPT k-vectors are not resident");
  CK(cudaSetDevice(ctx->d.device));
  return download_cplx(ctx, ctx->KV[which], kvec);
}
