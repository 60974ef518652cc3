// CPU block emulator for the kernel bodies of pinocchio_b200/csrc/kernels.cuh.  TEST ONLY.
//
// Each "block" is executed by NT fibers of one host thread that share a heap buffer as shared memory;
// __syncthreads() is a switch to the next fiber (see FiberLaunch); blocks run one after another.  The bodies are the very
// same templates that k_*.cu instantiate for sm_100a, so index math, plans, twiddles, the
// multi-rank scatter addressing and the collapse arithmetic are checked on a CPU-only box
// against the oracle (tests/test_emulator.py).  Several ranks are emulated one after the other
// in one process: "peer memory" is simply another host array.
// This library is never loaded by the product (pinocchio_b200/), which has no CPU path.
#include <ucontext.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../pinocchio_b200/csrc/kernels.cuh"
#include "../../pinocchio_b200/csrc/sort_cells.cuh"
#include "../../pinocchio_b200/csrc/scaledep_gm.cuh"

using namespace pinb;

namespace {
std::mutex g_atomic_mutex;

// A block's "threads" are fibers of ONE host thread (ucontext), scheduled round-robin: a fiber runs until its
// next barrier (or its end), then the next one runs; when the sweep is over every fiber stands at the same
// barrier, which is thereby complete.  (A first version used one OS thread per emulated thread and a pthread
// barrier: with a few hundred threads on a handful of cores two thirds of the run time of a 32^3 sweep went
// into futex system calls.)  Blocks run one after another because the callers share one "shared memory" buffer.
struct FiberLaunch {
  ucontext_t main_ctx;
  std::vector<ucontext_t> fibers;
  std::vector<char*> stacks;
  std::vector<char> done;
  int cur = 0;
  void yield_to_scheduler() { swapcontext(&fibers[cur], &main_ctx); }
};
thread_local FiberLaunch* g_launch = nullptr;
thread_local std::function<void(int)>* g_fiber_body = nullptr;

struct HostCtx {
  int tid_, bid_, nt_;
  FiberLaunch* launch;
  int tid() const { return tid_; }
  int bid() const { return bid_; }
  int nthreads() const { return nt_; }
  void sync() const { if (launch) launch->yield_to_scheduler(); }
  int warp_uniform(int v) const { return v; }
  // every line group calls sync_line the same number of times, so a block-wide barrier is a
  // valid (stronger) stand-in on the host
  void sync_line(int, int) const { sync(); }
  void async_copy16(void* dst, const void* src) const { std::memcpy(dst, src, 16); }
  void async_wait() const {}
  void prefetch_l2(const void*) const {}
  void mark(int) const {}
  void sched_fence() const {}
  template <int NT> void block_sum2(double* scratch, double& a, double& b) const { block_sum2_tree<NT>(*this, scratch, a, b); }
  void atomic_add(double* p, double v) const {
    std::lock_guard<std::mutex> lk(g_atomic_mutex);
    *p += v;
  }
};

extern "C" void emu_fiber_trampoline(int t) {
  (*g_fiber_body)(t);
  g_launch->done[t] = 1;
  // returning switches to uc_link (the scheduler)
}

template <class F> void run_blocks(long long nblocks, int nt, F body) {
  constexpr size_t STACK = 256 * 1024;
  FiberLaunch L;
  L.fibers.resize(nt);
  L.stacks.resize(nt);
  L.done.assign(nt, 0);
  for (int t = 0; t < nt; t++) L.stacks[t] = static_cast<char*>(std::malloc(STACK));
  FiberLaunch* prev_launch = g_launch;
  std::function<void(int)>* prev_body = g_fiber_body;
  long long block = 0;
  std::function<void(int)> fiber_body = [&](int t) {
    HostCtx ctx{t, (int)block, nt, &L};
    body(ctx);
  };
  g_launch = &L;
  g_fiber_body = &fiber_body;
  for (block = 0; block < nblocks; block++) {
    for (int t = 0; t < nt; t++) {
      getcontext(&L.fibers[t]);
      L.fibers[t].uc_stack.ss_sp = L.stacks[t];
      L.fibers[t].uc_stack.ss_size = STACK;
      L.fibers[t].uc_link = &L.main_ctx;
      makecontext(&L.fibers[t], (void (*)())emu_fiber_trampoline, 1, t);
      L.done[t] = 0;
    }
    int remaining = nt;
    while (remaining > 0) {
      remaining = 0;
      for (int t = 0; t < nt; t++) {
        if (L.done[t]) continue;
        L.cur = t;
        swapcontext(&L.main_ctx, &L.fibers[t]);
        if (!L.done[t]) remaining++;
      }
    }
  }
  g_launch = prev_launch;
  g_fiber_body = prev_body;
  for (int t = 0; t < nt; t++) std::free(L.stacks[t]);
}

Geom make_geom(int N, int rank, int nranks) {
  Geom g;
  g.N = N;
  g.M = N / 2;
  g.P = g.M + 8;
  g.lx = N / nranks;
  g.ly = N / nranks;
  g.x0 = rank * g.lx;
  g.y0 = rank * g.ly;
  g.knorm = 2. * PINB_PI / (double)N;
  return g;
}
int ilog2(int v) { int s = 0; while ((1 << s) < v) s++; return s; }
}  // namespace

#define EMU_GRIDS(X) X(32) X(64)
#define EMU_LINES(X) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048)

// ---- single-tile line FFTs for every supported length ---------------------------------------
// in/out: [L][TK] complex (tile of TK lines, element-major), tw: NROOT=L roots
template <int L, int DIR> static void strided_tile(const double2* in, double2* out, const double2* tw) {
  constexpr int TK = StridedCfg<L>::TK;
  constexpr int NT = Plan<L, false>::TPL * TK;
  std::vector<double2> smem((size_t)L * TK);
  run_blocks(1, NT, [&](HostCtx& ctx) {
    auto load = [&](int e, int tk) { return in[e * TK + tk]; };
    auto store = [&](int e, int tk, double2 v) { out[e * TK + tk] = v; };
    strided_tile_fft<L, TK, DIR>(ctx, smem.data(), tw, 1, load, store);
  });
}

extern "C" int emu_strided_tk(int L) {
  switch (L) {
#define X(LL) case LL: return StridedCfg<LL>::TK;
    EMU_LINES(X)
#undef X
  }
  return 0;
}

extern "C" int emu_strided_tile_fft(int L, int dir, const double* in, double* out, const double* tw) {
  switch (L) {
#define X(LL)                                                                                       \
  case LL:                                                                                          \
    if (dir > 0) strided_tile<LL, +1>((const double2*)in, (double2*)out, (const double2*)tw);        \
    else strided_tile<LL, -1>((const double2*)in, (double2*)out, (const double2*)tw);                \
    return 0;
    EMU_LINES(X)
#undef X
  }
  return 1;
}

// contiguous line of M complex in padded smem; tw has 2M roots
template <int M, int DIR> static void zline(const double2* in, double2* out, const double2* tw) {
  constexpr int NT = Plan<M, true>::TPL;
  std::vector<double2> smem(ZLine<M>::PITCH);
  for (int e = 0; e < M; e++) smem[zpad(e)] = in[e];
  run_blocks(1, NT, [&](HostCtx& ctx) { zline_fft_smem<M, DIR>(ctx, smem.data(), ctx.tid(), tw); });
  for (int e = 0; e < M; e++) out[e] = smem[zpad(e)];
}

extern "C" int emu_zline_fft(int M, int dir, const double* in, double* out, const double* tw) {
  switch (M) {
#define X(LL)                                                                                 \
  case LL:                                                                                    \
    if (dir > 0) zline<LL, +1>((const double2*)in, (double2*)out, (const double2*)tw);         \
    else zline<LL, -1>((const double2*)in, (double2*)out, (const double2*)tw);                 \
    return 0;
    EMU_LINES(X)
#undef X
  }
  return 1;
}

// ---- whole kernels on pitched arrays (K layout [N][ly][P], R layout [lx][N][P]) ---------------
// blocks of a strided pass: fewer than tiles, and not a divisor of their number, so that the blocks walk several tiles
// each (as the persistent launches of k_strided.cu do) and the last round is ragged
static int emu_grid(int ntiles) { return ntiles > 7 ? 7 : ntiles; }
template <int N, int DIR> static void xpass_run(XPassParams& p, int with_nyq) {
  if constexpr (DIR > 0 && XCfg<N, +1>::SPLIT) {
    if (p.dst_klayout == 2 && !p.kf.gk) {  // as xpass_launch (k_strided.cu): the all-local pass never splits its lines
      using C = XCfg<N, +1, true>;
      p.ntiles_z = (N / 2) / C::TK + (with_nyq ? 1 : 0);
      std::vector<double2> smem((size_t)C::LT * C::TK);
      p.nblocks = p.g.ly * p.ntiles_z;
      p.tile_stride = emu_grid(p.nblocks);
      run_blocks(p.tile_stride, C::NT, [&](HostCtx& ctx) { xpass_body<N, +1, false, HostCtx, false, true>(ctx, smem.data(), p); });
      return;
    }
  }
  using C = XCfg<N, DIR>;
  p.ntiles_z = (N / 2) / C::TK + (with_nyq ? 1 : 0);
  std::vector<double2> smem((size_t)C::LT * C::TK);
  p.nblocks = p.g.ly * p.ntiles_z;
  p.tile_stride = emu_grid(p.nblocks);
  if (p.kf.gk)
    run_blocks(p.tile_stride, C::NT, [&](HostCtx& ctx) { xpass_body<N, DIR, true, HostCtx, true>(ctx, smem.data(), p); });
  else
    run_blocks(p.tile_stride, C::NT, [&](HostCtx& ctx) { xpass_body<N, DIR, true>(ctx, smem.data(), p); });
}

// dsts: 3*nranks pointers, dsts[p*nranks + r] = destination field for power p on rank r
extern "C" int emu_xpass(int N, int dir, int rank, int nranks, const double* src, double** dsts, int dst_klayout,
                         int pmask, int with_nyq, const double* gauss, double scalar, int green, int times_i,
                         const double* tw);
// the same with the scale-dependent growth table gk[gk_n] (see KFactor); gk == nullptr: plain emu_xpass
extern "C" int emu_xpass_gk(int N, int dir, int rank, int nranks, const double* src, double** dsts, int dst_klayout,
                            int pmask, int with_nyq, const double* gauss, double scalar, int green, int times_i,
                            const double* tw, const double* gk, int gk_n, double gk_logkmin, double gk_dlogk, double gk_sign);
extern "C" int emu_xpass(int N, int dir, int rank, int nranks, const double* src, double** dsts, int dst_klayout,
                         int pmask, int with_nyq, const double* gauss, double scalar, int green, int times_i,
                         const double* tw) {
  return emu_xpass_gk(N, dir, rank, nranks, src, dsts, dst_klayout, pmask, with_nyq, gauss, scalar, green, times_i, tw, nullptr, 0,
                      0.0, 1.0, 1.0);
}
extern "C" int emu_xpass_gk(int N, int dir, int rank, int nranks, const double* src, double** dsts, int dst_klayout,
                            int pmask, int with_nyq, const double* gauss, double scalar, int green, int times_i,
                            const double* tw, const double* gk, int gk_n, double gk_logkmin, double gk_dlogk, double gk_sign) {
  XPassParams p{};
  p.kf.gk = gk;
  p.kf.gk_n = gk_n;
  p.kf.gk_logkmin = gk_logkmin;
  p.kf.gk_dlogk = gk_dlogk;
  p.kf.gk_sign = gk_sign;
  p.src = (const double2*)src;
  for (int pw = 0; pw < 3; pw++)
    for (int r = 0; r < nranks; r++) p.dst[pw].r[r] = (double2*)dsts[pw * nranks + r];
  p.dst_klayout = dst_klayout;
  p.pmask = pmask;
  p.kf.gauss = gauss;
  p.kf.scalar = scalar;
  p.kf.green = green;
  p.kf.times_i = times_i;
  p.g = make_geom(N, rank, nranks);
  p.lx_shift = ilog2(p.g.lx);
  p.tw = (const double2*)tw;
  switch (N) {
#define X(LL)                                                         \
  case LL:                                                            \
    if (dir > 0) xpass_run<LL, +1>(p, with_nyq); else xpass_run<LL, -1>(p, with_nyq); \
    return 0;
    EMU_GRIDS(X)
#undef X
  }
  return 1;
}

template <int N, int DIR> static void ypass_run(YPassParams& p, int with_nyq) {
  using C = YCfg<N>;
  p.ntiles_z = (N / 2) / C::TK + (with_nyq ? 1 : 0);
  std::vector<double2> smem((size_t)C::LT * C::TK);
  p.nblocks = p.g.lx * p.ntiles_z;
  p.tile_stride = emu_grid(p.nblocks);
  run_blocks(p.tile_stride, C::NT, [&](HostCtx& ctx) { ypass_body<N, DIR>(ctx, smem.data(), p); });
}

// srcs[3], dsts[6]: pointers (may be null); kdsts[nranks] (forward scatter); jobs: njobs x (src, q, dst)
extern "C" int emu_ypass(int N, int dir, int rank, int nranks, double** srcs, double** dsts, double** kdsts,
                         int dst_klayout, const int* jobs, int njobs, int with_nyq, const double* tw) {
  YPassParams p{};
  for (int i = 0; i < 3; i++) p.src[i] = (const double2*)srcs[i];
  for (int i = 0; i < 6; i++) p.dst[i] = dsts ? (double2*)dsts[i] : nullptr;
  if (kdsts)
    for (int r = 0; r < nranks; r++) p.kdst.r[r] = (double2*)kdsts[r];
  p.dst_klayout = dst_klayout;
  for (int j = 0; j < njobs; j++) p.job[j] = YJob{jobs[3 * j], jobs[3 * j + 1], jobs[3 * j + 2]};
  p.njobs = njobs;
  p.g = make_geom(N, rank, nranks);
  p.ly_shift = ilog2(p.g.ly);
  p.tw = (const double2*)tw;
  switch (N) {
#define X(LL)                                                         \
  case LL:                                                            \
    if (dir > 0) ypass_run<LL, +1>(p, with_nyq); else ypass_run<LL, -1>(p, with_nyq); \
    return 0;
    EMU_GRIDS(X)
#undef X
  }
  return 1;
}

template <int M> static void fill_pretw(ZSrc& zs) {
  constexpr int TPL = Plan<M, true>::TPL, RMAX = Plan<M, true>::RMAX;
  for (int s = 0; s < RMAX; s++) {
    const double a = 2.0 * PINB_PI * (double)TPL * s / (2.0 * M);
    zs.pretw[s] = make_double2(cos(a), sin(a));
  }
}

template <int N, bool TAB = false> static void collapse_run(CollapseParams p) {
  fill_pretw<N / 2>(p.zs);
  constexpr int M = N / 2, TL = ZCfg<M>::TL, CG = 6;
  using ZS = ZShape<M, TL, CG>;
  std::vector<double2> smem(ZS::fft_elems(6));
  std::vector<double> spl((size_t)p.spl_doubles + 2), scratch(2 * ZS::NT);
  run_blocks((long long)p.g.lx * N / TL, ZS::NT,
             [&](HostCtx& ctx) { zpass_collapse_body<M, TL, CG, HostCtx, 1, TAB>(ctx, smem.data(), spl.data(), scratch.data(), p); });
}

extern "C" int emu_zpass_collapse(int N, int nranks, double** srcs, const int* kzpow, int has_nyq, const double* dc,
                                  const double* spline, int nspl, int ismooth, float* fmax, int* rmax, double* sums,
                                  double** hdst, const double* tw) {
  CollapseParams p{};
  for (int k = 0; k < 6; k++) {
    p.zs.src[k] = (const double2*)srcs[k];
    p.zs.kzpow[k] = kzpow[k];
    p.hdst[k] = hdst ? (double2*)hdst[k] : nullptr;
  }
  p.zs.ncomp = 6;
  p.zs.has_nyq = has_nyq;
  p.zs.dc_add = dc;
  p.g = make_geom(N, 0, nranks);
  p.tw = (const double2*)tw;
  p.spline = spline;
  p.nspl = nspl;
  p.spl_doubles = (int)spline_table_doubles(nspl);
  p.ismooth = ismooth;
  p.Fmax = fmax;
  p.Rmax = rmax;
  p.sums = sums;
  switch (N) {
#define X(LL) case LL: collapse_run<LL>(p); return 0;
    EMU_GRIDS(X)
#undef X
  }
  return 1;
}

template <int N> static void zout_run(ZOutParams p) {
  fill_pretw<N / 2>(p.zs);
  constexpr int M = N / 2, TL = ZCfg<M>::TL;
  using ZS = ZShape<M, TL, 1>;
  std::vector<double2> smem(ZS::fft_elems(6));
  run_blocks((long long)p.g.lx * N / TL, ZS::NT, [&](HostCtx& ctx) { zpass_out_body<M, TL, 1>(ctx, smem.data(), p); });
}

extern "C" int emu_zpass_out(int N, int nranks, int ncomp, double** srcs, const int* kzpow, int has_nyq, const double* dc,
                             int mode, double** rdst, float** fdst, double** hsrc, const double* weight, double* acc,
                             const double* tw) {
  ZOutParams p{};
  for (int k = 0; k < ncomp; k++) {
    p.zs.src[k] = (const double2*)srcs[k];
    p.zs.kzpow[k] = kzpow[k];
    if (rdst) p.rdst[k] = (double2*)rdst[k];
    if (fdst) p.fdst[k] = fdst[k];
    if (hsrc) p.hsrc[k] = hsrc[k];
    if (weight) p.weight[k] = weight[k];
  }
  p.zs.ncomp = ncomp;
  p.zs.has_nyq = has_nyq;
  p.zs.dc_add = dc;
  p.g = make_geom(N, 0, nranks);
  p.tw = (const double2*)tw;
  p.mode = mode;
  p.acc = acc;
  switch (N) {
#define X(LL) case LL: zout_run<LL>(p); return 0;
    EMU_GRIDS(X)
#undef X
  }
  return 1;
}

template <int N> static void r2c_run(const ZR2CParams& p) {
  constexpr int M = N / 2, TL = ZCfg<M>::TL;
  using ZS = ZShape<M, TL, 1>;
  std::vector<double2> smem(ZS::fft_elems(1));
  run_blocks((long long)p.g.lx * N / TL, ZS::NT, [&](HostCtx& ctx) { zpass_r2c_body<M, TL>(ctx, smem.data(), p); });
}

extern "C" int emu_zpass_r2c(int N, int nranks, const double* src, double* dst, const double* tw) {
  ZR2CParams p{};
  p.src = (const double2*)src;
  p.dst = (double2*)dst;
  p.g = make_geom(N, 0, nranks);
  p.tw = (const double2*)tw;
  switch (N) {
#define X(LL) case LL: r2c_run<LL>(p); return 0;
    EMU_GRIDS(X)
#undef X
  }
  return 1;
}

extern "C" int emu_sources(int N, int nranks, double** h, double* s2, double* s31, double* s32, int lpt_order) {
  SourcesParams p{};
  Geom g = make_geom(N, 0, nranks);
  for (int k = 0; k < 6; k++) p.h[k] = h[k];
  p.s2 = s2;
  p.s31 = s31;
  p.s32 = s32;
  p.nrows = (size_t)g.lx * N;
  p.N = N;
  p.pitch = 2 * g.P;
  p.lpt_order = lpt_order;
  const int nt = 8, nb = 4;
  run_blocks(nb, nt, [&](HostCtx& ctx) { lpt_sources_body(ctx, nt * nb, p); });
  return 0;
}

extern "C" int emu_genic(int N, int rank, int nranks, const unsigned int* seeds, const double* pk, double box,
                         int fixed_ic, int paired_ic, double* kd) {
  GenicParams p{};
  p.seeds = seeds;
  p.pk = pk;
  p.kd = (double2*)kd;
  p.box = box;
  p.fixed_ic = fixed_ic;
  p.paired_ic = paired_ic;
  p.g = make_geom(N, rank, nranks);
  constexpr int NT = 8;
  std::vector<double> smem(2 * 12 * NT);
  run_blocks(((long long)N * p.g.ly + NT - 1) / NT, NT, [&](HostCtx& ctx) { genic_body<NT>(ctx, smem.data(), p); });
  return 0;
}

// ---- TABULATED_CT / ELL_SNG (collapse_table.cuh) ---------------------------------------------------
extern "C" int emu_ct_delta_vector(double* dv, int nd) {
  ct_default_delta_vector(dv, nd);
  return 0;
}
// the table of one smoothing radius: ct_build_body over all points (blocks of 8 host threads)
extern "C" int emu_ct_build(int model, const double* dv, int nd, int nxy, double bin_x, double ampl, const double* spline, int nspl,
                            double D_in, const double* cosmo4, int first, int npoints, double* table) {
  CTBuildParams p{};
  p.model = model;
  p.dv = dv;
  p.nd = nd;
  p.nxy = nxy;
  p.bin_x = bin_x;
  p.ampl = ampl;
  p.spline = spline;
  p.nspl = nspl;
  p.D_in = D_in;
  if (cosmo4) p.cosmo = SngCosmo{cosmo4[0], cosmo4[1], cosmo4[2], cosmo4[3], cosmo4[4], cosmo4[5], cosmo4[6]};  // 7 values
  p.table = table;
  p.npoints = first + npoints;
  // these bodies have no barrier: every (block, thread) pair is run as an independent call, spread over
  // the host cores
  constexpr int NT = 128;
  const int nw = (int)std::max(1u, std::thread::hardware_concurrency());
  std::vector<std::thread> th;
  for (int w = 0; w < nw; w++)
    th.emplace_back([&, w]() {
      for (int i = first + w; i < first + npoints; i += nw) {
        HostCtx ctx{i % NT, i / NT, NT, nullptr};
        ct_build_body(ctx, p);
      }
    });
  for (auto& x : th) x.join();
  return 0;
}
extern "C" long long emu_ct_knots_doubles(int nd) { return (long long)ct_knots_doubles(nd); }
extern "C" int emu_ct_pack_knots(const double* dv, int nd, double* out) {
  ct_pack_knots(dv, nd, out);
  return 0;
}
// spline records of all columns: coef holds ncols * (nd + 2) * 4 doubles, 32-byte aligned
extern "C" int emu_ct_spline(const double* dv, int nd, int ncols, const double* table, double* coef) {
  CTSplineParams p{dv, nd, ncols, table, reinterpret_cast<CTRec*>(coef)};
  constexpr int NT = 64;
  for (int col = 0; col < ncols; col++) {
    HostCtx ctx{col % NT, col / NT, NT, nullptr};
    ct_spline_body(ctx, p);
  }
  return 0;
}
static CTView make_ct_view(const double* knots, const double* coef, int nd, int nxy, double ampl, double bin_x) {
  CTView v{};
  v.knots = knots;
  v.coef = reinterpret_cast<const CTRec*>(coef);
  v.nd = nd;
  v.nxy = nxy;
  v.inv_ampl = 1.0 / ampl;
  v.inv_bin_x = 1.0 / bin_x;
  return v;
}
// per-cell evaluation on eigenvalues (which = 0) or on Hessians h6[6][n] (which = 1)
extern "C" int emu_ct_cells(int which, const double* in, long long n, const double* knots, const double* coef, int nd, int nxy,
                            double ampl, double bin_x, double* F) {
  const CTView v = make_ct_view(knots, coef, nd, nxy, ampl, bin_x);
  for (long long i = 0; i < n; i++) {
    if (which == 0) {
      F[i] = ct_interpolate(v, in[3 * i], in[3 * i + 1], in[3 * i + 2]);
    } else {
      double h[6];
      for (int c = 0; c < 6; c++) h[c] = in[c * n + i];
      F[i] = inverse_collapse_time_tab(h, v);
    }
  }
  return 0;
}
extern "C" double emu_ell_sng(double l1, double l2, double l3, double D_in, const double* cosmo4) {
  return ell_sng(l1, l2, l3, D_in, SngCosmo{cosmo4[0], cosmo4[1], cosmo4[2], cosmo4[3], cosmo4[4], cosmo4[5], cosmo4[6]});
}
// the collapse z pass with the table in place of ell_classic
extern "C" int emu_zpass_collapse_tab(int N, int nranks, double** srcs, const int* kzpow, int has_nyq, const double* dc,
                                      const double* knots, const double* coef, int nd, int nxy, double ampl, double bin_x, int ismooth,
                                      float* fmax, int* rmax, double* sums, double** hdst, const double* tw) {
  CollapseParams p{};
  for (int k = 0; k < 6; k++) {
    p.zs.src[k] = (const double2*)srcs[k];
    p.zs.kzpow[k] = kzpow[k];
    p.hdst[k] = hdst ? (double2*)hdst[k] : nullptr;
  }
  p.zs.ncomp = 6;
  p.zs.has_nyq = has_nyq;
  p.zs.dc_add = dc;
  p.g = make_geom(N, 0, nranks);
  p.tw = (const double2*)tw;
  p.ismooth = ismooth;
  p.Fmax = fmax;
  p.Rmax = rmax;
  p.sums = sums;
  p.ct = make_ct_view(knots, coef, nd, nxy, ampl, bin_x);
  switch (N) {
#define X(LL) case LL: collapse_run<LL, true>(p); return 0;
    EMU_GRIDS(X)
#undef X
  }
  return 1;
}

// the packed spline table, built by the same host code the engine uses (spline_pack.h)
extern "C" long long emu_spline_table_doubles(int n) { return (long long)spline_table_doubles(n); }
extern "C" int emu_pack_spline(const double* x, const double* y, int n, double* out) {
  std::vector<double> t;
  pack_spline(x, y, n, t);
  std::memcpy(out, t.data(), t.size() * sizeof(double));
  return 0;
}

extern "C" int emu_collapse_cells(const double* h6, long long n, const double* spline, int nspl, double* F) {
  SplineView sp{spline, nspl};
  for (long long i = 0; i < n; i++) {
    double h[6];
    for (int c = 0; c < 6; c++) h[c] = h6[c * n + i];
    F[i] = inverse_collapse_time(h, sp);
  }
  return 0;
}

// elementary functions of fastmath.cuh, exposed for direct comparison with libm
extern "C" int emu_fastmath(int which, const double* x, long long n, double* y) {
  for (long long i = 0; i < n; i++) {
    switch (which) {
      case 0: y[i] = fm_acos(x[i]); break;
      case 1: y[i] = fm_log10(x[i]); break;
      case 2: y[i] = fm_exp_neg(x[i]); break;
      case 3: y[i] = fm_exp10(x[i]); break;
      case 4: y[i] = fm_rcp(x[i]); break;
      case 5: y[i] = fm_div(x[2 * i], x[2 * i + 1]); break;  // x holds n (a, b) pairs
      case 6: y[i] = fm_sqrt(x[i]); break;
      case 7: y[i] = fm_sqrt_pair(x[i]).rs; break;
      case 8: y[i] = fm_cbrt_pair(x[i]).c; break;
      case 9: y[i] = fm_cbrt_pair(x[i]).rc; break;
      case 10: { double c0, c1, c2; cos_thirds(x[i], c0, c1, c2); y[3 * i] = c0; y[3 * i + 1] = c1; y[3 * i + 2] = c2; break; }  // y holds 3n
      default: return 1;
    }
  }
  return 0;
}

// ---- collapsed-cell filter + radix sort (sort_cells.cuh): the engine's schedule of
// pinb200_collapsed_cells on host arrays.  idx_out must hold n entries; returns the count.
extern "C" long long emu_collapsed_cells(const float* fmax, long long n, float f_last, unsigned int* idx_out) {
  const unsigned int ntiles_all = (unsigned int)((n + SORT_TILE - 1) / SORT_TILE);
  std::vector<unsigned int> counts((size_t)256 * ntiles_all), key[2], idx[2];
  std::vector<unsigned short> cnt(256 * SORT_ROW);
  std::vector<unsigned int> base(256), scratch(64);
  unsigned long long total = 0;
  SortPassParams p{};
  p.fmax = fmax;
  p.f_last = f_last;
  p.n = (unsigned long long)n;
  p.counts = counts.data();
  p.ntiles = ntiles_all;
  auto count_scan = [&](const SortPassParams& q) {
    run_blocks(q.ntiles, SORT_NT, [&](HostCtx& ctx) { sort_count_body(ctx, cnt.data(), q); });
    run_blocks(1, 64, [&](HostCtx& ctx) { sort_scan_body(ctx, scratch.data(), q.counts, (unsigned long long)256 * q.ntiles, &total); });
  };
  count_scan(p);
  const unsigned long long m = total;
  for (int b = 0; b < 2; b++) { key[b].assign(m ? m : 1, 0xdeadbeefu); idx[b].assign(m ? m : 1, 0xdeadbeefu); }
  int cur = 0;
  for (int pass = 0; pass < 4 && m > 0; pass++) {
    p.shift = 8 * pass;
    p.key_out = key[cur].data();
    p.idx_out = idx[cur].data();
    count_scan(p);
    run_blocks(p.ntiles, SORT_NT, [&](HostCtx& ctx) { sort_scatter_body(ctx, cnt.data(), base.data(), p); });
    p.fmax = nullptr;
    p.key_in = key[cur].data();
    p.idx_in = idx[cur].data();
    p.n = m;
    p.ntiles = (unsigned int)((m + SORT_TILE - 1) / SORT_TILE);
    cur ^= 1;
  }
  for (unsigned long long i = 0; i < m; i++) idx_out[i] = idx[cur ^ 1][i];
  return (long long)m;
}

// ---- batched set_scaledep_GM integrals (scaledep_gm.cuh): the two kernels of k_scaledep.cu, block by block ----
extern "C" int emu_scaledep_variances(int n, const double* logk, const double* a_dens, const double* a_disp, int nk, int nt,
                                      double logkmin, double dlogk, const double* lg, const double* fo, int ns,
                                      const double* r_dens, const double* r_disp, double* out) {
  constexpr int NT = 256, RC = 8;  // as k_scaledep.cu
  std::vector<double> T((size_t)2 * ns * n);
  SdgmParams p{};
  p.logk = logk; p.a_dens = a_dens; p.a_disp = a_disp; p.n = n;
  p.lg = lg; p.fo = fo; p.nk = nk; p.nt = nt; p.logkmin = logkmin; p.dlogk = dlogk;
  p.r_dens = r_dens; p.r_disp = r_disp; p.ns = ns; p.T = T.data(); p.out = out;
  const long long nw = 2ll * ns * n;
  run_blocks((nw + 255) / 256, 256, [&](HostCtx& ctx) { sdgm_window_body(ctx, p); });
  std::vector<double> scratch(NT);
  run_blocks(nt, NT, [&](HostCtx& ctx) { sdgm_integrate_body<NT, RC>(ctx, scratch.data(), p); });
  return 0;
}
