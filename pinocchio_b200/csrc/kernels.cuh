// Kernel bodies of the collapse-time hot path, written against an execution-context
// template parameter `Ctx` (tid / bid / sync / atomic_add) so that the identical source runs
// as sm_100a kernels (kernels.cu, DevCtx) and under the pthread block emulator used by the
// CPU-only unit tests (tests/host/emu.cpp, HostCtx).  No kernel here is a CPU fallback of the
// product: the shared library exports only the CUDA instantiations.
//
// Data layout (device, per rank):
//   half-complex field  : double2 [x][y][P]   P = N/2 + 8 (row pitch; kz = 0..N/2 valid)
//   real field          : double  [x][y][2P]  (z = 0..N-1 valid) -- same bytes, so the z pass
//                         is in place
//   K layout = k-space slab split along y : [x (N)][yl (ly)][P]      (x pass domain)
//   R layout = real-space slab split along x: [xl (lx)][y (N)][P]    (y and z pass domain)
// Reference counterparts: compute_derivative k loop (src/fmax-pfft.c:306-397), pfft_execute
// (src/fmax-pfft.c:191-228), compute_collapse_times cell loop (src/collapse_times.c:545-591),
// LPT sources/contraction (src/LPT.c:64-141), GenIC column loop (src/GenIC.c:188-411).
#pragma once
#include "collapse.cuh"
#include "collapse_table.cuh"
#include "fft_core.cuh"

#define PINB_NBINS_PDF 210 /* NBINS, src/pinocchio.h:65 */

namespace pinb {

struct Geom {
  int N, M, P;     // grid side, N/2, complex row pitch
  int lx, ly;      // local extents: x in R layout, y in K layout
  int x0, y0;      // global offsets of the local slabs
  double knorm;    // 2*pi/N  (k in grid units, src/fmax-pfft.c:290)
};

// Mode factor applied when the x pass loads delta_k (fused K2).
struct KFactor {
  const double* gauss;  // gauss[n] = exp(-0.5*(knorm*n)^2*Rs^2), n = 0..M; nullptr -> 1
  double scalar;        // 1/N^3 normalisation (src/fmax-pfft.c:224-225) times growth_rate
  int green;            // 1: divide by k^2, k=0 mode dropped here and re-added as a constant
  int times_i;          // 1: (re,im) -> (-im,re)  (first derivatives, src/fmax-pfft.c:389-394)
  // Scale-dependent growth (ScaleDep.order 1..4, src/fmax-pfft.c:344-364): growth_rate(|k|) =
  // gk_sign * 10^(interpolation, linear in log10 k, of gk[0..gk_n-1]) where gk[j] is
  // InterpolateGrowth's j-th k-bin spline evaluated at the segment redshift (src/cosmo.c:1728-1757),
  // |k| in grid units as the reference passes it.  nullptr: `scalar` already holds the growth rate.
  const double* gk;
  int gk_n;
  double gk_logkmin, gk_dlogk;  // LOGKMIN, DELTALOGK (src/def_splines.h:41-42)
  double gk_sign;               // -1 for GrowingMode_3LPT_1 (src/cosmo.c:1810)
};

// InterpolateGrowth (src/cosmo.c:1728-1757) at a fixed redshift, followed by the pow(10., .) of
// GrowingMode* (src/cosmo.c:1786-1819).  kmin = 10^LOGKMIN, kmax = 10^(LOGKMIN + (n-1) DELTALOGK)
// as src/cosmo.c:169-170 computes them.
// The reference evaluates log10(k) and pow(10., v) with libm for every mode; r02 measured the x pass with that
// loader at 35.8 ms against 9.6 ms without it (1024^3), i.e. the pass was bound by two libm calls per mode.  Here:
// log10 k = log10(k^2) / 2 (no square root; the clamps compare k^2 with kmin^2, kmax^2, set once per thread) and
// the constant-bank log10 / 10^x of fastmath.cuh (1e-16 relative, as libm to the last bits; the function is
// continuous across both clamps, so an ulp of difference at a boundary moves nothing).
struct GrowthRange {
  double kmin2 = 0.0, kmax2 = 0.0, inv_dlogk = 0.0;
  PINB_HD void set(const KFactor& kf) {
    const double kmin = fm_exp10(kf.gk_logkmin);
    const double kmax = fm_exp10(kf.gk_logkmin + (kf.gk_n - 1) * kf.gk_dlogk);
    kmin2 = kmin * kmin;
    kmax2 = kmax * kmax;
    inv_dlogk = 1.0 / kf.gk_dlogk;
  }
};
// k2 = |k|^2 (grid units, as the reference passes k_module, src/fmax-pfft.c:340)
PINB_HD double growth_of_k2(const KFactor& kf, const GrowthRange& gr, double k2) {
  double v;
  if (k2 < gr.kmin2) {
    v = kf.gk[0];
  } else if (k2 > gr.kmax2) {
    v = kf.gk[kf.gk_n - 1];
  } else {
    double dk = (0.5 * fm_log10(k2) - kf.gk_logkmin) * gr.inv_dlogk;
    int kk = (int)dk;
    kk = kk < 0 ? 0 : kk;                       // rounding of log10 just below kmin
    dk -= kk;
    // k == kmax gives kk = n-1 and dk = 0: the reference reads SPLINE[pointer+kk+1] there too
    // (the next table, times zero); here the weight-zero term is dropped instead
    v = (kk + 1 < kf.gk_n) ? dk * kf.gk[kk + 1] + (1 - dk) * kf.gk[kk < kf.gk_n ? kk : kf.gk_n - 1] : kf.gk[kf.gk_n - 1];
  }
  return kf.gk_sign * fm_exp10(v);
}

// Launch shapes shared by the CUDA instantiations and the host emulator.
// kz-tile width of the strided passes: 8 complex (128 B contiguous per line element) while the
// tile fits 128 KB of shared memory.
template <int L> struct StridedCfg { static constexpr int TK = (L <= 1024) ? 8 : 4; };
// Lines longer than PINB_SPLIT_ABOVE do not fit a TK = 8 tile.  With TK = 4 every line element
// is a 64-byte piece: the 2048^3 run on 8 GPUs scattered its x pass over NVLink at 280 GB/s (r01).
// Such passes run one decimation-in-frequency step while loading instead:
//   X[2k]   = FFT_{L/2}[ a[n] + a[n+L/2] ],   X[2k+1] = FFT_{L/2}[ (a[n] - a[n+L/2]) w_L^n ]
// i.e. two half-length jobs per output field on a TK = 8 tile (128-byte pieces on both sides);
// the second reading of the source tile comes from L2.  Not usable in place.
#ifndef PINB_SPLIT_ABOVE
#define PINB_SPLIT_ABOVE 1024
#endif
// LOCAL: the inverse x pass of the pipelined multi-GPU sweep (XPassParams::dst_klayout == 2) stores into LOCAL memory
// only -- the copy engines do the transpose -- so the 64-byte pieces of a TK = 4 tile never cross NVLink and the whole
// line fits one tile again: no decimation step, no second reading of the source (r02, slab of 2048^3 on one GPU:
// see profiles/r02_slabbench.txt).
template <int L, int DIR, bool LOCAL = false> struct XCfg {
  static constexpr bool SPLIT = (L > PINB_SPLIT_ABOVE) && DIR > 0 && !LOCAL;  // the forward x pass is in place (and local)
  static constexpr int LT = SPLIT ? L / 2 : L;
  static constexpr int TK = LT <= 1024 ? 8 : 4;
  static constexpr int CHUNK = SPLIT ? 4 : 0;  // see strided_tile_jobs (r01 slab benchmark: 41 -> 32 ms)
  using PL = XPlan<LT>;
  static constexpr int NT = PL::TPL * TK;
};
// The y pass stays local (no NVLink stores) and measured faster unsplit at N = 2048 (r01 slab
// benchmark: 26 ms with TK = 4 against 52 ms split), so its threshold is separate; the split y
// code is kept compiled and tested through the PINB_SPLIT_ABOVE_Y = 32 test build.
#ifndef PINB_SPLIT_ABOVE_Y
#define PINB_SPLIT_ABOVE_Y 4096
#endif
#ifndef PINB_YTK
#define PINB_YTK 8  // kz-tile width of the y pass for lines up to 1024 (tools/passbench explores 4)
#endif
template <int L> struct YCfg {
  static constexpr bool SPLIT = (L > PINB_SPLIT_ABOVE_Y);
  static constexpr int LT = SPLIT ? L / 2 : L;
  static constexpr int TK = LT <= 1024 ? PINB_YTK : 4;
  static constexpr int CHUNK = SPLIT ? 4 : 0;
  using PL = Plan<LT, false>;
  static constexpr int NT = PL::TPL * TK;
};
// rows per block of the z passes: >= 16 threads per component group on the tiny test grids,
// one row per block (three resident blocks per SM at N = 1024) for the production sizes.
template <int M> struct ZCfg { static constexpr int TL = M <= 32 ? 4 : (M <= 128 ? 2 : 1); };

PINB_HD int fold(int e, int N, int M) { return e > M ? e - N : e; }  // index N/2 stays +N/2
PINB_HD double ipow(double k, int p) { return p == 0 ? 1.0 : (p == 1 ? k : k * k); }

template <class T> PINB_HD T ld_ro(const T* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

// ---------------------------------------------------------------------------------------
// Strided tile FFT: TK adjacent lines (adjacent kz), line elements `stride` apart.
// Block = Plan::TPL * TK threads; shared memory = L*TK double2.
// ---------------------------------------------------------------------------------------
template <int L, int TK, int DIR, class Ctx, class LoadF, class StoreF>
PINB_HD void strided_tile_fft(Ctx& ctx, double2* s, const double2* __restrict__ tw, int twscale, LoadF load,
                              StoreF store) {
  using PL = Plan<L, false>;
  constexpr int TPL = PL::TPL, RMAX = PL::RMAX;
  const int tk = ctx.tid() % TK, jl = ctx.tid() / TK;
  double2 v[RMAX];
  auto s_in = [&](int e) { return s[e * TK + tk]; };
  auto s_out = [&](int e, double2 val) { s[e * TK + tk] = val; };
  auto g_in = [&](int e) { return load(e, tk); };
  auto g_out = [&](int e, double2 val) { store(e, tk, val); };
  stage_load<L, PL::R0, TPL, RMAX>(jl, v, g_in);
  stage_store<L, PL::R0, 1, DIR, TPL, RMAX>(jl, v, s_out, tw, twscale);
  ctx.sync();
  if constexpr (PL::NST == 2) {
    stage_load<L, PL::R1, TPL, RMAX>(jl, v, s_in);
    stage_store<L, PL::R1, PL::R0, DIR, TPL, RMAX>(jl, v, g_out, tw, twscale);
  } else {
    stage_load<L, PL::R1, TPL, RMAX>(jl, v, s_in);
    ctx.sync();
    stage_store<L, PL::R1, PL::R0, DIR, TPL, RMAX>(jl, v, s_out, tw, twscale);
    ctx.sync();
    stage_load<L, PL::R2, TPL, RMAX>(jl, v, s_in);
    stage_store<L, PL::R2, PL::R0 * PL::R1, DIR, TPL, RMAX>(jl, v, g_out, tw, twscale);
  }
  ctx.sync();  // shared memory free for the next job
}

// TMA descriptor of a pitched field seen as a 3-D tensor of doubles {2 P (kz, re/im interleaved), y, x} -- the K layout
// [N][ly][P] and the R layout [lx][N][P] alike -- with a box of one tile piece: 2 TK doubles of up to 256 consecutive
// line elements (x pass: along dimension 2, y pass: along dimension 1).  Opaque here (CUtensorMap, 128 bytes, filled by
// cuTensorMapEncodeTiled in k_strided.cu); the host emulator never touches it.
struct alignas(64) TileMap { unsigned long long opaque[16]; };
struct TileSrc {
  const TileMap* map;
  int c0, c1, c2;  // coordinates of the tile's first piece
  int adv;         // the dimension (1 or 2) along which the line runs
};
template <class Ctx, class = void> struct CtxHasTma { static constexpr bool value = false; };
template <class Ctx> struct CtxHasTma<Ctx, decltype((void)Ctx::kTma)> { static constexpr bool value = Ctx::kTma; };

// Software-pipelined sequence of jobs on one tile.  Job j takes the raw tile at src(j, e, tk)
// (pointer), transforms each element with xform(j, e, tk, raw) and stores with store(j, e, tk, v).
// The raw tile travels HBM/L2 -> shared memory with cp.async (no registers in flight); each
// thread copies exactly the elements its stage 0 will read, so the copies need no block barrier.
// As soon as the last stage of job j has pulled its operands into registers the shared-memory
// tile is free again and the copies of job j+1 are issued: they overlap the last butterflies and
// the global stores of job j (with one 128 KB tile per SM nothing else can hide that latency).
// CHUNK > 0: xform is applied in a separate rolled loop, CHUNK elements per iteration.
// use_tma (device contexts only): the tile is fetched by the SM's TMA unit instead -- one thread arms an mbarrier with
// the tile's byte count and issues L/256 cp.async.bulk.tensor copies (a box = 2 TK doubles x 256 line elements, the
// same dense [e][tk] layout in shared memory), everybody waits on the barrier's phase.  No per-thread address
// arithmetic or LDGSTS traffic through L1; tsrc(job) names the tensor map and the tile's coordinates.
template <int L, int TK, int DIR, class PL, int TWS, int CHUNK = 0, class Ctx, class BeginF, class SrcF, class XformF, class StoreF, class TsrcF>
PINB_HD void strided_tile_jobs(Ctx& ctx, double2* s, const double2* __restrict__ tw, int njobs, int tile0, int tile_stride, int ntiles,
                               BeginF begin_tile, SrcF src, XformF xform, StoreF store, bool use_tma, TsrcF tsrc) {
  constexpr int TPL = PL::TPL, RMAX = PL::RMAX;
  constexpr int T0 = L / PL::R0, NB0 = T0 / TPL;
  constexpr int ROWS = L < 256 ? L : 256, NBOX = L / ROWS;  // TMA boxes per tile
  if (tile0 >= ntiles) return;  // (uniform over the block)
  const int tk = ctx.tid() % TK, jl = ctx.tid() / TK;
  double2 v[RMAX];
  auto s_in = [&](int e) { return s[e * TK + tk]; };
  auto s_out = [&](int e, double2 val) { s[e * TK + tk] = val; };
  bool tma = false;
  unsigned long long* bar = nullptr;
  if constexpr (CtxHasTma<Ctx>::value) {
    tma = use_tma;
    if (tma) {
      bar = ctx.tile_barrier();
      if (ctx.tid() == 0) ctx.mbar_init(bar, 1);
      ctx.fence_async_proxy();
      ctx.sync();
    }
  }
  auto issue = [&](int tile, int job) {
    if constexpr (CtxHasTma<Ctx>::value) {
      if (tma) {
        if (ctx.tid() == 0) {
          const TileSrc t = tsrc(tile, job);
          ctx.mbar_expect_tx(bar, (unsigned int)(L * TK * sizeof(double2)));
#pragma unroll
          for (int b = 0; b < NBOX; b++)
            ctx.tma_load_3d(s + (size_t)b * ROWS * TK, t.map, t.c0, t.c1 + (t.adv == 1 ? b * ROWS : 0),
                            t.c2 + (t.adv == 2 ? b * ROWS : 0), bar);
        }
        return;
      }
    }
#pragma unroll
    for (int m = 0; m < NB0; m++)
#pragma unroll
      for (int r = 0; r < PL::R0; r++) {
        const int e = jl + m * TPL + r * T0;
        ctx.async_copy16(s + e * TK + tk, src(tile, job, e, tk));
      }
  };
  // The block walks the tiles tile0, tile0 + tile_stride, ... (< ntiles) and, on each, the jobs 0 .. njobs-1, as ONE
  // pipelined sequence: the copies of the next tile's first job are issued where those of the next job would be, so
  // only the very first load of a block is exposed (r02: with one tile per block that load -- 128 KB at one SM's share
  // of the bandwidth, 3.7 k cycles -- was 9 % of an x-pass tile and 4.5 % of a y-pass tile).  begin_tile(t) sets the
  // caller's per-tile state, which xform and store read; src / tsrc take the tile explicitly because they are called
  // for the NEXT tile while the current one is still being stored.
  int tile = tile0, job = 0;
  unsigned int phase = 0;
  issue(tile, 0);
  begin_tile(tile);
  while (true) {
    auto raw_in = [&](int e) { return xform(job, e, tk, s[e * TK + tk]); };
    auto g_out = [&](int e, double2 val) { store(job, e, tk, val); };
    int ntile = tile, njob = job + 1;
    if (njob == njobs) {
      njob = 0;
      ntile = tile + tile_stride;
    }
    const bool more = ntile < ntiles;
    ctx.mark(4 * job + 0);
    bool waited = false;
    if constexpr (CtxHasTma<Ctx>::value) {
      if (tma) {
        ctx.mbar_wait(bar, phase);  // the whole tile has landed
        waited = true;
      }
    }
    if (!waited) ctx.async_wait();  // my own elements have landed
    ctx.mark(4 * job + 1);
    if constexpr (CHUNK == 0) {
      stage_load<L, PL::R0, TPL, RMAX>(jl, v, raw_in);
    } else {
      // transform my own elements in place first, CHUNK at a time in a rolled loop, then load
      // them plainly: for xform functors that read global memory (the split passes) the fully
      // unrolled form spilled 256 B/thread to L2 and overflowed the instruction cache (r01 ncu)
#pragma unroll CHUNK
      for (int i = 0; i < NB0 * PL::R0; i++) {
        const int e = jl + (i / PL::R0) * TPL + (i % PL::R0) * T0;
        s[e * TK + tk] = xform(job, e, tk, s[e * TK + tk]);
      }
      stage_load<L, PL::R0, TPL, RMAX>(jl, v, s_in);
    }
    ctx.sync();  // every thread holds its raw elements: the tile may be overwritten
    stage_store<L, PL::R0, 1, DIR, TPL, RMAX>(jl, v, s_out, tw, TWS);
    ctx.sync();
    if constexpr (PL::NST == 3) {
      stage_load<L, PL::R1, TPL, RMAX>(jl, v, s_in);
      ctx.sync();
      stage_store<L, PL::R1, PL::R0, DIR, TPL, RMAX>(jl, v, s_out, tw, TWS);
      ctx.sync();
      stage_load<L, PL::R2, TPL, RMAX>(jl, v, s_in);
      if constexpr (CtxHasTma<Ctx>::value) { if (tma) ctx.fence_async_proxy(); }  // my accesses before the next TMA write
      ctx.sync();  // tile free
      ctx.mark(4 * job + 2);
      if (more) issue(ntile, njob);
      stage_store<L, PL::R2, PL::R0 * PL::R1, DIR, TPL, RMAX>(jl, v, g_out, tw, TWS);
      ctx.mark(4 * job + 3);
    } else {
      stage_load<L, PL::R1, TPL, RMAX>(jl, v, s_in);
      if constexpr (CtxHasTma<Ctx>::value) { if (tma) ctx.fence_async_proxy(); }
      ctx.sync();  // tile free
      ctx.mark(4 * job + 2);
      if (more) issue(ntile, njob);
      stage_store<L, PL::R1, PL::R0, DIR, TPL, RMAX, decltype(g_out), (PL::R1 >= 16)>(jl, v, g_out, tw, TWS);
      ctx.mark(4 * job + 3);
    }
    if (!more) break;
    if (njob == 0) begin_tile(ntile);
    tile = ntile;
    job = njob;
    phase ^= 1u;
  }
}

// ---------------------------------------------------------------------------------------
// X pass.  One block = (yl, kz tile).  Up to three jobs: dst[p] = FFT_x[ kx^p * fac * src ].
// ---------------------------------------------------------------------------------------
// Slab decomposition over the GPUs of one box (reference: 1-D slabs, src/initialization.c:1317-1325).
// The all-to-all transpose that PFFT performs with MPI inside every 3-D transform is fused into
// the pass that precedes it: the inverse x pass stores each output element straight into the
// R-layout buffer of the rank that owns that x (peer memory over NVLink, cudaIpc-mapped), and
// the forward y pass stores into the K-layout buffer of the rank that owns that y.
#define PINB_MAXR 8
struct PeerPtrs { double2* r[PINB_MAXR]; };

struct XPassParams {
  const double2* src;   // K layout (local)
  PeerPtrs dst[3];      // per power p of kx: base pointer of the destination field on every rank
  int dst_klayout;      // 0: scatter to the R layout of the owner rank (inverse transforms)
                        // 1: local K layout, dst[p].r[0] (forward transforms, in place allowed)
                        // 2: staged transpose -- own x range straight into the local R-layout field dst[p].r[1],
                        //    everything else into the local K-layout staging dst[p].r[0] (the copy engines move it)
  int lx_shift;         // log2(lx)
  int pmask;            // bit p set -> compute job p
  int ntiles_z;         // kz tiles per row that are processed (M/TK, +1 with the Nyquist tile)
  int nblocks;          // tiles of the pass, ly * ntiles_z (0: one tile per block, tile = block index)
  int tile_stride;      // blocks of the launch when they walk the tiles bid, bid + stride, ... (0: one tile per block)
  KFactor kf;
  Geom g;
  const double2* tw;    // N-th roots of unity
  int use_tma;          // the source tile comes through the TMA unit (tmap; set by the launcher, never by the caller)
  TileMap tmap;         // src as {2 P, ly, N} doubles, box {2 TK, 1, 256}
};

// MULTI = false: one rank; the owner look-up and the per-rank pointer table are compiled out
// (they cost registers: the 1024-point kernel spilled 128 bytes with them).
// GK = true: the scale-dependent growth rate is evaluated per mode (separate instantiation so that
// the Hessian pass keeps its register budget).
template <int L, int DIR, bool MULTI, class Ctx, bool GK = false, bool LOCAL = false>
PINB_HD void xpass_body(Ctx& ctx, double2* smem, const XPassParams& p) {
  using C = XCfg<L, DIR, LOCAL>;
  constexpr int TK = C::TK, LT = C::LT;
  constexpr bool SPLIT = C::SPLIT;
  const Geom& g = p.g;
  const size_t xstride = (size_t)g.ly * g.P;
  int jobs[3], npw = 0;
  for (int pw = 0; pw < 3; pw++)
    if ((p.pmask >> pw) & 1) jobs[npw++] = pw;
  // state of the tile being transformed and stored (begin_tile); a block walks several tiles, see strided_tile_jobs
  int yl = 0, kz0 = 0;
  const double2* src = p.src;
  size_t roff = 0;      // offset of (y, kz0) inside an R-layout x plane
  double kyz2 = 0.0, fyz = 0.0;  // per-thread invariants of the mode factor: everything that does not depend on x
  const int tk0 = ctx.tid() % TK;
  auto begin_tile = [&](int t) {
    yl = t / p.ntiles_z;
    kz0 = (t % p.ntiles_z) * TK;
    const int ny = fold(g.y0 + yl, g.N, g.M);
    const double ky = g.knorm * ny;
    src = p.src + (size_t)yl * g.P + kz0;
    roff = (size_t)(g.y0 + yl) * g.P + kz0;
    const double kz_t = g.knorm * (kz0 + tk0);  // kz index <= M, never folded
    kyz2 = ky * ky + kz_t * kz_t;
    fyz = p.kf.scalar;
    if (p.kf.gauss) fyz *= ld_ro(p.kf.gauss + (ny < 0 ? -ny : ny)) * ld_ro(p.kf.gauss + kz0 + tk0);
  };
  auto srcf = [&](int t, int, int e, int tk) {
    return p.src + (size_t)(t / p.ntiles_z) * g.P + (size_t)(t % p.ntiles_z) * TK + (size_t)e * xstride + tk;
  };
  GrowthRange grange;
  if constexpr (GK) grange.set(p.kf);
  auto mode = [&](int pw, int e, double2 c) {
    const int nx = fold(e, g.N, g.M);
    const double kx = g.knorm * nx;
    double f = fyz;
    if (p.kf.green) {
      const double k2 = kx * kx + kyz2;
      f = (k2 != 0.0) ? f / k2 : 0.0;  // (a MUFU seed + Newton reciprocal measured 30 % slower here)
    }
    if constexpr (GK) {
      const double k2 = kx * kx + kyz2;
      if (k2 != 0.0) f *= growth_of_k2(p.kf, grange, k2);  // k_module = sqrt(k2), src/fmax-pfft.c:340
    }
    if (p.kf.gauss) f *= ld_ro(p.kf.gauss + (nx < 0 ? -nx : nx));
    f *= ipow(kx, pw);
    c = cscale(c, f);
    if (p.kf.times_i) c = make_double2(-c.y, c.x);
    return c;
  };
  // SPLIT: job = 2*(index of the power) + h; h selects the even or the odd outputs
  auto xform = [&](int job, int e, int tk, double2 c) {
    if constexpr (!SPLIT) {
      return mode(jobs[job], e, c);
    } else {
      const int pw = jobs[job >> 1];
      const double2 a = mode(pw, e, c);
      const double2 b = mode(pw, e + LT, ld_ro(src + (size_t)(e + LT) * xstride + tk));
      if (!(job & 1)) return cadd(a, b);
      return cmul(csub(a, b), twiddle<DIR>(p.tw, e));
    }
  };
  auto storef = [&](int job, int k, int tk, double2 val) {
    const int e = SPLIT ? 2 * k + (job & 1) : k;
    const PeerPtrs& dp = p.dst[jobs[SPLIT ? (job >> 1) : job]];
    if (p.dst_klayout == 2 && e >= g.x0 && e < g.x0 + g.lx) {
      dp.r[1][(size_t)(e - g.x0) * ((size_t)g.N * g.P) + roff + tk] = val;
    } else if (p.dst_klayout) {
      dp.r[0][(size_t)e * xstride + (size_t)yl * g.P + kz0 + tk] = val;
    } else if (!MULTI) {
      dp.r[0][(size_t)e * ((size_t)g.N * g.P) + roff + tk] = val;
    } else {
      const int owner = e >> p.lx_shift, xl = e & (g.lx - 1);
      dp.r[owner][(size_t)xl * ((size_t)g.N * g.P) + roff + tk] = val;
    }
  };
  // GK: the loader (a log10, a table interpolation and a 10^x per mode on top of the division) is ~250 instructions;
  // unrolled over the 16 stage-0 elements of a thread it overflows the instruction cache (r02: 31 ms against 9.6 ms
  // without G(k), although the arithmetic is worth 5 ms), so it runs as a rolled pre-pass over the tile like the
  // split passes' combine step
  constexpr int CHUNK = (GK && C::CHUNK == 0) ? 4 : C::CHUNK;
  auto tsrcf = [&](int t, int) { return TileSrc{&p.tmap, 2 * (t % p.ntiles_z) * TK, t / p.ntiles_z, 0, 2}; };
  const int ntiles = p.nblocks > 0 ? p.nblocks : ctx.bid() + 1;
  strided_tile_jobs<LT, TK, DIR, typename C::PL, L / LT, CHUNK>(ctx, smem, p.tw, SPLIT ? 2 * npw : npw, ctx.bid(),
                                                                  p.tile_stride > 0 ? p.tile_stride : ntiles, ntiles, begin_tile, srcf,
                                                                  xform, storef, p.use_tma != 0, tsrcf);
}

// ---------------------------------------------------------------------------------------
// Y pass.  One block = (xl, kz tile).  Jobs: dst[j.dst] = FFT_y[ ky^q * src[j.src] ].
// ---------------------------------------------------------------------------------------
struct YJob { int src, q, dst; };
struct YPassParams {
  const double2* src[3];  // R layout (local)
  double2* dst[6];        // R layout (local) -- inverse transforms
  PeerPtrs kdst;          // forward transforms: K-layout destination field on every rank (job 0 only)
  int dst_klayout;        // 1: scatter job 0 to the K layout of the rank that owns each y
  int ly_shift;           // log2(ly)
  YJob job[6];
  int njobs;
  int ntiles_z;
  int nblocks;            // as in XPassParams: lx * ntiles_z, ...
  int tile_stride;        // ... and the blocks of the launch
  int nsrc;               // distinct source fields
  Geom g;
  const double2* tw;
  int use_tma;            // as in XPassParams
  TileMap tmap[3];        // src[i] as {2 P, N, lx} doubles, box {2 TK, 256, 1}
};

template <int L, int DIR, class Ctx>
PINB_HD void ypass_body(Ctx& ctx, double2* smem, const YPassParams& p) {
  using C = YCfg<L>;
  constexpr int TK = C::TK, LT = C::LT;
  constexpr bool SPLIT = C::SPLIT;
  const Geom& g = p.g;
  int xl = 0, kz0 = 0;
  size_t base = 0;  // state of the tile being transformed and stored
  auto begin_tile = [&](int t) {
    xl = t / p.ntiles_z;
    kz0 = (t % p.ntiles_z) * TK;
    base = (size_t)xl * g.N * g.P + kz0;
  };
  // SPLIT: job = 2*j + h, see XCfg
  auto srcf = [&](int t, int job, int e, int tk) {
    return p.src[p.job[SPLIT ? (job >> 1) : job].src] + (size_t)(t / p.ntiles_z) * g.N * g.P + (size_t)(t % p.ntiles_z) * TK +
           (size_t)e * g.P + tk;
  };
  // (r02: ky^q from an L1-resident table instead of fold + int -> double + multiply per element made the pass SLOWER,
  // 21.0 -> 26.4 ms at 1024^3 with bit-identical output, profiles/r02_slabbench_kpow.txt: the loads sit in the
  // dependency chain of stage 0 where the conversions do not)
  auto mode = [&](int q, int e, double2 c) {
    if (q) c = cscale(c, ipow(g.knorm * fold(e, g.N, g.M), q));
    return c;
  };
  auto xform = [&](int job, int e, int tk, double2 c) {
    if constexpr (!SPLIT) {
      return mode(p.job[job].q, e, c);
    } else {
      const YJob& yj = p.job[job >> 1];
      const double2 a = mode(yj.q, e, c);
      const double2 b = mode(yj.q, e + LT, ld_ro(p.src[yj.src] + base + (size_t)(e + LT) * g.P + tk));
      if (!(job & 1)) return cadd(a, b);
      return cmul(csub(a, b), twiddle<DIR>(p.tw, e));
    }
  };
  auto storef = [&](int job, int k, int tk, double2 val) {
    const int e = SPLIT ? 2 * k + (job & 1) : k;
    const int j = SPLIT ? (job >> 1) : job;
    if (p.dst_klayout) {
      const int owner = e >> p.ly_shift, yl = e & (g.ly - 1);
      p.kdst.r[owner][((size_t)(g.x0 + xl) * g.ly + yl) * g.P + kz0 + tk] = val;
    } else {
      p.dst[p.job[j].dst][base + (size_t)e * g.P + tk] = val;
    }
  };
  auto tsrcf = [&](int t, int job) {
    return TileSrc{&p.tmap[p.job[SPLIT ? (job >> 1) : job].src], 2 * (t % p.ntiles_z) * TK, 0, t / p.ntiles_z, 1};
  };
  const int ntiles = p.nblocks > 0 ? p.nblocks : ctx.bid() + 1;
  strided_tile_jobs<LT, TK, DIR, typename C::PL, L / LT, C::CHUNK>(ctx, smem, p.tw, SPLIT ? 2 * p.njobs : p.njobs, ctx.bid(),
                                                                     p.tile_stride > 0 ? p.tile_stride : ntiles, ntiles, begin_tile, srcf,
                                                                     xform, storef, p.use_tma != 0, tsrcf);
}

// ---------------------------------------------------------------------------------------
// Cross-GPU stream barrier: every rank publishes `epoch` into its slot of every peer's flag
// array (peer memory), then waits until all slots of its own array have reached `epoch`.
// One block, PINB_MAXR threads.  Kernels on a stream run in order, so this orders the
// remote stores of the pass before it against the readers after it.
// ---------------------------------------------------------------------------------------
struct BarrierParams {
  unsigned long long* flags[PINB_MAXR];  // flags[r] = flag array (PINB_MAXR slots) living on rank r
  int rank, nranks;
  unsigned long long epoch;
  int* error;  // set to 1 on timeout
  unsigned long long timeout_ns;  // how long a rank waits for its peers before it reports instead of hanging
};

#if defined(__CUDACC__)
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Ranks reach a collective call with the skew of whatever the host did before it (the reference's
// fragmentation is load-imbalanced by seconds), so the wait is long (default 10 minutes, pinb200_desc /
// PINB200_BARRIER_TIMEOUT_S); it exists only so that a dead peer turns into an error instead of a hung GPU.
// Once the sticky error flag is set every later barrier returns at once: the call fails at its next
// host-side check instead of waiting again.
__device__ __forceinline__ void barrier_body(const BarrierParams& p) {
  const int t = threadIdx.x;
  __threadfence_system();
  if (t < p.nranks) {
    volatile unsigned long long* remote = p.flags[t] + p.rank;
    *remote = p.epoch;
    __threadfence_system();
    volatile unsigned long long* mine = p.flags[p.rank] + t;
    const unsigned long long t0 = global_timer_ns();
    unsigned int spins = 0;
    while (*mine < p.epoch) {
      __nanosleep(200);
      if ((++spins & 1023u) == 0) {
        if (*reinterpret_cast<volatile int*>(p.error)) break;
        if (global_timer_ns() - t0 > p.timeout_ns) {  // a peer died; fail loudly instead of hanging the GPU
          *p.error = 1;
          break;
        }
      }
    }
  }
  __threadfence_system();
}
#endif

// ---------------------------------------------------------------------------------------
// Contiguous (z) lines in shared memory.  A line is owned by TPL threads (one or two warps, or a
// fraction of a warp on the tiny test grids); the barriers between the Stockham stages only
// involve those threads (`ctx.sync_line`: __syncwarp or a named bar.sync), so the components
// and rows of a block proceed independently instead of meeting at block-wide barriers.
// ---------------------------------------------------------------------------------------
// all stages of one line; `in0(e, s)` supplies the stage-0 inputs (s: see stage_load_s)
template <int M, int DIR, class Ctx, class In0>
PINB_HD void zline_fft_stages(Ctx& ctx, double2* ln, int jl, int line_id, const double2* __restrict__ tw, In0 in0,
                              bool in0_reads_smem) {
  using PL = Plan<M, true>;
  constexpr int TPL = PL::TPL, RMAX = PL::RMAX;
  double2 v[RMAX];
  auto s_in = [&](int e) { return ln[zpad(e)]; };
  auto s_out = [&](int e, double2 val) { ln[zpad(e)] = val; };
  stage_load_s<M, PL::R0, TPL, RMAX>(jl, v, in0);
  if (in0_reads_smem) ctx.sync_line(line_id, TPL);
  stage_store<M, PL::R0, 1, DIR, TPL, RMAX, decltype(s_out), true>(jl, v, s_out, tw, 2);
  ctx.sync_line(line_id, TPL);
  stage_load<M, PL::R1, TPL, RMAX>(jl, v, s_in);
  ctx.sync_line(line_id, TPL);
  stage_store<M, PL::R1, PL::R0, DIR, TPL, RMAX, decltype(s_out), true>(jl, v, s_out, tw, 2);
  ctx.sync_line(line_id, TPL);
  if constexpr (PL::NST == 3) {
    stage_load<M, PL::R2, TPL, RMAX>(jl, v, s_in);
    ctx.sync_line(line_id, TPL);
    stage_store<M, PL::R2, PL::R0 * PL::R1, DIR, TPL, RMAX, decltype(s_out), true>(jl, v, s_out, tw, 2);
    ctx.sync_line(line_id, TPL);
  }
}

// in-place FFT of a line that already sits in shared memory
template <int M, int DIR, class Ctx>
PINB_HD void zline_fft_smem(Ctx& ctx, double2* ln, int jl, const double2* __restrict__ tw, int line_id = 0) {
  auto s_in = [&](int e, int) { return ln[zpad(e)]; };
  zline_fft_stages<M, DIR>(ctx, ln, jl, line_id, tw, s_in, true);
}

struct ZSrc {
  const double2* src[6];  // R layout half-complex fields
  int kzpow[6];           // power of kz applied on load
  int ncomp;
  int has_nyq;            // 0: the kz = N/2 plane is known to be zero (delta_k-derived fields)
  const double* dc_add;   // device scalar added to every component in real space (or nullptr)
  double2 pretw[32];      // exp(2 pi i * TPL * s / N), s < RMAX (filled by the launcher): pre-twiddle
                          // w^e = w^jl * pretw[s] for the element e = jl + s*TPL of a thread
};

// c2r of a tile of TL rows x ncomp components into shared memory: on return
// smem[(c*TL + line)*PITCH + zpad(m)] = (x[2m], x[2m+1]).  Block = TPL*TL*CG threads.
// The row goes HBM -> shared memory with cp.async (no registers in flight, full memory-level
// parallelism); stage 0 then reads X[e] and its mirror X[M-e] from shared memory and applies the
// kz power and the half-complex -> packed-complex pre-twiddle on the fly.  Ends with a
// block-wide barrier.
template <int M, int TL, int CG, class Ctx>
PINB_HD void zpass_c2r_tile(Ctx& ctx, double2* smem, const ZSrc& zs, const Geom& g, size_t row0,
                            const double2* __restrict__ tw) {
  using PL = Plan<M, true>;
  constexpr int TPL = PL::TPL;
  constexpr int PITCH = ZLine<M>::PITCH;
  const int tid = ctx.tid();
  const int jl = tid % TPL, line = (tid / TPL) % TL, cg = tid / (TPL * TL);
  const int line_id = line + TL * cg;
  // issue the copies of every component group first, then transform group by group
  for (int c0 = 0; c0 < zs.ncomp; c0 += CG) {
    const int c = c0 + cg;
    double2* ln = smem + ((size_t)c * TL + line) * PITCH;
    const double2* src = zs.src[c] + (row0 + line) * g.P;
    for (int e = jl; e < M; e += TPL) ctx.async_copy16(ln + zpad(e), src + e);
    if (jl == 0) {
      if (zs.has_nyq) ctx.async_copy16(ln + zpad(M), src + M);
      else ln[zpad(M)] = make_double2(0.0, 0.0);
    }
  }
  ctx.async_wait();
  for (int c0 = 0; c0 < zs.ncomp; c0 += CG) {
    const int c = c0 + cg;
    double2* ln = smem + ((size_t)c * TL + line) * PITCH;
    const int pw = zs.kzpow[c];
    const double2 wj = ld_ro(tw + jl);
    auto s_in0 = [&](int e, int s) {
      double2 xk = ln[zpad(e)];
      double2 xmk = ln[zpad(M - e)];
      if (pw) {
        xk = cscale(xk, ipow(g.knorm * e, pw));
        xmk = cscale(xmk, ipow(g.knorm * (M - e), pw));
      }
      if (e == 0) return make_double2(xk.x + xmk.x, xk.x - xmk.x);  // only Re X[0], Re X[M] (App. A.5)
      double2 zk, zmk;
      c2r_pre_pair(xk, xmk, cmul(wj, zs.pretw[s]), zk, zmk);
      return zk;
    };
    ctx.sync_line(line_id, TPL);  // the copies of this line have landed
    zline_fft_stages<M, +1>(ctx, ln, jl, line_id, tw, s_in0, true);
  }
  ctx.sync();
}

template <int M, int TL, int CG> struct ZShape {
  static constexpr int TPL = Plan<M, true>::TPL;
  static constexpr int NT = TPL * TL * CG;
  static constexpr int PITCH = ZLine<M>::PITCH;
  static constexpr size_t fft_elems(int ncomp) { return (size_t)ncomp * TL * PITCH; }
};

// block-wide sum of two doubles through shared scratch (2*NT doubles): portable tree version
// (the device context uses warp shuffles when NT is a multiple of 32)
template <int NT, class Ctx> PINB_HD void block_sum2_tree(Ctx& ctx, double* scratch, double& a, double& b) {
  const int tid = ctx.tid();
  scratch[tid] = a;
  scratch[NT + tid] = b;
  ctx.sync();
  int n = NT;
  while (n > 1) {
    const int half = (n + 1) >> 1;
    if (tid < n - half) {
      scratch[tid] += scratch[tid + half];
      scratch[NT + tid] += scratch[NT + tid + half];
    }
    ctx.sync();
    n = half;
  }
  a = scratch[0];
  b = scratch[NT];
}

// ---------------------------------------------------------------------------------------
// Z pass + collapse (fused K4 last pass + K5 + K10).  Six components.
// ---------------------------------------------------------------------------------------
struct CollapseParams {
  ZSrc zs;
  Geom g;
  const double2* tw;
  const double* spline;  // packed table of spline_pack.h (global); staged to shared memory
  int nspl;              // knots
  int spl_doubles;       // doubles of the packed table
  int ismooth;
  float* Fmax;           // [lx][N][N]
  int* Rmax;
  double* sums;          // [2]: sum(delta), sum(delta^2)   (atomicAdd)
  double2* hdst[6];      // if hdst[0] != nullptr: store the six real fields (in place allowed)
  CTView ct;             // TABULATED_CT: the table of this radius (collapse_table.cuh); unused otherwise
};

// TAB = true: F comes from the collapse-time table (TABULATED_CT, src/collapse_times.c:751) instead of
// ell_classic + InverseGrowingMode; the spline is then neither staged nor read.
template <int M, int TL, int CG, class Ctx, int CPT = 1, bool TAB = false>
PINB_HD void zpass_collapse_body(Ctx& ctx, double2* smem, double* spl_s, double* scratch, const CollapseParams& p,
                                 const bool spline_global = false) {
  using ZS = ZShape<M, TL, CG>;
  constexpr int N = 2 * M, NT = ZS::NT, PITCH = ZS::PITCH;
  const int tid = ctx.tid();
  const size_t row0 = (size_t)ctx.bid() * TL;
  // spline_global (a compile-time constant at the call site): evaluate the table where it lies
  if (!spline_global && !TAB)
    for (int i = 2 * tid; i < p.spl_doubles; i += 2 * NT) ctx.async_copy16(spl_s + i, p.spline + i);  // even count
  zpass_c2r_tile<M, TL, CG>(ctx, smem, p.zs, p.g, row0, p.tw);  // waits for the copies, ends with a barrier
  SplineView sp{spline_global ? p.spline : spl_s, p.nspl};
  const double dc = p.zs.dc_add ? ld_ro(p.zs.dc_add) : 0.0;
  double sd = 0.0, sd2 = 0.0;
  // Warp-chunk form of the cell loop: warp w takes the 32-cell chunks w, w + NT/32, ...  The trip
  // count then depends on a value the compiler knows to be warp-uniform (ctx.warp_uniform), so
  // the loop is convergent control flow: the ~110 polynomial coefficients of the collapse
  // arithmetic can be fetched through the uniform datapath (LDCU.128 + UR operands) instead of
  // one LDC.64 into two vector registers each, in a kernel bound by issue slots and capped at 56
  // registers.
  constexpr bool CHUNKED = (CPT == 1) && (NT % 32 == 0) && ((TL * N) % 32 == 0);
  if constexpr (CHUNKED) {
    constexpr int NW = NT / 32, NCH = TL * N / 32;
    const int warp = ctx.warp_uniform(tid >> 5), lane = tid & 31;
    for (int ch = warp; ch < NCH; ch += NW) {
      const int idx = ch * 32 + lane;
      const int line = idx / N, z = idx % N, m = z >> 1;
      double h[6];
#pragma unroll
      for (int c = 0; c < 6; c++) {
        const double2 v = smem[((size_t)c * TL + line) * PITCH + zpad(m)];
        h[c] = ((z & 1) ? v.y : v.x) + dc;
      }
      const size_t cell = (row0 + line) * N + z;
      float fm = -10.0f;  // ismooth == 0: the -10 / -1 initialisation of src/collapse_times.c:468-469
      if (p.ismooth > 0) fm = p.Fmax[cell];
      const double delta = h[0] + h[1] + h[2];
      sd += delta;
      sd2 += delta * delta;
      double F;
#ifdef PINB_BENCH_NOEPI  // tools/passbench only: the z pass without the collapse arithmetic
      F = h[0] * h[1] + h[2] * h[3] + h[4] * h[5];
#else
      if constexpr (TAB) F = inverse_collapse_time_tab(h, p.ct);
      else F = inverse_collapse_time(h, sp);
#endif
      if ((double)fm < F) {  // running max, src/collapse_times.c:587-590
        p.Fmax[cell] = (float)F;
        p.Rmax[cell] = p.ismooth;
      } else if (p.ismooth == 0) {
        p.Fmax[cell] = fm;
        p.Rmax[cell] = -1;
      }
    }
  } else
  // CPT independent cells per thread and iteration (instruction-level parallelism for the long
  // FP64 dependency chains of the collapse arithmetic)
  for (int base = tid; base < TL * N; base += CPT * NT) {
    double F[CPT];
    float fm[CPT];
    size_t cell[CPT];
    bool valid[CPT];
#pragma unroll
    for (int u = 0; u < CPT; u++) {
      const int idx = base + u * NT;
      valid[u] = idx < TL * N;
      const int idc = valid[u] ? idx : base;
      const int line = idc / N, z = idc % N, m = z >> 1;
      double h[6];
#pragma unroll
      for (int c = 0; c < 6; c++) {
        const double2 v = smem[((size_t)c * TL + line) * PITCH + zpad(m)];
        h[c] = ((z & 1) ? v.y : v.x) + dc;
      }
      cell[u] = (row0 + line) * N + z;
      // ismooth == 0 performs the -10 / -1 initialisation of src/collapse_times.c:468-469
      fm[u] = -10.0f;
      if (p.ismooth > 0) fm[u] = p.Fmax[cell[u]];
      const double delta = h[0] + h[1] + h[2];
      if (valid[u]) {
        sd += delta;
        sd2 += delta * delta;
      }
      if constexpr (TAB) F[u] = inverse_collapse_time_tab(h, p.ct);
      else F[u] = inverse_collapse_time(h, sp);
    }
#pragma unroll
    for (int u = 0; u < CPT; u++) {
      if (!valid[u]) continue;
      // running max, src/collapse_times.c:587-590 (float Fmax promoted to double for the test)
      if ((double)fm[u] < F[u]) {
        p.Fmax[cell[u]] = (float)F[u];
        p.Rmax[cell[u]] = p.ismooth;
      } else if (p.ismooth == 0) {
        p.Fmax[cell[u]] = fm[u];
        p.Rmax[cell[u]] = -1;
      }
    }
  }
  ctx.template block_sum2<NT>(scratch, sd, sd2);
  if (tid == 0) {
    ctx.atomic_add(p.sums + 0, sd);
    ctx.atomic_add(p.sums + 1, sd2);
  }
  if (p.hdst[0]) {
    constexpr int TPLG = NT / TL;  // threads cooperating on one row in the copy-out
    const int line = tid / TPLG, jl = tid % TPLG;
    for (int c = 0; c < 6; c++) {
      double2* dst = p.hdst[c] + (row0 + line) * p.g.P;
      const double2* ln = smem + ((size_t)c * TL + line) * PITCH;
      for (int e = jl; e < M; e += TPLG) {
        double2 v = ln[zpad(e)];
        dst[e] = make_double2(v.x + dc, v.y + dc);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Z pass with generic epilogues.
// ---------------------------------------------------------------------------------------
struct ZOutParams {
  ZSrc zs;
  Geom g;
  const double2* tw;
  int mode;              // 0: store real fields (rdst)   1: float displacement (fdst)
                         // 2: contraction  acc -= sum_c w[c] * val_c * H_c   (src/LPT.c:134-137)
  double2* rdst[6];      // mode 0: real fields, R layout (row pitch P double2)
  float* fdst[6];        // mode 1: float fields [lx][N][N]
  const double* hsrc[6]; // mode 2: Hessian real fields (row pitch 2P doubles)
  double weight[6];      // mode 2
  double* acc;           // mode 2: real field (row pitch 2P doubles)
};

template <int M, int TL, int CG, class Ctx>
PINB_HD void zpass_out_body(Ctx& ctx, double2* smem, const ZOutParams& p) {
  using ZS = ZShape<M, TL, CG>;
  constexpr int N = 2 * M, NT = ZS::NT, PITCH = ZS::PITCH;
  const int tid = ctx.tid();
  const size_t row0 = (size_t)ctx.bid() * TL;
  zpass_c2r_tile<M, TL, CG>(ctx, smem, p.zs, p.g, row0, p.tw);
  const double dc = p.zs.dc_add ? ld_ro(p.zs.dc_add) : 0.0;
  const int nc = p.zs.ncomp;
  if (p.mode == 0) {
    constexpr int TPLG = NT / TL;
    const int line = tid / TPLG, jl = tid % TPLG;
    for (int c = 0; c < nc; c++) {
      double2* dst = p.rdst[c] + (row0 + line) * p.g.P;
      const double2* ln = smem + ((size_t)c * TL + line) * PITCH;
      for (int e = jl; e < M; e += TPLG) {
        double2 v = ln[zpad(e)];
        dst[e] = make_double2(v.x + dc, v.y + dc);
      }
    }
    return;
  }
  for (int idx = tid; idx < TL * N; idx += NT) {
    const int line = idx / N, z = idx % N, m = z >> 1;
    const size_t row = row0 + line;
    if (p.mode == 1) {
      for (int c = 0; c < nc; c++) {
        const double2 v = smem[((size_t)c * TL + line) * PITCH + zpad(m)];
        p.fdst[c][row * N + z] = (float)(((z & 1) ? v.y : v.x) + dc);
      }
    } else {
      const size_t rcell = row * (2 * (size_t)p.g.P) + z;
      double a = p.acc[rcell];
      for (int c = 0; c < nc; c++) {
        const double2 v = smem[((size_t)c * TL + line) * PITCH + zpad(m)];
        a -= p.weight[c] * (((z & 1) ? v.y : v.x) + dc) * ld_ro(p.hsrc[c] + rcell);
      }
      p.acc[rcell] = a;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Z pass forward (r2c), one component, in place allowed.
// ---------------------------------------------------------------------------------------
struct ZR2CParams {
  const double2* src;  // real field viewed as double2 (row pitch P)
  double2* dst;        // half-complex field
  Geom g;
  const double2* tw;
};

template <int M, int TL, class Ctx>
PINB_HD void zpass_r2c_body(Ctx& ctx, double2* smem, const ZR2CParams& p) {
  using PL = Plan<M, true>;
  constexpr int TPL = PL::TPL, PITCH = ZLine<M>::PITCH;
  const int tid = ctx.tid();
  const int jl = tid % TPL, line = tid / TPL;
  const size_t row = (size_t)ctx.bid() * TL + line;
  double2* ln = smem + (size_t)line * PITCH;
  const double2* src = p.src + row * p.g.P;
  for (int e = jl; e < M; e += TPL) ctx.async_copy16(ln + zpad(e), src + e);
  ctx.async_wait();
  ctx.sync_line(line, TPL);
  zline_fft_smem<M, -1>(ctx, ln, jl, p.tw, line);
  double2* dst = p.dst + row * p.g.P;
  for (int k = jl; k <= M / 2; k += TPL) {
    if (k == 0) {
      const double2 z0 = ln[zpad(0)];
      dst[0] = make_double2(z0.x + z0.y, 0.0);
      dst[M] = make_double2(z0.x - z0.y, 0.0);
    } else {
      const double2 zk = ln[zpad(k)], zmk = ln[zpad(M - k)];
      double2 xk, xmk;
      r2c_post_pair(zk, zmk, ld_ro(p.tw + k), xk, xmk);
      dst[k] = xk;
      if (k != M / 2) dst[M - k] = xmk;
    }
  }
}

// ---------------------------------------------------------------------------------------
// LPT sources (K6, src/LPT.c:64-93): element-wise over real fields with row pitch 2P.
// ---------------------------------------------------------------------------------------
struct SourcesParams {
  const double* h[6];
  double *s2, *s31, *s32;
  size_t nrows;  // lx*N
  int N, pitch;  // pitch = 2P doubles
  int lpt_order;
};

template <class Ctx> PINB_HD void lpt_sources_body(Ctx& ctx, int nthreads_total, const SourcesParams& p) {
  const size_t total = p.nrows * (size_t)p.N;
  for (size_t i = (size_t)ctx.bid() * ctx.nthreads() + ctx.tid(); i < total; i += (size_t)nthreads_total) {
    const size_t row = i / p.N;
    const int z = (int)(i % p.N);
    const size_t a = row * p.pitch + z;
    const double s0 = p.h[0][a], s1 = p.h[1][a], s2 = p.h[2][a], s3 = p.h[3][a], s4 = p.h[4][a], s5 = p.h[5][a];
    const double src2 = s0 * s1 + s0 * s2 + s1 * s2 - s3 * s3 - s4 * s4 - s5 * s5;
    p.s2[a] = src2;
    if (p.lpt_order >= 3) {
      p.s31[a] = 3.0 * (s0 * (s1 * s2 - s5 * s5) - s3 * (s3 * s2 - s4 * s5) + s4 * (s3 * s5 - s4 * s1));
      p.s32[a] = 2.0 * (s0 + s1 + s2) * src2;
    }
  }
}

// ---------------------------------------------------------------------------------------
// GenIC (K1, src/GenIC.c:188-411 + gsl ranlxd1).
// One thread per (kx, ky) column; RANLUX state lives in shared memory, [12][NT] doubles per
// generator, two generators (column + k=0-plane mirror).
// ---------------------------------------------------------------------------------------
struct Ranlxd1 {
  double* x;   // x[k*stride]
  int stride;
  double carry;
  int ir, jr, ir_old;

  PINB_HD void set(unsigned int seed) {
    const double one_bit = 1.0 / 281474976710656.0;
    if (seed == 0) seed = 1;
    // GSL reads the seed into a signed int; i%2, i/=2 then give the digits of |i|
    const long long si = (long long)(int)seed;
    unsigned long long mag = (unsigned long long)(si < 0 ? -si : si);
    unsigned int bits = (unsigned int)(mag & 0x7FFFFFFFull);  // 31 digits
    int ibit = 0, jbit = 18;
    for (int k = 0; k < 12; k++) {
      double v = 0.0;
      for (int m = 1; m <= 48; m++) {
        const unsigned int bi = (bits >> ibit) & 1u, bj = (bits >> jbit) & 1u;
        const double y = (double)((bi + 1u) & 1u);
        v += v + y;
        bits = (bits & ~(1u << ibit)) | ((bi ^ bj) << ibit);
        ibit = (ibit + 1) % 31;
        jbit = (jbit + 1) % 31;
      }
      x[k * stride] = one_bit * v;
    }
    carry = 0.0;
    ir = 11;
    jr = 7;
    ir_old = 0;
  }
  PINB_HD void step() {
    const double one_bit = 1.0 / 281474976710656.0;
    double y = x[jr * stride] - x[ir * stride] - carry;
    if (y < 0) { carry = one_bit; y += 1.0; } else carry = 0.0;
    x[ir * stride] = y;
    ir = (ir == 11) ? 0 : ir + 1;
    jr = (jr == 11) ? 0 : jr + 1;
  }
  PINB_HD double get_double() {
    ir = (ir == 11) ? 0 : ir + 1;
    if (ir == ir_old) {
      int k = 0;
      while (ir > 0) { step(); k++; }
      while (k < 202) { step(); k++; }  // ranlxd1 luxury level pr = 202
      ir_old = ir;
    }
    return x[ir * stride];
  }
};

struct GenicParams {
  const unsigned int* seeds;  // SEEDTABLE[jj*N + ii], whole plane (src/GenIC.c:234-235)
  const double* pk;           // P(k) at k = 2*pi*sqrt(m)/Box, m = |n|^2 <= (N/2)^2
  double2* kd;                // K layout, zero-initialised
  double box;                 // true Mpc (src/GenIC.c:82)
  int fixed_ic, paired_ic;
  Geom g;
};

template <int NT, class Ctx> PINB_HD void genic_body(Ctx& ctx, double* smem, const GenicParams& p) {
  const Geom& g = p.g;
  const int N = g.N, N2 = g.M;
  const long long col = (long long)ctx.bid() * NT + ctx.tid();
  if (col >= (long long)N * g.ly) return;
  const int ii = (int)(col / g.ly);
  const int jl = (int)(col % g.ly);
  const int jj = g.y0 + jl;
  if (ii == N2 || jj == N2) return;
  Ranlxd1 rng, k0;
  rng.x = smem + ctx.tid();
  rng.stride = NT;
  k0.x = smem + 12 * NT + ctx.tid();
  k0.stride = NT;
  rng.set(p.seeds[(size_t)jj * N + ii]);
  const double fac = pow(1. / p.box, 1.5);
  const double fac2 = pow((double)N, 3.0);
  const int nx = ii < N2 ? ii : ii - N, ny = jj < N2 ? jj : jj - N;
  double2* out = p.kd + ((size_t)ii * g.ly + jl) * g.P;
  for (int kk = 0; kk < N2; kk++) {
    double phase = rng.get_double() * 2 * PINB_PI;
    double ampl;
    do ampl = rng.get_double(); while (ampl == 0);
    if (ii == 0 && jj == 0 && kk == 0) continue;
    const long long m = (long long)nx * nx + (long long)ny * ny + (long long)kk * kk;
    if (m > (long long)N2 * N2) continue;  // |n| > N/2, src/GenIC.c:280
    double p_of_k = ld_ro(p.pk + m);
    double sign = 1.0;
    if (kk == 0) {
      if (ii == 0 && jj == N2) continue;
      if (ii > N2 || (ii == 0 && jj > N2)) {
        int jjj = N - jj;
        if (jjj == N) jjj = 0;
        const int iii = ii > N2 ? N - ii : ii;
        sign = -1.0;
        k0.set(p.seeds[(size_t)jjj * N + iii]);
        phase = k0.get_double() * 2 * PINB_PI;
        do ampl = k0.get_double(); while (ampl == 0);
      }
    }
    if (p.paired_ic) phase += PINB_PI;
    if (!p.fixed_ic) p_of_k *= -log(ampl);
    const double delta = fac * sqrt(p_of_k);
    out[kk] = make_double2(delta * cos(phase) * fac2, sign * delta * sin(phase) * fac2);
  }
}

}  // namespace pinb
