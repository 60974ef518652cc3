// sm_100a instantiations of the contiguous (z) passes: c2r + collapse, c2r + epilogues, r2c.
#include "devctx.cuh"
#include "launch.h"

namespace pinb {

// The collapse epilogue is a long serial FP64 dependency chain per cell: the kernel is latency
// bound and gains from occupancy (passbench: 1 block/SM 108 ms, 2 blocks 72 ms, 3 blocks 57 ms at
// 1024^3), so the register budget is capped to fit three 384-thread blocks per SM.
// M = 1024 (N = 2048): the six padded rows take 108 KB; with the spline table (10.5 KB) next to
// them only one block fits an SM (r01: 80 ms against 50 ms for the same cell count at M = 512).
// There the table stays in global memory (L1/L2 resident) and two blocks fit: 58.5 ms (80 registers,
// some spills in the FFT part) against 80.5 ms with one 162-register block.
#ifndef PINB_SPLINE_GLOBAL_FROM
#define PINB_SPLINE_GLOBAL_FROM 1024  // the "split" test build lowers it so that small grids cover the path
#endif
template <int M> struct CollapseCfg {
  static constexpr bool SPLINE_GLOBAL = (M >= PINB_SPLINE_GLOBAL_FROM);
  static constexpr int MINB = (M >= 256 && M <= 512) ? 3 : (M == 1024 ? 2 : 1);
};
template <int NT> constexpr size_t sum_scratch_bytes() { return (NT % 32 == 0 ? 64 : 2 * NT) * sizeof(double); }

template <int M, int TL, int CG>
__global__ void __launch_bounds__(ZShape<M, TL, CG>::NT, CollapseCfg<M>::MINB) zpass_collapse_kernel(const __grid_constant__ CollapseParams p) {
  extern __shared__ double2 smem[];
  using ZS = ZShape<M, TL, CG>;
  constexpr bool SG = CollapseCfg<M>::SPLINE_GLOBAL;
  double* after = reinterpret_cast<double*>(smem + ZS::fft_elems(6));
  double* spl = SG ? nullptr : after;
  double* scratch = SG ? after : after + p.spl_doubles;
  DevCtx ctx;
  zpass_collapse_body<M, TL, CG>(ctx, smem, spl, scratch, p, SG);
}

// TABULATED_CT variant: the epilogue reads F from the collapse-time table of this radius (four 32-byte
// gathers from L2 per cell, ~60 FP64 instructions instead of ~400), no spline in shared memory.
template <int M, int TL, int CG>
__global__ void __launch_bounds__(ZShape<M, TL, CG>::NT, CollapseCfg<M>::MINB) zpass_collapse_tab_kernel(const __grid_constant__ CollapseParams p) {
  extern __shared__ double2 smem[];
  using ZS = ZShape<M, TL, CG>;
  double* scratch = reinterpret_cast<double*>(smem + ZS::fft_elems(6));
  DevCtx ctx;
  zpass_collapse_body<M, TL, CG, DevCtx, 1, true>(ctx, smem, nullptr, scratch, p, true);
}

// all components of a row are transformed concurrently (CG = ncomp): a 64-thread block per row
// left the z epilogues latency bound (20.7 ms for the 39 GB float store at 1024^3)
// register budget: keep >= 768 resident threads per SM (the compiler otherwise takes 168 registers)
template <int NT> struct ZOutMinBlocks { static constexpr int V = NT >= 768 ? 1 : (768 / NT > 8 ? 8 : 768 / NT); };
template <int M, int TL, int CG>
__global__ void __launch_bounds__(ZShape<M, TL, CG>::NT, ZOutMinBlocks<ZShape<M, TL, CG>::NT>::V) zpass_out_kernel(const __grid_constant__ ZOutParams p) {
  extern __shared__ double2 smem[];
  DevCtx ctx;
  zpass_out_body<M, TL, CG>(ctx, smem, p);
}

template <int M, int TL>
__global__ void __launch_bounds__(ZShape<M, TL, 1>::NT) zpass_r2c_kernel(const __grid_constant__ ZR2CParams p) {
  extern __shared__ double2 smem[];
  DevCtx ctx;
  zpass_r2c_body<M, TL>(ctx, smem, p);
}

// pre-twiddle factors exp(2 pi i TPL s / N), s < RMAX (see ZSrc::pretw)
template <int M> static void fill_pretw(ZSrc& zs) {
  constexpr int TPL = Plan<M, true>::TPL, RMAX = Plan<M, true>::RMAX;
  static_assert(RMAX <= 32, "pretw too small");
  for (int s = 0; s < RMAX; s++) {
    const long double a = 2.0L * 3.141592653589793238462643383279502884L * (long double)TPL * s / (2.0L * M);
    zs.pretw[s] = make_double2((double)cosl(a), (double)sinl(a));
  }
}

template <int N> static cudaError_t collapse_launch(const CollapseParams& p_in, size_t nrows, cudaStream_t s) {
  CollapseParams p = p_in;
  fill_pretw<N / 2>(p.zs);
  constexpr int M = N / 2, TL = ZCfg<M>::TL, CG = 6;
  using ZS = ZShape<M, TL, CG>;
  if (p.ct.coef) {
    const size_t smem_tab = ZS::fft_elems(6) * sizeof(double2) + sum_scratch_bytes<ZS::NT>();
    cudaError_t e = allow_smem(zpass_collapse_tab_kernel<M, TL, CG>, smem_tab);
    if (e != cudaSuccess) return e;
    zpass_collapse_tab_kernel<M, TL, CG><<<(unsigned)(nrows / TL), ZS::NT, smem_tab, s>>>(p);
    return cudaGetLastError();
  }
  const size_t smem = ZS::fft_elems(6) * sizeof(double2) + (CollapseCfg<M>::SPLINE_GLOBAL ? 0 : (size_t)p.spl_doubles * sizeof(double)) +
                      sum_scratch_bytes<ZS::NT>();
  cudaError_t e = allow_smem(zpass_collapse_kernel<M, TL, CG>, smem);
  if (e != cudaSuccess) return e;
  zpass_collapse_kernel<M, TL, CG><<<(unsigned)(nrows / TL), ZS::NT, smem, s>>>(p);
  return cudaGetLastError();
}

// rows per block of the generic z epilogues: keep >= 256 threads per block when few components
// are transformed (a 64-thread block per row ran the 1-component contraction at 1.5 TB/s)
template <int M, int CG> struct ZOutCfg {
  static constexpr int T0 = ZCfg<M>::TL;
  static constexpr int TL = (M >= 256 && CG <= 2) ? 4 / CG : T0;
};

template <int N, int CG> static cudaError_t out_launch_cg(const ZOutParams& p_in, size_t nrows, cudaStream_t s) {
  ZOutParams p = p_in;
  fill_pretw<N / 2>(p.zs);
  constexpr int M = N / 2, TL = ZOutCfg<M, CG>::TL;
  using ZS = ZShape<M, TL, CG>;
  const size_t smem = ZS::fft_elems(p.zs.ncomp) * sizeof(double2);
  cudaError_t e = allow_smem(zpass_out_kernel<M, TL, CG>, smem);  // ncomp == CG in this dispatch
  if (e != cudaSuccess) return e;
  zpass_out_kernel<M, TL, CG><<<(unsigned)(nrows / TL), ZS::NT, smem, s>>>(p);
  return cudaGetLastError();
}

template <int N> static cudaError_t out_launch(const ZOutParams& p, size_t nrows, cudaStream_t s) {
  switch (p.zs.ncomp) {
    case 1: return out_launch_cg<N, 1>(p, nrows, s);
    case 2: return out_launch_cg<N, 2>(p, nrows, s);
    case 3: return out_launch_cg<N, 3>(p, nrows, s);
    case 6: return out_launch_cg<N, 6>(p, nrows, s);
  }
  return cudaErrorInvalidValue;
}

// rows per block of the forward z pass: four rows (256 threads) on the production grids
template <int M> struct ZR2CCfg { static constexpr int TL = M >= 256 ? 4 : ZCfg<M>::TL; };

template <int N> static cudaError_t r2c_launch(const ZR2CParams& p, size_t nrows, cudaStream_t s) {
  constexpr int M = N / 2, TL = ZR2CCfg<M>::TL;
  using ZS = ZShape<M, TL, 1>;
  const size_t smem = ZS::fft_elems(1) * sizeof(double2);
  cudaError_t e = allow_smem(zpass_r2c_kernel<M, TL>, smem);
  if (e != cudaSuccess) return e;
  zpass_r2c_kernel<M, TL><<<(unsigned)(nrows / TL), ZS::NT, smem, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_zpass_collapse(int N, const CollapseParams& p, size_t nrows, cudaStream_t s) {
  switch (N) {
#define X(L) case L: return collapse_launch<L>(p, nrows, s);
    PINB_FOR_EACH_GRID(X)
#undef X
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_zpass_out(int N, const ZOutParams& p, size_t nrows, cudaStream_t s) {
  switch (N) {
#define X(L) case L: return out_launch<L>(p, nrows, s);
    PINB_FOR_EACH_GRID(X)
#undef X
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_zpass_r2c(int N, const ZR2CParams& p, size_t nrows, cudaStream_t s) {
  switch (N) {
#define X(L) case L: return r2c_launch<L>(p, nrows, s);
    PINB_FOR_EACH_GRID(X)
#undef X
  }
  return cudaErrorInvalidValue;
}

}  // namespace pinb
