// sm_100a instantiations of the strided (x and y) FFT passes.  See kernels.cuh.
#include "devctx.cuh"
#include "launch.h"

namespace pinb {

template <int L, int TK, int DIR, bool MULTI>
__global__ void __launch_bounds__(XPlan<L>::TPL* TK) xpass_kernel(const __grid_constant__ XPassParams p) {
  extern __shared__ double2 smem[];
  DevCtx ctx;
  xpass_body<L, TK, DIR, MULTI>(ctx, smem, p);
}

template <int L, int TK, int DIR>
__global__ void __launch_bounds__(Plan<L, false>::TPL* TK) ypass_kernel(const __grid_constant__ YPassParams p) {
  extern __shared__ double2 smem[];
  DevCtx ctx;
  ypass_body<L, TK, DIR>(ctx, smem, p);
}

template <int L, int DIR, bool MULTI> static cudaError_t xpass_launch_m(const XPassParams& p, int nblocks_y, cudaStream_t s) {
  constexpr int TK = StridedCfg<L>::TK;
  constexpr int NT = XPlan<L>::TPL * TK;
  const size_t smem = (size_t)L * TK * sizeof(double2);
  cudaError_t e = allow_smem(xpass_kernel<L, TK, DIR, MULTI>, smem);
  if (e != cudaSuccess) return e;
  xpass_kernel<L, TK, DIR, MULTI><<<(unsigned)(nblocks_y * p.ntiles_z), NT, smem, s>>>(p);
  return cudaGetLastError();
}
template <int L, int DIR> static cudaError_t xpass_launch(const XPassParams& p, int nblocks_y, cudaStream_t s) {
  // forward transforms write the local K layout only; inverse ones scatter to the owner ranks
  if (DIR < 0 || p.g.lx == p.g.N) return xpass_launch_m<L, DIR, false>(p, nblocks_y, s);
  return xpass_launch_m<L, DIR, true>(p, nblocks_y, s);
}

template <int L, int DIR> static cudaError_t ypass_launch(const YPassParams& p, int nblocks_x, cudaStream_t s) {
  constexpr int TK = StridedCfg<L>::TK;
  constexpr int NT = Plan<L, false>::TPL * TK;
  const size_t smem = (size_t)L * TK * sizeof(double2);
  cudaError_t e = allow_smem(ypass_kernel<L, TK, DIR>, smem);
  if (e != cudaSuccess) return e;
  ypass_kernel<L, TK, DIR><<<(unsigned)(nblocks_x * p.ntiles_z), NT, smem, s>>>(p);
  return cudaGetLastError();
}

int xpass_tk(int N) {
  switch (N) {
#define X(L) case L: return StridedCfg<L>::TK;
    PINB_FOR_EACH_GRID(X)
#undef X
  }
  return 0;
}

bool grid_supported(int N) { return xpass_tk(N) != 0; }

cudaError_t launch_xpass(int N, int dir, const XPassParams& p, int nblocks_y, cudaStream_t s) {
  switch (N) {
#define X(L) case L: return dir > 0 ? xpass_launch<L, +1>(p, nblocks_y, s) : xpass_launch<L, -1>(p, nblocks_y, s);
    PINB_FOR_EACH_GRID(X)
#undef X
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_ypass(int N, int dir, const YPassParams& p, int nblocks_x, cudaStream_t s) {
  switch (N) {
#define X(L) case L: return dir > 0 ? ypass_launch<L, +1>(p, nblocks_x, s) : ypass_launch<L, -1>(p, nblocks_x, s);
    PINB_FOR_EACH_GRID(X)
#undef X
  }
  return cudaErrorInvalidValue;
}

}  // namespace pinb
