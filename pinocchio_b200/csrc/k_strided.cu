// sm_100a instantiations of the strided (x and y) FFT passes.  See kernels.cuh.
#include "devctx.cuh"
#include "launch.h"

namespace pinb {

template <int L, int DIR, bool MULTI>
__global__ void __launch_bounds__(XCfg<L, DIR>::NT) xpass_kernel(const __grid_constant__ XPassParams p) {
  extern __shared__ double2 smem[];
  DevCtx ctx;
  xpass_body<L, DIR, MULTI>(ctx, smem, p);
}

// inverse x pass with the scale-dependent growth rate evaluated per mode (displacement stage of
// SCALE_DEPENDENT runs only; the Hessian sweep uses xpass_kernel)
template <int L, bool MULTI>
__global__ void __launch_bounds__(XCfg<L, +1>::NT) xpass_growthk_kernel(const __grid_constant__ XPassParams p) {
  extern __shared__ double2 smem[];
  DevCtx ctx;
  xpass_body<L, +1, MULTI, DevCtx, true>(ctx, smem, p);
}

template <int L, int DIR>
__global__ void __launch_bounds__(YCfg<L>::NT) ypass_kernel(const __grid_constant__ YPassParams p) {
  extern __shared__ double2 smem[];
  DevCtx ctx;
  ypass_body<L, DIR>(ctx, smem, p);
}

template <int L, int DIR, bool MULTI> static cudaError_t xpass_launch_m(const XPassParams& p, int nblocks_y, cudaStream_t s) {
  using C = XCfg<L, DIR>;
  const size_t smem = (size_t)C::LT * C::TK * sizeof(double2);
  cudaError_t e = allow_smem(xpass_kernel<L, DIR, MULTI>, smem);
  if (e != cudaSuccess) return e;
  xpass_kernel<L, DIR, MULTI><<<(unsigned)(nblocks_y * p.ntiles_z), C::NT, smem, s>>>(p);
  return cudaGetLastError();
}
template <int L, bool MULTI> static cudaError_t xpass_growthk_launch_m(const XPassParams& p, int nblocks_y, cudaStream_t s) {
  using C = XCfg<L, +1>;
  const size_t smem = (size_t)C::LT * C::TK * sizeof(double2);
  cudaError_t e = allow_smem(xpass_growthk_kernel<L, MULTI>, smem);
  if (e != cudaSuccess) return e;
  xpass_growthk_kernel<L, MULTI><<<(unsigned)(nblocks_y * p.ntiles_z), C::NT, smem, s>>>(p);
  return cudaGetLastError();
}
template <int L, int DIR> static cudaError_t xpass_launch(const XPassParams& p, int nblocks_y, cudaStream_t s) {
  if constexpr (DIR > 0) {
    if (p.kf.gk) {
      if (p.g.lx == p.g.N) return xpass_growthk_launch_m<L, false>(p, nblocks_y, s);
      return xpass_growthk_launch_m<L, true>(p, nblocks_y, s);
    }
  }
  // forward transforms write the local K layout only; inverse ones scatter to the owner ranks
  if (DIR < 0 || p.g.lx == p.g.N) return xpass_launch_m<L, DIR, false>(p, nblocks_y, s);
  return xpass_launch_m<L, DIR, true>(p, nblocks_y, s);
}

template <int L, int DIR> static cudaError_t ypass_launch(const YPassParams& p, int nblocks_x, cudaStream_t s) {
  using C = YCfg<L>;
  const size_t smem = (size_t)C::LT * C::TK * sizeof(double2);
  cudaError_t e = allow_smem(ypass_kernel<L, DIR>, smem);
  if (e != cudaSuccess) return e;
  ypass_kernel<L, DIR><<<(unsigned)(nblocks_x * p.ntiles_z), C::NT, smem, s>>>(p);
  return cudaGetLastError();
}

// kz-tile width of the x pass in direction dir and of the y pass (they differ for N > 1024)
int xpass_tk(int N, int dir) {
  switch (N) {
#define X(L) case L: return dir > 0 ? XCfg<L, +1>::TK : XCfg<L, -1>::TK;
    PINB_FOR_EACH_GRID(X)
#undef X
  }
  return 0;
}

int ypass_tk(int N) {
  switch (N) {
#define X(L) case L: return YCfg<L>::TK;
    PINB_FOR_EACH_GRID(X)
#undef X
  }
  return 0;
}

bool grid_supported(int N) { return ypass_tk(N) != 0; }

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
