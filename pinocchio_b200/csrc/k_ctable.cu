// sm_100a instantiations of the collapse-time table kernels (collapse_table.cuh): table points
// (ELL_CLASSIC or the ELL_SNG batch ODE job), spline records per column, and the per-cell look-up as a
// stand-alone kernel for the cell-level parity entry point.  SURVEY.md section 8 row a19.
#include "devctx.cuh"
#include "launch.h"

namespace pinb {

// One thread per table point.  ELL_SNG: ~10^2..10^3 adaptive rkf45 steps of a 9-variable system per point,
// all state in registers / local memory; consecutive threads are consecutive delta knots of one (x, y)
// column, so the step counts within a warp are similar.  250 000 points = 1954 blocks of 128 threads
// (13 per SM on 148 SMs).
__global__ void __launch_bounds__(128) ct_build_kernel(const __grid_constant__ CTBuildParams p) {
  DevCtx ctx;
  ct_build_body(ctx, p);
}

// One thread per (x, y) column: a 100-knot tridiagonal solve in local memory (2500 threads in all).
__global__ void __launch_bounds__(64) ct_spline_kernel(const __grid_constant__ CTSplineParams p) {
  DevCtx ctx;
  ct_spline_body(ctx, p);
}

__global__ void __launch_bounds__(128) collapse_cells_tab_kernel(const double* __restrict__ h6, size_t n, const __grid_constant__ CTView v,
                                                                 double* __restrict__ F) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double h[6];
#pragma unroll
    for (int c = 0; c < 6; c++) h[c] = h6[c * n + i];
    F[i] = inverse_collapse_time_tab(h, v);
  }
}

cudaError_t launch_ct_build(const CTBuildParams& p, cudaStream_t s) {
  ct_build_kernel<<<(unsigned)((p.npoints + 127) / 128), 128, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_ct_spline(const CTSplineParams& p, cudaStream_t s) {
  ct_spline_kernel<<<(unsigned)((p.ncols + 63) / 64), 64, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_collapse_cells_tab(const double* h6, size_t n, const CTView& v, double* F, cudaStream_t s) {
  size_t nb = (n + 127) / 128;
  if (nb > 148 * 16) nb = 148 * 16;
  if (nb < 1) nb = 1;
  collapse_cells_tab_kernel<<<(unsigned)nb, 128, 0, s>>>(h6, n, v, F);
  return cudaGetLastError();
}

}  // namespace pinb
