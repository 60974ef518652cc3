// sm_100a kernels that are not FFT passes: GenIC, LPT sources, product packing, histogram,
// layout converters at the host boundary, and the cell-wise collapse entry used by tests.
#include "devctx.cuh"
#include "launch.h"

namespace pinb {

static constexpr int GENIC_NT = 128;

__global__ void __launch_bounds__(GENIC_NT) genic_kernel(const __grid_constant__ GenicParams p) {
  __shared__ double state[2 * 12 * GENIC_NT];
  DevCtx ctx;
  genic_body<GENIC_NT>(ctx, state, p);
}

cudaError_t launch_genic(const GenicParams& p, cudaStream_t s) {
  const long long ncol = (long long)p.g.N * p.g.ly;
  genic_kernel<<<(unsigned)((ncol + GENIC_NT - 1) / GENIC_NT), GENIC_NT, 0, s>>>(p);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) sources_kernel(const __grid_constant__ SourcesParams p) {
  DevCtx ctx;
  lpt_sources_body(ctx, (int)(gridDim.x * blockDim.x), p);
}

cudaError_t launch_sources(const SourcesParams& p, cudaStream_t s) {
  const size_t total = p.nrows * (size_t)p.N;
  size_t nb = (total + 255) / 256;
  if (nb > 148 * 64) nb = 148 * 64;  // grid-stride, whole multiples of the SM count
  sources_kernel<<<(unsigned)nb, 256, 0, s>>>(p);
  return cudaGetLastError();
}

// gauss[n] = exp(-0.5 * (knorm*n)^2 * Rs^2), n = 0..M  (window of src/fmax-pfft.c:372, separable)
__global__ void gauss_kernel(double* gauss, int M, double knorm, double rs) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n <= M) {
    const double k = knorm * n;
    gauss[n] = exp(-0.5 * (k * k) * rs * rs);
  }
}

cudaError_t launch_gauss_table(double* gauss, int M, double knorm, double rsmooth, cudaStream_t s) {
  gauss_kernel<<<(M + 1 + 127) / 128, 128, 0, s>>>(gauss, M, knorm, rsmooth);
  return cudaGetLastError();
}

// The reference leaves the k = 0 mode unscaled by the Green function (src/fmax-pfft.c:368);
// after the c2r it is the constant Re(c0)/N^3 (or -Im(c0)/N^3 after the i-swap) in real space.
__global__ void dc_kernel(const double2* src, double* out, double scale, int times_i) {
  const double2 c = src[0];
  *out = scale * (times_i ? -c.y : c.x);
}

cudaError_t launch_dc_scalar(const double2* src, double* out, double scale, int times_i, cudaStream_t s) {
  dc_kernel<<<1, 1, 0, s>>>(src, out, scale, times_i);
  return cudaGetLastError();
}

// does a half-complex field carry anything on its kz = N/2 plane?  (flag |= 1)
__global__ void __launch_bounds__(256) nyq_probe_kernel(const double2* __restrict__ f, size_t nrows, int pitch, int M, int* flag) {
  bool any = false;
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (size_t)gridDim.x * blockDim.x) {
    const double2 v = f[r * pitch + M];
    any |= (v.x != 0.0) || (v.y != 0.0);
  }
  if (any) *flag = 1;
}

cudaError_t launch_nyq_probe(const double2* f, size_t nrows, int pitch, int M, int* flag, cudaStream_t s) {
  size_t nb = (nrows + 255) / 256;
  if (nb > 148 * 8) nb = 148 * 8;
  nyq_probe_kernel<<<(unsigned)nb, 256, 0, s>>>(f, nrows, pitch, M, flag);
  return cudaGetLastError();
}

// Fmax_PDF, src/fmax.c:509-550
__global__ void __launch_bounds__(256) pdf_kernel(const float* __restrict__ fmax, size_t n, unsigned long long* counts) {
  __shared__ unsigned int h[PINB_NBINS_PDF];
  for (int i = threadIdx.x; i < PINB_NBINS_PDF; i += blockDim.x) h[i] = 0;
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int xF = (int)(fmax[i] * 10.);
    if (xF < 0) xF = 0;
    if (xF >= PINB_NBINS_PDF) xF = PINB_NBINS_PDF - 1;
    atomicAdd(&h[xF], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PINB_NBINS_PDF; i += blockDim.x)
    if (h[i]) atomicAdd(&counts[i], (unsigned long long)h[i]);
}

cudaError_t launch_fmax_pdf(const float* fmax, size_t n, unsigned long long* counts, cudaStream_t s) {
  size_t nb = (n + 256 * 64 - 1) / (256 * 64);
  if (nb > 148 * 8) nb = 148 * 8;
  if (nb < 1) nb = 1;
  pdf_kernel<<<(unsigned)nb, 256, 0, s>>>(fmax, n, counts);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(128) collapse_cells_kernel(const double* __restrict__ h6, size_t n, const double* __restrict__ spline,
                                                             int nspl, double* __restrict__ F) {
  extern __shared__ double spl[];
  const int nd = (PINB_SPLINE_HDR + 5 * nspl + PINB_SPLINE_NLUT / 4 + 1) & ~1;  // spline_table_doubles(nspl)
  for (int i = threadIdx.x; i < nd; i += blockDim.x) spl[i] = spline[i];
  __syncthreads();
  SplineView sp{spl, nspl};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double h[6];
#pragma unroll
    for (int c = 0; c < 6; c++) h[c] = h6[c * n + i];
    F[i] = inverse_collapse_time(h, sp);
  }
}

cudaError_t launch_collapse_cells(const double* h6, size_t n, const double* spline, int nspl, double* F, cudaStream_t s) {
  size_t nb = (n + 127) / 128;
  if (nb > 148 * 16) nb = 148 * 16;
  if (nb < 1) nb = 1;
  collapse_cells_kernel<<<(unsigned)nb, 128, spline_table_doubles(nspl) * sizeof(double), s>>>(h6, n, spline, nspl, F);
  return cudaGetLastError();
}

// SoA (device) -> AoS product_data records (K8, src/fmax-pfft.c:563-631)
__global__ void __launch_bounds__(256) pack_kernel(const __grid_constant__ PackParams p) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.ncells; i += (size_t)gridDim.x * blockDim.x) {
    const size_t cell = p.gather ? (size_t)p.gather[p.cell_begin + i] : p.cell_begin + i;
    unsigned char* rec = p.out + i * p.stride;
    if (p.off_rmax >= 0 && p.rmax) *reinterpret_cast<int*>(rec + p.off_rmax) = p.rmax[cell];
    if (p.prodfloat_bytes == 4) {
      if (p.off_fmax >= 0 && p.fmax) *reinterpret_cast<float*>(rec + p.off_fmax) = p.fmax[cell];
      for (int v = 0; v < 4; v++)
        if (p.off_vel[v] >= 0 && p.vel[3 * v])
          for (int a = 0; a < 3; a++) reinterpret_cast<float*>(rec + p.off_vel[v])[a] = p.vel[3 * v + a][cell];
    } else {
      if (p.off_fmax >= 0 && p.fmax) *reinterpret_cast<double*>(rec + p.off_fmax) = (double)p.fmax[cell];
      for (int v = 0; v < 4; v++)
        if (p.off_vel[v] >= 0 && p.vel[3 * v])
          for (int a = 0; a < 3; a++) reinterpret_cast<double*>(rec + p.off_vel[v])[a] = (double)p.vel[3 * v + a][cell];
    }
  }
}

cudaError_t launch_pack_products(const PackParams& p, cudaStream_t s) {
  size_t nb = (p.ncells + 255) / 256;
  if (nb > 148 * 32) nb = 148 * 32;
  if (nb < 1) nb = 1;
  pack_kernel<<<(unsigned)nb, 256, 0, s>>>(p);
  return cudaGetLastError();
}

template <class T>
__global__ void __launch_bounds__(256) repitch_kernel(const T* __restrict__ src, T* __restrict__ dst, size_t nrows, int ncols, int spitch, int dpitch) {
  const size_t total = nrows * (size_t)ncols;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / ncols;
    const int c = (int)(i % ncols);
    dst[r * dpitch + c] = src[r * spitch + c];
  }
}

cudaError_t launch_repitch_c(const double2* src, double2* dst, size_t nrows, int ncols, int spitch, int dpitch, cudaStream_t s) {
  size_t nb = (nrows * ncols + 255) / 256;
  if (nb > 148 * 32) nb = 148 * 32;
  repitch_kernel<double2><<<(unsigned)nb, 256, 0, s>>>(src, dst, nrows, ncols, spitch, dpitch);
  return cudaGetLastError();
}

cudaError_t launch_repitch_r(const double* src, double* dst, size_t nrows, int ncols, int spitch, int dpitch, cudaStream_t s) {
  size_t nb = (nrows * ncols + 255) / 256;
  if (nb > 148 * 32) nb = 148 * 32;
  repitch_kernel<double><<<(unsigned)nb, 256, 0, s>>>(src, dst, nrows, ncols, spitch, dpitch);
  return cudaGetLastError();
}

}  // namespace pinb

namespace pinb {
__global__ void __launch_bounds__(PINB_MAXR) barrier_kernel(const __grid_constant__ BarrierParams p) { barrier_body(p); }

cudaError_t launch_barrier(const BarrierParams& p, cudaStream_t s) {
  barrier_kernel<<<1, PINB_MAXR, 0, s>>>(p);
  return cudaGetLastError();
}
}  // namespace pinb
