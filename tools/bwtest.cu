// HBM bandwidth experiments (B200): read/write mix and the strided-tile access patterns of the
// x and y FFT passes (copy only, no FFT) -- separates "access pattern" from "kernel structure".
#include <cstdio>
#include <cuda_runtime.h>
__global__ void rw(const double2* __restrict__ in, double2* __restrict__ out, size_t n, int nr, int nw) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    double2 acc = make_double2(1.0, 2.0);
    for (int r = 0; r < nr; r++) { double2 v = in[(size_t)r * n + i]; acc.x += v.x; acc.y += v.y; }
    for (int w = 0; w < nw; w++) out[(size_t)w * n + i] = acc;
  }
}
// tile copy: block = (outer index o, kz tile); 512 threads; element e of the line at e*stride
// TK complex contiguous per element. nout outputs.
template <int TK>
__global__ void __launch_bounds__(512) tilecopy(const double2* __restrict__ in, double2* __restrict__ out, size_t nfield,
                                                int L, size_t estride, size_t ostride, int ntz, int nout, int twophase) {
  extern __shared__ double2 sm[];
  const int o = blockIdx.x / ntz, kz0 = (blockIdx.x % ntz) * TK;
  const int tk = threadIdx.x % TK, j = threadIdx.x / TK;
  const int TPL = 512 / TK;
  const size_t base = (size_t)o * ostride + kz0 + tk;
  double2 v[16];
  const int per = L / TPL;
#pragma unroll
  for (int r = 0; r < 16; r++) if (r < per) v[r] = in[base + (size_t)(j + r * TPL) * estride];
  if (twophase) {  // through shared memory with a barrier, like an FFT stage
#pragma unroll
    for (int r = 0; r < 16; r++) if (r < per) sm[(j + r * TPL) * TK + tk] = v[r];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; r++) if (r < per) v[r] = sm[((j * per + r)) * TK + tk];
  }
  for (int w = 0; w < nout; w++)
#pragma unroll
    for (int r = 0; r < 16; r++) if (r < per) out[(size_t)w * nfield + base + (size_t)(j + r * TPL) * estride] = v[r];
}
int main() {
  const int N = 1024, M = 512, P = 520;
  size_t n = (size_t)N * N * P;
  double2 *in, *out;
  cudaMalloc(&in, n * 16 * 3); cudaMalloc(&out, n * 16 * 6);
  cudaMemset(in, 0, n * 16 * 3);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto timeit = [&](auto f) { float best = 1e9; for (int it = 0; it < 3; it++) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; } return best; };
  int cfg[][2] = {{0, 3}, {1, 1}, {3, 6}, {1, 3}};
  for (auto& c : cfg) {
    float ms = timeit([&] { rw<<<148 * 16, 256>>>(in, out, n, c[0], c[1]); });
    printf("stream reads %d writes %d : %.2f ms  %.0f GB/s\n", c[0], c[1], ms, (double)(c[0] + c[1]) * n * 16 / 1e9 / ms * 1e3);
  }
  cudaFuncSetAttribute(tilecopy<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  cudaFuncSetAttribute(tilecopy<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const double gb1 = (double)N * N * M * 16 / 1e9;
  for (int two = 0; two < 2; two++)
    for (int nout : {1, 3}) {
      // y pattern: o = x plane, lines along y (stride P), in R layout
      float ms = timeit([&] { tilecopy<8><<<N * (M / 8), 512, two ? 128 * 1024 : 0>>>(in, out, n, N, P, (size_t)N * P, M / 8, nout, two); });
      printf("y-pattern TK=8 nout=%d twophase=%d : %.2f ms  %.0f GB/s\n", nout, two, ms, (1 + nout) * gb1 / ms * 1e3);
      // x pattern: o = y, lines along x (stride N*P)
      ms = timeit([&] { tilecopy<8><<<N * (M / 8), 512, two ? 128 * 1024 : 0>>>(in, out, n, N, (size_t)N * P, (size_t)P, M / 8, nout, two); });
      printf("x-pattern TK=8 nout=%d twophase=%d : %.2f ms  %.0f GB/s\n", nout, two, ms, (1 + nout) * gb1 / ms * 1e3);
    }
  float ms = timeit([&] { tilecopy<4><<<N * (M / 4), 512, 0>>>(in, out, n, N, P, (size_t)N * P, M / 4, 1, 0); });
  printf("y-pattern TK=4 (512 thr, 8 elems) nout=1 : %.2f ms  %.0f GB/s\n", ms, 2 * gb1 / ms * 1e3);
  ms = timeit([&] { tilecopy<4><<<N * (M / 4), 512, 0>>>(in, out, n, N, (size_t)N * P, (size_t)P, M / 4, 1, 0); });
  printf("x-pattern TK=4 nout=1 : %.2f ms  %.0f GB/s\n", ms, 2 * gb1 / ms * 1e3);
  return 0;
}
