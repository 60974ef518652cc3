// FP64 pipe micro-benchmark for the collapse epilogue's scheduling model (sm_100a):
//   dependent-issue latency of DFMA (one warp, chains of 1/2/4 independent accumulators) and the
//   per-SM throughput with W resident warps each running C independent chains.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64lat tools/fp64lat.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int C> __global__ void chain(double* out, long long* cyc, int iters, double a, double b) {
  double x[C];
#pragma unroll
  for (int c = 0; c < C; c++) x[c] = threadIdx.x * 1e-3 + c;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int c = 0; c < C; c++) x[c] = fma(x[c], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < C; c++) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int C> void run(int warps, double* out, long long* cyc) {
  const int iters = 2000;
  chain<C><<<1, 32 * warps>>>(out, cyc, iters, 0.999999, 1e-9);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_fma = (double)h / (iters * 8.0 * C);
  printf("warps/SM %2d  chains/thread %d : %.2f cycles per DFMA per warp  -> %.3f warp-DFMA/clk/SM (peak 2.0 = 4 sub-partitions x 1/2)\n",
         warps, C, per_fma, warps / per_fma);
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8 * 8); cudaMalloc(&cyc, 64);
  for (int w : {1, 4, 8, 12, 16, 24, 32}) { run<1>(w, out, cyc); }
  for (int w : {1, 4, 8, 12, 16, 24, 32}) { run<2>(w, out, cyc); }
  for (int w : {1, 4, 8, 16, 32}) { run<4>(w, out, cyc); }
  return 0;
}
