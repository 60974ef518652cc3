// Device execution context for the kernel bodies of kernels.cuh.
#pragma once
#include <cuda_runtime.h>

namespace pinb {
struct DevCtx {
  __device__ __forceinline__ int tid() const { return threadIdx.x; }
  __device__ __forceinline__ int bid() const { return blockIdx.x; }
  __device__ __forceinline__ int nthreads() const { return blockDim.x; }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  __device__ __forceinline__ void atomic_add(double* p, double v) const { atomicAdd(p, v); }
};

template <class K> inline cudaError_t allow_smem(K kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return cudaSuccess;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
}  // namespace pinb

#define PINB_FOR_EACH_GRID(X) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048)
