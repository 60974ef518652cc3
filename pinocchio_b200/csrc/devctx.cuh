// Device execution context for the kernel bodies of kernels.cuh.
#pragma once
#include <cuda_runtime.h>
#include "kernels.cuh"

namespace pinb {
struct DevCtx {
  __device__ __forceinline__ int tid() const { return threadIdx.x; }
  __device__ __forceinline__ int bid() const { return blockIdx.x; }
  __device__ __forceinline__ int nthreads() const { return blockDim.x; }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  // v has the same value in every lane of the warp (e.g. tid >> 5): routing it through a shuffle
  // from lane 0 lets the compiler's divergence analysis see that
  __device__ __forceinline__ int warp_uniform(int v) const { return __shfl_sync(0xffffffffu, v, 0); }
  // barrier among the n threads that own one z line: a warp-level sync when the line fits a warp
  // (several small lines may share a warp: every lane executes the same instruction stream),
  // else a named barrier (ids 1..15; n is a multiple of 32)
  __device__ __forceinline__ void sync_line(int line_id, int n) const {
    if (n <= 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + (line_id % 15)), "r"(n) : "memory");
  }
  // 16-byte asynchronous copy HBM -> shared memory (LDGSTS: no register staging) and its fence
  __device__ __forceinline__ void async_copy16(void* smem_dst, const void* gsrc) const {
    const unsigned int sa = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
  }
  __device__ __forceinline__ void mark(int) const {}  // timeline hook (tools/passbench only)
  // compiler scheduling fence: memory operations are not moved across it (bounds loads in flight)
  __device__ __forceinline__ void sched_fence() const { asm volatile("" ::: "memory"); }
  __device__ __forceinline__ void prefetch_l2(const void* g) const { asm volatile("prefetch.global.L2 [%0];" ::"l"(g)); }
  __device__ __forceinline__ void async_wait() const { asm volatile("cp.async.wait_all;" ::: "memory"); }
  __device__ __forceinline__ void atomic_add(double* p, double v) const { atomicAdd(p, v); }
  // block-wide sum of two doubles (result valid in thread 0): warp shuffles + one barrier
  template <int NT> __device__ __forceinline__ void block_sum2(double* scratch, double& a, double& b) const {
    if constexpr (NT % 32 != 0) {
      block_sum2_tree<NT>(*this, scratch, a, b);
    } else {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
      }
      const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
      if (l == 0) { scratch[w] = a; scratch[32 + w] = b; }
      __syncthreads();
      if (w == 0) {
        a = l < NT / 32 ? scratch[l] : 0.0;
        b = l < NT / 32 ? scratch[32 + l] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          a += __shfl_down_sync(0xffffffffu, a, o);
          b += __shfl_down_sync(0xffffffffu, b, o);
        }
      }
    }
  }
};

template <class K> inline cudaError_t allow_smem(K kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return cudaSuccess;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
}  // namespace pinb

#define PINB_FOR_EACH_GRID(X) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048)
