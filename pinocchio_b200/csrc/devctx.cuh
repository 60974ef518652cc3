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
  // ---- TMA tile loads (cp.async.bulk.tensor + mbarrier): one thread moves a whole strided tile, the copy
  // engine of the SM does the address arithmetic; see strided_tile_jobs
  static constexpr bool kTma = true;
  __device__ __forceinline__ unsigned long long* tile_barrier() const {
    __shared__ __align__(8) unsigned long long bar;
    return &bar;
  }
  __device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) const {
    const unsigned int a = (unsigned int)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) const {
    const unsigned int a = (unsigned int)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
  }
  // waits for the phase with the given parity; a transfer that never completes (wrong byte count, bad descriptor)
  // ends in a trap after a few seconds instead of hanging the GPU
  __device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity) const {
    const unsigned int a = (unsigned int)__cvta_generic_to_shared(bar);
    unsigned int done = 0, spins = 0;
    while (true) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.b32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(a), "r"(parity)
          : "memory");
      if (done) break;
      if (++spins > (1u << 26)) __trap();
    }
  }
  // generic-proxy accesses to shared memory before it -> async-proxy (TMA) accesses after it
  __device__ __forceinline__ void fence_async_proxy() const { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
  __device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, int c0, int c1, int c2,
                                              unsigned long long* bar) const {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    const unsigned int b = (unsigned int)__cvta_generic_to_shared(bar);
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(d), "l"(tmap), "r"(b), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
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
