// sm_100a instantiations of the collapsed-cell filter + radix sort (sort_cells.cuh).
#include "devctx.cuh"
#include "launch.h"
#include "sort_cells.cuh"

namespace pinb {

__global__ void __launch_bounds__(SORT_NT) sort_count_kernel(const __grid_constant__ SortPassParams p) {
  extern __shared__ unsigned short sort_smem[];
  DevCtx ctx;
  sort_count_body(ctx, sort_smem, p);
}

__global__ void __launch_bounds__(SORT_NT) sort_scatter_kernel(const __grid_constant__ SortPassParams p) {
  extern __shared__ unsigned short sort_smem[];
  DevCtx ctx;
  unsigned int* base = reinterpret_cast<unsigned int*>(sort_smem + 256 * SORT_ROW);
  sort_scatter_body(ctx, sort_smem, base, p);
}

static constexpr int SCAN_NT = 1024;
__global__ void __launch_bounds__(SCAN_NT) sort_scan_kernel(unsigned int* counts, unsigned long long len, unsigned long long* total) {
  __shared__ unsigned int scratch[SCAN_NT];
  DevCtx ctx;
  sort_scan_body(ctx, scratch, counts, len, total);
}

// count + scan only: *total (device) = number of elements that pass the filter
cudaError_t launch_sort_count(const SortPassParams& p, unsigned long long* total, cudaStream_t s) {
  cudaError_t e = allow_smem(sort_count_kernel, SORT_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  sort_count_kernel<<<p.ntiles, SORT_NT, SORT_SMEM_BYTES, s>>>(p);
  sort_scan_kernel<<<1, SCAN_NT, 0, s>>>(p.counts, (unsigned long long)256 * p.ntiles, total);
  return cudaGetLastError();
}

// one radix pass on stream s; *total (device) receives the number of elements written
cudaError_t launch_sort_pass(const SortPassParams& p, unsigned long long* total, cudaStream_t s) {
  cudaError_t e = allow_smem(sort_count_kernel, SORT_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = allow_smem(sort_scatter_kernel, SORT_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  sort_count_kernel<<<p.ntiles, SORT_NT, SORT_SMEM_BYTES, s>>>(p);
  sort_scan_kernel<<<1, SCAN_NT, 0, s>>>(p.counts, (unsigned long long)256 * p.ntiles, total);
  sort_scatter_kernel<<<p.ntiles, SORT_NT, SORT_SMEM_BYTES, s>>>(p);
  return cudaGetLastError();
}

}  // namespace pinb
