// Host-side launch interface between engine.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include "kernels.cuh"

namespace pinb {

bool grid_supported(int N);

cudaError_t launch_xpass(int N, int dir, const XPassParams& p, int nblocks_y, cudaStream_t s);
cudaError_t launch_ypass(int N, int dir, const YPassParams& p, int nblocks_x, cudaStream_t s);
cudaError_t launch_zpass_collapse(int N, const CollapseParams& p, size_t nrows, cudaStream_t s);
cudaError_t launch_zpass_out(int N, const ZOutParams& p, size_t nrows, cudaStream_t s);
cudaError_t launch_zpass_r2c(int N, const ZR2CParams& p, size_t nrows, cudaStream_t s);
// kz-tile width of the x pass (dir +1 inverse, -1 forward; dst_klayout as in XPassParams: the inverse pass that only
// stores locally, 2, never splits its lines) ...
int xpass_tk(int N, int dir, int dst_klayout = 0);
int ypass_tk(int N);           // ... and of the y pass for this grid
bool strided_tma_enabled();    // the strided passes fetch their tiles through the TMA unit (PINB200_TMA=0: cp.async)

cudaError_t launch_sources(const SourcesParams& p, cudaStream_t s);
cudaError_t launch_genic(const GenicParams& p, cudaStream_t s);

// cross-GPU stream barrier over peer-mapped flags
cudaError_t launch_barrier(const BarrierParams& p, cudaStream_t s);

// small helpers
cudaError_t launch_gauss_table(double* gauss, int M, double knorm, double rsmooth, cudaStream_t s);
cudaError_t launch_dc_scalar(const double2* src, double* out, double scale, int times_i, cudaStream_t s);
cudaError_t launch_nyq_probe(const double2* f, size_t nrows, int pitch, int M, int* flag, cudaStream_t s);
cudaError_t launch_fmax_pdf(const float* fmax, size_t n, unsigned long long* counts, cudaStream_t s);
cudaError_t launch_collapse_cells(const double* h6, size_t n, const double* spline, int nspl, double* F, cudaStream_t s);

// collapse-time tables (collapse_table.cuh, k_ctable.cu).  launch_zpass_collapse takes the tabulated
// variant of the kernel when CollapseParams::ct.coef is set.
cudaError_t launch_ct_build(const CTBuildParams& p, cudaStream_t s);
cudaError_t launch_ct_spline(const CTSplineParams& p, cudaStream_t s);
cudaError_t launch_collapse_cells_tab(const double* h6, size_t n, const CTView& v, double* F, cudaStream_t s);

struct PackParams {
  const float* fmax; const int* rmax; const float* vel[12];
  unsigned char* out; size_t stride; int prodfloat_bytes;
  int off_rmax, off_fmax, off_vel[4];
  size_t cell_begin, ncells;
  const unsigned int* gather;  // nullptr: records of cells cell_begin .. ; else of cells gather[cell_begin + i]
};
cudaError_t launch_pack_products(const PackParams& p, cudaStream_t s);

// collapsed-cell selection + radix sort (k_sort.cu)
size_t cell_sort_ntiles_select(unsigned long long ncells);
size_t cell_sort_ntiles_radix(unsigned long long n);
size_t cell_sort_scan_blocks(unsigned long long len);
int cell_sort_max_bins();
cudaError_t launch_scan_u32(unsigned int* a, unsigned long long len, unsigned int* block_sums, unsigned long long* total, cudaStream_t s);
cudaError_t launch_select_count(const float* fmax, unsigned long long n, float f_last, unsigned int* tile_counts, unsigned int* range, cudaStream_t s);
cudaError_t launch_select_write(const float* fmax, unsigned long long n, float f_last, const unsigned int* tile_base, const unsigned int* range,
                                unsigned int* key_out, unsigned int* idx_out, cudaStream_t s);
cudaError_t launch_radix_hist(const unsigned int* key, unsigned long long n, int shift, int bits, unsigned int* counts, cudaStream_t s);
cudaError_t launch_radix_scatter(const unsigned int* key_in, const unsigned int* idx_in, unsigned int* key_out, unsigned int* idx_out,
                                 unsigned long long n, int shift, int bits, const unsigned int* counts, cudaStream_t s);

// host <-> device layout converters (pitch P on the device, N/2+1 or N on the host)
cudaError_t launch_repitch_c(const double2* src, double2* dst, size_t nrows, int ncols, int spitch, int dpitch, cudaStream_t s);
cudaError_t launch_repitch_r(const double* src, double* dst, size_t nrows, int ncols, int spitch, int dpitch, cudaStream_t s);

}  // namespace pinb
