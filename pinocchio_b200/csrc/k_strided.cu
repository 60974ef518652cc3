// sm_100a instantiations of the strided (x and y) FFT passes.  See kernels.cuh.
#include <cuda.h>  // CUtensorMap and cuTensorMapEncodeTiled's types only: the entry point comes from the runtime

#include <cstdlib>
#include <cstring>

#include "devctx.cuh"
#include "launch.h"

namespace pinb {

// ---- TMA descriptors of the strided passes' source fields ------------------------------------------------------
// PINB200_TMA=0 keeps the cp.async (LDGSTS) loader.  The driver entry point is looked up through the runtime, so
// the library does not link libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    const char* env = getenv("PINB200_TMA");
    if (env && !atoi(env)) return (EncodeTiledFn) nullptr;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return (EncodeTiledFn) nullptr;
    return (EncodeTiledFn)ptr;
  }();
  return fn;
}
bool strided_tma_enabled() { return encode_tiled() != nullptr; }
// field: ny rows of `pitch` complex per x plane, nx planes; a box is 2 tk doubles of `rows` consecutive line elements
// along dimension `line_dim` (1: y, 2: x)
static bool make_tile_map(TileMap* out, const double2* field, int pitch, int ny, int nx, int tk, int rows, int line_dim) {
  static_assert(sizeof(CUtensorMap) == sizeof(TileMap), "TileMap must mirror CUtensorMap");
  EncodeTiledFn enc = encode_tiled();
  if (!enc || (reinterpret_cast<uintptr_t>(field) & 15)) return false;
  CUtensorMap m;
  const cuuint64_t dims[3] = {(cuuint64_t)2 * pitch, (cuuint64_t)ny, (cuuint64_t)nx};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch * sizeof(double2), (cuuint64_t)ny * pitch * sizeof(double2)};
  const cuuint32_t box[3] = {(cuuint32_t)(2 * tk), line_dim == 1 ? (cuuint32_t)rows : 1u, line_dim == 2 ? (cuuint32_t)rows : 1u};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double2*>(field), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  memcpy(out, &m, sizeof(m));
  return true;
}
// Measured (r02, tools/slabbench 1024 1, profiles/r02_slabbench_tma.txt): with several jobs per tile the TMA loader
// wins (x pass, 3 outputs: 17.8 -> 15.5 ms; y pass, 6 jobs: 20.2 -> 19.2 ms); a pass with ONE job per tile has no
// second load to overlap and does better when every thread waits for its own 16 cp.async pieces only
// (5.95 against 6.19 ms), so those keep the LDGSTS loader.
// Blocks of a strided pass: as many as are resident at once; each walks the tiles bid, bid + grid, ... so that the
// load of its next tile overlaps the stores of the current one (strided_tile_jobs).  PINB200_PERSISTENT=0: one tile
// per block.
// Passes that store into PEER memory keep one tile per block: on 8 GPUs the displacement stage, whose transposes are
// peer stores, took 2.3 s instead of 0.5 s with walking blocks (r02, profiles/r02_multi/bench_8gpu_2048_walk_everywhere.json)
// -- a block that walks on to its next tile while the NVLink stores of the last one drain stalls behind them, where a
// fresh block on another SM does not.
template <class K> static unsigned strided_grid(K kernel, int nthreads, size_t smem, unsigned ntiles, int* tile_stride, bool peer_stores) {
  static const bool persistent = [] { const char* e = getenv("PINB200_PERSISTENT"); return !(e && !atoi(e)); }();
  *tile_stride = 0;
  if (!persistent || peer_stores) return ntiles;
  int dev = 0, sms = 0, occ = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, nthreads, smem) != cudaSuccess || occ < 1 || sms < 1)
    return ntiles;
  const unsigned g = (unsigned)occ * (unsigned)sms;
  if (g >= ntiles) return ntiles;
  *tile_stride = (int)g;
  return g;
}
template <int LT> static void xpass_tma(XPassParams& p, int tk) {
  const int npw = ((p.pmask >> 0) & 1) + ((p.pmask >> 1) & 1) + ((p.pmask >> 2) & 1);
  p.use_tma = (npw >= 2 && make_tile_map(&p.tmap, p.src, p.g.P, p.g.ly, p.g.N, tk, LT < 256 ? LT : 256, 2)) ? 1 : 0;
}
template <int LT> static void ypass_tma(YPassParams& p, int tk) {
  p.use_tma = p.njobs >= 2 ? 1 : 0;
  if (!p.use_tma) return;
  bool used[3] = {false, false, false};
  for (int j = 0; j < p.njobs; j++) used[p.job[j].src] = true;
  for (int i = 0; i < 3; i++)
    if (used[i] && !make_tile_map(&p.tmap[i], p.src[i], p.g.P, p.g.N, p.g.lx, tk, LT < 256 ? LT : 256, 1)) p.use_tma = 0;
}

template <int L, int DIR, bool MULTI>
__global__ void __launch_bounds__(XCfg<L, DIR>::NT) xpass_kernel(const __grid_constant__ XPassParams p) {
  extern __shared__ __align__(128) double2 smem[];
  DevCtx ctx;
  xpass_body<L, DIR, MULTI>(ctx, smem, p);
}

// inverse x pass with the scale-dependent growth rate evaluated per mode (displacement stage of
// SCALE_DEPENDENT runs only; the Hessian sweep uses xpass_kernel)
template <int L, bool MULTI>
__global__ void __launch_bounds__(XCfg<L, +1>::NT) xpass_growthk_kernel(const __grid_constant__ XPassParams p) {
  extern __shared__ __align__(128) double2 smem[];
  DevCtx ctx;
  xpass_body<L, +1, MULTI, DevCtx, true>(ctx, smem, p);
}

// inverse x pass of the pipelined multi-GPU sweep for lines that the scattering kernel has to split (XCfg LOCAL):
// every store is local (dst_klayout == 2), the whole line is one TK = 4 tile
template <int L>
__global__ void __launch_bounds__(XCfg<L, +1, true>::NT) xpass_local_kernel(const __grid_constant__ XPassParams p) {
  extern __shared__ __align__(128) double2 smem[];
  DevCtx ctx;
  xpass_body<L, +1, false, DevCtx, false, true>(ctx, smem, p);
}

template <int L, int DIR>
__global__ void __launch_bounds__(YCfg<L>::NT) ypass_kernel(const __grid_constant__ YPassParams p) {
  extern __shared__ __align__(128) double2 smem[];
  DevCtx ctx;
  ypass_body<L, DIR>(ctx, smem, p);
}

template <int L, int DIR, bool MULTI> static cudaError_t xpass_launch_m(const XPassParams& p_in, int nblocks_y, cudaStream_t s) {
  using C = XCfg<L, DIR>;
  XPassParams p = p_in;
  xpass_tma<C::LT>(p, C::TK);
  const size_t smem = (size_t)C::LT * C::TK * sizeof(double2);
  cudaError_t e = allow_smem(xpass_kernel<L, DIR, MULTI>, smem);
  if (e != cudaSuccess) return e;
  p.nblocks = nblocks_y * p.ntiles_z;
  const unsigned grid = strided_grid(xpass_kernel<L, DIR, MULTI>, C::NT, smem, (unsigned)p.nblocks, &p.tile_stride, MULTI && p.dst_klayout == 0);
  xpass_kernel<L, DIR, MULTI><<<grid, C::NT, smem, s>>>(p);
  return cudaGetLastError();
}
template <int L, bool MULTI> static cudaError_t xpass_growthk_launch_m(const XPassParams& p_in, int nblocks_y, cudaStream_t s) {
  using C = XCfg<L, +1>;
  XPassParams p = p_in;
  xpass_tma<C::LT>(p, C::TK);
  const size_t smem = (size_t)C::LT * C::TK * sizeof(double2);
  cudaError_t e = allow_smem(xpass_growthk_kernel<L, MULTI>, smem);
  if (e != cudaSuccess) return e;
  p.nblocks = nblocks_y * p.ntiles_z;
  const unsigned grid = strided_grid(xpass_growthk_kernel<L, MULTI>, C::NT, smem, (unsigned)p.nblocks, &p.tile_stride, MULTI && p.dst_klayout == 0);
  xpass_growthk_kernel<L, MULTI><<<grid, C::NT, smem, s>>>(p);
  return cudaGetLastError();
}
template <int L> static cudaError_t xpass_local_launch(const XPassParams& p_in, int nblocks_y, cudaStream_t s) {
  using C = XCfg<L, +1, true>;
  XPassParams p = p_in;
  xpass_tma<C::LT>(p, C::TK);
  const size_t smem = (size_t)C::LT * C::TK * sizeof(double2);
  cudaError_t e = allow_smem(xpass_local_kernel<L>, smem);
  if (e != cudaSuccess) return e;
  p.nblocks = nblocks_y * p.ntiles_z;
  const unsigned grid = strided_grid(xpass_local_kernel<L>, C::NT, smem, (unsigned)p.nblocks, &p.tile_stride, false);
  xpass_local_kernel<L><<<grid, C::NT, smem, s>>>(p);
  return cudaGetLastError();
}
template <int L, int DIR> static cudaError_t xpass_launch(const XPassParams& p, int nblocks_y, cudaStream_t s) {
  if constexpr (DIR > 0 && XCfg<L, +1>::SPLIT) {
    if (p.dst_klayout == 2 && !p.kf.gk) return xpass_local_launch<L>(p, nblocks_y, s);
  }
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

template <int L, int DIR> static cudaError_t ypass_launch(const YPassParams& p_in, int nblocks_x, cudaStream_t s) {
  using C = YCfg<L>;
  YPassParams p = p_in;
  ypass_tma<C::LT>(p, C::TK);
  const size_t smem = (size_t)C::LT * C::TK * sizeof(double2);
  cudaError_t e = allow_smem(ypass_kernel<L, DIR>, smem);
  if (e != cudaSuccess) return e;
  p.nblocks = nblocks_x * p.ntiles_z;
  const unsigned grid = strided_grid(ypass_kernel<L, DIR>, C::NT, smem, (unsigned)p.nblocks, &p.tile_stride, p.dst_klayout != 0 && p.g.lx != p.g.N);
  ypass_kernel<L, DIR><<<grid, C::NT, smem, s>>>(p);
  return cudaGetLastError();
}

// kz-tile width of the x pass in direction dir and of the y pass (they differ for N > 1024)
int xpass_tk(int N, int dir, int dst_klayout) {
  switch (N) {
#define X(L) case L: return dir > 0 ? (dst_klayout == 2 ? XCfg<L, +1, true>::TK : XCfg<L, +1>::TK) : XCfg<L, -1>::TK;
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
