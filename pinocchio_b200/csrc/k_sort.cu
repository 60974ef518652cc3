// sm_100a kernels of the hand-off to the fragmentation: the cells with Fmax >= F_last in
// order of descending Fmax, ties in ascending cell index -- the selection of src/distribute.c:58-175,547-600 and
// the order of sort_and_organize (src/fragment.c:484-520) -- selected and ordered on the device.
//
// r01 ran the block-emulator-portable bodies of sort_cells.cuh here (one thread walking 128 consecutive elements
// with private 16-bit counters, a ONE-block scan): 807 ms at 1024^3, slower than the nine-radius sweep.  Those
// bodies remain the CPU-checkable statement of the algorithm (tests/host, emulated ABI).  This file is the
// device implementation; tests/test_zgpu_3_fragment_handoff.py checks it against a stable NumPy sort.
//
//   1. select : stable stream compaction of (key, cell index), key = ~bits(Fmax) - kmin.  Positive floats order
//               like their bit patterns, so ascending keys are descending Fmax.  Two sweeps over Fmax (count per
//               tile + global key range, then write) around one scan of the tile counts.
//   2. sort   : least-significant-digit radix sort of the pairs over the bits in which the keys actually differ
//               (kmax - kmin: 26 bits for 1 <= F < 32), digits of up to 9 bits => 3 passes instead of 4 x 8.
//               One pass = per-tile digit histogram, multi-block exclusive scan in (digit, tile) order, scatter.
//               Ranks inside a tile follow memory order (warp-level multi-split with match.any, per-warp
//               counters, no atomics), which makes every pass stable; the first pass reads cells in index order,
//               so ties end in ascending cell index.
// Traffic: 2 x 4 B/cell for the selection + 3 x 24 B/selected cell; about 50 GB at 1024^3.
#include "devctx.cuh"
#include "launch.h"

namespace pinb {

namespace {

constexpr int SEL_NT = 256;                 // threads of the selection kernels
constexpr int SEL_IPT = 16;                 // elements per thread and round
constexpr int SEL_ROUNDS = 4;               // rounds per tile
constexpr int SEL_TILE = SEL_NT * SEL_IPT * SEL_ROUNDS;  // 16384 cells

constexpr int RS_NT = 256;                  // threads of a radix tile (8 warps)
constexpr int RS_NW = RS_NT / 32;
constexpr int RS_IPT = 16;                  // elements per thread in one chunk (registers hold their ranks)
constexpr int RS_CHUNK = RS_NT * RS_IPT;    // 4096
constexpr int RS_CHUNKS = 4;                // chunks per tile
constexpr int RS_TILE = RS_CHUNK * RS_CHUNKS;  // 16384 pairs
constexpr int RS_MAXBITS = 9;
constexpr int RS_MAXBINS = 1 << RS_MAXBITS;

constexpr int SCAN_NT = 1024;
constexpr int SCAN_IPT = 16;
constexpr int SCAN_SEG = SCAN_NT * SCAN_IPT;   // 16384 counters per block

__device__ __forceinline__ unsigned int lanemask_lt() {
  unsigned int m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// ---- selection ---------------------------------------------------------------------------------
__device__ __forceinline__ bool passes(float f, float f_last) { return f >= f_last; }  // NaN never passes

// tile_counts[tile] = selected cells of the tile; range[0] = min, range[1] = max of ~bits(Fmax) over them
__global__ void __launch_bounds__(SEL_NT) select_count_kernel(const float* __restrict__ fmax, unsigned long long n, float f_last,
                                                              unsigned int* __restrict__ tile_counts, unsigned int* range) {
  __shared__ unsigned int s_cnt[SEL_NT / 32], s_min[SEL_NT / 32], s_max[SEL_NT / 32];
  const unsigned long long base = (unsigned long long)blockIdx.x * SEL_TILE;
  unsigned int cnt = 0, kmin = 0xffffffffu, kmax = 0u;
  for (int i = threadIdx.x; i < SEL_TILE; i += SEL_NT) {
    const unsigned long long e = base + i;
    if (e < n) {
      const float f = fmax[e];
      if (passes(f, f_last)) {
        const unsigned int k = ~__float_as_uint(f);
        cnt++;
        kmin = min(kmin, k);
        kmax = max(kmax, k);
      }
    }
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { s_cnt[w] = cnt; s_min[w] = kmin; s_max[w] = kmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < SEL_NT / 32; i++) { cnt += s_cnt[i]; kmin = min(kmin, s_min[i]); kmax = max(kmax, s_max[i]); }
    tile_counts[blockIdx.x] = cnt;
    if (cnt) { atomicMin(range, kmin); atomicMax(range + 1, kmax); }
  }
}

// writes the pairs of tile b at tile_base[b] (the scanned counts), in cell order
__global__ void __launch_bounds__(SEL_NT) select_write_kernel(const float* __restrict__ fmax, unsigned long long n, float f_last,
                                                              const unsigned int* __restrict__ tile_base, const unsigned int* range,
                                                              unsigned int* __restrict__ key_out, unsigned int* __restrict__ idx_out) {
  __shared__ unsigned int s_warp[SEL_NT / 32];
  __shared__ unsigned int s_run;
  const unsigned int kmin = range[0];
  const unsigned long long base = (unsigned long long)blockIdx.x * SEL_TILE;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) s_run = tile_base[blockIdx.x];
  __syncthreads();
  // rounds of SEL_NT consecutive cells: rank = cells selected before this one in the tile
  for (int i0 = 0; i0 < SEL_TILE; i0 += SEL_NT) {
    const unsigned long long e = base + i0 + threadIdx.x;
    float f = 0.0f;
    bool sel = false;
    if (e < n) {
      f = fmax[e];
      sel = passes(f, f_last);
    }
    const unsigned int bal = __ballot_sync(0xffffffffu, sel);
    if (lane == 0) s_warp[w] = __popc(bal);
    __syncthreads();
    unsigned int off = s_run;
    for (int i = 0; i < w; i++) off += s_warp[i];
    if (sel) {
      const unsigned int pos = off + __popc(bal & lanemask_lt());
      key_out[pos] = ~__float_as_uint(f) - kmin;
      idx_out[pos] = (unsigned int)e;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int t = 0;
      for (int i = 0; i < SEL_NT / 32; i++) t += s_warp[i];
      s_run += t;
    }
    __syncthreads();
  }
}

// ---- exclusive scan of a long u32 array (three kernels) -----------------------------------------------
__global__ void __launch_bounds__(SCAN_NT) scan_reduce_kernel(const unsigned int* __restrict__ a, unsigned long long len, unsigned int* __restrict__ block_sums) {
  __shared__ unsigned int s[SCAN_NT / 32];
  const unsigned long long b = (unsigned long long)blockIdx.x * SCAN_SEG;
  unsigned int v = 0;
  for (int i = threadIdx.x; i < SCAN_SEG; i += SCAN_NT)
    if (b + i < len) v += a[b + i];
  v = __reduce_add_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = 0;
    for (int i = 0; i < SCAN_NT / 32; i++) t += s[i];
    block_sums[blockIdx.x] = t;
  }
}

// one block: block_sums -> exclusive prefix in place, *total = grand total
__global__ void __launch_bounds__(SCAN_NT) scan_sums_kernel(unsigned int* block_sums, unsigned int nb, unsigned long long* total) {
  __shared__ unsigned int s[SCAN_NT];
  const unsigned int per = (nb + SCAN_NT - 1) / SCAN_NT;
  const unsigned int b = threadIdx.x * per, e = min(b + per, nb);
  unsigned int v = 0;
  for (unsigned int i = b; i < e; i++) v += block_sums[i];
  s[threadIdx.x] = v;
  __syncthreads();
  // Hillis-Steele over the 1024 partials
  for (int o = 1; o < SCAN_NT; o <<= 1) {
    const unsigned int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0u;
    __syncthreads();
    s[threadIdx.x] += t;
    __syncthreads();
  }
  unsigned int run = s[threadIdx.x] - v;  // exclusive
  for (unsigned int i = b; i < e; i++) {
    const unsigned int c = block_sums[i];
    block_sums[i] = run;
    run += c;
  }
  if (threadIdx.x == SCAN_NT - 1 && total) *total = s[SCAN_NT - 1];
}

// per segment: exclusive scan in place, offset by the segment's prefix
__global__ void __launch_bounds__(SCAN_NT) scan_apply_kernel(unsigned int* __restrict__ a, unsigned long long len, const unsigned int* __restrict__ block_sums) {
  __shared__ unsigned int s[SCAN_NT];
  const unsigned long long b = (unsigned long long)blockIdx.x * SCAN_SEG + (unsigned long long)threadIdx.x * SCAN_IPT;
  unsigned int v[SCAN_IPT], sum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_IPT; i++) {
    v[i] = (b + i < len) ? a[b + i] : 0u;
    sum += v[i];
  }
  s[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 1; o < SCAN_NT; o <<= 1) {
    const unsigned int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0u;
    __syncthreads();
    s[threadIdx.x] += t;
    __syncthreads();
  }
  unsigned int run = block_sums[blockIdx.x] + s[threadIdx.x] - sum;
#pragma unroll
  for (int i = 0; i < SCAN_IPT; i++) {
    if (b + i < len) a[b + i] = run;
    run += v[i];
  }
}

// ---- radix pass -------------------------------------------------------------------------------------
// counts[d * ntiles + tile] = keys of the tile whose digit is d
__global__ void __launch_bounds__(RS_NT) radix_hist_kernel(const unsigned int* __restrict__ key, unsigned long long n, int shift, int nbins,
                                                           unsigned int* __restrict__ counts, unsigned int ntiles) {
  __shared__ unsigned int h[RS_MAXBINS];
  for (int i = threadIdx.x; i < nbins; i += RS_NT) h[i] = 0;
  __syncthreads();
  const unsigned long long base = (unsigned long long)blockIdx.x * RS_TILE;
  const unsigned int mask = (unsigned int)nbins - 1u;
  for (int i = threadIdx.x; i < RS_TILE; i += RS_NT) {
    const unsigned long long e = base + i;
    const bool ok = e < n;
    const unsigned int d = ok ? ((key[e] >> shift) & mask) : (unsigned int)nbins;
    // lanes of a warp with the same digit add once
    const unsigned int peers = __match_any_sync(0xffffffffu, d);
    if (ok && (peers & lanemask_lt()) == 0) atomicAdd(&h[d], (unsigned int)__popc(peers));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nbins; i += RS_NT) counts[(size_t)i * ntiles + blockIdx.x] = h[i];
}

// scatter of one tile; counts hold the exclusive scan (global base of every (digit, tile) bucket)
__global__ void __launch_bounds__(RS_NT) radix_scatter_kernel(const unsigned int* __restrict__ key_in, const unsigned int* __restrict__ idx_in,
                                                              unsigned int* __restrict__ key_out, unsigned int* __restrict__ idx_out,
                                                              unsigned long long n, int shift, int nbins,
                                                              const unsigned int* __restrict__ counts, unsigned int ntiles) {
  __shared__ unsigned int wcnt[RS_NW][RS_MAXBINS];  // per-warp digit counters of the chunk, then their exclusive scan over the warps
  __shared__ unsigned int dbase[RS_MAXBINS];        // global position of the next key of each digit of this tile
  __shared__ unsigned int dtot[RS_MAXBINS];         // keys of each digit in the current chunk
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned int mask = (unsigned int)nbins - 1u;
  for (int i = threadIdx.x; i < nbins; i += RS_NT) dbase[i] = counts[(size_t)i * ntiles + blockIdx.x];
  const unsigned long long tile0 = (unsigned long long)blockIdx.x * RS_TILE;
  for (int c = 0; c < RS_CHUNKS; c++) {
    const unsigned long long chunk0 = tile0 + (unsigned long long)c * RS_CHUNK;
    if (chunk0 >= n) break;  // uniform over the block
    for (int i = threadIdx.x; i < RS_NW * RS_MAXBINS; i += RS_NT) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    // warp w owns the RS_IPT * 32 consecutive pairs at chunk0 + w * 512; round r = 32 consecutive ones
    unsigned int k[RS_IPT], v[RS_IPT];
    unsigned short rank[RS_IPT];
    const unsigned long long w0 = chunk0 + (unsigned long long)w * (RS_IPT * 32);
#pragma unroll
    for (int r = 0; r < RS_IPT; r++) {
      const unsigned long long e = w0 + r * 32 + lane;
      const bool ok = e < n;
      k[r] = ok ? key_in[e] : 0xffffffffu;
      v[r] = ok ? idx_in[e] : 0u;
    }
#pragma unroll
    for (int r = 0; r < RS_IPT; r++) {
      const bool ok = w0 + r * 32 + lane < n;
      // lanes past the end take a digit of their own (nbins): they match only each other and are not counted
      const unsigned int d = ok ? ((k[r] >> shift) & mask) : (unsigned int)nbins;
      const unsigned int peers = __match_any_sync(0xffffffffu, d);
      const int leader = __ffs(peers) - 1;
      unsigned int old = 0;
      if (ok && lane == leader) {
        old = wcnt[w][d];
        wcnt[w][d] = old + __popc(peers);
      }
      old = __shfl_sync(0xffffffffu, old, leader);
      rank[r] = (unsigned short)(old + __popc(peers & lanemask_lt()));
    }
    __syncthreads();
    // exclusive scan of every digit's counters over the warps
    for (int d = threadIdx.x; d < nbins; d += RS_NT) {
      unsigned int run = 0;
#pragma unroll
      for (int i = 0; i < RS_NW; i++) {
        const unsigned int t = wcnt[i][d];
        wcnt[i][d] = run;
        run += t;
      }
      dtot[d] = run;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_IPT; r++) {
      if (w0 + r * 32 + lane < n) {
        const unsigned int d = (k[r] >> shift) & mask;
        const unsigned int pos = dbase[d] + wcnt[w][d] + rank[r];
        key_out[pos] = k[r];
        idx_out[pos] = v[r];
      }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < nbins; d += RS_NT) dbase[d] += dtot[d];
    // (the next chunk's zeroing of wcnt and its barrier order these updates before the next scatter)
  }
}

}  // namespace

size_t cell_sort_ntiles_select(unsigned long long ncells) { return (size_t)((ncells + SEL_TILE - 1) / SEL_TILE); }
size_t cell_sort_ntiles_radix(unsigned long long n) { return (size_t)((n + RS_TILE - 1) / RS_TILE); }
size_t cell_sort_scan_blocks(unsigned long long len) { return (size_t)((len + SCAN_SEG - 1) / SCAN_SEG); }
int cell_sort_max_bins() { return RS_MAXBINS; }

// exclusive scan of a[0..len) in place; *total (device, may be null) = sum.  block_sums: cell_sort_scan_blocks(len) u32.
cudaError_t launch_scan_u32(unsigned int* a, unsigned long long len, unsigned int* block_sums, unsigned long long* total, cudaStream_t s) {
  const unsigned int nb = (unsigned int)cell_sort_scan_blocks(len);
  scan_reduce_kernel<<<nb, SCAN_NT, 0, s>>>(a, len, block_sums);
  scan_sums_kernel<<<1, SCAN_NT, 0, s>>>(block_sums, nb, total);
  scan_apply_kernel<<<nb, SCAN_NT, 0, s>>>(a, len, block_sums);
  return cudaGetLastError();
}

cudaError_t launch_select_count(const float* fmax, unsigned long long n, float f_last, unsigned int* tile_counts, unsigned int* range, cudaStream_t s) {
  select_count_kernel<<<(unsigned)cell_sort_ntiles_select(n), SEL_NT, 0, s>>>(fmax, n, f_last, tile_counts, range);
  return cudaGetLastError();
}

cudaError_t launch_select_write(const float* fmax, unsigned long long n, float f_last, const unsigned int* tile_base, const unsigned int* range,
                                unsigned int* key_out, unsigned int* idx_out, cudaStream_t s) {
  select_write_kernel<<<(unsigned)cell_sort_ntiles_select(n), SEL_NT, 0, s>>>(fmax, n, f_last, tile_base, range, key_out, idx_out);
  return cudaGetLastError();
}

cudaError_t launch_radix_hist(const unsigned int* key, unsigned long long n, int shift, int bits, unsigned int* counts, cudaStream_t s) {
  const unsigned int nt = (unsigned int)cell_sort_ntiles_radix(n);
  radix_hist_kernel<<<nt, RS_NT, 0, s>>>(key, n, shift, 1 << bits, counts, nt);
  return cudaGetLastError();
}

cudaError_t launch_radix_scatter(const unsigned int* key_in, const unsigned int* idx_in, unsigned int* key_out, unsigned int* idx_out,
                                 unsigned long long n, int shift, int bits, const unsigned int* counts, cudaStream_t s) {
  const unsigned int nt = (unsigned int)cell_sort_ntiles_radix(n);
  radix_scatter_kernel<<<nt, RS_NT, 0, s>>>(key_in, idx_in, key_out, idx_out, n, shift, 1 << bits, counts, nt);
  return cudaGetLastError();
}

}  // namespace pinb
