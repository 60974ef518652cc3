// Hand-off to the fragmentation: the cells the unchanged CPU code stores
// and sorts -- those with Fmax >= F_last (bitmap filter of src/distribute.c:58-175,547-600), in
// order of descending Fmax (sort_and_organize, src/fragment.c:484-520) -- selected and ordered
// on the device, so that the host neither scans 56-byte records of uncollapsed cells nor qsorts.
//
// Stable least-significant-digit radix sort of (key, cell index) pairs, 8 bits per pass, with
// the filter folded into the first pass.  key = ~bits(Fmax): Fmax >= F_last > 0, and positive
// floats order like their bit patterns, so ascending keys are descending Fmax; the sort being
// stable and the input being in cell order, ties come out in ascending cell index (the
// reference's qsort leaves their order unspecified).
//
// One pass = three kernels over tiles of SORT_NT * SORT_IPT consecutive elements:
//   count  : per tile, the number of elements of each digit          -> counts[digit][tile]
//   scan   : exclusive prefix sum of counts in (digit, tile) order   -> global base of each
//            (digit, tile) bucket; its last value is the number of elements that passed
//   scatter: recount, rank every element inside its tile, store it at base + rank
// Thread t of a tile owns the SORT_IPT consecutive elements t*IPT .. t*IPT+IPT-1 and walks them
// in order; its per-digit counters are a private column of a [256][NT] u16 table in shared
// memory (no atomics), rows padded to an odd number of 32-bit words so that the row-wise and the
// column-wise phases are both free of bank conflicts.  Ranks inside a tile follow (thread,
// element) order = global element order, which is what makes the pass stable.
// Bodies are written against the Ctx of kernels.cuh and run under the CPU block emulator too.
#pragma once
#include "fft_core.cuh"

namespace pinb {

constexpr int SORT_NT = 128;    // threads per tile (66 KB of counters: three tiles per SM)
constexpr int SORT_IPT = 128;   // elements per thread (u16 counters: tile of 16384 < 65536)
constexpr int SORT_TILE = SORT_NT * SORT_IPT;
constexpr int SORT_ROW = SORT_NT + 2;  // u16 per digit row: 65 words
constexpr size_t SORT_SMEM_BYTES = (size_t)256 * SORT_ROW * sizeof(unsigned short) + 256 * sizeof(unsigned int);

struct SortPassParams {
  // source: either the Fmax field itself (first pass: fmax != nullptr, key and index generated on
  // the fly, elements below f_last dropped) or the (key, index) pairs of the previous pass
  const float* fmax;
  float f_last;
  const unsigned int* key_in;
  const unsigned int* idx_in;
  unsigned int* key_out;
  unsigned int* idx_out;
  unsigned long long n;   // elements of the source
  int shift;              // bit position of this pass's digit: 0, 8, 16, 24
  unsigned int* counts;   // [256][ntiles]: per-tile digit counts, then (after the scan) bucket bases
  unsigned int ntiles;
};

PINB_HD unsigned int sort_float_bits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  union { float f; unsigned int u; } c;
  c.f = f;
  return c.u;
#endif
}

// element e of the source: false if it is filtered out
PINB_HD bool sort_fetch(const SortPassParams& p, unsigned long long e, unsigned int& key, unsigned int& idx) {
  if (p.fmax) {
    const float f = p.fmax[e];
    if (!(f >= p.f_last)) return false;  // NaN never passes, like the reference's comparison
    key = ~sort_float_bits(f);
    idx = (unsigned int)e;
    return true;
  }
  key = p.key_in[e];
  idx = p.idx_in[e];
  return true;
}

// phase shared by count and scatter: cnt[d*ROW + t] = elements of digit d in thread t's chunk
template <class Ctx> PINB_HD void sort_tile_count(Ctx& ctx, unsigned short* cnt, const SortPassParams& p) {
  const int t = ctx.tid();
  for (int d = 0; d < 256; d++) cnt[d * SORT_ROW + t] = 0;
  const unsigned long long e0 = (unsigned long long)ctx.bid() * SORT_TILE + (unsigned long long)t * SORT_IPT;
  for (int i = 0; i < SORT_IPT; i++) {
    const unsigned long long e = e0 + i;
    unsigned int key, idx;
    if (e < p.n && sort_fetch(p, e, key, idx)) cnt[((key >> p.shift) & 255u) * SORT_ROW + t]++;
  }
  ctx.sync();
}

template <class Ctx> PINB_HD void sort_count_body(Ctx& ctx, unsigned short* cnt, const SortPassParams& p) {
  sort_tile_count(ctx, cnt, p);
  // thread t sums the digit rows t, t + NT, ...
  for (int d = ctx.tid(); d < 256; d += SORT_NT) {
    unsigned int s = 0;
    for (int i = 0; i < SORT_NT; i++) s += cnt[d * SORT_ROW + i];
    p.counts[(size_t)d * p.ntiles + ctx.bid()] = s;
  }
}

template <class Ctx> PINB_HD void sort_scatter_body(Ctx& ctx, unsigned short* cnt, unsigned int* base, const SortPassParams& p) {
  sort_tile_count(ctx, cnt, p);
  const int t = ctx.tid();
  for (int d = t; d < 256; d += SORT_NT) {
    // digit row d becomes exclusive prefix sums over the threads; its global base comes from the scan
    unsigned int run = 0;
    for (int i = 0; i < SORT_NT; i++) {
      const unsigned int c = cnt[d * SORT_ROW + i];
      cnt[d * SORT_ROW + i] = (unsigned short)run;
      run += c;
    }
    base[d] = p.counts[(size_t)d * p.ntiles + ctx.bid()];
  }
  ctx.sync();
  const unsigned long long e0 = (unsigned long long)ctx.bid() * SORT_TILE + (unsigned long long)t * SORT_IPT;
  for (int i = 0; i < SORT_IPT; i++) {
    const unsigned long long e = e0 + i;
    unsigned int key, idx;
    if (e < p.n && sort_fetch(p, e, key, idx)) {
      const unsigned int d = (key >> p.shift) & 255u;
      const unsigned int pos = base[d] + cnt[d * SORT_ROW + t]++;
      p.key_out[pos] = key;
      p.idx_out[pos] = idx;
    }
  }
}

// exclusive prefix sum of `len` counters in place, one block; *total = their sum (fewer than 2^32
// elements, so partial sums fit u32).  scratch: nthreads u32 in shared memory.
template <class Ctx> PINB_HD void sort_scan_body(Ctx& ctx, unsigned int* scratch, unsigned int* counts, unsigned long long len,
                                                 unsigned long long* total) {
  const int t = ctx.tid(), nt = ctx.nthreads();
  const unsigned long long chunk = (len + nt - 1) / nt;
  const unsigned long long b = (unsigned long long)t * chunk;
  const unsigned long long e = b + chunk < len ? b + chunk : len;
  unsigned int s = 0;
  for (unsigned long long i = b; i < e; i++) s += counts[i];
  scratch[t] = s;
  ctx.sync();
  if (t == 0) {
    unsigned int run = 0;
    for (int i = 0; i < nt; i++) {
      const unsigned int c = scratch[i];
      scratch[i] = run;
      run += c;
    }
    *total = run;
  }
  ctx.sync();
  unsigned int run = scratch[t];
  for (unsigned long long i = b; i < e; i++) {
    const unsigned int c = counts[i];
    counts[i] = run;
    run += c;
  }
}

}  // namespace pinb
