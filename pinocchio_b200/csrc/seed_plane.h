// Host-side seed plane of GenIC (src/GenIC.c:482-990): the seed of column
// (i, j) is the m-th output of gsl_rng_mt19937(RandomSeed), m = get_map of the column's position on
// the square spiral around the origin.  Plain C++ (no CUDA): shared by the engine and by the
// test-only emulated ABI (tests/host/emu_abi.cpp).
#pragma once
#include <cstdint>
#include <vector>

namespace pinb {
struct MT19937 {
  uint32_t mt[624];
  int idx;
  explicit MT19937(uint32_t s) {
    if (s == 0) s = 4357;  // gsl default seed
    mt[0] = s;
    for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    idx = 624;
  }
  uint32_t next() {
    if (idx >= 624) {
      for (int k = 0; k < 624; k++) {
        const uint32_t yv = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
        mt[k] = mt[(k + 397) % 624] ^ (yv >> 1) ^ ((yv & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    uint32_t yv = mt[idx++];
    yv ^= (yv >> 11);
    yv ^= (yv << 7) & 0x9d2c5680u;
    yv ^= (yv << 15) & 0xefc60000u;
    yv ^= (yv >> 18);
    return yv;
  }
};

inline long long spiral_ordinal(long long px, long long py) {  // get_map, src/GenIC.c:840-855
  const long long mx = px < 0 ? -px : px, my = py < 0 ? -py : py;
  const long long l = 2 * (mx > my ? mx : my);
  const long long c = (py > px) + (px > 0) * (px == py);
  const long long dd = c ? l * 3 + px + py : l - px - py;
  return (l - 1) * (l - 1) + dd;
}

inline void build_seed_plane(int N, int random_seed, std::vector<unsigned int>& seeds) {
  const int N2 = N / 2;
  long long mmax = 0;
  std::vector<long long> ord((size_t)N * N);
  for (int j = 0; j < N; j++)
    for (int i = 0; i < N; i++) {
      const long long sx = i >= N2 ? i - N : i, sy = j >= N2 ? j - N : j;
      const long long m = spiral_ordinal(sx, sy);
      ord[(size_t)j * N + i] = m;
      if (m > mmax) mmax = m;
    }
  std::vector<unsigned int> out((size_t)mmax);
  MT19937 rng((uint32_t)random_seed);
  for (long long k = 0; k < mmax; k++) out[(size_t)k] = rng.next();
  seeds.resize((size_t)N * N);
  for (size_t k = 0; k < seeds.size(); k++) seeds[k] = out[(size_t)(ord[k] - 1)];
}

}  // namespace pinb
