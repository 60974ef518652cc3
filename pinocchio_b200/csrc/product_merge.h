// Host side of the products[] download when the record has bytes that are not ours.
//
// The reference fills product_data member by member (Fmax/Rmax in compute_collapse_times,
// src/collapse_times.c:587-590; Vel* in write_from_rvector_to_products, src/fmax-pfft.c:563-631)
// and leaves the other members alone: with -DRECOMPUTE_DISPLACEMENTS the four *_prev[3] members hold
// the displacements of the previous redshift segment (shift_all_displacements, src/fragment.c:834-845,
// runs right BEFORE compute_displacements(0,0,z)), with -DSNAPSHOT zacc and group_ID belong to the
// fragmentation.  So a download may overwrite whole records only when every byte of the record is
// one of the members it delivers (the default 56-byte record); otherwise the packed records are
// staged and only the members are copied.  Plain C++ (shared with tests/host/emu_abi.cpp).
#pragma once
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/pinb200.h"

namespace pinb {

struct MemberRange { int off, len; };

// byte ranges of the members a download delivers; has_fr: Fmax/Rmax computed; has_vel[v]: field v computed
inline std::vector<MemberRange> product_members(const pinb200_product_layout& L, bool has_fr, const bool has_vel[4]) {
  std::vector<MemberRange> m;
  const int pf = L.prodfloat_bytes;
  if (has_fr && L.off_Rmax >= 0) m.push_back({L.off_Rmax, 4});
  if (has_fr && L.off_Fmax >= 0) m.push_back({L.off_Fmax, pf});
  const int off[4] = {L.off_Vel, L.off_Vel_2LPT, L.off_Vel_3LPT_1, L.off_Vel_3LPT_2};
  for (int v = 0; v < 4; v++)
    if (has_vel[v] && off[v] >= 0) m.push_back({off[v], 3 * pf});
  return m;
}

// true when the members tile the whole record: the packed records may be copied over the caller's
inline bool members_cover_record(const std::vector<MemberRange>& m, size_t stride) {
  size_t bytes = 0;
  for (const MemberRange& r : m) bytes += (size_t)r.len;
  return bytes == stride;
}

// dst[i].member = src[i].member for every listed member, i < n; other bytes of dst are untouched
inline void merge_product_members(unsigned char* dst, const unsigned char* src, size_t stride, size_t n,
                                  const std::vector<MemberRange>& m, int nthreads = 8) {
  auto work = [&](size_t b, size_t e) {
    for (size_t i = b; i < e; i++)
      for (const MemberRange& r : m) std::memcpy(dst + i * stride + r.off, src + i * stride + r.off, (size_t)r.len);
  };
  if (n < (size_t)1 << 16 || nthreads <= 1) { work(0, n); return; }
  std::vector<std::thread> th;
  const size_t per = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; t++) {
    const size_t b = (size_t)t * per, e = b + per < n ? b + per : n;
    if (b < e) th.emplace_back(work, b, e);
  }
  for (auto& x : th) x.join();
}

}  // namespace pinb
