// Times the product launchers (libpinb200.so) on ONE rank's slab of an N^3 grid split over P
// ranks, on a single GPU: every "peer" destination is a disjoint block of local memory, so the
// numbers are the SM/HBM side of each pass without NVLink.  2048^3 over 8 ranks has the per-GPU
// volume of 1024^3 on one GPU and cannot be run through the engine on fewer than 8 GPUs.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I pinocchio_b200/csrc \
//        tools/slabbench.cu -L pinocchio_b200 -lpinb200 -Xlinker -rpath=$PWD/pinocchio_b200 -o tools/slabbench
//   tools/slabbench N P [reps]
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "launch.h"
#include "spline_pack.h"

using namespace pinb;

#define CKE(x)                                                                              \
  do {                                                                                      \
    cudaError_t e_ = (x);                                                                   \
    if (e_ != cudaSuccess) {                                                                \
      fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));  \
      exit(1);                                                                              \
    }                                                                                       \
  } while (0)

template <class F> static float timed(int reps, F f) {
  cudaEvent_t a, b;
  CKE(cudaEventCreate(&a));
  CKE(cudaEventCreate(&b));
  f();
  CKE(cudaDeviceSynchronize());
  CKE(cudaEventRecord(a));
  for (int i = 0; i < reps; i++) f();
  CKE(cudaEventRecord(b));
  CKE(cudaDeviceSynchronize());
  float ms;
  CKE(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 2048, P = argc > 2 ? atoi(argv[2]) : 8, reps = argc > 3 ? atoi(argv[3]) : 3;
  if (!grid_supported(N) || P < 1 || P > PINB_MAXR || N % P) { fprintf(stderr, "bad N/P\n"); return 1; }
  Geom g;
  g.N = N; g.M = N / 2; g.P = g.M + 8; g.lx = N / P; g.ly = N / P; g.x0 = 0; g.y0 = 0; g.knorm = 2 * M_PI / N;
  int sh = 0;
  while ((1 << sh) < g.lx) sh++;
  const size_t fe = (size_t)g.lx * g.N * g.P;  // elements of one slab field (either layout)
  const double gb = fe * 16 / 1e9;
  double2 *src, *A[3], *B[6], *tw;
  CKE(cudaMalloc(&src, fe * sizeof(double2)));
  for (auto& a : A) CKE(cudaMalloc(&a, fe * sizeof(double2)));
  for (auto& b : B) CKE(cudaMalloc(&b, fe * sizeof(double2)));
  {
    std::vector<double2> h(1 << 20);
    for (size_t i = 0; i < h.size(); i++) h[i] = make_double2(sin(0.37 * i) * 1e3, cos(0.91 * i) * 1e3);
    for (size_t off = 0; off < fe; off += h.size()) {
      const size_t n = std::min(h.size(), fe - off);
      CKE(cudaMemcpy(src + off, h.data(), n * sizeof(double2), cudaMemcpyHostToDevice));
    }
  }
  std::vector<double2> htw(N);
  for (int k = 0; k < N; k++) htw[k] = make_double2(cos(2 * M_PI * k / N), sin(2 * M_PI * k / N));
  CKE(cudaMalloc(&tw, N * sizeof(double2)));
  CKE(cudaMemcpy(tw, htw.data(), N * sizeof(double2), cudaMemcpyHostToDevice));
  std::vector<double> hg(g.M + 1);
  for (int n = 0; n <= g.M; n++) hg[n] = exp(-0.5 * pow(g.knorm * n * 2.0, 2));
  double* gauss;
  CKE(cudaMalloc(&gauss, hg.size() * 8));
  CKE(cudaMemcpy(gauss, hg.data(), hg.size() * 8, cudaMemcpyHostToDevice));
  const int nspl = 210;
  std::vector<double> sx(nspl), sy(nspl), spl;
  for (int i = 0; i < nspl; i++) {
    sy[i] = -4 + 0.02 * i;
    sx[i] = sy[i] - 0.15 * exp(3.0 * (sy[i] + 0.2)) / (1 + exp(3.0 * (sy[i] + 0.2)));
  }
  pack_spline(sx.data(), sy.data(), nspl, spl);
  double* dspl;
  CKE(cudaMalloc(&dspl, spl.size() * 8));
  CKE(cudaMemcpy(dspl, spl.data(), spl.size() * 8, cudaMemcpyHostToDevice));
  const size_t ncell = (size_t)g.lx * N * N;
  float* fmax; int* rmax; double* sums;
  CKE(cudaMalloc(&fmax, ncell * 4));
  CKE(cudaMalloc(&rmax, ncell * 4));
  CKE(cudaMalloc(&sums, 16));
  CKE(cudaMemset(fmax, 0, ncell * 4));
  CKE(cudaMemset(rmax, 0, ncell * 4));
  CKE(cudaMemset(sums, 0, 16));
  printf("slab of %d^3 over %d ranks: lx=%d, %.2f GB per field\n", N, P, g.lx, gb);

  // ---- inverse x pass, 3 powers; "owner o" = y-block o of the local R-layout field (disjoint stores)
  for (int pmask : {7, 1}) {
    XPassParams p{};
    p.src = src;
    for (int pw = 0; pw < 3; pw++)
      for (int o = 0; o < P; o++) p.dst[pw].r[o] = A[pw] + (size_t)o * g.ly * g.P;
    p.dst_klayout = 0; p.lx_shift = sh; p.pmask = pmask;
    p.ntiles_z = g.M / xpass_tk(N, +1);
    p.kf.gauss = gauss; p.kf.scalar = 1e-9; p.kf.green = 1; p.kf.times_i = 0; p.g = g; p.tw = tw;
    p.nblocks = g.ly * p.ntiles_z;
    const float ms = timed(reps, [&] { CKE(launch_xpass(N, +1, p, g.ly, 0)); });
    const int nj = pmask == 7 ? 3 : 1;
    printf("xpass inv pmask=%d tk=%d: %.2f ms  (%.0f GB/s algorithmic: 1 read + %d writes)\n", pmask, xpass_tk(N, +1), ms,
           (1 + nj) * gb / ms * 1e3, nj);
  }
  // ---- the same pass as the pipelined multi-GPU sweep runs it (dst_klayout = 2): own x planes into the local R-layout
  // field, everything else into local K-layout staging; lines longer than the tile are NOT split there (XCfg LOCAL).
  // Checked against the scattering kernel's result in the plain local K layout (dst_klayout = 1).
  {
    double2* S[3];
    for (auto& b : S) CKE(cudaMalloc(&b, fe * sizeof(double2)));
    XPassParams p{};
    p.src = src;
    p.lx_shift = sh; p.pmask = 7;
    p.kf.gauss = gauss; p.kf.scalar = 1e-9; p.kf.green = 1; p.kf.times_i = 0; p.g = g; p.tw = tw;
    // reference: every element in the K layout, by the kernel that scatters (split lines at N = 2048)
    for (int pw = 0; pw < 3; pw++) p.dst[pw].r[0] = B[pw];
    p.dst_klayout = 1;
    p.ntiles_z = g.M / xpass_tk(N, +1, 1);
    p.nblocks = g.ly * p.ntiles_z;
    const float ms1 = timed(reps, [&] { CKE(launch_xpass(N, +1, p, g.ly, 0)); });
    printf("xpass inv K-layout (scattering kernel, local stores) tk=%d: %.2f ms  (%.0f GB/s)\n", xpass_tk(N, +1, 1), ms1, 4 * gb / ms1 * 1e3);
    for (int pw = 0; pw < 3; pw++) { p.dst[pw].r[0] = S[pw]; p.dst[pw].r[1] = A[pw]; }
    p.dst_klayout = 2;
    p.ntiles_z = g.M / xpass_tk(N, +1, 2);
    p.nblocks = g.ly * p.ntiles_z;
    const float ms2 = timed(reps, [&] { CKE(launch_xpass(N, +1, p, g.ly, 0)); });
    printf("xpass inv staged (dst_klayout 2, all local) tk=%d: %.2f ms  (%.0f GB/s algorithmic: 1 read + 3 writes)\n", xpass_tk(N, +1, 2), ms2,
           4 * gb / ms2 * 1e3);
    // compare on the host, field by field, a few x planes at a time
    double worst = 0, scale = 0;
    const size_t plane = (size_t)g.ly * g.P;            // one x plane of the K-layout slab
    std::vector<double2> h1(plane), h2(plane);
    for (int pw = 0; pw < 3; pw++)
      for (int e = 0; e < N; e += 37) {
        CKE(cudaMemcpy(h1.data(), B[pw] + (size_t)e * plane, plane * sizeof(double2), cudaMemcpyDeviceToHost));
        const bool own = e >= g.x0 && e < g.x0 + g.lx;
        if (!own) CKE(cudaMemcpy(h2.data(), S[pw] + (size_t)e * plane, plane * sizeof(double2), cudaMemcpyDeviceToHost));
        else  // R layout: plane (e - x0) holds N rows of P; this rank's rows are y0 .. y0 + ly
          CKE(cudaMemcpy(h2.data(), A[pw] + (size_t)(e - g.x0) * g.N * g.P + (size_t)g.y0 * g.P, plane * sizeof(double2), cudaMemcpyDeviceToHost));
        for (size_t yl = 0; yl < (size_t)g.ly; yl++)
          for (int kz = 0; kz < g.M; kz++) {
            const double2 a = h1[yl * g.P + kz], b = h2[yl * g.P + kz];
            worst = std::max(worst, std::max(fabs(a.x - b.x), fabs(a.y - b.y)));
            scale = std::max(scale, std::max(fabs(a.x), fabs(a.y)));
          }
      }
    printf("staged vs scattering kernel: max |diff| %.3e of max |value| %.3e -> rel %.2e %s\n", worst, scale, worst / scale,
           worst <= 1e-13 * scale ? "OK" : "MISMATCH");
    for (auto& b : S) CKE(cudaFree(b));
  }
  // ---- inverse y pass, 6 jobs from 3 sources
  {
    YPassParams p{};
    for (int i = 0; i < 3; i++) p.src[i] = A[i];
    for (int i = 0; i < 6; i++) p.dst[i] = B[i];
    const YJob jobs[6] = {{0, 0, 2}, {0, 2, 1}, {2, 0, 0}, {1, 1, 3}, {1, 0, 4}, {0, 1, 5}};
    for (int i = 0; i < 6; i++) p.job[i] = jobs[i];
    p.njobs = 6; p.dst_klayout = 0; p.ly_shift = sh; p.ntiles_z = g.M / ypass_tk(N); p.g = g; p.tw = tw; p.nsrc = 3;
    p.nblocks = g.lx * p.ntiles_z;
    const float ms = timed(reps, [&] { CKE(launch_ypass(N, +1, p, g.lx, 0)); });
    printf("ypass inv 6 jobs tk=%d: %.2f ms  (%.0f GB/s algorithmic: 3 reads + 6 writes)\n", ypass_tk(N), ms, 9 * gb / ms * 1e3);
    p.njobs = 1; p.nsrc = 1; p.job[0] = YJob{0, 0, 0};
    const float ms1 = timed(reps, [&] { CKE(launch_ypass(N, +1, p, g.lx, 0)); });
    printf("ypass inv 1 job: %.2f ms  (%.0f GB/s)\n", ms1, 2 * gb / ms1 * 1e3);
    // forward y pass scattering to the K layout of P owners (local stand-ins: x-block o)
    for (int o = 0; o < P; o++) p.kdst.r[o] = src + (size_t)o * 0;  // same local K field: disjoint by (x0+xl, yl) only for o fixed
    p.dst_klayout = 1;
    const float ms2 = timed(reps, [&] { CKE(launch_ypass(N, -1, p, g.lx, 0)); });
    printf("ypass fwd 1 job (overlapping local stand-in stores): %.2f ms\n", ms2);
  }
  // ---- z pass with the collapse epilogue
  {
    CollapseParams c{};
    const int kz[6] = {0, 0, 2, 0, 1, 1};
    for (int i = 0; i < 6; i++) { c.zs.src[i] = B[i]; c.zs.kzpow[i] = kz[i]; }
    c.zs.ncomp = 6; c.zs.has_nyq = 0; c.zs.dc_add = nullptr;
    c.g = g; c.tw = tw; c.spline = dspl; c.nspl = nspl; c.spl_doubles = (int)spl.size(); c.ismooth = 1;
    c.Fmax = fmax; c.Rmax = rmax; c.sums = sums;
    const float ms = timed(reps, [&] { CKE(launch_zpass_collapse(N, c, (size_t)g.lx * N, 0)); });
    printf("zpass collapse: %.2f ms  (%.0f GB/s algorithmic: 6 half-complex reads + 8 B/cell)\n", ms,
           (6 * gb + ncell * 8 / 1e9) / ms * 1e3);
  }
  // ---- forward x pass (in place, local)
  {
    XPassParams p{};
    p.src = src; p.dst[0].r[0] = src; p.dst_klayout = 1; p.lx_shift = sh; p.pmask = 1;
    p.ntiles_z = g.M / xpass_tk(N, -1) + 1;
    p.kf.gauss = nullptr; p.kf.scalar = 1.0; p.g = g; p.tw = tw; p.nblocks = g.ly * p.ntiles_z;
    const float ms = timed(reps, [&] { CKE(launch_xpass(N, -1, p, g.ly, 0)); });
    printf("xpass fwd in place tk=%d: %.2f ms  (%.0f GB/s)\n", xpass_tk(N, -1), ms, 2 * gb / ms * 1e3);
  }
  return 0;
}
