// Stand-alone tuning harness: times variants of the FFT pass kernels on synthetic 1024^3 buffers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -o tools/passbench tools/passbench.cu
// Not part of the product library.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>
#include <cmath>
#include "../pinocchio_b200/csrc/devctx.cuh"
#include "../pinocchio_b200/csrc/kernels.cuh"

using namespace pinb;

#define CKE(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int L, int TK, int DIR, int MINB>
__global__ void __launch_bounds__(XPlan<L>::TPL* TK, MINB) xk(const __grid_constant__ XPassParams p) {
  extern __shared__ double2 smem[];
  DevCtx ctx;
  xpass_body<L, DIR, false>(ctx, smem, p);  // TK must equal XCfg<L, DIR>::TK
}
__device__ long long* g_dbg;
struct TimedCtx : DevCtx {
  __device__ __forceinline__ void mark(int k) const {
    if (threadIdx.x == 0 && blockIdx.x % 997 == 3) g_dbg[(blockIdx.x / 997) * 32 + k] = clock64();
  }
};
template <int L, int TK>
__global__ void __launch_bounds__(Plan<L, false>::TPL* TK, 1) yk_timed(const __grid_constant__ YPassParams p) {
  extern __shared__ double2 smem[];
  TimedCtx ctx;
  ypass_body<L, +1>(ctx, smem, p);
}
template <int L, int TK, int DIR, int MINB>
__global__ void __launch_bounds__(Plan<L, false>::TPL* TK, MINB) yk(const __grid_constant__ YPassParams p) {
  extern __shared__ double2 smem[];
  DevCtx ctx;
  ypass_body<L, DIR>(ctx, smem, p);
}
template <int M, int TL, int CG, int MINB, int CPT>
__global__ void __launch_bounds__(ZShape<M, TL, CG>::NT, MINB) zck(const __grid_constant__ CollapseParams p) {
  extern __shared__ double2 smem[];
  using ZS = ZShape<M, TL, CG>;
  double* spl = reinterpret_cast<double*>(smem + ZS::fft_elems(6));
  double* scratch = spl + p.spl_doubles;
  DevCtx ctx;
  zpass_collapse_body<M, TL, CG, DevCtx, CPT>(ctx, smem, spl, scratch, p);
}
template <int M, int TL, int CG>
__global__ void __launch_bounds__(ZShape<M, TL, CG>::NT, (768 / ZShape<M, TL, CG>::NT > 8 ? 8 : 768 / ZShape<M, TL, CG>::NT)) zok(const __grid_constant__ ZOutParams p) {
  extern __shared__ double2 smem[];
  DevCtx ctx;
  zpass_out_body<M, TL, CG>(ctx, smem, p);
}

template <int M> static void fill_pretw(ZSrc& zs) {
  constexpr int TPL = Plan<M, true>::TPL, RMAX = Plan<M, true>::RMAX;
  for (int s = 0; s < RMAX; s++) { const double a = 2.0 * M_PI * TPL * s / (2.0 * M); zs.pretw[s] = make_double2(cos(a), sin(a)); }
}

struct Timer {
  cudaEvent_t a, b;
  Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
  template <class F> float run(F f, int reps = 3) {
    f();  // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int i = 0; i < reps; i++) {
      cudaEventRecord(a);
      f();
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      if (ms < best) best = ms;
    }
    CKE(cudaGetLastError());
    return best;
  }
};

static Geom geom(int N) {
  Geom g; g.N = N; g.M = N / 2; g.P = g.M + 8; g.lx = N; g.ly = N; g.x0 = 0; g.y0 = 0; g.knorm = 2 * M_PI / N; return g;
}

template <int N, int TK, int MINB> void bench_x(const Geom& g, double2* src, double2** A, const double2* tw, const double* gauss, int pmask, const char* tag, int pf = 0) {
  constexpr int NT = XPlan<N>::TPL * TK;
  const size_t smem = (size_t)N * TK * sizeof(double2);
  CKE(cudaFuncSetAttribute(xk<N, TK, +1, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  XPassParams p{};
  p.src = src; p.dst[0].r[0] = A[0]; p.dst[1].r[0] = A[1]; p.dst[2].r[0] = A[2]; p.dst_klayout = 0; p.lx_shift = 10; p.pmask = pmask; p.ntiles_z = g.M / TK;
  p.kf.gauss = gauss; p.kf.scalar = 1e-9; p.kf.green = 1; p.kf.times_i = 0; p.g = g; p.tw = tw; p.nblocks = g.ly * p.ntiles_z;
  Timer t;
  float ms = t.run([&] { xk<N, TK, +1, MINB><<<g.ly * p.ntiles_z, NT, smem>>>(p); });
  int nout = __builtin_popcount(pmask);
  double gb = (1 + nout) * 16.0 * g.N * g.N * g.M / 1e9;
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, xk<N, TK, +1, MINB>);
  int nb; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, xk<N, TK, +1, MINB>, NT, smem);
  printf("xpass %-10s pf=%4d TK=%d minb=%d nout=%d regs=%d blocks/SM=%d : %.2f ms  %.0f GB/s\n", tag, pf, TK, MINB, nout, fa.numRegs, nb, ms, gb / ms * 1e3);
}

template <int N, int TK, int MINB> void bench_y(const Geom& g, double2** A, double2** B, const double2* tw, int njobs, const char* tag, int pf = 0) {
  constexpr int NT = Plan<N, false>::TPL * TK;
  const size_t smem = (size_t)N * TK * sizeof(double2);
  CKE(cudaFuncSetAttribute(yk<N, TK, +1, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  YPassParams p{};
  for (int i = 0; i < 3; i++) p.src[i] = A[i];
  for (int i = 0; i < 6; i++) p.dst[i] = B[i];
  static const YJob jobs[6] = {{2, 0, 0}, {0, 2, 1}, {0, 0, 2}, {1, 1, 3}, {1, 0, 4}, {0, 1, 5}};
  for (int i = 0; i < njobs; i++) p.job[i] = jobs[i];
  p.njobs = njobs; p.dst_klayout = 0; p.ly_shift = 10; p.ntiles_z = g.M / TK; p.g = g; p.tw = tw; p.nblocks = g.lx * p.ntiles_z; p.nsrc = njobs == 6 ? 3 : 1; if (njobs == 1) p.job[0] = YJob{0, 0, 0};
  Timer t;
  float ms = t.run([&] { yk<N, TK, +1, MINB><<<g.lx * p.ntiles_z, NT, smem>>>(p); });
  double gb = (njobs == 6 ? 9 : 2 * njobs) * 16.0 * g.N * g.N * g.M / 1e9;
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, yk<N, TK, +1, MINB>);
  int nb; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, yk<N, TK, +1, MINB>, NT, smem);
  if (njobs == 6 && TK == 8) {
    long long* dbg; CKE(cudaMalloc(&dbg, 80 * 32 * 8)); CKE(cudaMemset(dbg, 0, 80 * 32 * 8));
    CKE(cudaMemcpyToSymbol(g_dbg, &dbg, sizeof(dbg)));
    CKE(cudaFuncSetAttribute(yk_timed<N, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    yk_timed<N, 8><<<g.lx * p.ntiles_z, NT, smem>>>(p);
    CKE(cudaDeviceSynchronize());
    std::vector<long long> h(80 * 32);
    CKE(cudaMemcpy(h.data(), dbg, h.size() * 8, cudaMemcpyDeviceToHost));
    double acc[24] = {0}; int nb = 0;
    for (int b = 5; b < 60; b++) { if (!h[b * 32]) continue; nb++; for (int k = 1; k < 24; k++) acc[k] += (double)(h[b * 32 + k] - h[b * 32 + k - 1]); }
    printf("y-pass timeline (cycles, mean of %d blocks): per job [wait copies | stage0..last-load | issue+last stage+stores]\n", nb);
    for (int j = 0; j < 6; j++) printf("  job %d: gap %.0f | wait %.0f | stages %.0f | tail %.0f\n", j, j ? acc[4 * j] / nb : 0.0, acc[4 * j + 1] / nb, acc[4 * j + 2] / nb, acc[4 * j + 3] / nb);
  }
  printf("ypass %-10s pf=%4d TK=%d minb=%d njobs=%d regs=%d blocks/SM=%d : %.2f ms  %.0f GB/s\n", tag, pf, TK, MINB, njobs, fa.numRegs, nb, ms, gb / ms * 1e3);
}

template <int N, int TL, int CG, int MINB, int CPT> void bench_zc(const Geom& g, double2** B, const double2* tw, const double* spline, int nspl, float* fmax, int* rmax, double* sums, const char* tag) {
  constexpr int M = N / 2;
  using ZS = ZShape<M, TL, CG>;
  const size_t smem = ZS::fft_elems(6) * sizeof(double2) + spline_table_doubles(nspl) * sizeof(double) + 64 * sizeof(double);  // as the product kernel (k_zpass.cu): three blocks per SM
  CKE(cudaFuncSetAttribute(zck<M, TL, CG, MINB, CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CollapseParams p{};
  static const int kz[6] = {0, 0, 2, 0, 1, 1};
  for (int k = 0; k < 6; k++) { p.zs.src[k] = B[k]; p.zs.kzpow[k] = kz[k]; p.hdst[k] = nullptr; }
  p.zs.ncomp = 6; p.zs.has_nyq = 0; p.zs.dc_add = nullptr; p.g = g; p.tw = tw; p.spline = spline; p.nspl = nspl; p.spl_doubles = (int)spline_table_doubles(nspl);
  p.ismooth = 1; p.Fmax = fmax; p.Rmax = rmax; p.sums = sums; fill_pretw<M>(p.zs);
  Timer t;
  float ms = t.run([&] { zck<M, TL, CG, MINB, CPT><<<(unsigned)((size_t)g.lx * g.N / TL), ZS::NT, smem>>>(p); });
  double gb = (6 * 16.0 * g.N * g.N * g.M + 12.0 * g.N * g.N * g.N) / 1e9;
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, zck<M, TL, CG, MINB, CPT>);
  int nb; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, zck<M, TL, CG, MINB, CPT>, ZS::NT, smem);
  printf("zcollapse %-10s CPT=%d TL=%d CG=%d minb=%d regs=%d blocks/SM=%d smem=%zu : %.2f ms  %.0f GB/s\n", tag, CPT, TL, CG, MINB, fa.numRegs, nb, smem, ms, gb / ms * 1e3);
}

template <int N, int TL, int CG> void bench_zo(const Geom& g, double2** B, const double2* tw, float** fo, int mode, double* acc, const char* tag) {
  constexpr int M = N / 2;
  using ZS = ZShape<M, TL, CG>;
  const int ncomp = 3;
  const size_t smem = ZS::fft_elems(ncomp) * sizeof(double2);
  CKE(cudaFuncSetAttribute(zok<M, TL, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ZS::fft_elems(6) * sizeof(double2))));
  ZOutParams p{};
  for (int k = 0; k < ncomp; k++) { p.zs.src[k] = B[k]; p.zs.kzpow[k] = k == 2; p.fdst[k] = fo[k]; p.hsrc[k] = (const double*)B[3 + k]; p.weight[k] = 2.0; }
  p.zs.ncomp = ncomp; p.zs.has_nyq = 1; p.zs.dc_add = nullptr; p.g = g; p.tw = tw; p.mode = mode; p.acc = acc; fill_pretw<M>(p.zs);
  Timer t;
  float ms = t.run([&] { zok<M, TL, CG><<<(unsigned)((size_t)g.lx * g.N / TL), ZS::NT, smem>>>(p); });
  int nb; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, zok<M, TL, CG>, ZS::NT, smem);
  printf("zout mode%d %-8s TL=%d CG=%d blocks/SM=%d : %.2f ms\n", mode, tag, TL, CG, nb, ms);
}

int main(int argc, char** argv) {
  constexpr int N = 1024;
  Geom g = geom(N);
  const size_t fe = (size_t)g.N * g.N * g.P;
  double2 *src, *A[3], *B[6], *tw;
  CKE(cudaMalloc(&src, fe * sizeof(double2)));
  for (auto& a : A) CKE(cudaMalloc(&a, fe * sizeof(double2)));
  for (auto& b : B) CKE(cudaMalloc(&b, fe * sizeof(double2)));
  // fill src with a smooth random-ish field (values O(1)); B gets filled by the passes
  {
    std::vector<double2> h(1 << 20);
    for (size_t i = 0; i < h.size(); i++) h[i] = make_double2(sin(0.37 * i) * 1e3, cos(0.91 * i) * 1e3);
    for (size_t off = 0; off < fe; off += h.size()) {
      size_t n = std::min(h.size(), fe - off);
      CKE(cudaMemcpy(src + off, h.data(), n * sizeof(double2), cudaMemcpyHostToDevice));
    }
  }
  std::vector<double2> htw(N);
  for (int k = 0; k < N; k++) htw[k] = make_double2(cos(2 * M_PI * k / N), sin(2 * M_PI * k / N));
  CKE(cudaMalloc(&tw, N * sizeof(double2)));
  CKE(cudaMemcpy(tw, htw.data(), N * sizeof(double2), cudaMemcpyHostToDevice));
  std::vector<double> hg(g.M + 1);
  for (int n = 0; n <= g.M; n++) hg[n] = exp(-0.5 * pow(g.knorm * n * 2.0, 2));
  double* gauss; CKE(cudaMalloc(&gauss, hg.size() * 8)); CKE(cudaMemcpy(gauss, hg.data(), hg.size() * 8, cudaMemcpyHostToDevice));
  // a plausible inverse-growth spline: x = log10 D in [-4, 0.12], y = log10 a
  const int nspl = 210;
  std::vector<double> sx(nspl), sy(nspl), spl;
  for (int i = 0; i < nspl; i++) { sy[i] = -4 + 0.02 * i; sx[i] = sy[i] - 0.15 * exp(3.0 * (sy[i] + 0.2)) / (1 + exp(3.0 * (sy[i] + 0.2))); }
  pack_spline(sx.data(), sy.data(), nspl, spl);
  double* dspl; CKE(cudaMalloc(&dspl, spl.size() * 8)); CKE(cudaMemcpy(dspl, spl.data(), spl.size() * 8, cudaMemcpyHostToDevice));
  float* fmax; int* rmax; double* sums;
  CKE(cudaMalloc(&fmax, (size_t)N * N * N * 4)); CKE(cudaMalloc(&rmax, (size_t)N * N * N * 4)); CKE(cudaMalloc(&sums, 16));
  CKE(cudaMemset(fmax, 0, (size_t)N * N * N * 4));

  // tests selected by name on the command line (default: the strided passes)
  auto want = [&](const char* name) {
    if (argc < 2) return std::string(name) == "x" || std::string(name) == "y";
    for (int i = 1; i < argc; i++) if (std::string(argv[i]) == name) return true;
    return false;
  };
  constexpr int YTK = PINB_YTK;
#ifndef PB_MINB
#define PB_MINB 1
#endif
  if (want("x")) {
    bench_x<N, 8, 1>(g, src, A, tw, gauss, 7, "base", 0);
    bench_x<N, 8, 1>(g, src, A, tw, gauss, 1, "base", 0);
  } else {  // the later passes need their inputs
    bench_x<N, 8, 1>(g, src, A, tw, gauss, 7, "fill", 0);
  }
  if (want("y")) {
    bench_y<N, YTK, PB_MINB>(g, A, B, tw, 6, "ytk", 0);
    bench_y<N, YTK, PB_MINB>(g, A, B, tw, 1, "ytk", 0);
  } else {
    bench_y<N, YTK, PB_MINB>(g, A, B, tw, 6, "fill", 0);
  }
  if (want("z")) {
#ifdef PINB_BENCH_NOEPI
    const char* tag = "NOEPI";
#else
    const char* tag = "classic";
#endif
    bench_zc<N, 1, 6, 3, 1>(g, B, tw, dspl, nspl, fmax, rmax, sums, tag);
    bench_zc<N, 1, 6, 2, 1>(g, B, tw, dspl, nspl, fmax, rmax, sums, tag);
  }
  if (want("z2")) {  // two cells per thread and iteration
    bench_zc<N, 1, 6, 3, 2>(g, B, tw, dspl, nspl, fmax, rmax, sums, "cpt2");
    bench_zc<N, 1, 6, 2, 2>(g, B, tw, dspl, nspl, fmax, rmax, sums, "cpt2");
  }
  return 0;
}
