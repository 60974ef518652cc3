// sm_100a kernels and C entry point of the batched set_scaledep_GM integrals (scaledep_gm.cuh).
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/pinb200.h"
#include "devctx.cuh"
#include "scaledep_gm.cuh"

namespace pinb {

constexpr int SDGM_NT = 256, SDGM_RC = 8;

__global__ void __launch_bounds__(256) sdgm_window_kernel(const __grid_constant__ SdgmParams p) {
  DevCtx ctx;
  sdgm_window_body(ctx, p);
}
__global__ void __launch_bounds__(SDGM_NT) sdgm_integrate_kernel(const __grid_constant__ SdgmParams p) {
  __shared__ double scratch[SDGM_NT];
  DevCtx ctx;
  sdgm_integrate_body<SDGM_NT, SDGM_RC>(ctx, scratch, p);
}

void set_create_error(const std::string& s);  // engine.cu: what pinb200_last_error(NULL) returns

}  // namespace pinb

using namespace pinb;

#define SD_CK(x)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (x);                                                                                \
    if (e_ != cudaSuccess) {                                                                             \
      set_create_error(std::string("pinb200_scaledep_variances: ") + #x + ": " + cudaGetErrorString(e_)); \
      cleanup();                                                                                         \
      return 1;                                                                                          \
    }                                                                                                    \
  } while (0)

extern "C" int pinb200_scaledep_variances(const pinb200_sdgm_desc* d, double* out) {
  std::vector<void*> owned;
  auto cleanup = [&] { for (void* q : owned) cudaFree(q); };
  if (!d || !out || !d->logk || !d->a_dens || !d->a_disp || !d->log10_growth || !d->fomega || !d->radius_dens || !d->radius_disp) {
    set_create_error("pinb200_scaledep_variances: null argument");
    return 1;
  }
  if (d->nnodes < 1 || d->nkbins < 1 || d->ntimes < 1 || d->nsmooth < 1 || d->nsmooth > 64 || !(d->dlogk > 0.0) ||
      (long long)d->nnodes * d->nsmooth > (1ll << 28)) {
    set_create_error("pinb200_scaledep_variances: nnodes, nkbins, ntimes >= 1, 1 <= nsmooth <= 64, dlogk > 0 required");
    return 1;
  }
  SD_CK(cudaSetDevice(d->device));
  cudaStream_t st = nullptr;
  auto up = [&](const double* h, size_t n, const double** dev) -> cudaError_t {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(double));
    if (e != cudaSuccess) return e;
    owned.push_back(q);
    *dev = static_cast<const double*>(q);
    return cudaMemcpyAsync(q, h, n * sizeof(double), cudaMemcpyHostToDevice, st);
  };
  SdgmParams p{};
  p.n = d->nnodes; p.nk = d->nkbins; p.nt = d->ntimes; p.ns = d->nsmooth;
  p.logkmin = d->logkmin; p.dlogk = d->dlogk;
  const size_t n = p.n, tab = (size_t)p.nk * p.nt;
  SD_CK(up(d->logk, n, &p.logk));
  SD_CK(up(d->a_dens, n, &p.a_dens));
  SD_CK(up(d->a_disp, n, &p.a_disp));
  SD_CK(up(d->log10_growth, tab, &p.lg));
  SD_CK(up(d->fomega, tab, &p.fo));
  SD_CK(up(d->radius_dens, p.ns, &p.r_dens));
  SD_CK(up(d->radius_disp, p.ns, &p.r_disp));
  void* q = nullptr;
  SD_CK(cudaMalloc(&q, 2 * (size_t)p.ns * n * sizeof(double)));
  owned.push_back(q);
  p.T = static_cast<double*>(q);
  const size_t nout = 3 * (size_t)p.ns * p.nt;
  SD_CK(cudaMalloc(&q, nout * sizeof(double)));
  owned.push_back(q);
  p.out = static_cast<double*>(q);
  const long long nw = 2ll * p.ns * p.n;
  sdgm_window_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(p);
  SD_CK(cudaGetLastError());
  sdgm_integrate_kernel<<<(unsigned)p.nt, SDGM_NT, 0, st>>>(p);
  SD_CK(cudaGetLastError());
  SD_CK(cudaMemcpyAsync(out, p.out, nout * sizeof(double), cudaMemcpyDeviceToHost, st));
  SD_CK(cudaStreamSynchronize(st));
  cleanup();
  return 0;
}
