/* TEST INFRASTRUCTURE (oracle/_ref): the PFFT entry points that the reference's src/fmax-pfft.c
 * calls, implemented for ONE task on an in-repo FFT, so that the reference's own compute_fmax /
 * compute_derivative / LPT code can run verbatim in an image without PFFT, FFTW and MPI.
 *
 * Semantics reproduced (PFFT 1.0.8 / FFTW 3.3.10 manuals): r2c is the unnormalised forward DFT
 * (sign -1) of a real [N0][N1][N2] array into [N0][N1][N2/2+1] complex; c2r is the unnormalised
 * backward DFT (sign +1) that reads only the non-redundant half and, like FFTW, only the real
 * parts of the self-conjugate kz = 0 and kz = N2/2 entries of each z line.  One task owns the
 * whole box: local sizes = global sizes, starts = 0, non-transposed layout (flags ignored except
 * that TRANSPOSED_* is refused).  The c2r transform destroys its input, as FFTW's may.
 *
 * FFT: Stockham autosort, radix 4 (+ one radix-2 stage), power-of-two lengths, OpenMP over
 * blocks of lines.  This is NOT FFTW: timings of oracle/_ref are those of "reference C code on
 * shimmed MPI/PFFT/GSL" and are labelled so wherever they are reported.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

#include <pfft.h>

typedef struct { double re, im; } cpx;

struct pfft_plan_s {
  int n[3];
  int sign;
  double* real;
  cpx* cplx;
};

/* ---- Stockham FFT of `s` interleaved sequences of length n: x[q + s*k], q < s -------------- */
static cpx* tw_table(int n, int sign) {
  /* w[k] = exp(sign * 2 pi i k / n); cached per (n, sign) */
  static cpx* cache[2][32];
  int lg = 0;
  while ((1 << lg) < n) lg++;
  cpx** slot = &cache[sign > 0][lg];
  if (!*slot) { /* first use happens outside parallel regions (make_plan / test hooks) */
    cpx* w = malloc((size_t)n * sizeof(cpx));
    for (int k = 0; k < n; k++) {
      const double a = 2.0 * M_PI * k / n;
      w[k].re = cos(a);
      w[k].im = sign * sin(a);
    }
    *slot = w;
  }
  return *slot;
}

/* returns the buffer (x or y) that holds the result */
static cpx* stockham(int n, int s, int sign, cpx* x, cpx* y) {
  const int n_total = n;
  const cpx* w = tw_table(n_total, sign);
  int wstep = 1; /* twiddle of the current sub-length n is w[p * wstep] */
  while (n >= 4) {
    const int n1 = n / 4;
    for (int p = 0; p < n1; p++) {
      const cpx w1 = w[p * wstep], w2 = w[2 * p * wstep], w3 = w[3 * p * wstep];
      const cpx* xa = x + (size_t)s * p;
      const cpx* xb = x + (size_t)s * (p + n1);
      const cpx* xc = x + (size_t)s * (p + 2 * n1);
      const cpx* xd = x + (size_t)s * (p + 3 * n1);
      cpx* y0 = y + (size_t)s * (4 * p);
      cpx* y1 = y0 + s;
      cpx* y2 = y1 + s;
      cpx* y3 = y2 + s;
      for (int q = 0; q < s; q++) {
        const double apc_r = xa[q].re + xc[q].re, apc_i = xa[q].im + xc[q].im;
        const double amc_r = xa[q].re - xc[q].re, amc_i = xa[q].im - xc[q].im;
        const double bpd_r = xb[q].re + xd[q].re, bpd_i = xb[q].im + xd[q].im;
        /* sign * i * (b - d) */
        const double jr = -sign * (xb[q].im - xd[q].im), ji = sign * (xb[q].re - xd[q].re);
        y0[q].re = apc_r + bpd_r;
        y0[q].im = apc_i + bpd_i;
        const double t1r = amc_r + jr, t1i = amc_i + ji;
        y1[q].re = t1r * w1.re - t1i * w1.im;
        y1[q].im = t1r * w1.im + t1i * w1.re;
        const double t2r = apc_r - bpd_r, t2i = apc_i - bpd_i;
        y2[q].re = t2r * w2.re - t2i * w2.im;
        y2[q].im = t2r * w2.im + t2i * w2.re;
        const double t3r = amc_r - jr, t3i = amc_i - ji;
        y3[q].re = t3r * w3.re - t3i * w3.im;
        y3[q].im = t3r * w3.im + t3i * w3.re;
      }
    }
    n /= 4;
    s *= 4;
    wstep *= 4;
    cpx* t = x; x = y; y = t;
  }
  if (n == 2) {
    const cpx* xa = x;
    const cpx* xb = x + s;
    cpx* y0 = y;
    cpx* y1 = y + s;
    for (int q = 0; q < s; q++) {
      const double ar = xa[q].re, ai = xa[q].im, br = xb[q].re, bi = xb[q].im;
      y0[q].re = ar + br; y0[q].im = ai + bi;
      y1[q].re = ar - br; y1[q].im = ai - bi;
    }
    cpx* t = x; x = y; y = t;
  }
  return x;
}

/* transform along an axis with element stride `stride` (in complex elements) for `count`
 * consecutive positions q (contiguous in memory), in blocks of QB positions */
#define QB 32
static void fft_axis(cpx* base, int n, size_t stride, size_t count, int nouter, size_t outer_stride, int sign) {
  const size_t nblk = (count + QB - 1) / QB;
#pragma omp parallel
  {
    cpx* a = malloc((size_t)n * QB * sizeof(cpx));
    cpx* b = malloc((size_t)n * QB * sizeof(cpx));
#pragma omp for schedule(static)
    for (size_t job = 0; job < nblk * (size_t)nouter; job++) {
      cpx* plane = base + (job / nblk) * outer_stride;
      const size_t q0 = (job % nblk) * QB;
      const int qb = (int)(count - q0 < QB ? count - q0 : QB);
      for (int k = 0; k < n; k++) memcpy(a + (size_t)k * qb, plane + q0 + (size_t)k * stride, qb * sizeof(cpx));
      const cpx* r = stockham(n, qb, sign, a, b);
      for (int k = 0; k < n; k++) memcpy(plane + q0 + (size_t)k * stride, r + (size_t)k * qb, qb * sizeof(cpx));
    }
    free(a);
    free(b);
  }
}

static int is_pow2(int n) { return n >= 2 && (n & (n - 1)) == 0; }

static void exec_r2c(const struct pfft_plan_s* p) {
  const int N0 = p->n[0], N1 = p->n[1], N2 = p->n[2], NC = N2 / 2 + 1;
  const size_t rows = (size_t)N0 * N1;
  /* z: real line -> complex line (zero imaginary part), keep kz <= N2/2 */
#pragma omp parallel
  {
    cpx* a = malloc((size_t)N2 * sizeof(cpx));
    cpx* b = malloc((size_t)N2 * sizeof(cpx));
#pragma omp for schedule(static)
    for (size_t r = 0; r < rows; r++) {
      const double* in = p->real + r * N2;
      for (int k = 0; k < N2; k++) { a[k].re = in[k]; a[k].im = 0.0; }
      const cpx* res = stockham(N2, 1, -1, a, b);
      memcpy(p->cplx + r * NC, res, (size_t)NC * sizeof(cpx));
    }
    free(a);
    free(b);
  }
  /* y: stride NC inside every x plane */
  fft_axis(p->cplx, N1, NC, NC, N0, (size_t)N1 * NC, -1);
  /* x: stride N1*NC */
  fft_axis(p->cplx, N0, (size_t)N1 * NC, (size_t)N1 * NC, 1, 0, -1);
}

static void exec_c2r(const struct pfft_plan_s* p) {
  const int N0 = p->n[0], N1 = p->n[1], N2 = p->n[2], NC = N2 / 2 + 1;
  const size_t rows = (size_t)N0 * N1;
  fft_axis(p->cplx, N0, (size_t)N1 * NC, (size_t)N1 * NC, 1, 0, +1);
  fft_axis(p->cplx, N1, NC, NC, N0, (size_t)N1 * NC, +1);
#pragma omp parallel
  {
    cpx* a = malloc((size_t)N2 * sizeof(cpx));
    cpx* b = malloc((size_t)N2 * sizeof(cpx));
#pragma omp for schedule(static)
    for (size_t r = 0; r < rows; r++) {
      const cpx* in = p->cplx + r * NC;
      /* Hermitian extension; the real part of the result then ignores Im of kz = 0 and N2/2 */
      for (int k = 0; k < NC; k++) a[k] = in[k];
      for (int k = NC; k < N2; k++) { a[k].re = in[N2 - k].re; a[k].im = -in[N2 - k].im; }
      a[0].im = 0.0;
      a[N2 / 2].im = 0.0;
      const cpx* res = stockham(N2, 1, +1, a, b);
      double* out = p->real + r * N2;
      for (int k = 0; k < N2; k++) out[k] = res[k].re;
    }
    free(a);
    free(b);
  }
}

/* ---- PFFT entry points ---------------------------------------------------------------------- */
void pfft_init(void) {}
void pfft_cleanup(void) {}

ptrdiff_t pfft_local_size_dft_r2c_3d(const ptrdiff_t* n, MPI_Comm comm, unsigned flags, ptrdiff_t* local_ni, ptrdiff_t* local_i_start,
                                     ptrdiff_t* local_no, ptrdiff_t* local_o_start) {
  (void)comm;
  if (flags & (PFFT_TRANSPOSED_IN | PFFT_TRANSPOSED_OUT)) {
    fprintf(stderr, "oracle/_ref: transposed PFFT layouts are not provided\n");
    abort();
  }
  for (int i = 0; i < 3; i++) {
    local_ni[i] = n[i];
    local_no[i] = n[i];
    local_i_start[i] = 0;
    local_o_start[i] = 0;
  }
  local_no[2] = n[2] / 2 + 1;
  return n[0] * n[1] * (n[2] / 2 + 1); /* complex elements; the caller allocates 2x doubles */
}

static pfft_plan make_plan(const ptrdiff_t* n, double* real, pfft_complex* cplx, int sign, unsigned flags) {
  if ((flags & (PFFT_TRANSPOSED_IN | PFFT_TRANSPOSED_OUT)) || !is_pow2((int)n[0]) || !is_pow2((int)n[1]) || !is_pow2((int)n[2])) {
    fprintf(stderr, "oracle/_ref: only non-transposed power-of-two transforms are provided\n");
    abort();
  }
  pfft_plan p = malloc(sizeof(*p));
  for (int i = 0; i < 3; i++) {
    p->n[i] = (int)n[i];
    tw_table((int)n[i], -1); /* built here, single-threaded; only read inside the parallel loops */
    tw_table((int)n[i], +1);
  }
  p->sign = sign;
  p->real = real;
  p->cplx = (cpx*)cplx;
  return p;
}
pfft_plan pfft_plan_dft_r2c_3d(const ptrdiff_t* n, double* in, pfft_complex* out, MPI_Comm comm, int sign, unsigned flags) {
  (void)comm;
  return make_plan(n, in, out, sign, flags);
}
pfft_plan pfft_plan_dft_c2r_3d(const ptrdiff_t* n, pfft_complex* in, double* out, MPI_Comm comm, int sign, unsigned flags) {
  (void)comm;
  return make_plan(n, out, in, sign, flags);
}
void pfft_execute(const pfft_plan p) {
  if (p->sign < 0) exec_r2c(p); else exec_c2r(p);
}
void pfft_destroy_plan(pfft_plan p) { free(p); }
double* pfft_alloc_real(size_t n) { return aligned_alloc(64, ((n * sizeof(double) + 63) / 64) * 64); }
pfft_complex* pfft_alloc_complex(size_t n) { return aligned_alloc(64, ((n * sizeof(pfft_complex) + 63) / 64) * 64); }
void pfft_free(void* p) { free(p); }

/* test hooks: the two transforms on caller arrays (checked against numpy in tests/) */
int ref_fft_r2c(int N, const double* in, double* out) {
  struct pfft_plan_s p = {{N, N, N}, -1, (double*)in, (cpx*)out};
  tw_table(N, -1);
  exec_r2c(&p);
  return 0;
}
int ref_fft_c2r(int N, double* in_destroyed, double* out) {
  struct pfft_plan_s p = {{N, N, N}, +1, out, (cpx*)in_destroyed};
  tw_table(N, +1);
  exec_c2r(&p);
  return 0;
}
