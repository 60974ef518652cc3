// FP64 line-FFT building blocks for the 3-pass 3-D transforms (replaces pfft_execute,
// reference src/fmax-pfft.c:191-228).
//
// Everything here is written per-thread with the thread id passed in explicitly and the
// block barrier left to the caller, so the very same code runs (a) inside the sm_100a
// kernels and (b) in a host emulation (tests/host/fft_core_host.cpp) that executes each
// phase for all thread ids in turn.  That is how the index math is checked on a CPU-only
// box before any GPU time is spent.
//
// Algorithm: Stockham autosort, mixed radix {2,4,8,16}, at most three stages per line.
// A thread owns RMAX = max radix complex values in registers during a stage; a stage is
//   load (RMAX values)  ->  [barrier]  ->  twiddle + radix butterflies + scatter -> [barrier]
// Stage 0 loads straight from global memory and the last stage may store straight to global
// memory, so a 3-stage line costs two shared-memory round trips.
#pragma once

#if defined(__CUDACC__)
#define PINB_HD __host__ __device__ __forceinline__
#include <cuda_runtime.h>
#else
#define PINB_HD inline
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
#endif

namespace pinb {

PINB_HD double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
PINB_HD double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
PINB_HD double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
PINB_HD double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
PINB_HD double2 cscale(double2 a, double s) { return make_double2(a.x * s, a.y * s); }
// multiply by +i (DIR=+1) or -i (DIR=-1)
template <int DIR> PINB_HD double2 cmul_i(double2 a) {
  return DIR > 0 ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}

// ---------------------------------------------------------------------------------------
// Radix butterflies: in-place DFT of R values, natural order in and out.
// DIR = -1: forward (exp(-i...)),  DIR = +1: backward (exp(+i...)).
// ---------------------------------------------------------------------------------------
template <int R, int DIR> struct Radix;

template <int DIR> struct Radix<2, DIR> {
  static PINB_HD void run(double2* v) {
    double2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  }
};

template <int DIR> struct Radix<4, DIR> {
  static PINB_HD void run4(double2& a0, double2& a1, double2& a2, double2& a3) {
    double2 t0 = cadd(a0, a2), t1 = csub(a0, a2);
    double2 t2 = cadd(a1, a3), t3 = cmul_i<DIR>(csub(a1, a3));
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = cadd(t1, t3);
    a3 = csub(t1, t3);
  }
  static PINB_HD void run(double2* v) { run4(v[0], v[1], v[2], v[3]); }
};

template <int DIR> struct Radix<8, DIR> {
  static PINB_HD void run(double2* v) {
    const double h = 0.70710678118654752440;
    // even / odd 4-point DFTs
    Radix<4, DIR>::run4(v[0], v[2], v[4], v[6]);
    Radix<4, DIR>::run4(v[1], v[3], v[5], v[7]);
    // twiddles w8^k on the odd outputs: w8 = exp(DIR*i*pi/4)
    double2 o1 = v[3], o2 = v[5], o3 = v[7];
    o1 = DIR > 0 ? make_double2(h * (o1.x - o1.y), h * (o1.x + o1.y))
                 : make_double2(h * (o1.x + o1.y), h * (o1.y - o1.x));
    o2 = cmul_i<DIR>(o2);
    o3 = DIR > 0 ? make_double2(-h * (o3.x + o3.y), h * (o3.x - o3.y))
                 : make_double2(h * (o3.y - o3.x), -h * (o3.x + o3.y));
    double2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
    v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
  }
};

template <int DIR> struct Radix<16, DIR> {
  static PINB_HD void run(double2* v) {
    const double c1 = 0.92387953251128675613;  // cos(pi/8)
    const double s1 = 0.38268343236508977173;  // sin(pi/8)
    const double h = 0.70710678118654752440;
    // step 1: for each n2, 4-point DFT over n1 of v[4*n1 + n2]  ->  b[n2][k1] left in v[4*k1+n2]
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) Radix<4, DIR>::run4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
    // step 2: twiddle by w16^(n2*k1), w16 = exp(DIR*i*pi/8); element index 4*k1+n2
    const double sg = DIR > 0 ? 1.0 : -1.0;
    const double2 w1 = make_double2(c1, sg * s1), w2 = make_double2(h, sg * h), w3 = make_double2(s1, sg * c1);
    const double2 w6 = make_double2(-h, sg * h), w9 = make_double2(-c1, -sg * s1);
    v[4 + 1] = cmul(v[4 + 1], w1);
    v[4 + 2] = cmul(v[4 + 2], w2);
    v[4 + 3] = cmul(v[4 + 3], w3);
    v[8 + 1] = cmul(v[8 + 1], w2);
    v[8 + 2] = cmul_i<DIR>(v[8 + 2]);      // w4
    v[8 + 3] = cmul(v[8 + 3], w6);
    v[12 + 1] = cmul(v[12 + 1], w3);
    v[12 + 2] = cmul(v[12 + 2], w6);
    v[12 + 3] = cmul(v[12 + 3], w9);
    // step 3: for each k1, 4-point DFT over n2  ->  X[k1 + 4*k2] left in v[4*k1 + k2]
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) Radix<4, DIR>::run4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
    // step 4: transpose to natural order: X[k1 + 4*k2] must sit at v[k1 + 4*k2]
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = a + 1; b < 4; b++) {
        double2 t = v[4 * a + b];
        v[4 * a + b] = v[4 * b + a];
        v[4 * b + a] = t;
      }
  }
};

// Radix 32 = 4 x 8: n = 8*n1 + n2, k = k1 + 4*k2.  Used by the strided passes of the 512- and
// 1024-point lines: two stages instead of three means two trips through shared memory per
// element instead of four (r02 timeline: the 16x8x8 plan was shared-memory-bandwidth bound).
template <int DIR> struct Radix<32, DIR> {
  static PINB_HD void run(double2* v) {
    constexpr double C[22] = {1.0, 0.98078528040323043, 0.92387953251128674, 0.83146961230254524, 0.70710678118654752,
                              0.55557023301960218, 0.38268343236508977, 0.19509032201612825, 0.0, -0.19509032201612825,
                              -0.38268343236508977, -0.55557023301960218, -0.70710678118654752, -0.83146961230254524,
                              -0.92387953251128674, -0.98078528040323043, -1.0, -0.98078528040323043,
                              -0.92387953251128674, -0.83146961230254524, -0.70710678118654752, -0.55557023301960218};
    constexpr double S[22] = {0.0, 0.19509032201612825, 0.38268343236508977, 0.55557023301960218, 0.70710678118654752,
                              0.83146961230254524, 0.92387953251128674, 0.98078528040323043, 1.0, 0.98078528040323043,
                              0.92387953251128674, 0.83146961230254524, 0.70710678118654752, 0.55557023301960218,
                              0.38268343236508977, 0.19509032201612825, 0.0, -0.19509032201612825,
                              -0.38268343236508977, -0.55557023301960218, -0.70710678118654752, -0.83146961230254524};
    // step 1: for each n2, 4-point DFT over n1 of v[8*n1 + n2] -> b[n2][k1] left in v[8*k1 + n2]
#pragma unroll
    for (int n2 = 0; n2 < 8; n2++) Radix<4, DIR>::run4(v[n2], v[8 + n2], v[16 + n2], v[24 + n2]);
    // step 2: twiddle by w32^(n2*k1)
#pragma unroll
    for (int k1 = 1; k1 < 4; k1++)
#pragma unroll
      for (int n2 = 1; n2 < 8; n2++) {
        const int m = n2 * k1;
        v[8 * k1 + n2] = cmul(v[8 * k1 + n2], make_double2(C[m], DIR > 0 ? S[m] : -S[m]));
      }
    // step 3: for each k1, 8-point DFT over n2 -> X[k1 + 4*k2] left in v[8*k1 + k2]
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) Radix<8, DIR>::run(v + 8 * k1);
    // step 4: natural order
    double2 t[32];
#pragma unroll
    for (int i = 0; i < 32; i++) t[i] = v[i];
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++)
#pragma unroll
      for (int k2 = 0; k2 < 8; k2++) v[k1 + 4 * k2] = t[8 * k1 + k2];
  }
};

// ---------------------------------------------------------------------------------------
// Plans: radices per line length.  ZFIRST8 plans (used by the contiguous z pass) start with
// radix 8 so that the padded shared-memory layout e + (e >> 3) is bank-conflict free.
// ---------------------------------------------------------------------------------------
template <int L, bool ZFIRST8> struct Plan;
#define PINB_PLAN(LEN, Z, NSTAGES, A, B, C)                                     \
  template <> struct Plan<LEN, Z> {                                            \
    static constexpr int NST = NSTAGES, R0 = A, R1 = B, R2 = C;                \
    static constexpr int RMAX = (A > B ? (A > C ? A : C) : (B > C ? B : C));   \
    static constexpr int TPL = LEN / RMAX;                                     \
  };
PINB_PLAN(16, false, 2, 4, 4, 1)
PINB_PLAN(32, false, 2, 8, 4, 1)
PINB_PLAN(64, false, 2, 8, 8, 1)
PINB_PLAN(128, false, 2, 16, 8, 1)
PINB_PLAN(256, false, 2, 16, 16, 1)
PINB_PLAN(512, false, 2, 32, 16, 1)
PINB_PLAN(1024, false, 2, 32, 32, 1)
PINB_PLAN(2048, false, 3, 16, 16, 8)
PINB_PLAN(4096, false, 3, 16, 16, 16)
PINB_PLAN(16, true, 2, 8, 2, 1)
PINB_PLAN(32, true, 2, 8, 4, 1)
PINB_PLAN(64, true, 2, 8, 8, 1)
PINB_PLAN(128, true, 2, 8, 16, 1)
PINB_PLAN(256, true, 3, 8, 8, 4)
PINB_PLAN(512, true, 3, 8, 8, 8)
PINB_PLAN(1024, true, 3, 8, 8, 16)
PINB_PLAN(2048, true, 3, 8, 16, 16)
#undef PINB_PLAN

// The x pass keeps three-stage plans with <= 16 values per thread: its loader (the fused k-space
// Green/window factor) is heavy, and unrolled 32x it overflowed the instruction cache
// (r02: "no_instructions" was the top stall, 43 ms instead of 15 ms).
template <int L> struct XPlan : Plan<L, false> {};
template <> struct XPlan<512> {
  static constexpr int NST = 3, R0 = 8, R1 = 8, R2 = 8, RMAX = 8, TPL = 64;
};
template <> struct XPlan<1024> {
  static constexpr int NST = 3, R0 = 16, R1 = 8, R2 = 8, RMAX = 16, TPL = 64;
};

// Twiddle table: tw[k] = exp(+2*pi*i*k/NROOT), k in [0, NROOT).  A line of length L uses
// every (NROOT/L)-th entry; DIR=-1 conjugates.
template <int DIR> PINB_HD double2 twiddle(const double2* __restrict__ tw, int idx) {
#if defined(__CUDA_ARCH__)
  double2 w = __ldg(tw + idx);
#else
  double2 w = tw[idx];
#endif
  return DIR > 0 ? w : cconj(w);
}

// ---------------------------------------------------------------------------------------
// One Stockham stage, split at the barrier.
//   L    line length            R   radix of this stage        NS  product of earlier radices
//   TPL  threads per line       RMAX registers (complex) per thread;  (L/R)/TPL * R == RMAX
// jl = thread index within the line, 0 <= jl < TPL.
// ---------------------------------------------------------------------------------------
template <int L, int R, int TPL, int RMAX, class In>
PINB_HD void stage_load(int jl, double2* v, In in) {
  constexpr int T = L / R;
  constexpr int NB = T / TPL;
  static_assert(NB * R == RMAX && NB * TPL == T, "bad stage shape");
#pragma unroll
  for (int m = 0; m < NB; m++) {
    const int j = jl + m * TPL;
#pragma unroll
    for (int r = 0; r < R; r++) v[m * R + r] = in(j + r * T);
  }
}

// same, the functor also receives s = (e - jl)/TPL = m + r*NB (a compile-time constant after
// unrolling): element e = jl + s*TPL, so exp(i*phi*e) = exp(i*phi*jl) * exp(i*phi*TPL)^s
template <int L, int R, int TPL, int RMAX, class In>
PINB_HD void stage_load_s(int jl, double2* v, In in) {
  constexpr int T = L / R;
  constexpr int NB = T / TPL;
  static_assert(NB * R == RMAX && NB * TPL == T, "bad stage shape");
#pragma unroll
  for (int m = 0; m < NB; m++) {
    const int j = jl + m * TPL;
#pragma unroll
    for (int r = 0; r < R; r++) v[m * R + r] = in(j + r * T, m + r * NB);
  }
}

// TWPOW: fetch only w^k and build w^(rk) by (shallow) repeated multiplication.  Used by the z
// passes, whose shared-memory footprint leaves no L1 for the twiddle table: every table read
// is an L2 round trip there (r02 profile: the twiddle multiply was the top stall of the FFT part).
template <int L, int R, int NS, int DIR, int TPL, int RMAX, class Out, bool TWPOW = false>
PINB_HD void stage_store(int jl, double2* v, Out out, const double2* __restrict__ tw, int twscale) {
  constexpr int T = L / R;
  constexpr int NB = T / TPL;
#pragma unroll
  for (int m = 0; m < NB; m++) {
    const int j = jl + m * TPL;
    const int k = j & (NS - 1);
    if (NS > 1) {
      const int base = k * (L / (NS * R)) * twscale;
      if (TWPOW) {
        double2 w[R];
        w[1] = twiddle<DIR>(tw, base);
#pragma unroll
        for (int r = 2; r < R; r++) w[r] = (r & 1) ? cmul(w[r - 1], w[1]) : cmul(w[r / 2], w[r / 2]);
#pragma unroll
        for (int r = 1; r < R; r++) v[m * R + r] = cmul(v[m * R + r], w[r]);
      } else {
#pragma unroll
        for (int r = 1; r < R; r++) v[m * R + r] = cmul(v[m * R + r], twiddle<DIR>(tw, r * base));
      }
    }
    Radix<R, DIR>::run(v + m * R);
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; r++) out(j0 + r * NS, v[m * R + r]);
  }
}

// ---------------------------------------------------------------------------------------
// Real <-> half-complex glue for the contiguous (z) pass.  M = N/2 complex points per line.
// c2r:  Z[k] = (X[k] + conj X[M-k]) + i w^k (X[k] - conj X[M-k]),  w = exp(+2 pi i/N);
//       z = IDFT_M(Z);  x[2m] = Re z[m], x[2m+1] = Im z[m].   Only Re X[0], Re X[M] are used
//       (FFTW/PFFT c2r semantics).
// r2c:  Z = DFT_M(z);  X[k] = 1/2 [(Z[k] + conj Z[M-k]) - i w^-k (Z[k] - conj Z[M-k])].
// Each call handles the pair (k, M-k), 0 <= k <= M/2; `tw` has NROOT = N entries.
// ---------------------------------------------------------------------------------------
PINB_HD void c2r_pre_pair(double2 xk, double2 xmk, double2 w, double2& zk, double2& zmk) {
  // k -> uses w = w^k ; (M-k) -> uses w^(M-k) = -conj(w^k)
  const double2 a = xk, b = cconj(xmk);
  const double2 e = cadd(a, b), o = cmul(csub(a, b), w);
  zk = make_double2(e.x - o.y, e.y + o.x);                       // e + i o
  const double2 e2 = cconj(e);                                   // X[M-k] + conj X[k]
  const double2 o2 = cmul(csub(xmk, cconj(xk)), make_double2(-w.x, w.y));
  zmk = make_double2(e2.x - o2.y, e2.y + o2.x);
}

PINB_HD void r2c_post_pair(double2 zk, double2 zmk, double2 w /* = w^k, +sign table */, double2& xk, double2& xmk) {
  // forward transform uses conj(w^k) = exp(-2 pi i k/N)
  const double2 wc = cconj(w);
  const double2 a = zk, b = cconj(zmk);
  const double2 e = cadd(a, b), o = cmul(csub(a, b), wc);
  xk = make_double2(0.5 * (e.x + o.y), 0.5 * (e.y - o.x));       // (e - i o)/2
  // k' = M-k: w^-(M-k) = -conj(w^-k) = -w
  const double2 e2 = cconj(e);
  const double2 o2 = cmul(csub(zmk, cconj(zk)), make_double2(-w.x, -w.y));
  xmk = make_double2(0.5 * (e2.x + o2.y), 0.5 * (e2.y - o2.x));
}

// padded shared-memory index of element e of a contiguous line (16-byte elements)
PINB_HD int zpad(int e) { return e + (e >> 3); }
template <int M> struct ZLine { static constexpr int PITCH = M + (M >> 3) + 2; };

}  // namespace pinb
