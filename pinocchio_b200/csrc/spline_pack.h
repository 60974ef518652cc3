// Host-side packing of a natural cubic spline for the device (shared by engine.cu and the CPU
// emulator so that both evaluate exactly the same table).
//
// Coefficients follow gsl_interp_cspline (GSL 2.7 interpolation/cspline.c: natural boundary,
// tridiagonal system, b and d derived per interval).  On top of them a
// uniform look-up table over [x0, xlast] gives the interval index in O(1): the reference's
// gsl_interp_accel/bsearch (src/cosmo.c:2016-2027) costs ~10 dependent shared-memory loads per
// cell on the GPU and was the top stall of the collapse kernel.
//
// Layout (doubles): [0] x0  [1] xlast  [2] 1/h  [3] slope_lo  [4] slope_hi  [5] y0  [6] ylast  [7] n
//                   [8 .. 8+5n)      knots {x, y, b, c, d}
//                   [8+5n .. )       PINB_SPLINE_NLUT uint16 interval indices (bin j starts at x0 + j*h)
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#define PINB_SPLINE_NLUT 1024
#define PINB_SPLINE_HDR 8

namespace pinb {

// rounded up to an even count: the table is staged with 16-byte asynchronous copies
inline size_t spline_table_doubles(int n) { return ((size_t)PINB_SPLINE_HDR + 5 * (size_t)n + PINB_SPLINE_NLUT / 4 + 1) & ~(size_t)1; }

inline void pack_spline(const double* x, const double* y, int n, std::vector<double>& t) {
  t.assign(spline_table_doubles(n), 0.0);
  std::vector<double> c(n, 0.0);
  const int m = n - 2;
  if (m > 0) {
    std::vector<double> diag(m), off(m), rhs(m), cp(m, 0.0), dp(m, 0.0);
    for (int i = 0; i < m; i++) {
      const double h_i = x[i + 1] - x[i], h_ip1 = x[i + 2] - x[i + 1];
      const double ydiff_i = y[i + 1] - y[i], ydiff_ip1 = y[i + 2] - y[i + 1];
      off[i] = h_ip1;
      diag[i] = 2.0 * (h_ip1 + h_i);
      rhs[i] = 3.0 * (ydiff_ip1 / h_ip1 - ydiff_i / h_i);
    }
    cp[0] = m > 1 ? off[0] / diag[0] : 0.0;
    dp[0] = rhs[0] / diag[0];
    for (int i = 1; i < m; i++) {
      const double den = diag[i] - off[i - 1] * cp[i - 1];
      if (i < m - 1) cp[i] = off[i] / den;
      dp[i] = (rhs[i] - off[i - 1] * dp[i - 1]) / den;
    }
    c[m] = dp[m - 1];
    for (int i = m - 2; i >= 0; i--) c[i + 1] = dp[i] - cp[i] * c[i + 2];
  }
  double* k = t.data() + PINB_SPLINE_HDR;
  for (int i = 0; i < n; i++) {
    k[5 * i + 0] = x[i];
    k[5 * i + 1] = y[i];
    k[5 * i + 3] = c[i];
  }
  for (int i = 0; i < n - 1; i++) {
    const double dx = x[i + 1] - x[i], dy = y[i + 1] - y[i];
    k[5 * i + 2] = dy / dx - dx * (c[i + 1] + 2.0 * c[i]) / 3.0;
    k[5 * i + 4] = (c[i + 1] - c[i]) / (3.0 * dx);
  }
  const double h = (x[n - 1] - x[0]) / PINB_SPLINE_NLUT;
  t[0] = x[0];
  t[1] = x[n - 1];
  t[2] = 1.0 / h;
  t[3] = (y[1] - y[0]) / (x[1] - x[0]);                      // my_spline_eval's secant below the table
  t[4] = (y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2]);      // ... and above it
  t[5] = y[0];
  t[6] = y[n - 1];
  t[7] = (double)n;
  std::vector<uint16_t> lut(PINB_SPLINE_NLUT);
  int i = 0;
  for (int j = 0; j < PINB_SPLINE_NLUT; j++) {
    const double xl = x[0] + j * h;
    while (i < n - 2 && x[i + 1] <= xl) i++;
    lut[j] = (uint16_t)i;  // the evaluator steps forward (and, for rounding at a bin edge, one step back)
  }
  std::memcpy(t.data() + PINB_SPLINE_HDR + 5 * (size_t)n, lut.data(), PINB_SPLINE_NLUT * sizeof(uint16_t));
}

}  // namespace pinb
