/* TEST INFRASTRUCTURE (oracle/_ref): the two GSL 2.7.1 random number generators that the
 * reference's src/GenIC.c draws from, restated from their published algorithms because GSL is
 * not installed here (third-party arithmetic absent from /root/reference):
 *   gsl_rng_ranlxd1  -- Luescher's double-precision RANLUX, luxury level 1 (GSL rng/ranlxd.c:
 *                       12 doubles of state, subtract-with-borrow r = 12, s = 5 walked as
 *                       (ir, jr) = (11, 7), 202 updates between blocks of 12 outputs; seeding by
 *                       the 31-bit shift register on the seed read as a signed int)
 *   gsl_rng_mt19937  -- Matsumoto & Nishimura's MT19937 with the 2002 initialisation
 *                       (GSL rng/mt.c; seed 0 -> 4357)
 * Pinned by the known answers of GSL's own rng/test.c (tests/test_reference_oracle.py):
 * mt19937 seed 4357 -> 1000th output 1186927261; ranlxd1 seed 1 -> 10000th output 1998227290.
 * Nothing in the product links this file. */
#include <stdlib.h>
#include <string.h>

#include <gsl/gsl_rng.h>

struct gsl_rng_type_s { int kind; };
static const gsl_rng_type type_ranlxd1 = {1}, type_mt19937 = {2};
const gsl_rng_type* gsl_rng_ranlxd1 = &type_ranlxd1;
const gsl_rng_type* gsl_rng_mt19937 = &type_mt19937;

struct gsl_rng_s {
  int kind;
  /* ranlxd */
  double xdbl[12], carry;
  unsigned int ir, jr, ir_old, pr;
  /* mt19937 */
  unsigned long mt[624];
  int mti;
};

static const double one_bit = 1.0 / 281474976710656.0; /* 2^-48 */

static void ranlxd_increment_state(gsl_rng* s) {
  int k, kmax;
  double y1;
  unsigned int ir = s->ir, jr = s->jr;
  double carry = s->carry;
  /* bring ir to 0, then further single updates up to pr in total (GSL unrolls the middle part
     in blocks of twelve; the sequence of updates is the same) */
  for (k = 0; ir > 0; ++k) {
    y1 = s->xdbl[jr] - s->xdbl[ir] - carry;
    if (y1 < 0) { carry = one_bit; y1 += 1; } else carry = 0;
    s->xdbl[ir] = y1;
    ir = (ir + 1) % 12;
    jr = (jr + 1) % 12;
  }
  kmax = (int)s->pr;
  for (; k < kmax; ++k) {
    y1 = s->xdbl[jr] - s->xdbl[ir] - carry;
    if (y1 < 0) { carry = one_bit; y1 += 1; } else carry = 0;
    s->xdbl[ir] = y1;
    ir = (ir + 1) % 12;
    jr = (jr + 1) % 12;
  }
  s->ir = ir;
  s->ir_old = ir;
  s->jr = jr;
  s->carry = carry;
}

static double ranlxd_get_double(gsl_rng* s) {
  unsigned int ir = s->ir;
  s->ir = (ir + 1) % 12;
  ir = s->ir;
  if (ir == s->ir_old) ranlxd_increment_state(s);
  return s->xdbl[s->ir];
}

static void ranlxd_set(gsl_rng* s, unsigned long int seed_in) {
  int ibit, jbit, i, k, l, xbit[31];
  double x, y;
  long int seed;
  if (seed_in == 0) seed_in = 1; /* default seed is 1 */
  seed = (long int)(int)seed_in; /* GSL keeps the seed in a signed int: i%2 and i/=2 below see |i| */
  i = (int)(seed & 0x7FFFFFFFUL);
  if (seed < 0) i = (int)((-seed) & 0x7FFFFFFFL);
  for (k = 0; k < 31; ++k) { xbit[k] = i % 2; i /= 2; }
  ibit = 0;
  jbit = 18;
  for (k = 0; k < 12; ++k) {
    x = 0;
    for (l = 1; l <= 48; ++l) {
      y = (double)((xbit[ibit] + 1) % 2);
      x += x + y;
      xbit[ibit] = (xbit[ibit] + xbit[jbit]) % 2;
      ibit = (ibit + 1) % 31;
      jbit = (jbit + 1) % 31;
    }
    s->xdbl[k] = one_bit * x;
  }
  s->carry = 0;
  s->ir = 11;
  s->jr = 7;
  s->ir_old = 0;
  s->pr = 202; /* luxury level 1 */
}

static void mt_set(gsl_rng* s, unsigned long int seed) {
  int i;
  if (seed == 0) seed = 4357;
  s->mt[0] = seed & 0xffffffffUL;
  for (i = 1; i < 624; i++) {
    s->mt[i] = (1812433253UL * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (unsigned long)i);
    s->mt[i] &= 0xffffffffUL;
  }
  s->mti = 624;
}

static unsigned long mt_get(gsl_rng* s) {
  unsigned long k;
  unsigned long* const mt = s->mt;
  if (s->mti >= 624) {
    int kk;
    for (kk = 0; kk < 624; kk++) {
      const unsigned long y = (mt[kk] & 0x80000000UL) | (mt[(kk + 1) % 624] & 0x7fffffffUL);
      mt[kk] = mt[(kk + 397) % 624] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
    }
    s->mti = 0;
  }
  k = mt[s->mti++];
  k ^= (k >> 11);
  k ^= (k << 7) & 0x9d2c5680UL;
  k ^= (k << 15) & 0xefc60000UL;
  k ^= (k >> 18);
  return k & 0xffffffffUL;
}

gsl_rng* gsl_rng_alloc(const gsl_rng_type* T) {
  gsl_rng* r = calloc(1, sizeof(gsl_rng));
  r->kind = T->kind;
  gsl_rng_set(r, 0); /* gsl_rng_alloc seeds with the default seed */
  return r;
}
void gsl_rng_set(const gsl_rng* r, unsigned long int seed) {
  gsl_rng* s = (gsl_rng*)r;
  if (s->kind == 1) ranlxd_set(s, seed); else mt_set(s, seed);
}
double gsl_rng_uniform(const gsl_rng* r) {
  gsl_rng* s = (gsl_rng*)r;
  if (s->kind == 1) return ranlxd_get_double(s);
  return mt_get(s) / 4294967296.0;
}
unsigned long int gsl_rng_get(const gsl_rng* r) {
  gsl_rng* s = (gsl_rng*)r;
  if (s->kind == 1) return (unsigned long int)(ranlxd_get_double(s) * 4294967296.0); /* 2^32 */
  return mt_get(s);
}
void gsl_rng_free(gsl_rng* r) { free(r); }
