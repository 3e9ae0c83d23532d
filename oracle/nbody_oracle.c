/* nbody_oracle.c — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the hot path of UoB-HPC/stdpar-nbody (force + leapfrog for all-pairs,
 * all-pairs-collapsed, octree and bvh, plus the galaxy initial conditions), each function citing the
 * reference file:line it follows (paths relative to /root/reference/). It is the checker for the CUDA
 * product in stdpar-nbody_b200/: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it. The product never links or calls it and has no CPU fallback.
 *
 * Parity status: PINNED. The reference ships no tests or golden vectors, so this oracle is
 * pinned against outputs of the reference itself: oracle/_ref/refdump_d{2,3} (the unmodified reference
 * headers compiled with the same -O2 -ffp-contract=off flags, oracle/Makefile) — bit-for-bit in
 * tests/test_oracle_vs_ref.py where /root/reference or the prebuilt _ref binaries exist — and against
 * the committed fixtures under tests/golden/ that were generated from those binaries
 * (tests/golden/make_golden.py).
 *
 * Exported symbols are suffixed _f2,_f3,_d2,_d3 (REAL = float|double, D = 2|3).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef NBO_FAST
  #define NBO_FAST 0
#endif

/* ---- std::mt19937{42} + std::uniform_real_distribution<double> as libstdc++ implements them (system.h:22-25).
 * generate_canonical<double,53> with a 32-bit engine draws k=2 values: (g1 + g2*2^32) / 2^64, clamped below 1. */
typedef struct { uint32_t mt[624]; int idx; } nbo_rng;
static void nbo_rng_seed(nbo_rng* g, uint32_t s) {
  g->mt[0] = s;
  for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}
static uint32_t nbo_rng_next(nbo_rng* g) {
  if (g->idx >= 624) {
    for (int i = 0; i < 624; ++i) {
      uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
      g->mt[i]   = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}
static double nbo_canonical(nbo_rng* g) {
  double sum = 0, tmp = 1;
  for (int k = 0; k < 2; ++k) {
    sum += (double)nbo_rng_next(g) * tmp;
    tmp *= 4294967296.0;
  }
  double r = sum / tmp;
  if (r >= 1.0) r = nextafter(1.0, 0.0);
  return r;
}
static double nbo_unit(nbo_rng* g) { return (1.0 - 0.0) * nbo_canonical(g) + 0.0; }
static double nbo_angle(nbo_rng* g) { return (2 * 3.141592653589793238462643383279502884 - 0.0) * nbo_canonical(g) + 0.0; }
static double nbo_sym(nbo_rng* g) { return (1.0 - -1.0) * nbo_canonical(g) + -1.0; }

/* ---- vec.h:266-293 interleave_bits */
static uint64_t nbo_split2(uint64_t x) {
  x = (x | x << 16) & 0xffff0000ffffull;
  x = (x | x << 8) & 0xff00ff00ff00ffull;
  x = (x | x << 4) & 0xf0f0f0f0f0f0f0full;
  x = (x | x << 2) & 0x3333333333333333ull;
  x = (x | x << 1) & 0x5555555555555555ull;
  return x;
}
static uint64_t nbo_split3(uint64_t x) {
  x &= 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}

/* ---- vec.h:299-356 hilbert<N>: Skilling's transform over n=2 axes with `bits` bits, then Gray encode, then
 * interleave. In 3-D the reference also uses n=2 (vec.h:328): axis 2 is interleaved untransformed. */
static void nbo_skilling2(uint32_t* x, int bits) {
  const uint32_t M = 1u << (bits - 1);
  for (uint32_t Q = M; Q > 1; Q >>= 1) {
    uint32_t P = Q - 1;
    for (int i = 0; i < 2; ++i) {
      if (x[i] & Q) x[0] ^= P;
      else {
        uint32_t t = (x[0] ^ x[i]) & P;
        x[0] ^= t;
        x[i] ^= t;
      }
    }
  }
  x[1] ^= x[0];
  uint32_t t = 0;
  for (uint32_t Q = M; Q > 1; Q >>= 1)
    if (x[1] & Q) t ^= Q - 1;
  x[0] ^= t;
  x[1] ^= t;
}
uint64_t nbo_hilbert2(uint32_t c0, uint32_t c1) {
  uint32_t x[2] = {c0, c1};
  nbo_skilling2(x, 32);
  return nbo_split2(x[1]) | (nbo_split2(x[0]) << 1);
}
uint64_t nbo_hilbert3(uint32_t c0, uint32_t c1, uint32_t c2) {
  uint32_t x[2] = {c0, c1};
  nbo_skilling2(x, 21);
  return nbo_split3(c2) | (nbo_split3(x[1]) << 1) | (nbo_split3(x[0]) << 2);
}

/* ---- bvh.h:62-69: sort (key, index) by key. The reference's std::sort is unstable (ties undefined, SURVEY §9
 * Q5); this oracle is STABLE (ties keep index order) — LSD radix over the 8 key bytes. perm[i] = original index. */
void nbo_sort_perm(uint32_t n, const uint64_t* keys, uint32_t* perm) {
  uint32_t* a = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
  uint32_t* b = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
  for (uint32_t i = 0; i < n; ++i) a[i] = i;
  for (int pass = 0; pass < 8; ++pass) {
    size_t cnt[257] = {0};
    for (uint32_t i = 0; i < n; ++i) cnt[((keys[a[i]] >> (8 * pass)) & 0xff) + 1]++;
    for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
    for (uint32_t i = 0; i < n; ++i) b[cnt[(keys[a[i]] >> (8 * pass)) & 0xff]++] = a[i];
    uint32_t* t = a; a = b; b = t;
  }
  memcpy(perm, a, sizeof(uint32_t) * (size_t)n);
  free(a);
  free(b);
}

#define REAL float
#define REAL_IS_FLOAT 1
#define D 2
#define SUF f2
#include "nbody_oracle_impl.inc"
#undef D
#undef SUF
#define D 3
#define SUF f3
#include "nbody_oracle_impl.inc"
#undef D
#undef SUF
#undef REAL
#undef REAL_IS_FLOAT

#define REAL double
#define REAL_IS_FLOAT 0
#define D 2
#define SUF d2
#include "nbody_oracle_impl.inc"
#undef D
#undef SUF
#define D 3
#define SUF d3
#include "nbody_oracle_impl.inc"
#undef D
#undef SUF
#undef REAL
#undef REAL_IS_FLOAT
