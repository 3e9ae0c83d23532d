/* bvh_walk_sim.c — CPU emulation of the warp-cooperative BVH walk (csrc/nbx_bvh.cu) to SIZE design variants before
 * spending GPU time: counts warp steps, active lanes and accepts of (A) the one-node-per-step key walk and (B) the
 * sibling-pair walk (one record = both children of a node; a lane that accepts the left child tests the right one in the
 * same step; lanes already waiting at the right child are served too). Development tool only: links the test oracle
 * (oracle/libnbody_oracle.so) for the galaxy, keys, sort and the bit-exact tree.
 *   gcc -O2 -ffp-contract=off -fopenmp tools/bvh_walk_sim.c -o tools/bvh_walk_sim -Loracle -lnbody_oracle -Wl,-rpath,$PWD/oracle -lm
 *   tools/bvh_walk_sim <n> [stride of sampled warps] */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifndef LANES
#define LANES 32
#endif

uint32_t nbo_galaxy_f3(uint32_t, float*, float*, float*);
void nbo_bbox_f3(uint32_t, const float*, float*, float*);
void nbo_keys_f3(uint32_t, const float*, const float*, const float*, uint64_t*);
void nbo_sort_perm(uint32_t, const uint64_t*, uint32_t*);
void nbo_permute_f3(uint32_t, const uint32_t*, float*, float*, float*, float*, float*);
uint32_t nbo_bvh_levels_f3(uint32_t);
void nbo_bvh_build_f3(uint32_t, const float*, const float*, float*, float*, float*);

static inline int accept(const float* xs, const float* nm, float w2, float th2) {
  float dx = xs[0] - nm[0], dy = xs[1] - nm[1], dz = xs[2] - nm[2];
  float d2 = dx * dx + dy * dy;
  d2       = d2 + dz * dz;
  return w2 < th2 * d2;
}

static int cmp_u32(const void* a, const void* b) {
  uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
  return x < y ? -1 : x > y;
}

int main(int argc, char** argv) {
  uint32_t n      = argc > 1 ? (uint32_t)atol(argv[1]) : 1000000;
  uint32_t stride = argc > 2 ? (uint32_t)atol(argv[2]) : 16;
  float theta     = argc > 3 ? (float)atof(argv[3]) : 0.5f;
  float *m = calloc(n, 4), *x = calloc((size_t)n * 3, 4), *v = calloc((size_t)n * 3, 4), *a = calloc((size_t)n * 3, 4),
        *ao = calloc((size_t)n * 3, 4);
  nbo_galaxy_f3(n, m, x, v);
  float lo[3], hi[3];
  nbo_bbox_f3(n, x, lo, hi);
  uint64_t* keys = malloc(8 * (size_t)n);
  uint32_t* perm = malloc(4 * (size_t)n);
  nbo_keys_f3(n, x, lo, hi, keys);
  nbo_sort_perm(n, keys, perm);
  nbo_permute_f3(n, perm, m, x, v, a, ao);
  uint32_t levels = nbo_bvh_levels_f3(n);
  uint64_t nn     = ((uint64_t)1 << levels) - 1;
  float *node_m = calloc(nn * 4, 4), *bw = calloc(nn, 4), *b = calloc(nn * 6, 4);
  nbo_bvh_build_f3(n, m, x, node_m, bw, b);
  const float th2 = theta * theta;
  /* 1-based heap index kk = k + 1 */
#define NM(kk) (node_m + ((size_t)(kk)-1) * 4)
#define W2(kk) (bw[(kk)-1] * bw[(kk)-1])
  const uint32_t nlim = ((n + (n & 1u)) << 4) - 4u, sent = 16u << levels, step0 = 16u << levels;
  uint64_t A_steps = 0, A_act = 0, A_anytake = 0, A_body = 0, A_left = 0, A_left_alltake = 0;
  uint64_t B_steps = 0, B_tests = 0, B_rpart = 0, B_body = 0, B_any_take = 0, B_testsL = 0, B_testsR = 0;
  uint64_t warps = 0;
  const int64_t nwarps_all = (int64_t)((n + LANES - 1) / LANES);
  uint32_t* per_warp = calloc((size_t)nwarps_all, 4); /* A's steps of every sampled warp: the walk ends with the heaviest one */
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : A_steps, A_act, A_anytake, A_body, A_left, A_left_alltake, B_steps, B_tests, B_rpart, B_body, B_any_take, B_testsL, B_testsR, warps)
  for (int64_t w = 0; w < (int64_t)((n + LANES - 1) / LANES); w += stride) {
    ++warps;
    uint32_t key[LANES];
    const uint32_t i0 = (uint32_t)w * LANES;
    /* ---- A: one node per step ---- */
    for (int l = 0; l < LANES; ++l) key[l] = i0 + l < n ? 0u : 0xffffffffu;
    for (;;) {
      uint32_t kmin = 0xffffffffu;
      for (int l = 0; l < LANES; ++l) kmin = key[l] < kmin ? key[l] : kmin;
      if (kmin >= nlim) break;
      ++A_steps;
      ++per_warp[w];
      const uint32_t cl = kmin & 31u;
      if (cl == levels) {
        ++A_body;
        for (int l = 0; l < LANES; ++l)
          if (key[l] == kmin) { key[l] = kmin + (cl ? 31u : 32u); ++A_act; }
        continue;
      }
      const uint32_t kk = (kmin | sent) >> (levels + 4u - cl);
      const uint32_t cand_take = kmin + (step0 >> cl) - (kk & 1u), cand_open = kmin + 1u;
      const float* nm = NM(kk);
      const float w2  = W2(kk);
      int any = 0, all = 1, nact = 0;
      for (int l = 0; l < LANES; ++l) {
        if (key[l] != kmin) continue;
        ++nact;
        const int t = accept(x + (size_t)(i0 + l) * 3, nm, w2, th2);
        any |= t; all &= t;
        key[l] = t ? cand_take : cand_open;
      }
      A_act += nact;
      A_anytake += any;
      if (!(kk & 1u)) { ++A_left; A_left_alltake += all; }
    }
    /* ---- B: sibling pair per step ---- */
    for (int l = 0; l < LANES; ++l) key[l] = i0 + l < n ? 0u : 0xffffffffu;
    for (;;) {
      uint32_t kmin = 0xffffffffu;
      for (int l = 0; l < LANES; ++l) kmin = key[l] < kmin ? key[l] : kmin;
      if (kmin >= nlim) break;
      ++B_steps;
      const uint32_t cl = kmin & 31u;
      if (cl == levels) {
        ++B_body;
        for (int l = 0; l < LANES; ++l)
          if (key[l] == kmin) key[l] = kmin + (cl ? 31u : 32u);
        continue;
      }
      const uint32_t kk   = (kmin | sent) >> (levels + 4u - cl);
      const uint32_t size = step0 >> cl;
      const uint32_t keyL = (kk & 1u) ? 0xfffffff0u /* never */ : kmin, keyR = (kk & 1u) ? kmin : kmin + size;
      const uint32_t kL = kk & ~1u, kR = kk | 1u;
      int rpart = 0, any = 0;
      for (int l = 0; l < LANES; ++l) {
        const float* xs = x + (size_t)(i0 + l) * 3;
        int atR = key[l] == keyR;
        if (key[l] == keyL) {
          ++B_tests; ++B_testsL;
          if (accept(xs, NM(kL), W2(kL), th2)) { atR = 1; any = 1; }
          else key[l] = keyL + 1u;
        }
        if (atR) {
          ++B_tests; ++B_testsR;
          rpart = 1;
          if (accept(xs, NM(kR), W2(kR), th2)) { key[l] = keyR + size - 1u; any = 1; }
          else key[l] = keyR + 1u;
        }
      }
      B_rpart += rpart;
      B_any_take += any;
    }
  }
  printf("n=%u levels=%u theta=%.2f sampled warps=%lu\n", n, levels, theta, (unsigned long)warps);
  printf("A: steps/warp=%.0f  tests/body=%.0f  lane util=%.3f  body-level steps=%.3f  steps with any take=%.3f  left-node steps=%.3f of which all active lanes take=%.3f\n",
         (double)A_steps / warps, (double)A_act / (warps * (double)LANES), (double)A_act / ((double)LANES * A_steps), (double)A_body / A_steps,
         (double)A_anytake / A_steps, (double)A_left / A_steps, (double)A_left_alltake / (A_left ? A_left : 1));
  {
    uint32_t* q = malloc((size_t)warps * 4);
    size_t c = 0;
    for (int64_t w = 0; w < nwarps_all; w += stride) q[c++] = per_warp[w];
    qsort(q, c, 4, cmp_u32);
    printf("A: steps per warp: median %u  p90 %u  p99 %u  max %u  (max / mean = %.2f)\n", q[c / 2], q[c * 9 / 10], q[c * 99 / 100], q[c - 1],
           q[c - 1] / ((double)A_steps / warps));
  }
  printf("B: steps/warp=%.0f (%.3f of A)  tests/body=%.0f (L %.0f R %.0f)  steps needing the R half=%.3f  body-level=%.3f any take=%.3f\n",
         (double)B_steps / warps, (double)B_steps / A_steps, (double)B_tests / (warps * (double)LANES), (double)B_testsL / (warps * (double)LANES),
         (double)B_testsR / (warps * (double)LANES), (double)B_rpart / B_steps, (double)B_body / B_steps, (double)B_any_take / B_steps);
  return 0;
}
