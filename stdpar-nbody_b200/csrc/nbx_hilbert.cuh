// nbx_hilbert.cuh — a TRUE D-dimensional Hilbert index, used only to assign the bodies of a tree walk to warp lanes
// (locality of the 32 bodies a warp walks for together; never part of a parity artefact — the reference's own key, with
// its two-axis transform in 3-D, is restated bit-for-bit in nbx_bvh.cu).
#pragma once
#include <stdint.h>

namespace nbx {

// Skilling, "Programming the Hilbert curve" (AIP Conf. Proc. 707, 2004), AxestoTranspose over D axes of HB bits, then the
// transpose is interleaved (axis 0 first) into one HB*D-bit index. Consecutive indices are face-adjacent cells.
template <int D>
__device__ __forceinline__ uint64_t hilbert_index(uint32_t (&X)[D], int HB) {
  const uint32_t M = 1u << (HB - 1);
  for (uint32_t Q = M; Q > 1; Q >>= 1) {
    const uint32_t P = Q - 1;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      if (X[k] & Q) X[0] ^= P;
      else {
        const uint32_t t = (X[0] ^ X[k]) & P;
        X[0] ^= t;
        X[k] ^= t;
      }
    }
  }
#pragma unroll
  for (int k = 1; k < D; ++k) X[k] ^= X[k - 1];
  // t_j = parity of the bits of X[D-1] above j (Skilling's loop `if (X[D-1] & Q) t ^= Q - 1` in closed form)
  uint32_t t = X[D - 1] & ((M << 1) - 1u);
  t ^= t >> 1;
  t ^= t >> 2;
  t ^= t >> 4;
  t ^= t >> 8;
  t ^= t >> 16;
  t >>= 1;
  uint64_t h = 0;
  for (int j = HB - 1; j >= 0; --j)
#pragma unroll
    for (int k = 0; k < D; ++k) h = (h << 1) | (((X[k] ^ t) >> j) & 1u);
  return h;
}

}  // namespace nbx
