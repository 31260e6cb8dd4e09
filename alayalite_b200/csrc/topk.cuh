// Warp-collective sorted top-k list ordered by (distance, id) -- shared by the exact brute-force
// kernels and the merge kernel.
#pragma once
#include "common.cuh"

namespace ab {

__device__ __forceinline__ bool pair_less(float d1, uint32_t i1, float d2, uint32_t i2) {
  return d1 < d2 || (d1 == d2 && i1 < i2);
}

// td/ti: sorted list in shared memory, `size` entries of capacity k; arguments warp-uniform
__device__ __forceinline__ void topk_insert(float *td, uint32_t *ti, int &size, int k, float d, uint32_t id, int lane) {
  constexpr uint32_t kFullMask = 0xFFFFFFFFu;
  if (size == k && !pair_less(d, id, td[k - 1], ti[k - 1])) return;
  int pos = 0;
  for (int base = 0; base < size; base += 32) {
    const int i = base + lane;
    const uint32_t b = __ballot_sync(kFullMask, i < size && pair_less(td[i], ti[i], d, id));
    pos += __popc(b);
    if (b != kFullMask) break;
  }
  const int nsize = min(size + 1, k);
  for (int base = (nsize - 1) & ~31; base >= 0 && base + 31 > pos; base -= 32) {
    const int i = base + lane;
    const bool mv = i > pos && i < nsize;
    float xd = 0.f;
    uint32_t xi = 0;
    if (mv) {
      xd = td[i - 1];
      xi = ti[i - 1];
    }
    __syncwarp();
    if (mv) {
      td[i] = xd;
      ti[i] = xi;
    }
    __syncwarp();
  }
  if (lane == 0) {
    td[pos] = d;
    ti[pos] = id;
  }
  __syncwarp();
  size = nsize;
}

}  // namespace ab
