// Per-pixel part of the equivalence-aware metrics (validate.py:88-139), shared by the stand-alone histogram kernel
// (eval_hist.cu) and the fused top-k + histogram kernel (eval_topk_umma.cu) so that both count exactly the same things.
#pragma once
#include "common.cuh"

namespace rc {

// one atomic per distinct bin per warp (labels are spatially coherent, so usually 1-2 per warp); every lane of the warp
// must call this
template <typename CounterT>
__device__ __forceinline__ void warp_agg_add(CounterT* bins, int bin, bool pred) {
  const unsigned active = __ballot_sync(0xffffffffu, pred);
  if (!pred) return;
  const unsigned peers = __match_any_sync(active, bin);
  const int leader = __ffs(peers) - 1;
  if ((int)(threadIdx.x & 31) == leader) atomicAdd(&bins[bin], (CounterT)__popc(peers));
}

struct PixelMetric {
  bool ok;                  // gt and top-1 are valid class ids
  int ge, p1, orc;          // class of the ground truth, of the top-1 prediction, of the "oracle" prediction
  bool top1_same, orc_same;
  bool any_eq1, any_eqk;    // E[gt, top1], any_j E[gt, topk_j]
};

// `id_of(j)` returns the j-th predicted id (int64) of the pixel, j = 0 .. k-1
template <typename IdOf>
__device__ __forceinline__ PixelMetric pixel_metric(int64_t g, int k, IdOf id_of, const uint8_t* __restrict__ E,
                                                    const int64_t* __restrict__ cmap, int C) {
  PixelMetric m;
  m.ge = 0; m.p1 = 0; m.orc = 0; m.top1_same = false; m.orc_same = false; m.any_eq1 = false; m.any_eqk = false;
  const int64_t t1 = id_of(0);
  m.ok = (uint64_t)g < (uint64_t)C && (uint64_t)t1 < (uint64_t)C;
  if (!m.ok) return m;
  m.ge = (int)cmap[g];
  m.p1 = (int)cmap[t1];
  const uint8_t* Erow = E + (int64_t)g * C;
  bool any_eq = Erow[t1] != 0;
  bool hit = (m.p1 == m.ge);
  m.any_eq1 = any_eq;
  for (int j = 1; j < k; ++j) {
    const int64_t tj = id_of(j);
    if ((uint64_t)tj < (uint64_t)C) {
      any_eq |= Erow[tj] != 0;
      hit |= ((int)cmap[tj] == m.ge);
    }
  }
  m.any_eqk = any_eq;
  m.orc = hit ? m.ge : (int)t1;   // validate.py:122 -- oracle_pred starts from RAW top-1 ids
  m.top1_same = (m.p1 == m.ge);
  m.orc_same = (m.orc == m.ge);
  return m;
}

}  // namespace rc
