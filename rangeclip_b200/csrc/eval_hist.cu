// K8: equivalence-aware accuracy / mIoU histograms.
// Replaces the Python per-label loops of RangeCLIP/src/depth_segmentation_model/validate.py:88-139
// (and the top-1 accuracy lines of evaluation.py:94-104) with one pass over (gt, top-k) that
// builds five per-class histograms with warp-aggregated atomics.  Integer only -> bit-exact.
#include "common.cuh"
#include "eval_metrics.cuh"

namespace rc {

template <bool kSmem>
__global__ void __launch_bounds__(256)
eval_hist_kernel(const int64_t* __restrict__ gt, const int64_t* __restrict__ topk, int B, int64_t HW, int k,
                 const uint8_t* __restrict__ E, const int64_t* __restrict__ cmap, int C,
                 unsigned long long* __restrict__ hist, unsigned long long* __restrict__ counters) {
  extern __shared__ unsigned int sh[];  // [5][C] when kSmem
  if (kSmem) {
    for (int i = threadIdx.x; i < 5 * C; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
  }
  const int64_t n = (int64_t)B * HW;
  unsigned int c1 = 0, ck = 0, tot = 0;
  // grid-stride in whole warps so that every lane of a warp reaches the warp collectives
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n_round = (n + 31) / 32 * 32;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
    bool ok = i < n;
    int ge = 0, p1 = 0, orc = 0;
    bool top1_same = false, orc_same = false;
    if (ok) {
      const int64_t b = i / HW, p = i - b * HW;
      const int64_t* tk = topk + (b * k) * HW + p;
      const PixelMetric m = pixel_metric(gt[i], k, [&](int j) { return tk[(int64_t)j * HW]; }, E, cmap, C);
      ok = m.ok;
      if (ok) {
        ge = m.ge; p1 = m.p1; orc = m.orc; top1_same = m.top1_same; orc_same = m.orc_same;
        c1 += m.any_eq1 ? 1u : 0u;
        ck += m.any_eqk ? 1u : 0u;
        tot += 1u;
      }
    }
    if (kSmem) {
      warp_agg_add(sh + 0 * C, ge, ok);
      warp_agg_add(sh + 1 * C, p1, ok);
      warp_agg_add(sh + 2 * C, ge, ok && top1_same);
      warp_agg_add(sh + 3 * C, orc, ok);
      warp_agg_add(sh + 4 * C, ge, ok && orc_same);
    } else {
      warp_agg_add(hist + 0 * (int64_t)C, ge, ok);
      warp_agg_add(hist + 1 * (int64_t)C, p1, ok);
      warp_agg_add(hist + 2 * (int64_t)C, ge, ok && top1_same);
      warp_agg_add(hist + 3 * (int64_t)C, orc, ok);
      warp_agg_add(hist + 4 * (int64_t)C, ge, ok && orc_same);
    }
  }
  // counters: warp reduce then one atomic per warp
  for (int o = 16; o > 0; o >>= 1) {
    c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    ck += __shfl_xor_sync(0xffffffffu, ck, o);
    tot += __shfl_xor_sync(0xffffffffu, tot, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (c1) atomicAdd(&counters[0], (unsigned long long)c1);
    if (ck) atomicAdd(&counters[1], (unsigned long long)ck);
    if (tot) atomicAdd(&counters[2], (unsigned long long)tot);
  }
  if (kSmem) {
    __syncthreads();
    for (int i = threadIdx.x; i < 5 * C; i += blockDim.x) {
      const unsigned int v = sh[i];
      if (v) atomicAdd(&hist[i], (unsigned long long)v);
    }
  }
}

__global__ void eval_fold_kernel(const long long* __restrict__ h, int C, int batch_index,
                                 long long* __restrict__ acc, int* __restrict__ first_seen) {
  const int L = blockIdx.x * blockDim.x + threadIdx.x;
  if (L >= C) return;
  const long long ge = h[L], p1 = h[C + L], i1 = h[2 * C + L], orc = h[3 * C + L], ik = h[4 * C + L];
  if (ge + p1 > 0) {  // validate.py:108 -- only labels in unique(gt_equiv U pred_equiv_top1)
    acc[L] += i1;
    acc[C + L] += ge + p1 - i1;
    acc[2 * C + L] += ik;
    acc[3 * C + L] += ge + orc - ik;
    if (batch_index < first_seen[L]) first_seen[L] = batch_index;
  }
}

}  // namespace rc

extern "C" int rc_eval_hist(const int64_t* gt, const int64_t* topk, int B, int64_t HW, int k,
                            const uint8_t* E, const int64_t* cmap, int C, int64_t* hist,
                            int64_t* counters, void* stream) {
  RC_REQUIRE(gt && topk && E && cmap && hist && counters, "rc_eval_hist: null pointer");
  RC_REQUIRE(B >= 0 && HW >= 0 && k >= 1 && C >= 1, "rc_eval_hist: bad shape B=%d HW=%lld k=%d C=%d", B, (long long)HW, k, C);
  const int64_t n = (int64_t)B * HW;
  if (n == 0) return RC_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int threads = 256;
  int64_t want = (n + threads - 1) / threads;
  int grid = (int)(want < (int64_t)rc::num_sms() * 4 ? want : (int64_t)rc::num_sms() * 4);
  const size_t smem = (size_t)5 * C * sizeof(unsigned int);
  if (smem <= 160 * 1024) {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(rc::eval_hist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return rc::fail(RC_ERR_CUDA, "rc_eval_hist: smem opt-in: %s", cudaGetErrorString(e));
      if (grid > rc::num_sms()) grid = rc::num_sms();
    }
    rc::eval_hist_kernel<true><<<grid, threads, smem, s>>>(gt, topk, B, HW, k, E, cmap, C,
        (unsigned long long*)hist, (unsigned long long*)counters);
  } else {
    rc::eval_hist_kernel<false><<<grid, threads, 0, s>>>(gt, topk, B, HW, k, E, cmap, C,
        (unsigned long long*)hist, (unsigned long long*)counters);
  }
  return rc::check_launch("rc_eval_hist");
}

extern "C" int rc_eval_fold(const int64_t* batch_hist, int C, int32_t batch_index, int64_t* acc,
                            int32_t* first_seen, void* stream) {
  RC_REQUIRE(batch_hist && acc && first_seen && C >= 1, "rc_eval_fold: bad argument");
  rc::eval_fold_kernel<<<(C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      (const long long*)batch_hist, C, batch_index, (long long*)acc, first_seen);
  return rc::check_launch("rc_eval_fold");
}
