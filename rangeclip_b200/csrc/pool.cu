// K3/K4: segment-masked average pooling of pixel embeddings.
// Replaces the per-object Python loops of dataloader.py:286-304 (per image, per label) and
// model.py:36-54 (batch-wide per label), each of which materialises a full [D,H,W] `where`
// temporary per object, with ONE read of X: every warp owns 256 consecutive pixels of one image
// for a group of channels, derives the run structure of the slot ids once, then per channel does
// a 16-byte coalesced load per lane, a lane-local sum and a warp-shuffle segmented reduction;
// only run heads issue a (no-return) atomic add.
#include "common.cuh"
#include <type_traits>

namespace rc {

constexpr int kPoolThreads = 256;
constexpr int kPoolDGroup = 64;   // channels per warp task
constexpr int kPoolPx = 256;      // pixels per warp task (8 per lane)

__device__ __forceinline__ int pool_slot(const int64_t* __restrict__ seg, const int32_t* __restrict__ lut,
                                         int64_t lut_off, int C, int64_t p, int64_t HW) {
  if (p >= HW) return -1;
  const int64_t lab = seg[p];
  if ((uint64_t)lab >= (uint64_t)C) return -1;
  return lut[lut_off + lab];
}

// Run structure of one warp's 256 pixels: key per lane (-1 = no slot or mixed lane), and for the
// 5 shuffle steps whether lane+o continues the same run.
struct RunInfo {
  int key;          // slot id shared by the lane's 8 pixels, or -1
  bool mixed;       // the lane's 8 pixels carry different slots (slow path)
  unsigned steps;   // bit i: lane + (1<<i) is in the same run
  bool head;        // first lane of its run
  int run_px;       // pixels in the run (valid on head lanes)
};

__device__ __forceinline__ RunInfo pool_runs(const int (&slot)[8]) {
  RunInfo r;
  bool uni = true;
#pragma unroll
  for (int j = 1; j < 8; ++j) uni &= (slot[j] == slot[0]);
  r.mixed = !uni;
  r.key = uni ? slot[0] : -1;
  const int lane = threadIdx.x & 31;
  const int key_prev = __shfl_up_sync(0xffffffffu, r.key, 1);
  const bool starts = (lane == 0) || (key_prev != r.key);
  // run ids are monotone along the warp, so equal ids at distance o imply a contiguous run
  const unsigned heads = __ballot_sync(0xffffffffu, starts);
  const int run_id = __popc(heads & (0xffffffffu >> (31 - lane)));
  r.head = starts && (r.key >= 0);
  r.steps = 0;
  int len = (r.key >= 0) ? 8 : 0;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const int o = 1 << i;
    const int id_up = __shfl_down_sync(0xffffffffu, run_id, o);
    const int len_up = __shfl_down_sync(0xffffffffu, len, o);
    if (lane + o < 32 && id_up == run_id) { r.steps |= (1u << i); len += len_up; }
  }
  r.run_px = len;
  return r;
}

// A warp task is TWO consecutive 256-pixel chunks of one image x 64 channels.  When both chunks carry the same slot
// in every pixel position (consecutive image rows inside object masks: the common case) they are summed lane-locally
// first and share ONE segmented reduction and ONE set of atomics per channel -- half the shuffles and half the
// atomics per byte read, which is what bounded the bf16 kernel (8 atomics per 512 bytes).  Otherwise the two chunks
// are processed one after the other.
template <typename T>
__global__ void __launch_bounds__(kPoolThreads, 4)
pool_fwd_kernel(const T* __restrict__ x, int B, int D, int64_t HW, const int64_t* __restrict__ seg,
                const int32_t* __restrict__ lut, int64_t lut_ld, int C, float* __restrict__ sum,
                int* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = kPoolThreads / 32;
  const int64_t chunks = (HW + kPoolPx - 1) / kPoolPx;
  const int64_t cpairs = (chunks + 1) / 2;
  const int dgroups = (D + kPoolDGroup - 1) / kPoolDGroup;
  const int64_t n_tasks = (int64_t)B * cpairs * dgroups;
  constexpr int kVec = sizeof(T) == 2 ? 1 : 2;          // 16-byte vectors per lane per channel (8 pixels)
  constexpr int kLoads = sizeof(T) == 2 ? 8 : 4;        // (channel, chunk) rows whose loads are issued before any is consumed:
                                                         // 4 KB in flight per warp for either dtype
  for (int64_t task = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); task < n_tasks;
       task += (int64_t)gridDim.x * warps_per_block) {
    const int dg = (int)(task % dgroups);
    const int64_t rest = task / dgroups;
    const int64_t cp = rest % cpairs;
    const int b = (int)(rest / cpairs);
    const int64_t pA = (2 * cp) * kPoolPx + lane * 8;
    const int64_t pB = pA + kPoolPx;
    const bool hasB = (2 * cp + 1) < chunks;
    const int64_t* segb = seg + (int64_t)b * HW;
    const int d0 = dg * kPoolDGroup;
    const int d1 = min(D, d0 + kPoolDGroup);
    // the lane's 8 pixels of one channel row as floats (raw 16-byte vectors stay packed until here)
    auto unpack8 = [&](const uint4 (&raw)[kVec], float (&v)[8]) {
      if (sizeof(T) == 2) {
        const uint32_t u[4] = {raw[0].x, raw[0].y, raw[0].z, raw[0].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(u[i] << 16); v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u); }
      } else {
        const uint4 a = raw[0], c = raw[kVec - 1];
        v[0] = __uint_as_float(a.x); v[1] = __uint_as_float(a.y); v[2] = __uint_as_float(a.z); v[3] = __uint_as_float(a.w);
        v[4] = __uint_as_float(c.x); v[5] = __uint_as_float(c.y); v[6] = __uint_as_float(c.z); v[7] = __uint_as_float(c.w);
      }
    };
    // their sum: bf16 words are added with two FHADD.BF16 chains, no unpack
    auto lane_sum = [&](const uint4 (&raw)[kVec]) -> float {
      if (sizeof(T) == 2) {
        const uint32_t u[4] = {raw[0].x, raw[0].y, raw[0].z, raw[0].w};
        float sa = addacc_bf16x2_lo(0.f, u[0]), sb = addacc_bf16x2_hi(0.f, u[0]);
#pragma unroll
        for (int i = 1; i < 4; ++i) { sa = addacc_bf16x2_lo(sa, u[i]); sb = addacc_bf16x2_hi(sb, u[i]); }
        return sa + sb;
      }
      float v[8];
      unpack8(raw, v);
      return ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    };
    // one pass over the channel group for `rows` (1 or 2) chunks starting at pixel p0 that share `slot` / `ri`
    auto process = [&](const int (&slot)[8], const RunInfo& ri, int64_t p0, auto rows_c) {
      constexpr int rows = decltype(rows_c)::value;
      if (dg == 0) {
        if (ri.head) atomicAdd(&count[ri.key], ri.run_px * rows);
        if (ri.mixed) {
#pragma unroll
          for (int j = 0; j < 8; ++j) if (slot[j] >= 0) atomicAdd(&count[slot[j]], rows);
        }
      }
      const bool in_range = p0 < HW;   // HW % 8 == 0 on this path, so the lane's 8 px are all in or out
      const T* src = x + ((int64_t)b * D + d0) * HW + p0;
      constexpr int cpb = kLoads / rows;   // channels per batch of loads
      for (int dbase = d0; dbase < d1; dbase += cpb, src += (int64_t)cpb * HW) {
        uint4 raw[kLoads][kVec];
#pragma unroll
        for (int q = 0; q < kLoads; ++q) {
          const int dq = rows == 2 ? (q >> 1) : q;                  // channel of load slot q
          const int64_t off = (int64_t)dq * HW + ((rows == 2 && (q & 1)) ? kPoolPx : 0);
#pragma unroll
          for (int h = 0; h < kVec; ++h)
            raw[q][h] = (in_range && dbase + dq < d1 && p0 + (off - (int64_t)dq * HW) < HW)      // second chunk may end early
                            ? __ldg(reinterpret_cast<const uint4*>(src + off) + h) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int q = 0; q < kLoads; ++q) {
          if (rows == 2 && (q & 1)) continue;                       // consumed together with its even partner
          const int d = dbase + (rows == 2 ? (q >> 1) : q);
          if (d >= d1) break;
          float sl = 0.f;
          if (!ri.mixed) {
            sl = lane_sum(raw[q]);
            if (rows == 2) sl += lane_sum(raw[q | 1]);
          } else {             // lane with several slots among its 8 pixels (mask borders): per-pixel atomics
#pragma unroll
            for (int rr = 0; rr < rows; ++rr) {
              float v[8];
              unpack8(raw[q | rr], v);
#pragma unroll
              for (int j = 0; j < 8; ++j) if (slot[j] >= 0) atomicAdd(&sum[(int64_t)slot[j] * D + d], v[j]);
            }
          }
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const float up = __shfl_down_sync(0xffffffffu, sl, 1 << i);
            if (ri.steps & (1u << i)) sl += up;
          }
          if (ri.head) atomicAdd(&sum[(int64_t)ri.key * D + d], sl);
        }
      }
    };
    int slot[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) slot[j] = pool_slot(segb, lut, (int64_t)b * lut_ld, C, pA + j, HW);
    bool same = hasB;
    if (hasB) {
#pragma unroll
      for (int j = 0; j < 8; ++j) same &= (pool_slot(segb, lut, (int64_t)b * lut_ld, C, pB + j, HW) == slot[j]);
    }
    const bool merged = __all_sync(0xffffffffu, same) != 0;
    {
      const RunInfo ri = pool_runs(slot);
      if (merged) process(slot, ri, pA, std::integral_constant<int, 2>{});
      else process(slot, ri, pA, std::integral_constant<int, 1>{});
    }
    if (hasB && !merged) {
#pragma unroll
      for (int j = 0; j < 8; ++j) slot[j] = pool_slot(segb, lut, (int64_t)b * lut_ld, C, pB + j, HW);
      const RunInfo ri = pool_runs(slot);
      process(slot, ri, pB, std::integral_constant<int, 1>{});
    }
  }
}

// generic fallback (HW % 8 != 0 or unaligned base): one thread per (b, d, p) element strip
template <typename T>
__global__ void pool_fwd_generic_kernel(const T* __restrict__ x, int B, int D, int64_t HW,
                                        const int64_t* __restrict__ seg, const int32_t* __restrict__ lut,
                                        int64_t lut_ld, int C, float* __restrict__ sum, int* __restrict__ count) {
  const int64_t n = (int64_t)B * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / HW, p = i - b * HW;
    const int s = pool_slot(seg + b * HW, lut, b * lut_ld, C, p, HW);
    if (s < 0) continue;
    atomicAdd(&count[s], 1);
    const T* src = x + b * (int64_t)D * HW + p;
    for (int d = 0; d < D; ++d) atomicAdd(&sum[(int64_t)s * D + d], ElemIO<T>::ld(src + (int64_t)d * HW));
  }
}

__global__ void pool_finish_kernel(float* __restrict__ sum, const int* __restrict__ count, int n, int D) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n * D) return;
  const int c = count[i / D];
  sum[i] = c > 0 ? sum[i] / (float)c : 0.f;   // dataloader.py:300-304: zeros when the mask is empty
}

template <typename T>
__global__ void __launch_bounds__(kPoolThreads)
pool_bwd_kernel(const float* __restrict__ g, const int* __restrict__ count, int B, int D, int64_t HW,
                const int64_t* __restrict__ seg, const int32_t* __restrict__ lut, int64_t lut_ld, int C,
                T* __restrict__ dx, int accumulate, bool vec_ok) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = kPoolThreads / 32;
  const int64_t chunks = (HW + kPoolPx - 1) / kPoolPx;
  const int dgroups = (D + kPoolDGroup - 1) / kPoolDGroup;
  const int64_t n_tasks = (int64_t)B * chunks * dgroups;
  for (int64_t task = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); task < n_tasks;
       task += (int64_t)gridDim.x * warps_per_block) {
    const int dg = (int)(task % dgroups);
    const int64_t rest = task / dgroups;
    const int64_t ch = rest % chunks;
    const int b = (int)(rest / chunks);
    const int64_t p0 = ch * kPoolPx + lane * 8;
    int slot[8];
    float inv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      slot[j] = pool_slot(seg + (int64_t)b * HW, lut, (int64_t)b * lut_ld, C, p0 + j, HW);
      inv[j] = 0.f;
      if (slot[j] >= 0) { const int c = count[slot[j]]; inv[j] = c > 0 ? 1.f / (float)c : 0.f; }
    }
    const int d0 = dg * kPoolDGroup;
    const int d1 = min(D, d0 + kPoolDGroup);
    T* dst = dx + ((int64_t)b * D + d0) * HW + p0;
    bool uni = true;
#pragma unroll
    for (int j = 1; j < 8; ++j) uni &= (slot[j] == slot[0]);
    if (uni && vec_ok && p0 + 8 <= HW && (D & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
      // the lane's 8 pixels share one slot (the common case away from mask borders): one 16-byte gather of four
      // channels of that slot's gradient row instead of eight scalar gathers per channel
      const bool any = slot[0] >= 0;
      const float4* grow = reinterpret_cast<const float4*>(g + (int64_t)(any ? slot[0] : 0) * D);
      for (int d = d0; d < d1; d += 4, dst += 4 * HW) {
        float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (any) gv = __ldg(grow + (d >> 2));
        const float gq[4] = {gv.x * inv[0], gv.y * inv[0], gv.z * inv[0], gv.w * inv[0]};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (d + q >= d1) break;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = gq[q];
          T* o_ = dst + (int64_t)q * HW;
          if (accumulate) {
            float o[8];
            load8(o_, o);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += o[j];
          }
          store8(o_, v);
        }
      }
      continue;
    }
    for (int d = d0; d < d1; ++d, dst += HW) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = slot[j] >= 0 ? __ldg(&g[(int64_t)slot[j] * D + d]) * inv[j] : 0.f;
      if (vec_ok && p0 + 8 <= HW) {
        if (accumulate) {
          float o[8];
          load8(dst, o);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += o[j];
        }
        store8(dst, v);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (p0 + j < HW) {
            const float o = accumulate ? ElemIO<T>::ld(dst + j) : 0.f;
            ElemIO<T>::st(dst + j, v[j] + o);
          }
      }
    }
  }
}

}  // namespace rc

extern "C" int rc_pool_fwd(const void* x, rc_dtype x_dtype, int B, int D, int64_t HW, const int64_t* seg,
                           const int32_t* lut, int64_t lut_ld, int C, int n_slots, float* sum, int32_t* count,
                           void* stream) {
  RC_REQUIRE(x && seg && lut && sum && count, "rc_pool_fwd: null pointer");
  RC_REQUIRE(B >= 0 && D >= 1 && HW >= 0 && C >= 1 && n_slots >= 0, "rc_pool_fwd: bad shape");
  if (B == 0 || HW == 0 || n_slots == 0) return RC_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec = (HW % 8 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (vec) {
    const int64_t chunk_pairs = ((HW + rc::kPoolPx - 1) / rc::kPoolPx + 1) / 2;      // a warp task = two 256-pixel chunks
    const int64_t tasks = (int64_t)B * chunk_pairs * ((D + rc::kPoolDGroup - 1) / rc::kPoolDGroup);
    const int64_t blocks = (tasks + 7) / 8;
    const int64_t cap = (int64_t)rc::num_sms() * 8;
    const int grid = (int)(blocks < cap ? blocks : cap);
    if (x_dtype == RC_F32)
      rc::pool_fwd_kernel<float><<<grid, rc::kPoolThreads, 0, s>>>((const float*)x, B, D, HW, seg, lut, lut_ld, C, sum, count);
    else
      rc::pool_fwd_kernel<__nv_bfloat16><<<grid, rc::kPoolThreads, 0, s>>>((const __nv_bfloat16*)x, B, D, HW, seg, lut, lut_ld, C, sum, count);
  } else {
    const int64_t n = (int64_t)B * HW;
    const int grid = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
    if (x_dtype == RC_F32)
      rc::pool_fwd_generic_kernel<float><<<grid, 256, 0, s>>>((const float*)x, B, D, HW, seg, lut, lut_ld, C, sum, count);
    else
      rc::pool_fwd_generic_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, B, D, HW, seg, lut, lut_ld, C, sum, count);
  }
  return rc::check_launch("rc_pool_fwd");
}

extern "C" int rc_pool_finish(float* sum_inout, const int32_t* count, int n_slots, int D, void* stream) {
  RC_REQUIRE(sum_inout && count && n_slots >= 0 && D >= 1, "rc_pool_finish: bad argument");
  if (n_slots == 0) return RC_OK;
  const int64_t n = (int64_t)n_slots * D;
  rc::pool_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(sum_inout, count, n_slots, D);
  return rc::check_launch("rc_pool_finish");
}

extern "C" int rc_pool_bwd(const float* g, const int32_t* count, int B, int D, int64_t HW, const int64_t* seg,
                           const int32_t* lut, int64_t lut_ld, int C, int n_slots, void* dx, rc_dtype x_dtype,
                           int accumulate, void* stream) {
  RC_REQUIRE(g && count && seg && lut && dx, "rc_pool_bwd: null pointer");
  RC_REQUIRE(B >= 0 && D >= 1 && HW >= 0 && C >= 1, "rc_pool_bwd: bad shape");
  (void)n_slots;
  if (B == 0 || HW == 0) return RC_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec = (HW % 8 == 0) && ((reinterpret_cast<uintptr_t>(dx) & 15) == 0);
  const int64_t tasks = (int64_t)B * ((HW + rc::kPoolPx - 1) / rc::kPoolPx) * ((D + rc::kPoolDGroup - 1) / rc::kPoolDGroup);
  const int64_t blocks = (tasks + 7) / 8;
  const int64_t cap = (int64_t)rc::num_sms() * 8;
  const int grid = (int)(blocks < cap ? blocks : cap);
  if (x_dtype == RC_F32)
    rc::pool_bwd_kernel<float><<<grid, rc::kPoolThreads, 0, s>>>(g, count, B, D, HW, seg, lut, lut_ld, C, (float*)dx, accumulate, vec);
  else
    rc::pool_bwd_kernel<__nv_bfloat16><<<grid, rc::kPoolThreads, 0, s>>>(g, count, B, D, HW, seg, lut, lut_ld, C, (__nv_bfloat16*)dx, accumulate, vec);
  return rc::check_launch("rc_pool_bwd");
}
