// dText of the pixel-text InfoNCE on the tensor cores:  dT[k][d] = sum_p G[p][k] x[d][p]
//
// G[p][k] = rs_p (e_pk - sum_p [k = y_p]) is the pre-scaled softmax-minus-onehot tile that the CTA-pair kernel
// (infonce_umma2.cu) writes as bf16 [B][HW][Kp] when dText is requested (rs_p carries w_p / sum(w), 1/tau, 1/|x_p|
// and the upstream gradient), so the gradient with respect to the normalised text rows (model.py:272-291 and its
// autograd: cross_entropy backward, matmul backward) is ONE GEMM with the pixel index as the contraction
// dimension: M = Kp text rows, N = D channels, K = B*HW pixels.
//
// Split-K over persistent CTAs: CTA c owns channel half c % (D/256) and every (gridDim / (D/256))-th 64-pixel
// slab.  Per slab the TMA brings G [64 px][Kp] (MN-major A operand: the pixel rows are the K index) and
// X [256 d][64 px] (K-major B operand, straight from NCHW); the fp32 accumulators [Kp][256] stay in TMEM (all 512
// columns at Kp = 256) for the whole kernel and are added to dt with vector reductions at the end.
// HBM bound: G is read D/256 times (2 x 2.1 GB at the headline size), X once.
#include "common.cuh"
#include "umma.cuh"
#include <stdlib.h>

namespace rc {
using namespace umma;

namespace dtk {

constexpr int kThreads = 192;            // warps: 0 TMA producer, 1 MMA issuer (+ TMEM alloc), 2-5 final reduction
constexpr int kSlabPx = 64;
constexpr int kStages = 3;
constexpr int kGBytes = 4 * 8192;        // up to four [64 px][64 k] sub-tiles
constexpr int kXBytes = 256 * 128;       // [256 d][64 px]
constexpr int kStageBytes = kGBytes + kXBytes;
constexpr int kTmemCols = 512;

struct __align__(8) Bars {
  uint64_t full[kStages], empty[kStages];
  uint64_t done;
  uint32_t tmem_base, pad;
};
constexpr int kOffBars = kStages * kStageBytes;
constexpr int kSmemBytes = kOffBars + (int)sizeof(Bars);
static_assert(kSmemBytes <= 232448, "shared-memory budget of one SM (227 KB)");

struct Params {
  int B, D, K, Kp;
  int64_t HW;
  int slabs_per_img, n_slabs, n_dh;
  float* dt;
};

__global__ void __launch_bounds__(kThreads, 1)
infonce_dt_umma_kernel(const __grid_constant__ CUtensorMap map_g,    // G [B][HW][Kp], box (64 k, 64 px, 1)
                       const __grid_constant__ CUtensorMap map_x,    // X [B][D][HW],  box (64 px, 256 d, 1)
                       const Params prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Bars* bars = reinterpret_cast<Bars*>(smem + kOffBars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dh = blockIdx.x % prm.n_dh;
  const int range = blockIdx.x / prm.n_dh, n_ranges = gridDim.x / prm.n_dh;
  const int n_mb = (prm.Kp + 127) / 128;          // 128-row blocks of text rows (M of the MMA)
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_g); tma_prefetch_desc(&map_x);
    for (int i = 0; i < kStages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    mbar_init(&bars->done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(&bars->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0 && lane == 0) {
    // =============================== TMA producer ===============================
    uint32_t it = 0;
    for (int sl = range; sl < prm.n_slabs; sl += n_ranges, ++it) {
      const int b = sl / prm.slabs_per_img;
      const int px0 = (sl - b * prm.slabs_per_img) * kSlabPx;
      const int st = it % kStages;
      mbar_wait(&bars->empty[st], ((it / kStages) & 1) ^ 1, 1);
      uint8_t* sb = smem + st * kStageBytes;
      mbar_arrive_expect_tx(&bars->full[st], n_mb * 2 * 8192 + kXBytes);     // out-of-range boxes arrive as zeros
      for (int j = 0; j < n_mb * 2; ++j) tma_load_3d(sb + j * 8192, &map_g, &bars->full[st], j * 64, px0, b);
      tma_load_3d(sb + kGBytes, &map_x, &bars->full[st], px0, dh * 256, b);
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (converged warp, one elected lane) ===============================
    const uint32_t idesc = make_idesc_bf16(128, 256, /*A MN-major*/ 1, /*B K-major*/ 0);
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t dsc_g = desc_mnmajor_sw128(0, 8192);      // two 64-row groups of text rows, 8 KB apart
    const uint64_t dsc_x = desc_kmajor_sw128(0);
    uint32_t it = 0;
    for (int sl = range; sl < prm.n_slabs; sl += n_ranges, ++it) {
      const int st = it % kStages;
      mbar_wait(&bars->full[st], (it / kStages) & 1, 2);
      tc_fence_after();
      const uint32_t sb = smem_base + st * kStageBytes;
      if (elect_one()) {
        for (int mb = 0; mb < n_mb; ++mb) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)       // 16 pixels per MMA
            mma_bf16_ss(tmem + mb * 256, dsc_g + ((sb + mb * 16384 + ks * 2048) >> 4), dsc_x + ((sb + kGBytes + ks * 32) >> 4),
                        idesc, (it | ks) != 0);
        }
        mma_commit(&bars->empty[st]);
      }
      __syncwarp();
    }
    if (elect_one()) mma_commit(&bars->done);
    __syncwarp();
  } else if (warp >= 2) {
    // ======================= final reduction: accumulators -> dt (fp32, vector reductions) =======================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    mbar_wait(&bars->done, 0, 3);
    tc_fence_after();
    const bool any = range < prm.n_slabs;       // a CTA without work has undefined accumulators
    for (int mb = 0; mb < n_mb && any; ++mb) {
      const int k = mb * 128 + row;
      for (int c = 0; c < 8; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem + ((uint32_t)(quarter * 32) << 16) + mb * 256 + c * 32, r);
        tmem_ld_wait();
        if (k < prm.K) {
          float* dst = prm.dt + (int64_t)k * prm.D + dh * 256 + c * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(__uint_as_float(r[i])),
                         "f"(__uint_as_float(r[i + 1])), "f"(__uint_as_float(r[i + 2])), "f"(__uint_as_float(r[i + 3]))
                         : "memory");
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<kTmemCols>(tmem); }
}

// ------------------------------------------------------------------------------------------------------------------
// CTA-pair form (Kp > 128): the single-CTA kernel above reads G once per 256-channel half, i.e. twice.  Here a cluster of two
// CTAs owns every (gridDim / 2)-th 64-pixel slab and ALL 512 channels: `tcgen05.mma.cta_group::2` with M = 256 text rows (CTA r
// stages and accumulates rows [128 r, 128 r + 128)) and N = 256 channels per MMA, two MMAs per 16 pixels (channels [0,256) and
// [256,512) -> TMEM columns [0,256) / [256,512) of each CTA); the B operand is split over the pair (CTA r stages channel rows
// [256 n + 128 r, +128) of block n).  G and X are each read ONCE: 2.1 + 4.3 GB at the headline size instead of 4.3 + 4.3.
// ------------------------------------------------------------------------------------------------------------------
namespace pairk {
constexpr int kStages = 4;
constexpr int kGBytes = 2 * 8192;        // own 128 text rows: two [64 px][64 k] sub-tiles
constexpr int kXBytes = 2 * 16384;       // own 128 channel rows of each of the two 256-channel blocks: 2 x [128 d][64 px]
constexpr int kStageBytes = kGBytes + kXBytes;
constexpr int kOffBars = kStages * kStageBytes;
constexpr int kSmemBytes = kOffBars + (int)sizeof(Bars) + 64;
static_assert(kStages <= 4 && kSmemBytes <= 232448, "shared-memory budget of one SM (227 KB)");
}  // namespace pairk

struct BarsP {
  uint64_t full[pairk::kStages], empty[pairk::kStages];
  uint64_t done;
  uint32_t tmem_base, pad;
};

__global__ void __launch_bounds__(kThreads, 1)
infonce_dt_umma_pair_kernel(const __grid_constant__ CUtensorMap map_g,    // G [B][HW][Kp], box (64 k, 64 px, 1)
                            const __grid_constant__ CUtensorMap map_x,    // X [B][D][HW],  box (64 px, 128 d, 1)
                            const Params prm) {
  using namespace pairk;
  extern __shared__ __align__(1024) uint8_t smem[];
  BarsP* bars = reinterpret_cast<BarsP*>(smem + pairk::kOffBars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int range = blockIdx.x >> 1, n_ranges = gridDim.x >> 1;
  const int n_nb = prm.D / 256;                   // 256-channel blocks (N of the MMA): 1 or 2
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_g); tma_prefetch_desc(&map_x);
    for (int i = 0; i < pairk::kStages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    mbar_init(&bars->done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm<kTmemCols>(&bars->tmem_base);
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0 && lane == 0) {
    // =============================== TMA producer (both CTAs) ===============================
    uint32_t it = 0;
    for (int sl = range; sl < prm.n_slabs; sl += n_ranges, ++it) {
      const int b = sl / prm.slabs_per_img;
      const int px0 = (sl - b * prm.slabs_per_img) * kSlabPx;
      const int st = it % pairk::kStages;
      mbar_wait(&bars->empty[st], ((it / pairk::kStages) & 1) ^ 1, 1);
      uint8_t* sb = smem + st * pairk::kStageBytes;
      if (leader) mbar_arrive_expect_tx(&bars->full[st], 2 * (pairk::kGBytes + n_nb * 16384));      // both CTAs' bytes; out-of-range boxes arrive as zeros
      for (int j = 0; j < 2; ++j) tma_load_3d_2sm(sb + j * 8192, &map_g, &bars->full[st], (int)rank * 128 + j * 64, px0, b);
      for (int n = 0; n < n_nb; ++n)
        tma_load_3d_2sm(sb + pairk::kGBytes + n * 16384, &map_x, &bars->full[st], px0, n * 256 + (int)rank * 128, b);
    }
  } else if (warp == 1 && leader) {
    // =============================== MMA issuer (leader CTA; converged warp, one elected lane) ===============================
    const uint32_t idesc = make_idesc_bf16(256, 256, /*A MN-major*/ 1, /*B K-major*/ 0);
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t dsc_g = desc_mnmajor_sw128(0, 8192);      // two 64-row groups of text rows, 8 KB apart
    const uint64_t dsc_x = desc_kmajor_sw128(0);
    uint32_t it = 0;
    for (int sl = range; sl < prm.n_slabs; sl += n_ranges, ++it) {
      const int st = it % pairk::kStages;
      mbar_wait(&bars->full[st], (it / pairk::kStages) & 1, 2);
      tc_fence_after();
      const uint32_t sb = smem_base + st * pairk::kStageBytes;
      if (elect_one()) {
        for (int n = 0; n < n_nb; ++n) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)       // 16 pixels per MMA
            mma_bf16_ss_2sm(tmem + n * 256, dsc_g + ((sb + ks * 2048) >> 4), dsc_x + ((sb + pairk::kGBytes + n * 16384 + ks * 32) >> 4),
                            idesc, (it | ks) != 0);
        }
        mma_commit_2sm(&bars->empty[st]);
      }
      __syncwarp();
    }
    if (elect_one()) mma_commit_2sm(&bars->done);
    __syncwarp();
  } else if (warp >= 2) {
    // ======================= final reduction: own 128 text rows x all channels -> dt (fp32, vector reductions) =======================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    mbar_wait(&bars->done, 0, 3);
    tc_fence_after();
    const bool any = range < prm.n_slabs;       // a pair without work has undefined accumulators
    const int k = (int)rank * 128 + row;
    for (int c = 0; c < n_nb * 8 && any; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem + ((uint32_t)(quarter * 32) << 16) + c * 32, r);
      tmem_ld_wait();
      if (k < prm.K) {
        float* dst = prm.dt + (int64_t)k * prm.D + c * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(__uint_as_float(r[i])),
                       "f"(__uint_as_float(r[i + 1])), "f"(__uint_as_float(r[i + 2])), "f"(__uint_as_float(r[i + 3]))
                       : "memory");
      }
    }
  }
  tc_fence_before();
  cluster_sync();
  if (warp == 1) { tc_fence_after(); tmem_dealloc_2sm<kTmemCols>(tmem); }
}

}  // namespace dtk

int launch_infonce_dt(const void* g, const void* xsrc, int B, int D, int64_t HW, int K, float* dt, cudaStream_t s) {
  using namespace dtk;
  if (D % 256 != 0 || D > 512) return fail(RC_ERR_UNSUPPORTED, "rc_infonce_bf16(dText): D=%d must be 256 or 512", D);
  RC_REQUIRE((reinterpret_cast<uintptr_t>(dt) & 15) == 0, "rc_infonce_bf16(dText): dt must be 16-byte aligned");
  const int Kp = (K + 63) / 64 * 64;
  CUtensorMap m_g, m_x;
  int rcode;
  {
    const uint64_t gdims[3] = {(uint64_t)Kp, (uint64_t)HW, (uint64_t)B};
    const uint64_t gstr[3] = {2, (uint64_t)Kp * 2, (uint64_t)HW * Kp * 2};
    const uint32_t gbox[3] = {64, (uint32_t)kSlabPx, 1};
    if ((rcode = make_tmap_bf16(&m_g, g, 3, gdims, gstr, gbox, "dt map_g"))) return rcode;
    const uint64_t dims[3] = {(uint64_t)HW, (uint64_t)D, (uint64_t)B};
    const uint64_t str[3] = {2, (uint64_t)HW * 2, (uint64_t)D * HW * 2};
    const uint32_t box[3] = {(uint32_t)kSlabPx, 256, 1};
    if ((rcode = make_tmap_bf16(&m_x, xsrc, 3, dims, str, box, "dt map_x"))) return rcode;
  }
  Params prm;
  prm.B = B; prm.D = D; prm.K = K; prm.Kp = Kp; prm.HW = HW;
  prm.slabs_per_img = (int)((HW + kSlabPx - 1) / kSlabPx);
  if ((int64_t)B * prm.slabs_per_img > 0x7fffffff) return fail(RC_ERR_UNSUPPORTED, "rc_infonce_bf16(dText): too many slabs");
  prm.n_slabs = B * prm.slabs_per_img;
  prm.n_dh = D / 256;
  prm.dt = dt;
  bool use_pair = Kp > 128 && prm.n_slabs >= 2;
#ifdef RC_BRINGUP
  { const char* v = getenv("RANGECLIP_B200_DT"); if (v != nullptr && v[0] == '1') use_pair = false; }      // A/B: "1cta"
#endif
  if (use_pair) {
    // CTA pairs: every pair takes all channels, each CTA half of the text rows -- G and X are read once
    CUtensorMap m_xp;
    const uint64_t dims[3] = {(uint64_t)HW, (uint64_t)D, (uint64_t)B};
    const uint64_t str[3] = {2, (uint64_t)HW * 2, (uint64_t)D * HW * 2};
    const uint32_t box[3] = {(uint32_t)kSlabPx, 128, 1};
    if ((rcode = make_tmap_bf16(&m_xp, xsrc, 3, dims, str, box, "dt map_x (pair)"))) return rcode;
    int n_clusters = num_sms() / 2;
    if (n_clusters > prm.n_slabs) n_clusters = prm.n_slabs;
    cudaError_t e = cudaFuncSetAttribute(infonce_dt_umma_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pairk::kSmemBytes);
    if (e != cudaSuccess) return fail(RC_ERR_CUDA, "rc_infonce_bf16(dText): smem opt-in: %s", cudaGetErrorString(e));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * n_clusters);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = pairk::kSmemBytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, infonce_dt_umma_pair_kernel, m_g, m_xp, prm);
    if (e != cudaSuccess) return fail(RC_ERR_CUDA, "rc_infonce_bf16(dText): cluster launch: %s", cudaGetErrorString(e));
    return check_launch("rc_infonce_bf16(dText)");
  }
  int grid = num_sms() / prm.n_dh * prm.n_dh;
  if (grid > prm.n_slabs * prm.n_dh) grid = prm.n_slabs * prm.n_dh;
  cudaError_t e = cudaFuncSetAttribute(infonce_dt_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e != cudaSuccess) return fail(RC_ERR_CUDA, "rc_infonce_bf16(dText): smem opt-in: %s", cudaGetErrorString(e));
  infonce_dt_umma_kernel<<<grid, kThreads, kSmemBytes, s>>>(m_g, m_x, prm);
  return check_launch("rc_infonce_bf16(dText)");
}

}  // namespace rc
