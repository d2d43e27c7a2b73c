// K1/K2, CTA-pair version: the fused pixel-text InfoNCE forward + backward of infonce_umma.cu on
// `tcgen05.mma.cta_group::2`.
//
// Two CTAs on neighbouring SMs form a cluster and work on two 128-pixel tiles at a time.  Every MMA
// spans both SMs (M = 256): each CTA stages only ITS half of the streamed operand and the tensor cores
// read the other half from the peer's shared memory, which halves the shared-memory fill traffic and
// the operand read traffic per SM -- the two limits the single-CTA kernel runs into
// (profiles/r1_infonce_umma_ncu.txt):
//   S   = X^T T^T   M = 256 px (128 per CTA, A = own X tile, MN-major), N = Kp text rows, B split: Kp/2 rows per CTA
//   dX^T = T^T P^T   M = 256 channels (128 per CTA, A = own rows of T^T), N = 128 px (64 of each CTA's tile,
//                    B = own P rows), four [128 ch x 128 px] accumulators per tile pair per CTA
// The softmax/CE epilogue is per CTA on its own tile; the dX epilogue of a CTA covers its 128 channels
// of every 256-channel block for the pixels of BOTH tiles, so the per-pixel row scales are exchanged
// through distributed shared memory.  Only the leader CTA (cluster rank 0) issues MMAs; completion is
// multicast to the mbarriers of both CTAs, consumer-release barriers live in the leader and receive
// remote arrivals from the peer.
#include "common.cuh"
#include "umma.cuh"
#include <float.h>

namespace rc {
using namespace umma;

namespace pair {

constexpr int kTilePx = 128;
constexpr int kThreads = 640;          // warps: 0 TMA, 1 MMA (leader CTA only), 2 TMEM alloc, 3 staging DMA, 4-11 softmax, 12-19 dX epilogue
constexpr int kStages = 4;
constexpr int kStageBytes = 32 * 1024; // two own X chunks | two text half-chunks [Kp/2][64 d] | own T^T rows [128 d][<=128 k]
constexpr int kPBytes = 64 * 1024;
constexpr int kStgBufs = 3;
constexpr int kStgPx = 32;
constexpr int kStgBytes = 128 * kStgPx * 2;
constexpr int kTmemCols = 512;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct __align__(8) Bars {
  uint64_t full[kStages], empty[kStages];
  uint64_t s_full, s_empty, p_full, p_empty;
  uint64_t acc_full[2], acc_empty[2];
  uint64_t sc_full[2];
  uint64_t stg_full[kStgBufs], stg_done[kStgBufs];
  uint32_t tmem_base, pad;
};

constexpr int kOffP = kStages * kStageBytes;
constexpr int kOffStg = kOffP + kPBytes;
constexpr int kScaleBufs = 3;          // the softmax warps run up to two tiles ahead of the dX epilogue warps
constexpr int kOffScale = kOffStg + kStgBufs * kStgBytes;   // rs, cs: [kScaleBufs tiles][2 owner CTAs][128] floats each
constexpr int kOffXch = kOffScale + 2 * kScaleBufs * 2 * 128 * 4;
constexpr int kOffBars = kOffXch + 4 * 2 * 128 * 4;
constexpr int kSmemBytes = kOffBars + (int)sizeof(Bars);
static_assert(kSmemBytes <= 232448, "shared-memory budget of one SM (227 KB)");

#ifdef RC_TIMING
#define RC_T0(name) const long long name = clock64()
#define RC_TACC(idx, name) wt[idx] += clock64() - name
#define RC_WAIT(fn, bar, par, tag) do { const long long t0_ = clock64(); fn(bar, par, tag); wt[tag] += clock64() - t0_; } while (0)
#else
#define RC_T0(name)
#define RC_TACC(idx, name)
#define RC_WAIT(fn, bar, par, tag) fn(bar, par, tag)
#endif

struct Params {
  long long* dbg;
  int B, D, K, Kp;
  int64_t HW;
  int tiles_per_img, n_tiles, n_pairs;
  const float* inv_norm;
  const int32_t* y;
  const float* w;
  float inv_tau;
  const float* grad_scale;
  const double* w_sum_in;
  float* lse;
  double* loss_sum;
  double* w_sum;
  double* dlogtau;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// r[i] for a per-thread dynamic i (registers cannot be indexed dynamically): binary select tree
__device__ __forceinline__ float select32(const uint32_t (&r)[32], int i) {
  float a[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) a[j] = __uint_as_float((i & 1) ? r[2 * j + 1] : r[2 * j]);
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = (i & 2) ? a[2 * j + 1] : a[2 * j];
#pragma unroll
  for (int j = 0; j < 4; ++j) a[j] = (i & 4) ? a[2 * j + 1] : a[2 * j];
#pragma unroll
  for (int j = 0; j < 2; ++j) a[j] = (i & 8) ? a[2 * j + 1] : a[2 * j];
  return (i & 16) ? a[1] : a[0];
}

// tile -> (image, first pixel); tiles past the end map to image index B, which is out of bounds for every
// tensor map (TMA loads return zeros, TMA stores write nothing)
__device__ __forceinline__ void tile_coords(const Params& prm, int tile, int& b, int& px0) {
  if (tile < prm.n_tiles) {
    b = tile / prm.tiles_per_img;
    px0 = (tile - b * prm.tiles_per_img) * kTilePx;
  } else {
    b = prm.B;
    px0 = 0;
  }
}

template <bool kBwd>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
infonce_umma_pair_kernel(const __grid_constant__ CUtensorMap map_x_s,   // X [B][D][HW], box (64 px, 64 d, 1)
                         const __grid_constant__ CUtensorMap map_t,     // T [Kp][D],    box (64 d, Kp/2 rows)
                         const __grid_constant__ CUtensorMap map_tt,    // T^T [D][Kp],  box (64 k, 128 d)
                         const __grid_constant__ CUtensorMap map_x_e,   // X,            box (32 px, 128 d, 1)
                         const __grid_constant__ CUtensorMap map_dx,    // dX,           box (32 px, 128 d, 1)
                         const Params prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Bars* bars = reinterpret_cast<Bars*>(smem + kOffBars);
  float* rs_s = reinterpret_cast<float*>(smem + kOffScale);   // [kScaleBufs][2][128]
  float* cs_s = rs_s + kScaleBufs * 2 * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader_cta = rank == 0;
  const int n_cp = prm.D / 128;          // X chunk pairs per tile
  const int n_blk = prm.D / 256;         // 256-channel blocks of the dX GEMM
  const int n_kchunks = prm.Kp / 64;
  const int n_units = (n_kchunks + 1) / 2;
  const int Nh = prm.Kp / 2;             // text rows staged by each CTA
  const int n_clusters = gridDim.x / 2;
  const int cluster_id = blockIdx.x / 2;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_x_s); tma_prefetch_desc(&map_t);
    if (kBwd) { tma_prefetch_desc(&map_tt); tma_prefetch_desc(&map_x_e); tma_prefetch_desc(&map_dx); }
    for (int i = 0; i < kStages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    mbar_init(&bars->s_full, 1); mbar_init(&bars->s_empty, 512);
    mbar_init(&bars->p_full, 512); mbar_init(&bars->p_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->acc_full[i], 1); mbar_init(&bars->acc_empty[i], 512); }
    mbar_init(&bars->sc_full[0], 512); mbar_init(&bars->sc_full[1], 512);
    for (int i = 0; i < kStgBufs; ++i) { mbar_init(&bars->stg_full[i], 1); mbar_init(&bars->stg_done[i], 256); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm<kTmemCols>(&bars->tmem_base);
  if (kBwd) {
    for (int i = threadIdx.x; i < kPBytes / 16; i += kThreads) reinterpret_cast<uint4*>(smem + kOffP)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  cluster_sync();            // both CTAs' barriers are initialised before any remote arrive / 2-SM TMA credit
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
#ifdef RC_TIMING
  long long wt[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) wt[i] = 0;
  const long long t_start = clock64();
#endif
  const uint32_t idesc_s = make_idesc_bf16(256, prm.Kp, /*A MN-major*/ 1, /*B K-major*/ 0);
  const uint32_t idesc_d = make_idesc_bf16(256, 128, 0, 0);
  // consumer-release barriers live in the leader CTA
  auto arrive_leader = [&](uint64_t* bar) {
    if (leader_cta) mbar_arrive(bar);
    else mbar_arrive_remote(map_to_cta(bar, 0));
  };

  if (warp == 0 && lane == 0) {
    // =============================== TMA producer (both CTAs) ===============================
    uint32_t it = 0;
    for (int pj = cluster_id; pj < prm.n_pairs; pj += n_clusters) {
      int b, px0;
      tile_coords(prm, 2 * pj + (int)rank, b, px0);
      for (int cp = 0; cp < n_cp; ++cp) {
        {   // own X chunks 2cp, 2cp+1
          const int st = it % kStages;
          RC_WAIT(mbar_wait, &bars->empty[st], ((it / kStages) & 1) ^ 1, 1);
          uint8_t* sb = smem + st * kStageBytes;
          if (leader_cta) mbar_arrive_expect_tx(&bars->full[st], 2 * 4 * 8192);
          for (int cc = 0; cc < 2; ++cc) {
            tma_load_3d_2sm(sb + cc * 16384, &map_x_s, &bars->full[st], px0, (2 * cp + cc) * 64, b);
            tma_load_3d_2sm(sb + cc * 16384 + 8192, &map_x_s, &bars->full[st], px0 + 64, (2 * cp + cc) * 64, b);
          }
          ++it;
        }
        {   // own half (Nh rows) of the text chunks 2cp, 2cp+1
          const int st = it % kStages;
          RC_WAIT(mbar_wait, &bars->empty[st], ((it / kStages) & 1) ^ 1, 1);
          uint8_t* sb = smem + st * kStageBytes;
          if (leader_cta) mbar_arrive_expect_tx(&bars->full[st], 2 * 2 * Nh * 128);
          for (int cc = 0; cc < 2; ++cc)
            tma_load_2d_2sm(sb + cc * 16384, &map_t, &bars->full[st], (2 * cp + cc) * 64, (int)rank * Nh);
          ++it;
        }
      }
      if (kBwd) {
        for (int blk = 0; blk < n_blk; ++blk)
          for (int u = 0; u < n_units; ++u, ++it) {   // own 128 rows of T^T for this 256-channel block
            const int st = it % kStages;
            RC_WAIT(mbar_wait, &bars->empty[st], ((it / kStages) & 1) ^ 1, 2);
            uint8_t* sb = smem + st * kStageBytes;
            const int nb = min(2, n_kchunks - 2 * u);
            if (leader_cta) mbar_arrive_expect_tx(&bars->full[st], 2 * nb * 16384);
            for (int jj = 0; jj < nb; ++jj)
              tma_load_2d_2sm(sb + jj * 16384, &map_tt, &bars->full[st], (2 * u + jj) * 64, blk * 256 + (int)rank * 128);
          }
      }
    }
  } else if (warp == 1 && lane == 0 && leader_cta) {
    // =============================== MMA issuer (leader CTA) ================================
    uint32_t it = 0, lt = 0, uc = 0;
    for (int pj = cluster_id; pj < prm.n_pairs; pj += n_clusters, ++lt) {
      RC_WAIT(mbar_wait_cluster, &bars->s_empty, (lt & 1) ^ 1, 3);
      tc_fence_after();
      for (int cp = 0; cp < n_cp; ++cp, it += 2) {
        const int sa = it % kStages, sb_ = (it + 1) % kStages;
        RC_WAIT(mbar_wait_cluster, &bars->full[sa], (it / kStages) & 1, 4);
        RC_WAIT(mbar_wait_cluster, &bars->full[sb_], ((it + 1) / kStages) & 1, 4);
        tc_fence_after();
        const uint32_t xa = smem_u32(smem + sa * kStageBytes);
        const uint32_t tb = smem_u32(smem + sb_ * kStageBytes);
        for (int cc = 0; cc < 2; ++cc) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t a = desc_mnmajor_sw128(xa + cc * 16384 + ks * 2048, 8192);
            const uint64_t bdesc = desc_kmajor_sw128(tb + cc * 16384 + ks * 32);
            mma_bf16_ss_2sm(tmem, a, bdesc, idesc_s, (cp | cc | ks) != 0);
          }
        }
        mma_commit_2sm(&bars->empty[sa]);
        mma_commit_2sm(&bars->empty[sb_]);
      }
      mma_commit_2sm(&bars->s_full);
      if (kBwd) {
        RC_WAIT(mbar_wait_cluster, &bars->p_full, lt & 1, 5);
        tc_fence_after();
        const uint32_t pb = smem_u32(smem + kOffP);
        for (int blk = 0; blk < n_blk; ++blk) {
          for (int u = 0; u < n_units; ++u) {
            const uint32_t jt = it + u;
            RC_WAIT(mbar_wait_cluster, &bars->full[jt % kStages], (jt / kStages) & 1, 7);
          }
          tc_fence_after();
          for (int pxh = 0; pxh < 2; ++pxh, ++uc) {
            const int ab = uc & 1;
            RC_WAIT(mbar_wait_cluster, &bars->acc_empty[ab], ((uc >> 1) & 1) ^ 1, 6);
            tc_fence_after();
            const uint32_t dcol = tmem + 256 + ab * 128;
            for (int u = 0; u < n_units; ++u) {
              const uint32_t sb = smem_u32(smem + ((it + u) % kStages) * kStageBytes);
              const int nb = min(2, n_kchunks - 2 * u);
              for (int jj = 0; jj < nb; ++jj) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  const uint64_t a = desc_kmajor_sw128(sb + jj * 16384 + ks * 32);                              // own T^T rows [128 d][64 k]
                  const uint64_t bdesc = desc_kmajor_sw128(pb + (2 * u + jj) * 16384 + pxh * 8192 + ks * 32);   // own P rows [64 px][64 k]
                  mma_bf16_ss_2sm(dcol, a, bdesc, idesc_d, (u | jj | ks) != 0);
                }
              }
            }
            mma_commit_2sm(&bars->acc_full[ab]);
          }
          for (int u = 0; u < n_units; ++u) mma_commit_2sm(&bars->empty[(it + u) % kStages]);
          it += n_units;
        }
        mma_commit_2sm(&bars->p_empty);
      }
    }
  } else if (kBwd && warp == 3 && lane == 0) {
    // ================= staging DMA: X loads for the dX epilogue, dX stores (both CTAs) =================
    const int steps_per_pair = n_blk * 2 * 4;
    const int my_pairs = (prm.n_pairs > cluster_id) ? (prm.n_pairs - cluster_id + n_clusters - 1) / n_clusters : 0;
    const int64_t total_steps = (int64_t)my_pairs * steps_per_pair;
    struct Cursor { int pj, rem, buf, b[2], px0[2]; };
    auto load_pair = [&](Cursor& c) {
      tile_coords(prm, 2 * c.pj, c.b[0], c.px0[0]);
      tile_coords(prm, 2 * c.pj + 1, c.b[1], c.px0[1]);
    };
    auto init = [&](Cursor& c) { c.pj = cluster_id; c.rem = 0; c.buf = 0; load_pair(c); };
    auto advance = [&](Cursor& c) {
      c.buf = (c.buf + 1 == kStgBufs) ? 0 : c.buf + 1;
      if (++c.rem == steps_per_pair) { c.rem = 0; c.pj += n_clusters; load_pair(c); }
    };
    // step `rem` = (blk, pxh, h): owner CTA = h >> 1, pixels [pxh*64 + (h&1)*32, +32) of the owner's tile,
    // channels [blk*256 + rank*128, +128)
    auto coords = [&](const Cursor& c, int& cx, int& cd, int& cb) {
      const int h = c.rem & 3, pxh = (c.rem >> 2) & 1, blk = c.rem >> 3, owner = h >> 1;
      cx = c.px0[owner] + pxh * 64 + (h & 1) * kStgPx;
      cd = blk * 256 + (int)rank * 128;
      cb = c.b[owner];
    };
    Cursor ld, stc;
    init(ld); init(stc);
    auto issue_next_load = [&]() {
      int cx, cd, cb;
      coords(ld, cx, cd, cb);
      mbar_arrive_expect_tx(&bars->stg_full[ld.buf], kStgBytes);
      tma_load_3d(smem + kOffStg + ld.buf * kStgBytes, &map_x_e, &bars->stg_full[ld.buf], cx, cd, cb);
      advance(ld);
    };
    for (int i = 0; i < kStgBufs - 1; ++i)
      if (i < total_steps) issue_next_load();
    uint32_t par = 0;
    for (int64_t s = 0; s < total_steps; ++s) {
      RC_WAIT(mbar_wait, &bars->stg_done[stc.buf], par, 13);
      int cx, cd, cb;
      coords(stc, cx, cd, cb);
      tma_store_3d(&map_dx, smem + kOffStg + stc.buf * kStgBytes, cx, cd, cb);
      tma_store_commit();
      if (stc.buf + 1 == kStgBufs) par ^= 1;
      advance(stc);
      if (s + kStgBufs - 1 < total_steps) {
        tma_store_wait_read0_keep1();
        issue_next_load();
      }
    }
    tma_store_wait_all0();
  } else if (warp >= 4 && warp < 12) {
    // ======================= softmax / CE warps: own tile (two warps per TMEM lane quarter) =======================
    const int half = warp >= 8 ? 1 : 0;
    const int row = (warp & 3) * 32 + lane;                 // softmax: pixel of the own tile; epilogue: own channel row
    const int Kh = prm.Kp >> 1;
    const int cb = half * Kh;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + cb;
    uint8_t* prow = smem + kOffP + row * 128;
    const int sw = row & 7;
    float* xch = reinterpret_cast<float*>(smem + kOffXch);
    float loss_acc = 0.f, w_acc = 0.f, dlt_acc = 0.f;
    float inv_wsum = 0.f, gscale = 1.f;
    if (kBwd) {
      const double ws = prm.w_sum_in[0];
      inv_wsum = ws > 0.0 ? (float)(1.0 / ws) : 0.f;
      if (prm.grad_scale) gscale = prm.grad_scale[0];
    }
    float nx_inv_n = 0.f, nx_w = 0.f;
    int nx_y = -1;
    auto load_pixel_scalars = [&](int pj) {
      nx_inv_n = 0.f; nx_w = 0.f; nx_y = -1;
      const int t = 2 * pj + (int)rank;
      if (pj < prm.n_pairs && t < prm.n_tiles) {
        const int tb = t / prm.tiles_per_img;
        const int tpx = (t - tb * prm.tiles_per_img) * kTilePx + row;
        if (tpx < prm.HW) {
          const int64_t tm = (int64_t)tb * prm.HW + tpx;
          nx_inv_n = __ldg(prm.inv_norm + tm);
          nx_y = __ldg(prm.y + tm);
          nx_w = __ldg(prm.w + tm);
        }
      }
    };
    load_pixel_scalars(cluster_id);
    const bool use_bound = prm.inv_tau * (2.02f * kLog2e) < 100.f;
    const float ml_bound = prm.inv_tau * (1.01f * kLog2e);
    // peer copies of the row-scale arrays (same offsets in the other CTA's shared memory)
    const uint32_t rs_peer = map_to_cta(rs_s, rank ^ 1), cs_peer = map_to_cta(cs_s, rank ^ 1);
    const uint32_t sc_peer0 = map_to_cta(&bars->sc_full[0], rank ^ 1), sc_peer1 = map_to_cta(&bars->sc_full[1], rank ^ 1);
    uint32_t lt = 0;
    for (int pj = cluster_id; pj < prm.n_pairs; pj += n_clusters, ++lt) {
      const int tile = 2 * pj + (int)rank;
      const bool tile_ok = tile < prm.n_tiles;
      const int b = tile_ok ? tile / prm.tiles_per_img : 0;
      const int px = tile_ok ? (tile - b * prm.tiles_per_img) * kTilePx + row : 0;
      const bool valid = tile_ok && px < prm.HW;
      const int64_t m = (int64_t)b * prm.HW + px;
      const float inv_n = nx_inv_n;
      const int yi = nx_y;
      const float wi = yi >= 0 ? nx_w : 0.f;
      load_pixel_scalars(pj + n_clusters);
      const float zs = inv_n * prm.inv_tau;
      const float zl = zs * kLog2e;
      RC_WAIT(mbar_wait, &bars->s_full, lt & 1, 8);
      tc_fence_after();
      RC_T0(tsm);
      float mx = -FLT_MAX;
      for (int c = 0; !use_bound && c * 32 < Kh; ++c) {
        const int nvalid = prm.K - (cb + c * 32);
        if (nvalid <= 0) break;
        uint32_t r[32];
        tmem_ld_32x32(trow + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) if (i < nvalid) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
      float ml = ml_bound;
      if (!use_bound) {
        xch[(0 * 2 + half) * 128 + row] = mx;
        named_bar_sync(2, 256);
        mx = fmaxf(mx, xch[(0 * 2 + (half ^ 1)) * 128 + row]);
        ml = mx * zl;
      }
      if (kBwd) RC_WAIT(mbar_wait, &bars->p_empty, (lt & 1) ^ 1, 9);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f, sy = 0.f;
      for (int c = 0; c * 32 < Kh; ++c) {
        const int k0 = cb + c * 32;
        const int nvalid = prm.K - k0;
        if (nvalid <= 0) break;
        uint32_t r[32];
        tmem_ld_32x32(trow + c * 32, r);
        tmem_ld_wait();
        const int yrel = yi - k0;
        if ((unsigned)yrel < 32u) sy = select32(r, yrel);      // target logit: once per row, not per column
        uint32_t pk[16];
        if (nvalid >= 32) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float a0 = __uint_as_float(r[i]), a1 = __uint_as_float(r[i + 1]);
            const float a2 = __uint_as_float(r[i + 2]), a3 = __uint_as_float(r[i + 3]);
            const float e0 = fast_exp2(fmaf(a0, zl, -ml)), e1 = fast_exp2(fmaf(a1, zl, -ml));
            const float e2 = fast_exp2(fmaf(a2, zl, -ml)), e3 = fast_exp2(fmaf(a3, zl, -ml));
            s0 += e0; s1 += e1; s2 += e2; s3 += e3;
            q0 = fmaf(e0, a0, q0); q1 = fmaf(e1, a1, q1); q2 = fmaf(e2, a2, q2); q3 = fmaf(e3, a3, q3);
            pk[i >> 1] = pack_bf16x2(e0, e1);
            pk[(i >> 1) + 1] = pack_bf16x2(e2, e3);
          }
        } else {          // the one chunk that straddles K
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float a0 = __uint_as_float(r[i]), a1 = __uint_as_float(r[i + 1]);
            const float e0 = (i < nvalid) ? fast_exp2(fmaf(a0, zl, -ml)) : 0.f;
            const float e1 = (i + 1 < nvalid) ? fast_exp2(fmaf(a1, zl, -ml)) : 0.f;
            s0 += e0; s1 += e1;
            q0 = fmaf(e0, a0, q0); q1 = fmaf(e1, a1, q1);
            pk[i >> 1] = pack_bf16x2(e0, e1);
          }
        }
        if (kBwd) {
          uint8_t* sub = prow + (k0 >> 6) * 16384;
          const int cbase = (k0 & 32) >> 3;
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(sub + (((cbase + g) ^ sw) << 4)) = make_uint4(pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
        }
      }
      tc_fence_before();
      arrive_leader(&bars->s_empty);
      float sum = (s0 + s1) + (s2 + s3);
      float sez = (q0 + q1) + (q2 + q3);
      const bool mine_y = yi >= cb && yi < cb + Kh;
      xch[(1 * 2 + half) * 128 + row] = sum;
      xch[(2 * 2 + half) * 128 + row] = sez;
      xch[(3 * 2 + half) * 128 + row] = mine_y ? sy : 0.f;
      named_bar_sync(2, 256);
      sum += xch[(1 * 2 + (half ^ 1)) * 128 + row];
      sez += xch[(2 * 2 + (half ^ 1)) * 128 + row];
      sy = mine_y ? sy : xch[(3 * 2 + (half ^ 1)) * 128 + row];
      const float zy = sy * zs;
      if (half == 0) {
        const float lse = (ml + __log2f(sum)) * kLn2;
        loss_acc += wi * (lse - zy);
        w_acc += wi;
        if (valid && prm.lse) prm.lse[m] = lse;
      }
      if (!kBwd) {
        named_bar_sync(2, 256);     // exchange buffers are rewritten by the next tile
        continue;
      }
      {
        const float coef = gscale * wi * inv_wsum;
        const float inv_sum = 1.f / sum;
        if (mine_y) {
          const float ey = fast_exp2(fmaf(sy, zl, -ml));
          const int kk = yi & 63;
          uint8_t* sub = prow + (yi >> 6) * 16384;
          *reinterpret_cast<__nv_bfloat16*>(sub + (((kk >> 3) ^ sw) << 4) + (kk & 7) * 2) = __float2bfloat16_rn(ey - sum);
        }
        if (half == 0) {
          const float cj = coef * (sez * zs * inv_sum - zy);
          const float rsv = inv_n * prm.inv_tau * coef * inv_sum, csv = inv_n * inv_n * cj;
          const int idx = ((lt % kScaleBufs) * 2 + (int)rank) * 128 + row;     // [tile buffer][owner = this CTA][pixel]
          rs_s[idx] = rsv; cs_s[idx] = csv;
          st_remote_f32(rs_peer + idx * 4, rsv);
          st_remote_f32(cs_peer + idx * 4, csv);
          dlt_acc -= cj;
        }
        fence_proxy_async_smem();                 // P is read by the tensor cores (async proxy)
        arrive_leader(&bars->p_full);
        mbar_arrive(&bars->sc_full[lt & 1]);              // row scales of this tile: visible here ...
        mbar_arrive_remote((lt & 1) ? sc_peer1 : sc_peer0);   // ... and in the peer CTA
        named_bar_sync(2, 256);                   // exchange buffers are rewritten by the next tile
      }
      RC_TACC(1, tsm);
    }
    if (half == 0) {
      loss_acc = warp_sum(loss_acc); w_acc = warp_sum(w_acc); dlt_acc = warp_sum(dlt_acc);
      if (lane == 0) {
        if (prm.loss_sum) atomicAdd(prm.loss_sum, (double)loss_acc);
        if (prm.w_sum) atomicAdd(prm.w_sum, (double)w_acc);
        if (kBwd && prm.dlogtau) atomicAdd(prm.dlogtau, (double)dlt_acc);
      }
    }
  } else if (kBwd && warp >= 12) {
    // ================ dX epilogue warps: own channel rows, pixels of both tiles (two warps per lane quarter) ================
    const int half = warp >= 16 ? 1 : 0;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t trow_acc = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + half * 16;
    int sb = 0;
    uint32_t sb_par = 0, uc = 0, lt = 0;
    for (int pj = cluster_id; pj < prm.n_pairs; pj += n_clusters, ++lt) {
      // ------------------------------ dX epilogue: own channels, pixels of both tiles ------------------------------
      RC_WAIT(mbar_wait_cluster, &bars->sc_full[lt & 1], (lt >> 1) & 1, 10);
      const float* rs = rs_s + (lt % kScaleBufs) * 256;
      const float* cs = cs_s + (lt % kScaleBufs) * 256;
      for (int blk = 0; blk < n_blk; ++blk) {
        for (int pxh = 0; pxh < 2; ++pxh, ++uc) {
          const int ab = uc & 1;
          RC_WAIT(mbar_wait, &bars->acc_full[ab], (uc >> 1) & 1, 11);
          tc_fence_after();
          for (int h = 0; h < 4; ++h) {
            // accumulator columns [h*32, +32) = pixels [pxh*64 + (h&1)*32, +32) of the tile owned by CTA (h >> 1)
            uint32_t acc[16];
            tmem_ld_32x16(trow_acc + ab * 128 + h * kStgPx, acc);
            tmem_ld_wait();
            if (h == 3) { tc_fence_before(); arrive_leader(&bars->acc_empty[ab]); }
            RC_WAIT(mbar_wait, &bars->stg_full[sb], sb_par, 12);
            RC_T0(tep);
            uint8_t* srow = smem + kOffStg + sb * kStgBytes + row * 64;
            const int sw64 = (row >> 1) & 3;
            const int pbase = (h >> 1) * 128 + pxh * 64 + (h & 1) * kStgPx + half * 16;
            const float* rsp = rs + pbase;
            const float* csp = cs + pbase;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              uint4* p = reinterpret_cast<uint4*>(srow + (((half * 2 + g) ^ sw64) << 4));
              const uint4 xv = *p;
              const uint32_t xu[4] = {xv.x, xv.y, xv.z, xv.w};
              const float4 r0 = *reinterpret_cast<const float4*>(rsp + g * 8);
              const float4 r1 = *reinterpret_cast<const float4*>(rsp + g * 8 + 4);
              const float4 c0 = *reinterpret_cast<const float4*>(csp + g * 8);
              const float4 c1 = *reinterpret_cast<const float4*>(csp + g * 8 + 4);
              const float rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
              const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
              float o[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float xval = (i & 1) ? __uint_as_float(xu[i >> 1] & 0xffff0000u) : __uint_as_float(xu[i >> 1] << 16);
                o[i] = fmaf(rr[i], __uint_as_float(acc[g * 8 + i]), -cc[i] * xval);
              }
              uint4 ov;
              ov.x = pack_bf16x2(o[0], o[1]); ov.y = pack_bf16x2(o[2], o[3]);
              ov.z = pack_bf16x2(o[4], o[5]); ov.w = pack_bf16x2(o[6], o[7]);
              *p = ov;
            }
            fence_proxy_async_smem();
            mbar_arrive(&bars->stg_done[sb]);
            RC_TACC(2, tep);
            if (++sb == kStgBufs) { sb = 0; sb_par ^= 1; }
          }
        }
      }
    }
  }
#ifdef RC_TIMING
  if (prm.dbg != nullptr && blockIdx.x < 2 && lane == 0 && (warp == 0 || warp == 1 || warp == 4 || warp == 12)) {
    wt[0] = clock64() - t_start;
    const int role = warp == 0 ? 0 : (warp == 1 ? 1 : (warp == 4 ? 2 : 3));
#pragma unroll
    for (int i = 0; i < 16; ++i) prm.dbg[(blockIdx.x * 4 + role) * 16 + i] = wt[i];
  }
#endif
  tc_fence_before();
  cluster_sync();            // no CTA leaves while its peer may still touch its barriers / shared memory
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm<kTmemCols>(tmem);
  }
}

}  // namespace pair

bool infonce_pair_supported(int D) { return D == 256 || D == 512; }

// launch helper used by rc_infonce_bf16 (infonce_umma.cu owns argument checking, the pre-pass and text maps)
int launch_infonce_pair(const void* xsrc, void* dx, const void* t_bf16, const void* tt_bf16, int B, int D, int64_t HW, int K,
                        const float* inv_norm, const int32_t* y, const float* w, float inv_tau, const float* grad_scale,
                        const double* w_sum_in, float* lse, double* loss_sum, double* w_sum, double* dlogtau, cudaStream_t s) {
  using namespace pair;
  const bool bwd = dx != nullptr;
  const int Kp = (K + 63) / 64 * 64;
  CUtensorMap m_xs, m_t, m_tt, m_xe, m_dx;
  int rcode;
  {
    const uint64_t dims[3] = {(uint64_t)HW, (uint64_t)D, (uint64_t)B};
    const uint64_t str[3] = {2, (uint64_t)HW * 2, (uint64_t)D * HW * 2};
    const uint32_t box_s[3] = {64, 64, 1}, box_e[3] = {kStgPx, 128, 1};
    if ((rcode = make_tmap_bf16(&m_xs, xsrc, 3, dims, str, box_s, "pair map_x_s"))) return rcode;
    if ((rcode = make_tmap_bf16(&m_xe, xsrc, 3, dims, str, box_e, "pair map_x_e"))) return rcode;
    if ((rcode = make_tmap_bf16(&m_dx, bwd ? dx : xsrc, 3, dims, str, box_e, "pair map_dx"))) return rcode;
    const uint64_t tdims[2] = {(uint64_t)D, (uint64_t)Kp}, tstr[2] = {2, (uint64_t)D * 2};
    const uint32_t tbox[2] = {64, (uint32_t)(Kp / 2)};
    if ((rcode = make_tmap_bf16(&m_t, t_bf16, 2, tdims, tstr, tbox, "pair map_t"))) return rcode;
    const uint64_t ttdims[2] = {(uint64_t)Kp, (uint64_t)D}, ttstr[2] = {2, (uint64_t)Kp * 2};
    const uint32_t ttbox[2] = {64, 128};
    if ((rcode = make_tmap_bf16(&m_tt, bwd ? tt_bf16 : t_bf16, 2, bwd ? ttdims : tdims, bwd ? ttstr : tstr,
                                bwd ? ttbox : tbox, "pair map_tt"))) return rcode;
  }
  Params prm;
  prm.dbg = debug_timing_buffer();
  prm.B = B; prm.D = D; prm.K = K; prm.Kp = Kp; prm.HW = HW;
  prm.tiles_per_img = (int)((HW + kTilePx - 1) / kTilePx);
  if ((int64_t)B * prm.tiles_per_img > 0x3fffffff) return fail(RC_ERR_UNSUPPORTED, "rc_infonce_bf16: too many tiles");
  prm.n_tiles = B * prm.tiles_per_img;
  prm.n_pairs = (prm.n_tiles + 1) / 2;
  prm.inv_norm = inv_norm; prm.y = y; prm.w = w; prm.inv_tau = inv_tau; prm.grad_scale = grad_scale;
  prm.w_sum_in = w_sum_in; prm.lse = lse; prm.loss_sum = loss_sum; prm.w_sum = w_sum; prm.dlogtau = dlogtau;
  int n_clusters = num_sms() / 2;
  if (n_clusters > prm.n_pairs) n_clusters = prm.n_pairs;
  const int grid = 2 * n_clusters;
  cudaError_t e;
  if (bwd) {
    e = cudaFuncSetAttribute(infonce_umma_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return fail(RC_ERR_CUDA, "rc_infonce_bf16(pair): smem opt-in: %s", cudaGetErrorString(e));
    infonce_umma_pair_kernel<true><<<grid, kThreads, kSmemBytes, s>>>(m_xs, m_t, m_tt, m_xe, m_dx, prm);
  } else {
    e = cudaFuncSetAttribute(infonce_umma_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return fail(RC_ERR_CUDA, "rc_infonce_bf16(pair): smem opt-in: %s", cudaGetErrorString(e));
    infonce_umma_pair_kernel<false><<<grid, kThreads, kSmemBytes, s>>>(m_xs, m_t, m_tt, m_xe, m_dx, prm);
  }
  return check_launch("rc_infonce_bf16(pair)");
}

}  // namespace rc
