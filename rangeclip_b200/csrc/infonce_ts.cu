// K1/K2 with the softmax tile as a TENSOR-MEMORY operand (TS-mode tcgen05.mma) and TWO tile pairs in flight per CTA pair.
//
// Same problem as infonce_umma2.cu (fused pixel-text InfoNCE forward + backward, model.py:272-291 and its autograd).
// That kernel computes dX^T = T^T P^T with both operands in shared memory (SS mode) and keeps ONE S accumulator: the tensor
// pipe, the softmax warps and the dX epilogue wait for one another in a chain S(i) -> exp(i) -> P(i) -> dX(i), and the SS
// operand reads (96-128 B/clk) saturate the 128 B/clk shared-memory pipe.  Here
//   S  = X^T T^T    M = 256 px (128 per CTA; A = own X chunk [64 d][128 px], MN-major, shared memory), N = Kp,
//                   B = text rows split Kp/2 per CTA; fp32 accumulators in the 256 tensor-memory columns of a SLOT
//   P               the softmax threads overwrite the first 128 columns of the slot with the scaled softmax-minus-onehot
//                   tile as packed bf16 (tcgen05.st): P never touches shared memory
//   dX = P T        M = 256 px (the SAME TMEM lanes), A = P from tensor memory (TS mode), N = 64 channels per block,
//                   B = own 32 rows of T^T [D][Kp] (K-major, shared memory: 32 B/clk of operand reads); two 64-column fp32
//                   accumulators in the remaining 128 columns of the slot
// Tensor memory holds TWO slots (2 x 256 columns); tile pairs alternate between them.  MMA issue order
//   S(0) S(1) | dX(0) S(2) | dX(1) S(3) | dX(2) S(4) | ...
// so the exp pass of pair j+1 runs while the tensor pipe works on dX(j) and S(j+2) of the OTHER slot: no role waits for a
// hand-off it has just produced.  Pixels stay on the TMEM lanes for both GEMMs, so a CTA's softmax and dX epilogue work on
// its own tile only (no row-scale exchange between the CTAs).  The dX epilogue thread owns a pixel: 32 channels per step
// come out of TMEM, the x values of the same [32 ch][32 px] box arrive by TMA in a small per-warp ring (the box doubles as
// the staging tile of the TMA store), dx = acc - cs x is formed in place and the box goes back with one TMA store.
// The row norms 1/|x_p| (model.py:272 F.normalize) are computed by the relay warp from the X chunks in the operand ring.
#include "common.cuh"
#include "umma.cuh"
#include <float.h>
#include <stdlib.h>

namespace rc {
using namespace umma;

namespace ts {

#ifndef RC_TS_X_HINT
#define RC_TS_X_HINT 0      // A/B: evict_last on the X loads of the S GEMM
#endif

constexpr int kTilePx = 128;
constexpr int kThreads = 640;          // warps: 0 text TMA, 1 MMA (leader CTA), 2 relay + row norms + TMEM alloc, 3 X TMA, 4-11 softmax, 12-19 dX epilogue
constexpr int kXStages = 6;            // X ring: own X chunks [64 d][128 px]; deep, so that most of the next tile is resident early
constexpr int kTStages = 3;            // text ring: text half-chunks [Kp/2][64 d] for S, own T^T rows [32 ch][Kp] of a dX block
constexpr int kStageBytes = 16 * 1024;
constexpr int kEpiBufs = 4;            // per epilogue warp: ring of [32 ch][32 px] bf16 boxes (x in, dX out)
constexpr int kEpiAhead = 2;           // x boxes requested this many steps ahead
constexpr int kEpiBufBytes = 2048;
constexpr int kTmemCols = 512;
constexpr int kSlotCols = 256;         // one slot: S [0,256) -> P [0,128) + two dX accumulators [128,192), [192,256)
constexpr int kColAcc = 128, kAccCols = 64;
constexpr int kRegsCtl = 48, kRegsSoftmax = 120;     // epilogue warps keep the entry allocation (96); 48 + 2*120 + 2*96 = 5*96
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct __align__(8) Bars {
  uint64_t xf[kXStages];        // own X chunk has landed (CTA-local; relayed to the leader's xfull)
  uint64_t xfull[kXStages], xempty[kXStages];
  uint64_t tfull[kTStages], tempty[kTStages];
  uint64_t s_full[2], p_full[2];                  // per slot
  uint64_t acc_full[2][2], acc_empty[2][2];       // per slot, per accumulator
  uint64_t sc_full[4];                            // projection coefficients of a tile (by pair index & 3)
  uint64_t ebar[8][kEpiBufs];                     // x box of an epilogue warp has landed
  uint32_t tmem_base, pad;
};

constexpr int kScaleBufs = 4;          // the softmax warps run up to two tiles ahead of the dX epilogue warps
constexpr int kOffT = kXStages * kStageBytes;
constexpr int kOffEpi = kOffT + kTStages * kStageBytes;
constexpr int kOffScale = kOffEpi + 8 * kEpiBufs * kEpiBufBytes;      // -cs per pixel: [kScaleBufs][128] float
constexpr int kOffXch = kOffScale + kScaleBufs * 128 * 4;
constexpr int kOffPart = kOffXch + 2 * 4 * 2 * 128 * 4;               // exchange: [2 tile parities][max, sum, sez, sy][2 halves][128]
constexpr int kOffBars = kOffPart + 8 * 128 * 4;                      // row-norm partial sums of squares [8 softmax warps][128 px]
constexpr int kSmemBytes = kOffBars + (int)sizeof(Bars);
static_assert(kSmemBytes <= 232448, "shared-memory budget of one SM (227 KB)");

struct Params {
  int B, D, K, Kp;
  int64_t HW;
  int tiles_per_img, n_tiles, n_pairs;
  uint32_t tpi_magic;       // floor(2^32 / tiles_per_img)
  int keep_w;               // K-blocked launches, see infonce_umma2.cu
  int acc_dx;               // dX += this launch's gradient (TMA reduce-add store)
  const float* lse_in;
  int kb;
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

__device__ __forceinline__ int div_tiles(const Params& prm, int tile) {
  int q = (int)__umulhi((uint32_t)tile, prm.tpi_magic);
  if (tile - q * prm.tiles_per_img >= prm.tiles_per_img) ++q;
  return q;
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float select16(const uint32_t (&r)[16], int i) {
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = __uint_as_float((i & 1) ? r[2 * j + 1] : r[2 * j]);
#pragma unroll
  for (int j = 0; j < 4; ++j) a[j] = (i & 2) ? a[2 * j + 1] : a[2 * j];
#pragma unroll
  for (int j = 0; j < 2; ++j) a[j] = (i & 4) ? a[2 * j + 1] : a[2 * j];
  return (i & 8) ? a[1] : a[0];
}

// tile -> (image, first pixel); tiles past the end map to image index B (out of bounds for every tensor map)
__device__ __forceinline__ void tile_coords(const Params& prm, int tile, int& b, int& px0) {
  if (tile < prm.n_tiles) {
    b = div_tiles(prm, tile);
    px0 = (tile - b * prm.tiles_per_img) * kTilePx;
  } else {
    b = prm.B;
    px0 = 0;
  }
}

// R = targets per embedding row (1, or 4 for the shared 2x2 form); kKB: K-blocked launches (keep_w / lse_in / kb)
template <int R, bool kKB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
infonce_ts_kernel(const __grid_constant__ CUtensorMap map_x_s,   // X [B][D][HW], box (64 px, 64 d, 1), 128B swizzle
                  const __grid_constant__ CUtensorMap map_t,     // T [Kp][D],    box (64 d, Kp/2 rows)
                  const __grid_constant__ CUtensorMap map_tt,    // T^T [D][Kp],  box (64 k, 32 d)
                  const __grid_constant__ CUtensorMap map_xe,    // X [B][D][HW], box (32 px, 32 d, 1), 64B swizzle
                  const __grid_constant__ CUtensorMap map_dx,    // dX [B][D][HW], box (32 px, 32 d, 1), 64B swizzle
                  const Params prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Bars* bars = reinterpret_cast<Bars*>(smem + kOffBars);
  float* sc_s = reinterpret_cast<float*>(smem + kOffScale);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader_cta = rank == 0;
  const int n_dchunks = prm.D / 64;      // 64-channel chunks of the S GEMM
  const int n_blk = prm.D / 64;          // 64-channel blocks of the dX GEMM
  const int n_kchunks = prm.Kp / 64;
  const int Nh = prm.Kp / 2;             // text rows staged by each CTA
  const int n_clusters = gridDim.x / 2;
  const int cluster_id = blockIdx.x / 2;
  const int my_pairs = cluster_id < prm.n_pairs ? (prm.n_pairs - cluster_id + n_clusters - 1) / n_clusters : 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_x_s); tma_prefetch_desc(&map_t); tma_prefetch_desc(&map_tt);
    tma_prefetch_desc(&map_xe); tma_prefetch_desc(&map_dx);
    // X ring: xfull = the relays of both CTAs, xempty = MMA commit + the eight softmax warps (row norms)
    for (int i = 0; i < kXStages; ++i) { mbar_init(&bars->xf[i], 1); mbar_init(&bars->xfull[i], 2); mbar_init(&bars->xempty[i], 9); }
    for (int i = 0; i < kTStages; ++i) { mbar_init(&bars->tfull[i], 1); mbar_init(&bars->tempty[i], 1); }
    for (int sl = 0; sl < 2; ++sl) {
      mbar_init(&bars->s_full[sl], 1); mbar_init(&bars->p_full[sl], 16);     // consumer releases: one arrival per warp
      for (int ab = 0; ab < 2; ++ab) { mbar_init(&bars->acc_full[sl][ab], 1); mbar_init(&bars->acc_empty[sl][ab], 16); }
    }
    for (int i = 0; i < kScaleBufs; ++i) mbar_init(&bars->sc_full[i], 4);
    for (int i = 0; i < 8; ++i)
      for (int j = 0; j < kEpiBufs; ++j) mbar_init(&bars->ebar[i][j], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm<kTmemCols>(&bars->tmem_base);
  tc_fence_before();
  cluster_sync();            // both CTAs' barriers are initialised before any remote arrive / 2-SM TMA credit
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const uint32_t idesc_s = make_idesc_bf16(256, prm.Kp, /*A MN-major*/ 1, /*B K-major*/ 0);
  const uint32_t idesc_d = make_idesc_bf16(256, kAccCols, 0, 0);
  auto arrive_leader = [&](uint64_t* bar) {
    if (leader_cta) mbar_arrive(bar);
    else mbar_arrive_remote(map_to_cta(bar, 0));
  };
  auto arrive_leader_warp = [&](uint64_t* bar) {
    __syncwarp();
    if (elect_one()) arrive_leader(bar);
  };
  // l-th tile pair of this cluster
  auto pair_of = [&](int l) -> int { return cluster_id + l * n_clusters; };
  auto koff_of = [&](int pj) -> int { return (kKB && prm.kb > 0) ? div_tiles(prm, 2 * pj) * 256 : 0; };

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsCtl));
    if (warp == 0 && lane == 0) {
      // =============================== text producer (both CTAs) ===============================
      // ring order == MMA issue order: S(0) S(1) | dX(0) S(2) | dX(1) S(3) | ...
      uint32_t it = 0;
      auto load_s = [&](int l) {
        const int koff = koff_of(pair_of(l));
        for (int c = 0; c < n_dchunks; ++c, ++it) {   // own half (Nh rows) of text chunk c
          const int st = it % kTStages;
          mbar_wait(&bars->tempty[st], ((it / kTStages) & 1) ^ 1, 1);
          if (leader_cta) mbar_arrive_expect_tx(&bars->tfull[st], 2 * Nh * 128);
          tma_load_2d_2sm(smem + kOffT + st * kStageBytes, &map_t, &bars->tfull[st], c * 64, koff + (int)rank * Nh);
        }
      };
      auto load_dx = [&](int l) {          // per 64-channel block: own 32 rows of T^T, all Kp columns (4 KB per 64 k)
        const int koff = koff_of(pair_of(l));
        for (int blk = 0; blk < n_blk; ++blk, ++it) {
          const int st = it % kTStages;
          mbar_wait(&bars->tempty[st], ((it / kTStages) & 1) ^ 1, 2);
          if (leader_cta) mbar_arrive_expect_tx(&bars->tfull[st], 2 * n_kchunks * 4096);
          for (int kc = 0; kc < n_kchunks; ++kc)
            tma_load_2d_2sm(smem + kOffT + st * kStageBytes + kc * 4096, &map_tt, &bars->tfull[st], koff + kc * 64,
                            blk * 64 + (int)rank * 32);
        }
      };
      if (my_pairs > 0) load_s(0);
      if (my_pairs > 1) load_s(1);
      for (int l = 0; l < my_pairs; ++l) {
        load_dx(l);
        if (l + 2 < my_pairs) load_s(l + 2);
      }
    } else if (warp == 3 && lane == 0) {
      // =============================== X producer (both CTAs) ===============================
      uint32_t xit = 0;
      const uint64_t pol_keep = RC_TS_X_HINT ? l2_policy_evict_last() : 0;
      for (int l = 0; l < my_pairs; ++l) {
        int b, px0;
        tile_coords(prm, 2 * pair_of(l) + (int)rank, b, px0);
        for (int c = 0; c < n_dchunks; ++c, ++xit) {
          const int st = xit % kXStages;
          mbar_wait(&bars->xempty[st], ((xit / kXStages) & 1) ^ 1, 3);
          uint8_t* sb = smem + st * kStageBytes;
          mbar_arrive_expect_tx(&bars->xf[st], 2 * 8192);          // CTA-local: the row-norm warp reads the chunk too
          const int bx = (kKB && prm.kb > 0) ? (b < prm.B ? 0 : 1) : b;      // kb mode: the one image of X (1 = out of bounds)
          if (RC_TS_X_HINT) {       // the tile is read again by the dX epilogue two iterations later: keep it in the L2
            tma_load_3d_hint(sb, &map_x_s, &bars->xf[st], px0, c * 64, bx, pol_keep);
            tma_load_3d_hint(sb + 8192, &map_x_s, &bars->xf[st], px0 + 64, c * 64, bx, pol_keep);
          } else {
            tma_load_3d(sb, &map_x_s, &bars->xf[st], px0, c * 64, bx);
            tma_load_3d(sb + 8192, &map_x_s, &bars->xf[st], px0 + 64, c * 64, bx);
          }
        }
      }
    } else if (warp == 1 && leader_cta) {
      // =============================== MMA issuer (leader CTA) ================================
      // whole warp converged (addresses / descriptors in uniform registers), one elected lane issues
      uint32_t it = 0, xit = 0;
      uint32_t acc_uses[2][2] = {{0, 0}, {0, 0}};      // per slot, per accumulator: launches so far
      const uint32_t smem_base = smem_u32(smem);
      const uint64_t dsc_x = desc_mnmajor_sw128(0, 8192);
      const uint64_t dsc_k = desc_kmajor_sw128(0);
      // the slot's accumulators have been drained by the epilogue warps (all launches so far)
      auto wait_acc_free = [&](int sl, int ab) {
        mbar_wait(&bars->acc_empty[sl][ab], (acc_uses[sl][ab] & 1u) ^ 1u, 7);
      };
      auto issue_s = [&](int l) {
        const int sl = l & 1;
        const uint32_t s_tmem = tmem + sl * kSlotCols;
        // S overwrites the slot: its P has been consumed (the dX MMAs were issued before, the pipe runs in order) and
        // both dX accumulators must have been read out
        wait_acc_free(sl, 0);
        wait_acc_free(sl, 1);
        tc_fence_after();
        for (int c = 0; c < n_dchunks; ++c, ++it, ++xit) {
          const int sa = xit % kXStages, sb_ = it % kTStages;
          mbar_wait(&bars->xfull[sa], (xit / kXStages) & 1, 5);
          mbar_wait(&bars->tfull[sb_], (it / kTStages) & 1, 6);
          tc_fence_after();
          const uint64_t xa = dsc_x + ((smem_base + sa * kStageBytes) >> 4);
          const uint64_t tb = dsc_k + ((smem_base + kOffT + sb_ * kStageBytes) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma_bf16_ss_2sm(s_tmem, xa + ((ks * 2048) >> 4), tb + ((ks * 32) >> 4), idesc_s, (c | ks) != 0);
            mma_commit_2sm(&bars->xempty[sa]);
            mma_commit_2sm(&bars->tempty[sb_]);
            if (c + 1 == n_dchunks) mma_commit_2sm(&bars->s_full[sl]);
          }
          __syncwarp();
        }
      };
      auto issue_dx = [&](int l) {
        const int sl = l & 1;
        const uint32_t base = tmem + sl * kSlotCols;
        mbar_wait(&bars->p_full[sl], (l >> 1) & 1, 9);           // P(l) is in tensor memory
        tc_fence_after();
        for (int blk = 0; blk < n_blk; ++blk, ++it) {
          const int ab = blk & 1;
          wait_acc_free(sl, ab);
          const int st = it % kTStages;
          mbar_wait(&bars->tfull[st], (it / kTStages) & 1, 8);
          tc_fence_after();
          const uint64_t sb = dsc_k + ((smem_base + kOffT + st * kStageBytes) >> 4);
          const uint32_t dcol = base + kColAcc + ab * kAccCols;
          if (elect_one()) {
            for (int kc = 0; kc < n_kchunks; ++kc) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)     // A: P [256 px][16 k] in TMEM (8 columns); B: own T^T rows [32 ch][16 k]
                mma_bf16_ts_2sm(dcol, base + (kc * 4 + ks) * 8, sb + ((kc * 4096 + ks * 32) >> 4), idesc_d, (kc | ks) != 0);
            }
            mma_commit_2sm(&bars->tempty[st]);
            mma_commit_2sm(&bars->acc_full[sl][ab]);
          }
          __syncwarp();
          ++acc_uses[sl][ab];
        }
      };
      if (my_pairs > 0) issue_s(0);
      if (my_pairs > 1) issue_s(1);
      for (int l = 0; l < my_pairs; ++l) {
        issue_dx(l);
        if (l + 2 < my_pairs) issue_s(l + 2);
      }
    } else if (warp == 2 && lane == 0) {
      // ========== relay (both CTAs): "own X chunk has landed" (CTA-local xf) -> the leader's full barrier ==========
      uint32_t xit = 0;
      for (int l = 0; l < my_pairs; ++l)
        for (int c = 0; c < n_dchunks; ++c, ++xit) {
          const int st = xit % kXStages;
          mbar_wait(&bars->xf[st], (xit / kXStages) & 1, 10);
          arrive_leader(&bars->xfull[st]);
        }
    }
  } else if (warp < 12) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsSoftmax));
    // ======================= softmax / CE warps: own tile (two warps per TMEM lane quarter) =======================
    const int half = warp >= 8 ? 1 : 0;
    const int row = (warp & 3) * 32 + lane;                 // pixel of the own tile == TMEM lane
    const int Kh = prm.Kp >> 1;
    const int cb = half * Kh;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float* xch_base = reinterpret_cast<float*>(smem + kOffXch);
    float loss_acc = 0.f, w_acc = 0.f, dlt_acc = 0.f;
    float inv_wsum = 0.f, gscale = 1.f;
    {
      const double ws = prm.w_sum_in[0];
      inv_wsum = ws > 0.0 ? (float)(1.0 / ws) : 0.f;
      if (prm.grad_scale) gscale = prm.grad_scale[0];
    }
    float nx_ok = 0.f, nx_w = 0.f;
    int nx_y = -1;
    auto load_pixel_scalars = [&](int l) {
      nx_ok = 0.f; nx_w = 0.f; nx_y = -1;
      if (l >= my_pairs) return;
      const int t = 2 * pair_of(l) + (int)rank;
      if (t < prm.n_tiles) {
        const int tb = div_tiles(prm, t);
        const int tpx = (t - tb * prm.tiles_per_img) * kTilePx + row;
        if (tpx < prm.HW) {
          const int64_t tm = (int64_t)tb * prm.HW + tpx;
          nx_ok = 1.f;
          if (R == 1) {
            nx_y = __ldg(prm.y + tm);
            nx_w = __ldg(prm.w + tm);
          } else {               // four targets, 8 bits each (K <= 256)
            const int4 y4 = __ldg(reinterpret_cast<const int4*>(prm.y) + tm);
            nx_y = (y4.x & 255) | ((y4.y & 255) << 8) | ((y4.z & 255) << 16) | ((y4.w & 255) << 24);
          }
        }
      }
    };
    load_pixel_scalars(0);
    // Row norms 1/|x_p| (model.py:272 F.normalize) of a tile from its X chunks where they sit in the operand ring
    // ([64 d][2 x 64 px], 128-byte swizzle): thread = (8-pixel group, row phase); sums of squares meet in shared memory.
    // The X ring holds most of a tile, so the chunks are normally resident when this runs (right before the tile's exp pass).
    float* part_s = reinterpret_cast<float*>(smem + kOffPart);   // [8 warps][128 px] partial sums of squares
    const int st_ = threadIdx.x - 128;                           // 0..255
    const int ng = st_ & 15, nr = st_ >> 4;
    uint32_t nit = 0;
    auto norm_tile = [&]() -> float {
      float ss[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) ss[j] = 0.f;
      for (int c = 0; c < n_dchunks; ++c, ++nit) {
        const int st = nit % kXStages;
        mbar_wait(&bars->xf[st], (nit / kXStages) & 1, 11);
        const uint8_t* base = smem + st * kStageBytes + (ng >> 3) * 8192 + nr * 128;
        const int ch = ng & 7;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const int rowd = nr + rr * 16;
          const uint4 v = *reinterpret_cast<const uint4*>(base + rr * 2048 + ((ch ^ (rowd & 7)) << 4));
          const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {       // FHFMA.BF16 on the word's halves: no unpack instructions
            ss[2 * j] = sqacc_bf16x2_lo(ss[2 * j], u[j]);
            ss[2 * j + 1] = sqacc_bf16x2_hi(ss[2 * j + 1], u[j]);
          }
        }
        __syncwarp();
        if (elect_one()) mbar_arrive(&bars->xempty[st]);
      }
      // fixed-order reduction (bit-reproducible): lane pairs, then the eight warps' partials through shared memory
#pragma unroll
      for (int j = 0; j < 8; ++j) ss[j] += __shfl_xor_sync(0xffffffffu, ss[j], 16);
      if (lane < 16) {
        float4* dst = reinterpret_cast<float4*>(part_s + (warp - 4) * 128 + ng * 8);
        dst[0] = make_float4(ss[0], ss[1], ss[2], ss[3]);
        dst[1] = make_float4(ss[4], ss[5], ss[6], ss[7]);
      }
      named_bar_sync(5, 256);
      float q = 0.f;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) q += part_s[wv * 128 + row];
      return 1.f / fmaxf(sqrtf(q), 1e-12f);
    };
    const bool use_bound = prm.inv_tau * (2.02f * kLog2e) < 100.f;
    const float ml_bound = prm.inv_tau * (1.01f * kLog2e);
    for (int l = 0; l < my_pairs; ++l) {
      const int sl = l & 1;
      const int tile = 2 * pair_of(l) + (int)rank;
      const bool tile_ok = tile < prm.n_tiles;
      const int b = tile_ok ? div_tiles(prm, tile) : 0;
      const int px = tile_ok ? (tile - b * prm.tiles_per_img) * kTilePx + row : 0;
      const bool valid = tile_ok && px < prm.HW;
      const int64_t m = (int64_t)b * prm.HW + px;
      const int Kt = (kKB && prm.kb > 0) ? min(256, prm.K - 256 * b) : prm.K;
      const bool px_ok = nx_ok != 0.f;
      const int yi = nx_y;
      const float wi = (R == 1 && (yi >= 0 || (kKB && prm.keep_w))) ? nx_w : 0.f;
      load_pixel_scalars(l + 1);
      float* xch = xch_base + (l & 1) * (4 * 2 * 128);
      const uint32_t trow = lane_base + sl * kSlotCols + cb;
      const float inv_n_tile = norm_tile();
      const float inv_n = px_ok ? inv_n_tile : 0.f;
      const float zs = inv_n * prm.inv_tau;
      const float zl = zs * kLog2e;
      mbar_wait(&bars->s_full[sl], (l >> 1) & 1, 12);
      tc_fence_after();
      float ml = ml_bound;
      if (!use_bound) {
        float mx = -FLT_MAX;
        for (int c = 0; c * 32 < Kh; ++c) {
          const int nvalid = Kt - (cb + c * 32);
          if (nvalid <= 0) break;
          uint32_t r[32];
          tmem_ld_32x32(trow + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) if (i < nvalid) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
        xch[(0 * 2 + half) * 128 + row] = mx;
        named_bar_sync(3, 256);
        mx = fmaxf(mx, xch[(0 * 2 + (half ^ 1)) * 128 + row]);
        ml = mx * zl;
      }
      uint32_t pk[64];
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
      float sy[R];
#pragma unroll
      for (int j = 0; j < R; ++j) sy[j] = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c * 16 < Kh) {
          uint32_t r[16];
          tmem_ld_32x16(trow + c * 16, r);
          tmem_ld_wait();
          const int k0 = cb + c * 16;
          const int nvalid = Kt - k0;
#pragma unroll
          for (int j = 0; j < R; ++j) {                          // target logit(s): once per row, not per column
            const int yrel = (R == 1 ? yi : (int)(((uint32_t)yi >> (8 * j)) & 255u)) - k0;
            if ((unsigned)yrel < 16u) sy[j] = select16(r, yrel);
          }
          if (nvalid >= 16) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float a0 = __uint_as_float(r[i]), a1 = __uint_as_float(r[i + 1]);
              const float a2 = __uint_as_float(r[i + 2]), a3 = __uint_as_float(r[i + 3]);
              const float e0 = fast_exp2(fmaf(a0, zl, -ml)), e1 = fast_exp2(fmaf(a1, zl, -ml));
              const float e2 = fast_exp2(fmaf(a2, zl, -ml)), e3 = fast_exp2(fmaf(a3, zl, -ml));
              s0 += e0; s1 += e1; s2 += e2; s3 += e3;
              q0 = fmaf(e0, a0, q0); q1 = fmaf(e1, a1, q1); q2 = fmaf(e2, a2, q2); q3 = fmaf(e3, a3, q3);
              pk[c * 8 + (i >> 1)] = pack_bf16x2(e0, e1);
              pk[c * 8 + (i >> 1) + 1] = pack_bf16x2(e2, e3);
            }
          } else {          // the step that straddles K, or padding columns only
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              const float a0 = __uint_as_float(r[i]), a1 = __uint_as_float(r[i + 1]);
              const float e0 = (i < nvalid) ? fast_exp2(fmaf(a0, zl, -ml)) : 0.f;
              const float e1 = (i + 1 < nvalid) ? fast_exp2(fmaf(a1, zl, -ml)) : 0.f;
              s0 += e0; s1 += e1;
              q0 = fmaf(e0, a0, q0); q1 = fmaf(e1, a1, q1);
              pk[c * 8 + (i >> 1)] = pack_bf16x2(e0, e1);
            }
          }
        }
      }
      float sum = (s0 + s1) + (s2 + s3);
      float sez = (q0 + q1) + (q2 + q3);
      int yj[R];
      float wj[R];
      if (R == 1) {
        yj[0] = yi; wj[0] = wi;
      } else {
        int4 y4 = make_int4(-1, -1, -1, -1);
        float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (px_ok) {
          y4 = __ldg(reinterpret_cast<const int4*>(prm.y) + m);
          w4 = __ldg(reinterpret_cast<const float4*>(prm.w) + m);
        }
        const int ya[4] = {y4.x, y4.y, y4.z, y4.w};
        const float wa[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int j = 0; j < R; ++j) { yj[j] = ya[j]; wj[j] = ya[j] >= 0 ? wa[j] : 0.f; }
      }
      float wtot = 0.f, tz = 0.f;
      bool mine[R];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        mine[j] = yj[j] >= cb && yj[j] < cb + Kh;
        wtot += wj[j];
        tz += mine[j] ? wj[j] * sy[j] : 0.f;
      }
      xch[(1 * 2 + half) * 128 + row] = sum;
      xch[(2 * 2 + half) * 128 + row] = sez;
      xch[(3 * 2 + half) * 128 + row] = tz;
      // every softmax warp has read its S columns (tcgen05.wait::ld above): after this barrier the slot's columns may be
      // overwritten with P
      tc_fence_before();
      named_bar_sync(2, 256);
      tc_fence_after();
      sum += xch[(1 * 2 + (half ^ 1)) * 128 + row];
      sez += xch[(2 * 2 + (half ^ 1)) * 128 + row];
      tz += xch[(3 * 2 + (half ^ 1)) * 128 + row];
      if (kKB && prm.lse_in != nullptr && valid) sum = fast_exp2(fmaf(__ldg(prm.lse_in + (prm.kb > 0 ? (int64_t)px : m)), kLog2e, -ml));   // 1 / sum = exp(m - lse)
      float lse = 0.f;
      if (half == 0) {
        lse = (ml + __log2f(sum)) * kLn2;
        loss_acc += wtot * lse - tz * zs;
        w_acc += wtot;
      }
      const float coefb = gscale * inv_wsum;
      const float coef = coefb * wtot;
      const float inv_sum = 1.f / sum;
      // P is stored pre-scaled: G[p][k] = rs_p (e_pk - sum_p [k = y_p]) = d loss / d(xhat_p . that_k) / |x_p|
      const float rsv = inv_n * prm.inv_tau * coef * inv_sum;
      if (half == 0) {
        const float cj = coef * sez * zs * inv_sum - coefb * tz * zs;
        sc_s[(l & (kScaleBufs - 1)) * 128 + row] = -(inv_n * inv_n * cj);      // dx = acc + (-cs) x
        dlt_acc -= cj;
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->sc_full[l & (kScaleBufs - 1)]);
      }
      {
        const uint32_t prow = lane_base + sl * kSlotCols + (cb >> 1);
        const uint32_t rs2 = pack_bf16x2(rsv, rsv);
        // G[row][y_j] = rs (e_y - sum * (weight of target y_j) / (weight of the row)), formed in fp32 before rounding
        uint32_t g16[R];
        int grel[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
          grel[j] = -1;
          g16[j] = 0;
          if (mine[j]) {
            const float ey = fast_exp2(fmaf(sy[j], zl, -ml));
            float gv;
            if (R == 1) {
              gv = (ey - sum) * rsv;
            } else {
              float wk = 0.f;
#pragma unroll
              for (int i = 0; i < R; ++i) wk += (yj[i] == yj[j]) ? wj[i] : 0.f;
              gv = (ey - sum * (wtot > 0.f ? wk / wtot : 0.f)) * rsv;
            }
            grel[j] = yj[j] - cb;
            g16[j] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(gv));
          }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c * 32 < Kh) {
            uint32_t v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = bf2_mul(pk[c * 16 + i], rs2);
#pragma unroll
            for (int j = 0; j < R; ++j) {
              if (grel[j] >= 0 && (grel[j] >> 5) == c) {
                const int jj = (grel[j] >> 1) & 15;
                const bool hi = grel[j] & 1;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (i == jj) v[i] = hi ? ((v[i] & 0xffffu) | (g16[j] << 16)) : ((v[i] & 0xffff0000u) | g16[j]);
              }
            }
            tmem_st_32x16(prow + c * 16, v);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        arrive_leader_warp(&bars->p_full[sl]);
      }
      if (half == 0 && valid && prm.lse && !(kKB && prm.lse_in != nullptr)) prm.lse[m] = lse;
    }
    if (half == 0) {
      loss_acc = warp_sum(loss_acc); w_acc = warp_sum(w_acc); dlt_acc = warp_sum(dlt_acc);
      if (lane == 0) {
        if (prm.loss_sum) atomicAdd(prm.loss_sum, (double)loss_acc);
        if (prm.w_sum) atomicAdd(prm.w_sum, (double)w_acc);
        if (prm.dlogtau) atomicAdd(prm.dlogtau, (double)dlt_acc);
      }
    }
  } else {
    // ================ dX epilogue warps: own tile, thread = pixel, one 64-channel block = one step of 32 channels ================
    // warp (q, h): TMEM lane quarter q = pixels [32 q, +32) of the tile, accumulator columns [32 h, +32) of every block
    const int ew = warp - 12;
    const int q = warp & 3, h = ew >> 2;
    uint8_t* ebuf = smem + kOffEpi + ew * (kEpiBufs * kEpiBufBytes);
    uint64_t* ebar = &bars->ebar[ew][0];
    const int total = my_pairs * n_blk;
    const uint64_t pol_first = l2_policy_evict_first();
    // cursor over (pair, block): coordinates of a step's box = (pixel, channel, image of x, image of dX)
    struct Cursor { int l, u, px, bo, bx; };
    auto cursor_pair = [&](Cursor& cu) {
      int b, px0;
      tile_coords(prm, 2 * pair_of(cu.l) + (int)rank, b, px0);
      cu.px = px0 + q * 32;
      cu.bo = b;
      cu.bx = (kKB && prm.kb > 0) ? (b < prm.B ? 0 : 1) : b;
    };
    auto cursor_next = [&](Cursor& cu) {
      if (++cu.u == n_blk) { cu.u = 0; ++cu.l; cursor_pair(cu); }
    };
    Cursor cs{0, 0, 0, 0, 0}, cr{0, 0, 0, 0, 0};      // step being stored / step being requested
    cursor_pair(cs);
    cursor_pair(cr);
    int s_req = 0;
    auto request_x = [&]() {                         // lane 0: the x box of step s_req
      const int bi = s_req % kEpiBufs;
      mbar_arrive_expect_tx(&ebar[bi], kEpiBufBytes);
      tma_load_3d(ebuf + bi * kEpiBufBytes, &map_xe, &ebar[bi], cr.px, cr.u * 64 + h * 32, cr.bx);
    };
    for (; s_req < kEpiAhead && s_req < total; ++s_req) {
      if (lane == 0) request_x();
      cursor_next(cr);
    }
    // The step works on 8x8 b16 matrices: tcgen05.ld.16x256b hands out the accumulators in the mma fragment layout (thread t:
    // pixel t/4 (+8), channel pair 2(t%4)), ldmatrix.trans reads x from the [32 ch][32 px] box (64-byte swizzle) in the SAME
    // layout and stmatrix.trans writes dX back over it -- ~40 instructions per thread and step instead of ~150 two-byte ones.
    // matrix (g, pb): channels [8g, +8) x pixels [8 pb, +8); one ldmatrix / stmatrix .x4 = the four pixel blocks of one g;
    // thread t addresses row (channel) 8g + (t & 7) of pixel block t >> 3
    const uint32_t moff = (uint32_t)((lane & 7) * 64 + ((((lane >> 3) ^ ((lane >> 1) & 3)) & 3) << 4));
    int s = 0;
    uint32_t acc_seen[2][2] = {{0, 0}, {0, 0}};
    for (int l = 0; l < my_pairs; ++l) {
      const int sl = l & 1;
      mbar_wait(&bars->sc_full[l & (kScaleBufs - 1)], (l >> 2) & 1, 14);
      uint32_t ncs2[4];                                // -cs of the thread's pixel in each of the four pixel blocks
#pragma unroll
      for (int pb = 0; pb < 4; ++pb) {
        const float v = sc_s[(l & (kScaleBufs - 1)) * 128 + q * 32 + pb * 8 + (lane >> 2)];
        ncs2[pb] = pack_bf16x2(v, v);
      }
      const uint32_t tacc = tmem + ((uint32_t)(q * 32) << 16) + sl * kSlotCols + kColAcc + h * 32;
      for (int blk = 0; blk < n_blk; ++blk, ++s) {
        const int ab = blk & 1;
        mbar_wait(&bars->acc_full[sl][ab], acc_seen[sl][ab] & 1u, 15);
        ++acc_seen[sl][ab];
        tc_fence_after();
        // accumulators of pixels [0,16) and [16,32) of the warp's lane quarter, 32 channel columns each, rounded to packed
        // bf16 channel pairs; the accumulator is handed back before any of it is processed
        uint32_t pa[4][4];                             // [g][pb]
        {
          uint32_t r[16];
          tmem_ld_16x256b_x4(tacc + ab * kAccCols, r);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            pa[g][0] = pack_bf16x2(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]));
            pa[g][1] = pack_bf16x2(__uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3]));
          }
          tmem_ld_16x256b_x4(tacc + ab * kAccCols + (16u << 16), r);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            pa[g][2] = pack_bf16x2(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]));
            pa[g][3] = pack_bf16x2(__uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3]));
          }
        }
        tc_fence_before();
        arrive_leader_warp(&bars->acc_empty[sl][ab]);
        const int bi = s % kEpiBufs;
        mbar_wait(&ebar[bi], (s / kEpiBufs) & 1, 16);
        // dx = acc - cs x in place on the box (same rounding as infonce_umma2.cu: the accumulator is rounded to bf16, then
        // one fused multiply-add on packed pairs)
        const uint32_t bx0 = smem_u32(ebuf + bi * kEpiBufBytes) + moff;
        uint32_t xm[4][4];
#pragma unroll
        for (int g = 0; g < 4; ++g) ldmatrix_x4_trans(bx0 + g * 512, xm[g]);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
          for (int pb = 0; pb < 4; ++pb) xm[g][pb] = bf2_fma(ncs2[pb], xm[g][pb], pa[g][pb]);
          stmatrix_x4_trans(bx0 + g * 512, xm[g]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (kKB && prm.acc_dx) tma_reduce_add_3d(&map_dx, ebuf + bi * kEpiBufBytes, cs.px, cs.u * 64 + h * 32, cs.bo);
          else tma_store_3d_hint(&map_dx, ebuf + bi * kEpiBufBytes, cs.px, cs.u * 64 + h * 32, cs.bo, pol_first);
          tma_store_commit();
          if (s_req < total) {
            // the box to refill was stored kEpiBufs - kEpiAhead steps ago: that store must have read it
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kEpiBufs - kEpiAhead) : "memory");
            request_x();
          }
        }
        cursor_next(cs);
        if (s_req < total) { cursor_next(cr); ++s_req; }
        __syncwarp();
      }
    }
    if (lane == 0) tma_store_wait_all0();       // shared memory must stay valid until the last bulk store has read it
  }
  tc_fence_before();
  cluster_sync();            // no CTA leaves while its peer may still touch its barriers / shared memory
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm<kTmemCols>(tmem);
  }
}

}  // namespace ts

// launch helper used by rc_infonce_bf16 for backward launches without dText (infonce_umma.cu owns argument checking)
int launch_infonce_ts(const void* xsrc, void* dx, const void* t_bf16, const void* tt_bf16, int B, int D, int64_t HW, int K,
                      const int32_t* y, const float* w, float inv_tau, const float* grad_scale, const double* w_sum_in,
                      float* lse, double* loss_sum, double* w_sum, double* dlogtau, int rep, int keep_w, const float* lse_in,
                      int kb, int acc_dx, cudaStream_t s) {
  using namespace ts;
  const int Kp = kb > 0 ? 256 : (K + 63) / 64 * 64;
  const int Kall = kb > 0 ? kb * 256 : Kp;          // rows of the text matrices
  CUtensorMap m_xs, m_t, m_tt, m_xe, m_dx;
  int rcode;
  {
    const uint64_t dims[3] = {(uint64_t)HW, (uint64_t)D, (uint64_t)B};
    const uint64_t xdims[3] = {(uint64_t)HW, (uint64_t)D, (uint64_t)(kb > 0 ? 1 : B)};
    const uint64_t str[3] = {2, (uint64_t)HW * 2, (uint64_t)D * HW * 2};
    const uint32_t box_s[3] = {64, 64, 1};
    if ((rcode = make_tmap_bf16(&m_xs, xsrc, 3, xdims, str, box_s, "ts map_x_s"))) return rcode;
    const uint32_t box_o[3] = {32, 32, 1};
    if ((rcode = make_tmap_bf16(&m_xe, xsrc, 3, xdims, str, box_o, "ts map_xe"))) return rcode;
    if ((rcode = make_tmap_bf16(&m_dx, dx, 3, dims, str, box_o, "ts map_dx"))) return rcode;
    const uint64_t tdims[2] = {(uint64_t)D, (uint64_t)Kall}, tstr[2] = {2, (uint64_t)D * 2};
    const uint32_t tbox[2] = {64, (uint32_t)(Kp / 2)};
    if ((rcode = make_tmap_bf16(&m_t, t_bf16, 2, tdims, tstr, tbox, "ts map_t"))) return rcode;
    const uint64_t ttdims[2] = {(uint64_t)Kall, (uint64_t)D}, ttstr[2] = {2, (uint64_t)Kall * 2};
    const uint32_t ttbox[2] = {64, 32};
    if ((rcode = make_tmap_bf16(&m_tt, tt_bf16, 2, ttdims, ttstr, ttbox, "ts map_tt"))) return rcode;
  }
  Params prm;
  prm.B = B; prm.D = D; prm.K = K; prm.Kp = Kp; prm.HW = HW;
  prm.tiles_per_img = (int)((HW + kTilePx - 1) / kTilePx);
  if ((int64_t)B * prm.tiles_per_img > 0x3fffffff) return fail(RC_ERR_UNSUPPORTED, "rc_infonce_bf16: too many tiles");
  prm.n_tiles = B * prm.tiles_per_img;
  prm.tpi_magic = prm.tiles_per_img == 1 ? 0xffffffffu : (uint32_t)(0x100000000ull / (uint64_t)prm.tiles_per_img);
  prm.n_pairs = (prm.n_tiles + 1) / 2;
  prm.y = y; prm.w = w; prm.inv_tau = inv_tau; prm.grad_scale = grad_scale;
  prm.w_sum_in = w_sum_in; prm.lse = lse; prm.loss_sum = loss_sum; prm.w_sum = w_sum; prm.dlogtau = dlogtau;
  prm.keep_w = keep_w; prm.lse_in = lse_in; prm.kb = kb; prm.acc_dx = acc_dx;
  if (kb > 0 && (prm.tiles_per_img & 1)) return fail(RC_ERR_UNSUPPORTED, "rc_infonce_bf16_kblocks: HW must be a multiple of 256");
  int n_clusters = num_sms() / 2;
  if (n_clusters > prm.n_pairs) n_clusters = prm.n_pairs;
  const int grid = 2 * n_clusters;
  auto launch = [&](auto kernel) -> int {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return fail(RC_ERR_CUDA, "rc_infonce_bf16(ts): smem opt-in: %s", cudaGetErrorString(e));
    kernel<<<grid, kThreads, kSmemBytes, s>>>(m_xs, m_t, m_tt, m_xe, m_dx, prm);
    return check_launch("rc_infonce_bf16(ts)");
  };
  if (rep == 4) return launch(infonce_ts_kernel<4, false>);
  if (keep_w || lse_in != nullptr || kb > 0 || acc_dx) return launch(infonce_ts_kernel<1, true>);
  return launch(infonce_ts_kernel<1, false>);
}

}  // namespace rc

// ------------------------------------------------------------------------------------------------
// bring-up kernel for the TS-mode building blocks: C[256][N] = A[256][Kd] B[N][Kd]^T with A written to tensor memory by
// the threads (tcgen05.st, packed bf16 pairs) and B rows split N/2 / N/2 over the two CTAs' shared memory
// ------------------------------------------------------------------------------------------------
namespace rc {

struct __align__(8) DebugTsBars { uint64_t full, done; uint32_t tmem_base, pad; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
debug_umma_gemm_ts_2sm_kernel(const __grid_constant__ CUtensorMap map_b, const __nv_bfloat16* __restrict__ a, int N, int Kd,
                              float* __restrict__ c) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sbm = smem;              // Kd/64 chunks of own N/2 rows of B (<= 4 x 16 KB)
  DebugTsBars* bars = reinterpret_cast<DebugTsBars*>(smem + 65536);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int nh = N / 2;
  if (threadIdx.x == 0) {
    mbar_init(&bars->full, 1);
    mbar_init(&bars->done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm<512>(&bars->tmem_base);
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const int row = warp * 32 + lane;
  {
    // own row of A as packed pairs: column 256 + j holds k = 2j (low half), 2j + 1 (high half)
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a + ((int64_t)rank * 128 + row) * Kd);
    for (int cc = 0; cc < Kd / 32; ++cc) {
      uint32_t v[16];
      for (int i = 0; i < 16; ++i) v[i] = src[cc * 16 + i];
      tmem_st_32x16(tmem + ((uint32_t)(warp * 32) << 16) + 256 + cc * 16, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const int nck = Kd / 64;
    if (rank == 0) mbar_arrive_expect_tx(&bars->full, 2 * nck * nh * 128);
    for (int ck = 0; ck < nck; ++ck) tma_load_2d_2sm(sbm + ck * nh * 128, &map_b, &bars->full, ck * 64, (int)rank * nh);
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(256, N, 0, 0);
      mbar_wait(&bars->full, 0, 210);
      tc_fence_after();
      for (int ck = 0; ck < nck; ++ck)
        for (int ks = 0; ks < 4; ++ks)
          mma_bf16_ts_2sm(tmem, tmem + 256 + (ck * 4 + ks) * 8, desc_kmajor_sw128(smem_u32(sbm) + ck * nh * 128 + ks * 32), idesc,
                          (ck | ks) != 0);
      mma_commit_2sm(&bars->done);
    }
    mbar_wait(&bars->done, 0, 211);
  }
  __syncthreads();
  tc_fence_after();
  for (int cc = 0; cc < N / 32; ++cc) {
    uint32_t r[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + cc * 32, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) c[((int64_t)rank * 128 + row) * N + cc * 32 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  cluster_sync();
  if (warp == 1) { tc_fence_after(); tmem_dealloc_2sm<512>(tmem); }
}

}  // namespace rc

#ifdef RC_BRINGUP
extern "C" int rc_debug_umma_gemm_ts_2sm(const void* a_bf16, const void* b_bf16, int N, int Kd, float* c, void* stream) {
  RC_REQUIRE(a_bf16 && b_bf16 && c, "rc_debug_umma_gemm_ts_2sm: null pointer");
  RC_REQUIRE(N >= 64 && N <= 256 && N % 64 == 0 && Kd >= 64 && Kd <= 256 && Kd % 64 == 0, "rc_debug_umma_gemm_ts_2sm: bad shape N=%d Kd=%d", N, Kd);
  CUtensorMap mb;
  int rcode;
  {
    const uint64_t bdims[2] = {(uint64_t)Kd, (uint64_t)N}, str[2] = {2, (uint64_t)Kd * 2};
    const uint32_t bbox[2] = {64, (uint32_t)(N / 2)};
    if ((rcode = rc::make_tmap_bf16(&mb, b_bf16, 2, bdims, str, bbox, "debug ts B"))) return rcode;
  }
  const int smem = 65536 + 64;
  cudaError_t e = cudaFuncSetAttribute(rc::debug_umma_gemm_ts_2sm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return rc::fail(RC_ERR_CUDA, "rc_debug_umma_gemm_ts_2sm: smem opt-in: %s", cudaGetErrorString(e));
  rc::debug_umma_gemm_ts_2sm_kernel<<<2, 128, smem, (cudaStream_t)stream>>>(mb, (const __nv_bfloat16*)a_bf16, N, Kd, c);
  return rc::check_launch("rc_debug_umma_gemm_ts_2sm");
}
#endif  // RC_BRINGUP
