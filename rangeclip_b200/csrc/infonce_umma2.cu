// K1/K2, CTA-pair version: the fused pixel-text InfoNCE forward + backward on
// `tcgen05.mma.cta_group::2`, software-pipelined across tiles.
//
// Two CTAs on neighbouring SMs form a cluster and work on two 128-pixel tiles at a time.  Every MMA
// spans both SMs (M = 256): each CTA stages only ITS half of the streamed operand and the tensor cores
// read the other half from the peer's shared memory, which halves the shared-memory fill traffic and
// the operand read traffic per SM:
//   S   = X^T T^T   M = 256 px (128 per CTA, A = own X tile, MN-major), N = Kp text rows, B split: Kp/2 rows per CTA
//   dX^T = T^T P^T   M = 256 channels (128 per CTA, A = own rows of T^T), N = 128 px (64 of each CTA's tile,
//                    B = own P rows), four [128 ch x 128 px] accumulators per tile pair per CTA
//
// Pipeline (per tile pair i; TMEM = S 256 columns + two 128-column dX accumulators, all 512 in use):
//   tensor pipe : S(0) | S(1) dX(0) | S(2) dX(1) | ...      S(i+1) is issued BEFORE dX(i)
//   softmax     : reads S(i) from TMEM while dX(i-1) runs, keeps P(i) = exp(z - m) as packed bf16 in
//                 REGISTERS, releases the S columns at once (so S(i+1) can start), and stores P(i) to the
//                 single shared-memory P buffer as soon as the dX(i-1) MMAs have finished reading P(i-1)
//   dX epilogue : own 128 channel rows x 64 contiguous pixels per accumulator: x is read straight from
//                 global memory (L2-resident: the S GEMM just streamed it) with 256-bit loads, prefetched one
//                 accumulator ahead, and dX goes back with 256-bit stores -- no shared-memory staging.
// Register budget by role (setmaxnreg; the sum over the five warpgroups must stay within the 5 x 96 the CTA is
// launched with): control warps 48, softmax warps 120, epilogue warps 96.
// Only the leader CTA (cluster rank 0) issues MMAs; completion is multicast to the mbarriers of both CTAs,
// consumer-release barriers live in the leader and receive remote arrivals from the peer.  The per-pixel
// row scales of a tile are exchanged through distributed shared memory (the dX epilogue of a CTA covers
// pixels of both tiles).
#include "common.cuh"
#include "umma.cuh"
#include <float.h>
#include <stdlib.h>
#include <stdio.h>

namespace rc {
using namespace umma;

namespace pair {

constexpr int kTilePx = 128;
constexpr int kThreads = 640;          // warps: 0 TMA, 1 MMA (leader CTA only), 2 TMEM alloc, 3 idle, 4-11 softmax, 12-19 dX epilogue
// RC_PAIR_STG2: two dX staging buffers per epilogue warp (a chunk's TMA store reads its buffer while the next chunk is staged)
// paid for with one X-ring stage (shared memory is full).  Measured, not adopted: 2.83 ms against 2.41 ms -- the X ring is the
// first-touch HBM path and three stages starve the S GEMM (with the dX stores ablated: 2.51 against 2.13 ms).
#ifndef RC_PAIR_STG2
#define RC_PAIR_STG2 0
#endif
// RC_PAIR_NORM_WARP: the row norms 1/|x_p| are taken by the relay warp (warp 2, otherwise one busy lane) as soon as a chunk
// lands, instead of by the eight softmax warps between their hand-offs: an X-ring slot is then released by the MMA commit and
// ONE early reader (was: eight readers that get to it only after their exp pass), and the softmax warps lose the pass.
// Measured, not adopted: 2.54 ms against 2.41 ms -- one warp's 64 dependent-latency LDS rounds per chunk release the slot
// later than eight warps' four, and warp 2 issues last on its scheduler.
// A/B knobs (tools/ablate_pair.py ab ...): L2 eviction hint of the text operand loads (0 none, 1 evict_last) and of the epilogue's
// x loads (1 evict_first, 0 none)
#ifndef RC_PAIR_TEXT_HINT
#define RC_PAIR_TEXT_HINT 0
#endif
#ifndef RC_PAIR_X_HINT
#define RC_PAIR_X_HINT 0      // evict_last on the X tile loads of the S GEMM (kBwd launches)
#endif
#ifndef RC_EPI_X_HINT
#define RC_EPI_X_HINT 1
#endif
// RC_PAIR_SMX_UNIT: the dX epilogue is the role that paces the launch (four accumulator units per tile pair, one after the other
// on eight warps), while the softmax warps idle between their P hand-off and the next S.  With this switch the softmax warps
// drain ONE of the four units (unit 1: channel block 0, pixels [64,128) of both tiles) in that window -- TMEM load, projection,
// direct 256-bit global stores (they have no staging tile) -- and the epilogue warps skip it.  Measured, not adopted (same box,
// base 2.43 ms): after the row norms (1) 3.18 ms, before them (2) 2.90 ms.  The softmax warps' wait for the next S is not slack:
// they sit on the S -> exp -> P -> dX dependency cycle, and whatever they do in between delays the exp pass of the next tile.
#ifndef RC_PAIR_SMX_UNIT
#define RC_PAIR_SMX_UNIT 0
#endif
#ifndef RC_EPI_FETCH_EARLY
#define RC_EPI_FETCH_EARLY 0      // measured (same box): early 2.50 ms, late 2.45 ms
#endif
#ifndef RC_PAIR_NORM_WARP
#define RC_PAIR_NORM_WARP 0
#endif
// RC_PAIR_CTL_HIGH: the four control warps (text TMA, MMA issuer, relay, X TMA) are the HIGHEST warp ids of the CTA.  The
// schedulers pick the highest eligible warp id first; as warps 0-3 the producers and the MMA issuer issue after the four busy
// softmax / epilogue warps that share their scheduler.
// Measured: 2.64 ms against 2.41 ms -- the control warps' barrier polling then takes issue slots from the workers.
// RC_PAIR_CTL_HIGH == 2: control warps stay lowest, but the SOFTMAX warps (the exp pass is on the S -> P -> dX dependency cycle)
// get the highest warp ids and the epilogue warps the middle ones: 2.56 ms.  The shipped order (control < softmax < epilogue)
// is the best of the three: the epilogue is the role with the least slack.
#ifndef RC_PAIR_CTL_HIGH
#define RC_PAIR_CTL_HIGH 0
#endif
constexpr int kXStages = RC_PAIR_STG2 ? 3 : 4;   // X ring: own X chunks [64 d][128 px] (first touch: HBM latency; also read by the row norms)
constexpr int kStgBufs = RC_PAIR_STG2 ? 2 : 1;
constexpr int kTStages = 4;            // text ring: text half-chunks [Kp/2][64 d] for S, own T^T rows [128 d][64 k] for dX (L2 hits)
constexpr int kStageBytes = 16 * 1024;
static_assert(kTStages >= 4, "a dX block keeps Kp/64 <= 4 slots of the text ring at once");
constexpr int kPBytes = 64 * 1024;
constexpr int kTmemCols = 512;
// Measured, not adopted: TMEM load of softmax step c+1 in flight while step c is computed (needs 128 softmax registers,
// taken from the control warps): 2.62 ms against 2.49 ms; forward-only launches, which have the registers without
// spilling, also lose (1.61 ms against 1.47 ms) -- the exposed TMEM latency is not what bounds the exp pass.
#ifndef RC_SMX_DOUBLE_BUFFER
#define RC_SMX_DOUBLE_BUFFER 0
#endif
#if RC_SMX_DOUBLE_BUFFER
constexpr int kRegsCtl = 32, kRegsSoftmax = 128;     // epilogue warps keep the entry allocation (96); 32 + 2*128 + 2*96 = 5*96
#else
constexpr int kRegsCtl = 48, kRegsSoftmax = 120;     // epilogue warps keep the entry allocation (96); 48 + 2*120 + 2*96 = 5*96
#endif
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct __align__(8) Bars {
  uint64_t xf[kXStages];        // own X chunk has landed (CTA-local; relayed to the leader's xfull)
  uint64_t xfull[kXStages], xempty[kXStages];
  uint64_t tfull[kTStages], tempty[kTStages];
  uint64_t s_full[2], s_empty[2], p_full, p_empty;   // S: one TMEM buffer with a backward, two (columns 0 / 256) forward-only
  uint64_t acc_full[2], acc_empty[2];
  uint64_t sc_full[2];
  uint64_t n_full[2], n_empty[2];   // row norms of a tile are in shared memory / have been read (RC_PAIR_NORM_WARP; by tile parity)
  uint32_t tmem_base, pad;
};

constexpr int kOffT = kXStages * kStageBytes;
constexpr int kOffP = kOffT + kTStages * kStageBytes;
constexpr int kScaleBufs = 3;          // the softmax warps run up to two tiles ahead of the dX epilogue warps
constexpr int kOffScale = kOffP + kPBytes;                  // {rs, -cs} bf16x2 pairs: [kScaleBufs tiles][2 owner CTAs][64 px pairs] uint2
constexpr int kOffXch = kOffScale + 2 * kScaleBufs * 2 * 128 * 4;
constexpr int kOffStg = kOffXch + 2 * 4 * 2 * 128 * 4;      // exchange: [2 tile parities][max, sum, sez, sy][2 halves][128]
constexpr int kStgBytes = 32 * 32 * 2;                      // dX staging of one epilogue warp: [32 d][32 px] bf16, 64-byte swizzle
constexpr int kOffPart = kOffStg + 8 * kStgBufs * kStgBytes;           // kStgBufs staging buffers per epilogue warp
constexpr int kOffBars = kOffPart + 8 * 128 * 4;            // row-norm partial sums of squares [8 softmax warps][128 px]
constexpr int kSmemBytes = kOffBars + (int)sizeof(Bars);
static_assert(kSmemBytes <= 232448, "shared-memory budget of one SM (227 KB)");

#ifdef RC_TIMING
__device__ __forceinline__ unsigned long long globaltimer_ns() {     // one clock for both CTAs of the pair
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define RC_T0(name) const long long name = clock64()
#define RC_TACC(idx, name) wt[idx] += clock64() - name
#define RC_WAIT(fn, bar, par, tag) do { const long long t0_ = clock64(); fn(bar, par, tag); wt[tag] += clock64() - t0_; } while (0)
// event trace of CTA 0 (tools/timeline_pair.py): clock of event `id` of tile-pair iteration `iter` in [40, 44)
#define RC_EV(iter, id) do { if (prm.dbg != nullptr && blockIdx.x < 2 && lane == 0 && (iter) >= 40u && (iter) < 44u) \
    prm.dbg[256 + blockIdx.x * 192 + ((iter) - 40u) * 48 + (id)] = (long long)globaltimer_ns(); } while (0)
#else
#define RC_T0(name)
#define RC_TACC(idx, name)
#define RC_WAIT(fn, bar, par, tag) fn(bar, par, tag)
#define RC_EV(iter, id)
#endif

// Measured, not adopted (RC_ROLE_WAIT): a whole role (eight warps) waits for one mbarrier phase with ONE polling warp and
// the others blocked in a named barrier.  A polling warp spends ~12 issue slots per probe and the epilogue's probes are
// a quarter of all instructions the SM issues (ncu source view), yet coupling the epilogue warps costs more than the
// probes (2.65 ms against 2.52 ms) and coupling the softmax warps is within noise (2.49 against 2.52 ms).
#ifndef RC_ROLE_WAIT
#define RC_ROLE_WAIT 0      // bit 0: epilogue warps, bit 1: softmax warps
#endif
#define ROLE_WAIT_ON(poller, bar, par, tag, bar_id) do { if (poller) RC_WAIT(mbar_wait, bar, par, tag); named_bar_sync(bar_id, 256); } while (0)
#if RC_ROLE_WAIT & 1
#define ROLE_WAIT_EPI ROLE_WAIT_ON
#else
#define ROLE_WAIT_EPI(poller, bar, par, tag, bar_id) RC_WAIT(mbar_wait, bar, par, tag)
#endif
#if RC_ROLE_WAIT & 2
#define ROLE_WAIT_SMX ROLE_WAIT_ON
#else
#define ROLE_WAIT_SMX(poller, bar, par, tag, bar_id) RC_WAIT(mbar_wait, bar, par, tag)
#endif

struct Params {
  long long* dbg;
  int B, D, K, Kp;
  int64_t HW;
  int tiles_per_img, n_tiles, n_pairs;
  uint32_t tpi_magic;       // floor(2^32 / tiles_per_img): tile / tiles_per_img as a multiply-high + one correction
  int ablate;               // bring-up only (RANGECLIP_B200_ABLATE): 1 no epilogue x loads, 2 no dX stores, 64 no row-norm reads, 128 no dX staging, 256 no text reloads
  int store_g;              // 1: also write G = rs (P - sum onehot) (bf16 [B][HW][Kp]) for the dText GEMM
  int wide;                 // 1: rows of X / dX are 32-byte aligned (256-bit global accesses allowed)
  int split_c, split_b;     // bring-up (RANGECLIP_B200_SPLIT="c,b"): S chunks / dX blocks issued in the first half of an iteration; -1 = default
  const __nv_bfloat16* x;
  __nv_bfloat16* dx;
  const float* inv_norm;
  // K-blocked use (more than 256 candidates, e.g. the area-image loss at thousands of objects): the candidates are split
  // into launches of <= 256 rows.  keep_w: a target outside this launch's rows keeps its weight (only the one-hot term is
  // dropped).  lse_in: the row's logsumexp over ALL candidates, from a first round of forward launches; the softmax of
  // this block is then exp(z - lse_in) instead of exp(z - m) / (block sum).
  int keep_w;
  int acc_dx;               // K-blocked backward launches after the first: dX += this block's gradient (TMA reduce-add store)
  const float* lse_in;
  // kb > 0: ALL candidate blocks in one launch.  The "images" of the tile index are the kb blocks of 256 candidate rows:
  // X (one image) is read with image coordinate 0, everything else (y, w, lse, dX, text rows) is indexed by the block.
  // K is then the total number of candidates; lse_in is indexed by the pixel alone.
  int kb;
  const int32_t* y;
  const float* w;
  float inv_tau;
  // sync-free callers (device-side contrast-set builder): the number of valid candidate rows and the log
  // temperature are read from device memory at kernel start (nullable: K / inv_tau above are used)
  const int32_t* k_dev;
  const float* log_tau_dev;
  const float* grad_scale;
  const double* w_sum_in;
  float* lse;
  double* loss_sum;
  double* w_sum;
  double* dlogtau;
};

// tile / tiles_per_img without the ~20-instruction runtime division (exact for tile < 2^32: the estimate is low by at most 1)
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

// 16 pixels (32 bytes) of one channel row, straight from / to global memory.  `n8` = number of valid
// 8-pixel groups (0, 1 or 2); the 256-bit form needs 32-byte alignment (`wide`).
__device__ __forceinline__ void ldg_px16(const __nv_bfloat16* p, bool wide, int n8, uint32_t* r, uint64_t pol = 0) {
  if (wide && n8 == 2 && pol != 0) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p), "l"(pol));
  } else if (wide && n8 == 2) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
  } else {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (g < n8) v = __ldg(reinterpret_cast<const uint4*>(p) + g);
      r[g * 4] = v.x; r[g * 4 + 1] = v.y; r[g * 4 + 2] = v.z; r[g * 4 + 3] = v.w;
    }
  }
}
__device__ __forceinline__ void stg_px16(__nv_bfloat16* p, bool wide, int n8, const uint32_t* r) {
  if (wide && n8 == 2) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                 "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
  } else {
#pragma unroll
    for (int g = 0; g < 2; ++g)
      if (g < n8) reinterpret_cast<uint4*>(p)[g] = make_uint4(r[g * 4], r[g * 4 + 1], r[g * 4 + 2], r[g * 4 + 3]);
  }
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

// tile -> (image, first pixel); tiles past the end map to image index B, which is out of bounds for every
// tensor map (TMA loads return zeros)
__device__ __forceinline__ void tile_coords(const Params& prm, int tile, int& b, int& px0) {
  if (tile < prm.n_tiles) {
    b = div_tiles(prm, tile);
    px0 = (tile - b * prm.tiles_per_img) * kTilePx;
  } else {
    b = prm.B;
    px0 = 0;
  }
}

// R = targets per embedding row: 1 = one pixel per row (y, w are [B*HW]); 4 = every row stands for the four pixels of a
// 2x2 block that share one embedding (decoder.py:113 nearest x2): y, w are [B*HW][4], the loss of the row is
// sum_j w_j (lse - z[y_j]) and dX is the gradient with respect to the shared embedding (the sum over the block).
// kKB: K-blocked launches (Params::keep_w / lse_in); a template flag so that the headline instantiation carries none of it
template <bool kBwd, int R, bool kKB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
infonce_umma_pair_kernel(const __grid_constant__ CUtensorMap map_x_s,   // X [B][D][HW], box (64 px, 64 d, 1)
                         const __grid_constant__ CUtensorMap map_t,     // T [Kp][D],    box (64 d, Kp/2 rows)
                         const __grid_constant__ CUtensorMap map_tt,    // T^T [D][Kp],  box (64 k, 128 d)
                         const __grid_constant__ CUtensorMap map_dx,    // dX [B][D][HW], box (32 px, 32 d, 1)
                         const __grid_constant__ CUtensorMap map_g,     // G [B][HW][Kp],  box (64 k, 128 px, 1) (dText only)
                         const Params prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Bars* bars = reinterpret_cast<Bars*>(smem + kOffBars);
  uint32_t* sc_s = reinterpret_cast<uint32_t*>(smem + kOffScale);   // -cs as bf16x2 of a pixel PAIR: [kScaleBufs tiles][2 owner CTAs][64 pairs]
  // `warp` is the ROLE index (0-3 control, 4-11 softmax, 12-19 epilogue); with RC_PAIR_CTL_HIGH the control roles run on the
  // hardware warps 16-19 (the TMEM lane quarter of a softmax / epilogue warp, warp & 3, is the same either way)
  const int lane = threadIdx.x & 31;
  const int pwarp = (int)(threadIdx.x >> 5);
  const int warp = RC_PAIR_CTL_HIGH == 1 ? (pwarp + 4) % (kThreads / 32)
                 : RC_PAIR_CTL_HIGH == 2 ? (pwarp < 4 ? pwarp : (pwarp < 12 ? pwarp + 8 : pwarp - 8)) : pwarp;
  const int rtid = warp * 32 + lane;          // thread index in role order
  const uint32_t rank = cluster_ctarank();
  const bool leader_cta = rank == 0;
  const int n_dchunks = prm.D / 64;      // 64-channel chunks of the S GEMM
  const int n_blk = prm.D / 256;         // 256-channel blocks of the dX GEMM
  const int n_kchunks = prm.Kp / 64;
  // MMA issue order per tile pair: the S GEMM of the next pair is split in two and wrapped around the dX blocks of
  // this pair, so that the dX epilogue drains accumulators while the tensor pipe works on S
  const int c_half = prm.split_c >= 0 ? min(prm.split_c, n_dchunks) : n_dchunks / 2;
  const int b_half = prm.split_b >= 0 ? min(prm.split_b, n_blk) : (n_blk + 1) / 2;
  const int Nh = prm.Kp / 2;             // text rows staged by each CTA
  const int n_clusters = gridDim.x / 2;
  const int cluster_id = blockIdx.x / 2;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_x_s); tma_prefetch_desc(&map_t);
    if (kBwd) { tma_prefetch_desc(&map_tt); tma_prefetch_desc(&map_dx); }
    // X ring: xfull = the relays of both CTAs, xempty = MMA commit + the eight softmax warps (row norms);
    // text ring: tfull = the leader's expect_tx (bytes of both CTAs), tempty = MMA commit
    for (int i = 0; i < kXStages; ++i) { mbar_init(&bars->xf[i], 1); mbar_init(&bars->xfull[i], 2); mbar_init(&bars->xempty[i], RC_PAIR_NORM_WARP ? 2 : 9); }
    mbar_init(&bars->n_full[0], 1); mbar_init(&bars->n_full[1], 1); mbar_init(&bars->n_empty[0], 8); mbar_init(&bars->n_empty[1], 8);
    for (int i = 0; i < kTStages; ++i) { mbar_init(&bars->tfull[i], 1); mbar_init(&bars->tempty[i], 1); }
    // consumer releases are ONE arrival per warp (after __syncwarp), not one per thread: an mbarrier arrive is a
    // shared-memory atomic, and 512 of them per hand-off (half of them remote) cost the leader's shared-memory pipe
    // more wavefronts than the P stores
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->s_full[i], 1); mbar_init(&bars->s_empty[i], 16); }
    mbar_init(&bars->p_full, 16); mbar_init(&bars->p_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->acc_full[i], 1); mbar_init(&bars->acc_empty[i], 16); }
    mbar_init(&bars->sc_full[0], 5); mbar_init(&bars->sc_full[1], 5);   // 4 local writer warps + 1 expect_tx (peer: st.async)
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
  // one arrival for the whole warp: every lane has fenced its own accesses, __syncwarp orders them before lane 0's arrive
  auto arrive_leader_warp = [&](uint64_t* bar) {
    __syncwarp();
    if (elect_one()) arrive_leader(bar);
  };

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsCtl));
    if (warp == 0 && lane == 0) {
      // =============================== TMA producer (both CTAs) ===============================
      // ring order == MMA issue order: S(first), then per tile pair [S(next)] [dX(this)]
      // text ring, in MMA issue order: S(first), then per tile pair
      //   [S(next) first half] [dX(this) first blocks] [S(next) second half] [dX(this) remaining blocks]
      uint32_t it = 0;
      const uint64_t pol_text = RC_PAIR_TEXT_HINT ? l2_policy_evict_last() : 0;
      // first candidate row of the pair's block (kb mode: both tiles of a pair lie in one block, tiles_per_img is even)
      auto koff_of = [&](int pj) -> int { return (kKB && prm.kb > 0) ? div_tiles(prm, 2 * pj) * 256 : 0; };
      auto load_s = [&](int c_begin, int c_end, int koff) {
        for (int c = c_begin; c < c_end; ++c, ++it) {   // own half (Nh rows) of text chunk c
          const int st = it % kTStages;
          RC_WAIT(mbar_wait, &bars->tempty[st], ((it / kTStages) & 1) ^ 1, 1);
          if ((prm.ablate & 256) && it >= (uint32_t)kTStages) {      // bring-up: no text traffic after the first ring pass (stale operands)
            if (leader_cta) mbar_arrive(&bars->tfull[st]);
            continue;
          }
          if (leader_cta) mbar_arrive_expect_tx(&bars->tfull[st], 2 * Nh * 128);
          if (RC_PAIR_TEXT_HINT) tma_load_2d_2sm_hint(smem + kOffT + st * kStageBytes, &map_t, &bars->tfull[st], c * 64, koff + (int)rank * Nh, pol_text);
          else tma_load_2d_2sm(smem + kOffT + st * kStageBytes, &map_t, &bars->tfull[st], c * 64, koff + (int)rank * Nh);
        }
      };
      auto load_dx = [&](int b_begin, int b_end, int koff) {
        for (int blk = b_begin; blk < b_end; ++blk)
          for (int kc = 0; kc < n_kchunks; ++kc, ++it) {   // own 128 rows of T^T for this 256-channel block, 64 k at a time
            const int st = it % kTStages;
            RC_WAIT(mbar_wait, &bars->tempty[st], ((it / kTStages) & 1) ^ 1, 2);
            if ((prm.ablate & 256) && it >= (uint32_t)kTStages) {
              if (leader_cta) mbar_arrive(&bars->tfull[st]);
              continue;
            }
            if (leader_cta) mbar_arrive_expect_tx(&bars->tfull[st], 2 * 16384);
            if (RC_PAIR_TEXT_HINT) tma_load_2d_2sm_hint(smem + kOffT + st * kStageBytes, &map_tt, &bars->tfull[st], koff + kc * 64, blk * 256 + (int)rank * 128, pol_text);
            else tma_load_2d_2sm(smem + kOffT + st * kStageBytes, &map_tt, &bars->tfull[st], koff + kc * 64, blk * 256 + (int)rank * 128);
          }
      };
      if (cluster_id < prm.n_pairs) load_s(0, n_dchunks, koff_of(cluster_id));
      for (int pj = cluster_id; pj < prm.n_pairs; pj += n_clusters) {
        const bool has_next = pj + n_clusters < prm.n_pairs;
        const int k_this = koff_of(pj), k_next = has_next ? koff_of(pj + n_clusters) : 0;
        if (has_next) load_s(0, c_half, k_next);
        if (kBwd) load_dx(0, b_half, k_this);
        if (has_next) load_s(c_half, n_dchunks, k_next);
        if (kBwd) load_dx(b_half, n_blk, k_this);
      }
    } else if (warp == 3 && lane == 0) {
      // =============================== X producer (both CTAs) ===============================
      // own X chunks of every tile, as far ahead as the X ring allows (the next tile's first chunks are in flight
      // while the dX GEMM of the previous pair runs)
      uint32_t xit = 0;
      const uint64_t pol_keep = RC_PAIR_X_HINT ? l2_policy_evict_last() : 0;
      for (int pj = cluster_id; pj < prm.n_pairs; pj += n_clusters) {
        int b, px0;
        tile_coords(prm, 2 * pj + (int)rank, b, px0);
        for (int c = 0; c < n_dchunks; ++c, ++xit) {
          const int st = xit % kXStages;
          RC_WAIT(mbar_wait, &bars->xempty[st], ((xit / kXStages) & 1) ^ 1, 1);
          uint8_t* sb = smem + st * kStageBytes;
          mbar_arrive_expect_tx(&bars->xf[st], 2 * 8192);          // CTA-local: the softmax warps read the chunk too
          const int bx = (kKB && prm.kb > 0) ? (b < prm.B ? 0 : 1) : b;      // kb mode: the one image of X (1 = out of bounds)
          if (RC_PAIR_X_HINT) {       // the dX epilogue reads the tile again about one and a half iterations later
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
      // The whole warp runs the loop converged, so stage indices, addresses and descriptors live in uniform
      // registers; one elected lane issues the tcgen05 instructions.
      uint32_t it = 0, xit = 0, uc = 0;
      const uint32_t smem_base = smem_u32(smem);
      // descriptor templates: only the 14-bit start-address field changes (+ bytes/16 per step)
      const uint64_t dsc_x = desc_mnmajor_sw128(0, 8192);
      const uint64_t dsc_k = desc_kmajor_sw128(0);
      // S(n) goes to TMEM columns [0,256); forward-only launches alternate with [256,512) (no dX accumulators there),
      // so that the S GEMM of the next pair overlaps the softmax of this one
      auto issue_s = [&](uint32_t n, int c_begin, int c_end) {
        const uint32_t sidx = kBwd ? 0u : (n & 1u);
        const uint32_t s_tmem = tmem + sidx * 256;
        if (c_begin == 0) {
          RC_WAIT(mbar_wait, &bars->s_empty[sidx], (((kBwd ? n : (n >> 1)) & 1u) ^ 1u), 3);   // previous user has read it
          tc_fence_after();
        }
        for (int c = c_begin; c < c_end; ++c, ++it, ++xit) {
          const int sa = xit % kXStages, sb_ = it % kTStages;
          RC_WAIT(mbar_wait, &bars->xfull[sa], (xit / kXStages) & 1, 4);
          RC_WAIT(mbar_wait, &bars->tfull[sb_], (it / kTStages) & 1, 12);
          tc_fence_after();
          const uint64_t xa = dsc_x + ((smem_base + sa * kStageBytes) >> 4);
          const uint64_t tb = dsc_k + ((smem_base + kOffT + sb_ * kStageBytes) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma_bf16_ss_2sm(s_tmem, xa + ((ks * 2048) >> 4), tb + ((ks * 32) >> 4), idesc_s, (c | ks) != 0);
            mma_commit_2sm(&bars->xempty[sa]);
            mma_commit_2sm(&bars->tempty[sb_]);
            if (c + 1 == n_dchunks) mma_commit_2sm(&bars->s_full[sidx]);
          }
          __syncwarp();
        }
      };
      const uint64_t pb = dsc_k + ((smem_base + kOffP) >> 4);
      uint32_t lt = 0;
      auto issue_dx = [&](int b_begin, int b_end) {
        for (int blk = b_begin; blk < b_end; ++blk) {
          for (int pxh = 0; pxh < 2; ++pxh, ++uc) {
            const int ab = uc & 1;
            RC_WAIT(mbar_wait, &bars->acc_empty[ab], ((uc >> 1) & 1) ^ 1, 6);
            tc_fence_after();
            RC_EV(lt, 3 + blk * 2 + pxh);     // dX unit issue starts (accumulator free)
            const uint32_t dcol = tmem + 256 + ab * 128;
            for (int kc = 0; kc < n_kchunks; ++kc) {
              const int st = (it + kc) % kTStages;
              if (pxh == 0) {          // the block's last text slot was requested only when the preceding S chunk retired:
                RC_WAIT(mbar_wait, &bars->tfull[st], ((it + kc) / kTStages) & 1, 7);     // wait per slot, not up front
                tc_fence_after();
              }
              const uint64_t sb = dsc_k + ((smem_base + kOffT + st * kStageBytes) >> 4);
              const uint64_t pk_ = pb + ((kc * 16384 + pxh * 8192) >> 4);
              if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)     // A: own T^T rows [128 d][64 k]; B: own P rows [64 px][64 k]
                  mma_bf16_ss_2sm(dcol, sb + ((ks * 32) >> 4), pk_ + ((ks * 32) >> 4), idesc_d, (kc | ks) != 0);
                if (pxh == 1) mma_commit_2sm(&bars->tempty[st]);     // second (last) reader: the slot refills at once
              }
              __syncwarp();
            }
            if (elect_one()) mma_commit_2sm(&bars->acc_full[ab]);
            __syncwarp();
          }
          it += n_kchunks;
        }
      };
      if (cluster_id < prm.n_pairs) issue_s(0, 0, n_dchunks);
      for (int pj = cluster_id; pj < prm.n_pairs; pj += n_clusters, ++lt) {
        const bool has_next = pj + n_clusters < prm.n_pairs;
        if (has_next) {
          issue_s(lt + 1, 0, c_half);          // waits until the softmax warps have read the previous S out of these columns
          RC_EV(lt + 1, 0);        // S(lt+1) issue started
        }
        if (kBwd) {
          RC_WAIT(mbar_wait, &bars->p_full, lt & 1, 5);
          tc_fence_after();
          RC_EV(lt, 2);            // dX(lt) issue starts (P(lt) is in shared memory)
          issue_dx(0, b_half);
        }
        if (has_next) {
          issue_s(lt + 1, c_half, n_dchunks);
          RC_EV(lt + 1, 1);        // S(lt+1) issued
        }
        if (kBwd) {
          issue_dx(b_half, n_blk);
          if (elect_one()) mma_commit_2sm(&bars->p_empty);
          __syncwarp();
          RC_EV(lt, 7);            // dX(lt) issued
        }
      }
    }
    else if (warp == 2 && (RC_PAIR_NORM_WARP || lane == 0)) {
      // ========== relay (both CTAs): "own X chunk has landed" (CTA-local xf) -> the leader's full barrier ==========
      // + (RC_PAIR_NORM_WARP) the row norms 1/|x_p| (model.py:272 F.normalize) of the tile, read from the chunk where it sits in
      // the operand ring ([64 d][2 x 64 px], 128-byte swizzle): lane = 4 consecutive pixels (8 bytes of every channel row),
      // channel rows and chunks summed in a fixed order (bit-reproducible); the slot is released right after the read
      uint32_t xit = 0, lt = 0;
      float* nrm = reinterpret_cast<float*>(smem + kOffPart);          // [2 tile parities][128 px]
      const uint8_t* lane_base = smem + (lane >> 4) * 8192 + (lane & 1) * 8;
      const int gran = (lane & 15) >> 1;                               // 16-byte granule of the row (before the swizzle)
      for (int pj = cluster_id; pj < prm.n_pairs; pj += n_clusters, ++lt) {
        float ss0 = 0.f, ss1 = 0.f, ss2 = 0.f, ss3 = 0.f;
        for (int c = 0; c < n_dchunks; ++c, ++xit) {
          const int st = xit % kXStages;
          RC_WAIT(mbar_wait, &bars->xf[st], (xit / kXStages) & 1, 14);
          if (lane == 0) arrive_leader(&bars->xfull[st]);
          if (RC_PAIR_NORM_WARP) {
            const uint8_t* base = lane_base + st * kStageBytes;
            if (!(prm.ablate & 64)) {
#pragma unroll 8
              for (int rowd = 0; rowd < 64; ++rowd) {
                const uint2 v = *reinterpret_cast<const uint2*>(base + rowd * 128 + ((gran ^ (rowd & 7)) << 4));
                ss0 = sqacc_bf16x2_lo(ss0, v.x); ss1 = sqacc_bf16x2_hi(ss1, v.x);
                ss2 = sqacc_bf16x2_lo(ss2, v.y); ss3 = sqacc_bf16x2_hi(ss3, v.y);
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->xempty[st]);
          }
        }
        if (RC_PAIR_NORM_WARP) {
          // nrm[lt & 1] was last read by the softmax warps of tile lt - 2
          if (lt >= 2) RC_WAIT(mbar_wait, &bars->n_empty[lt & 1], ((lt >> 1) & 1) ^ 1, 13);
          float4 o;
          o.x = 1.f / fmaxf(sqrtf(ss0), 1e-12f); o.y = 1.f / fmaxf(sqrtf(ss1), 1e-12f);
          o.z = 1.f / fmaxf(sqrtf(ss2), 1e-12f); o.w = 1.f / fmaxf(sqrtf(ss3), 1e-12f);
          reinterpret_cast<float4*>(nrm + (lt & 1) * 128)[lane] = o;
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->n_full[lt & 1]);
        }
      }
    }
  } else if (warp < 12) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsSoftmax));
    // ======================= softmax / CE warps: own tile (two warps per TMEM lane quarter) =======================
    const int half = warp >= 8 ? 1 : 0;
    const int row = (warp & 3) * 32 + lane;                 // pixel of the own tile == TMEM lane
    const int Kh = prm.Kp >> 1;
    const int cb = half * Kh;
    const uint32_t trow0 = tmem + ((uint32_t)((warp & 3) * 32) << 16) + cb;
    uint8_t* prow = smem + kOffP + row * 128;
    const int sw = row & 7;
    float* xch_base = reinterpret_cast<float*>(smem + kOffXch);
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
        const int tb = div_tiles(prm, t);
        const int tpx = (t - tb * prm.tiles_per_img) * kTilePx + row;
        if (tpx < prm.HW) {
          const int64_t tm = (int64_t)tb * prm.HW + tpx;
          nx_inv_n = 1.f;        // valid pixel (the value itself comes from the norm warps)
          if (R == 1) {
            nx_y = __ldg(prm.y + tm);
            nx_w = __ldg(prm.w + tm);
          } else {               // four targets, 8 bits each (K <= 256); ignored targets are sorted out when w is read
            const int4 y4 = __ldg(reinterpret_cast<const int4*>(prm.y) + tm);
            nx_y = (y4.x & 255) | ((y4.y & 255) << 8) | ((y4.z & 255) << 16) | ((y4.w & 255) << 24);
          }
        }
      }
    };
    load_pixel_scalars(cluster_id);
    // Row norms 1/|x_p| (model.py:272 F.normalize) of the NEXT tile, computed by these warps while they would
    // otherwise wait for its S GEMM: the X chunks are read where they sit in the operand ring ([64 d][2 x 64 px],
    // 128-byte swizzle).  thread = (8-pixel group g, row phase r); sums of squares meet in shared memory.
    float* part_s = reinterpret_cast<float*>(smem + kOffPart);   // [8 warps][128 px] partial sums of squares
    const int st_ = rtid - 128;                                  // 0..255
    const int ng = st_ & 15, nr = st_ >> 4;
    uint32_t nit = 0;
    float inv_n_next = 0.f;
    auto norm_tile = [&]() {
      float ss[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) ss[j] = 0.f;
      for (int c = 0; c < n_dchunks; ++c, ++nit) {
        const int st = nit % kXStages;
        ROLE_WAIT_SMX(warp == 4, &bars->xf[st], (nit / kXStages) & 1, 15, 8);
        const uint8_t* base = smem + st * kStageBytes + (ng >> 3) * 8192 + nr * 128;
        const int ch = ng & 7;
#pragma unroll
        for (int rr = 0; rr < ((prm.ablate & 64) ? 0 : 4); ++rr) {
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
        if (elect_one()) mbar_arrive(&bars->xempty[st]);     // (elect.sync: no lane id to re-materialise in this hot loop)
      }
      // fixed-order reduction (bit-reproducible): lane pairs, then the eight warps' partials through shared memory
#pragma unroll
      for (int j = 0; j < 8; ++j) ss[j] += __shfl_xor_sync(0xffffffffu, ss[j], 16);     // lanes l and l^16: same pixels, other rows
      if (lane < 16) {
        float4* dst = reinterpret_cast<float4*>(part_s + (warp - 4) * 128 + ng * 8);
        dst[0] = make_float4(ss[0], ss[1], ss[2], ss[3]);
        dst[1] = make_float4(ss[4], ss[5], ss[6], ss[7]);
      }
      named_bar_sync(5, 256);
      float q = 0.f;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) q += part_s[wv * 128 + row];
      inv_n_next = 1.f / fmaxf(sqrtf(q), 1e-12f);
    };
    if (!RC_PAIR_NORM_WARP && cluster_id < prm.n_pairs) norm_tile();
    const float* nrm = reinterpret_cast<const float*>(smem + kOffPart);
    const float inv_tau = prm.log_tau_dev != nullptr ? expf(-__ldg(prm.log_tau_dev)) : prm.inv_tau;
    const int K_valid = prm.k_dev != nullptr ? max(1, min(__ldg(prm.k_dev), prm.K)) : prm.K;      // rows past it are zero pads
    const bool use_bound = inv_tau * (2.02f * kLog2e) < 100.f;
    const float ml_bound = inv_tau * (1.01f * kLog2e);
    // peer copies of the row-scale arrays (same offsets in the other CTA's shared memory)
    const uint32_t sc_peer = map_to_cta(sc_s, rank ^ 1);
    const uint32_t sc_peer0 = map_to_cta(&bars->sc_full[0], rank ^ 1), sc_peer1 = map_to_cta(&bars->sc_full[1], rank ^ 1);
    uint32_t lt = 0;
    for (int pj = cluster_id; pj < prm.n_pairs; pj += n_clusters, ++lt) {
      const int tile = 2 * pj + (int)rank;
      const bool tile_ok = tile < prm.n_tiles;
      const int b = tile_ok ? div_tiles(prm, tile) : 0;
      const int px = tile_ok ? (tile - b * prm.tiles_per_img) * kTilePx + row : 0;
      const bool valid = tile_ok && px < prm.HW;
      // dText mode: the bulk store of the previous tile's G must have read the P buffer before anyone rewrites it
      // (every thread passes the named barrier below before its next P store)
      if (kBwd && prm.store_g && rtid == 128) tma_store_wait_read0();
      const int64_t m = (int64_t)b * prm.HW + px;
      // candidates in this tile's block (kb mode: the last block may be short; its zero pad rows are masked like any pad)
      const int Kt = (kKB && prm.kb > 0) ? min(256, prm.K - 256 * b) : K_valid;
      const bool px_ok = nx_inv_n != 0.f;
      const int yi = nx_y;
      const float wi = (R == 1 && (yi >= 0 || (kKB && prm.keep_w))) ? nx_w : 0.f;
      load_pixel_scalars(pj + n_clusters);
      float* xch = xch_base + (lt & 1) * (4 * 2 * 128);      // double-buffered: one named barrier per tile suffices
      if (RC_PAIR_NORM_WARP) {
        RC_WAIT(mbar_wait, &bars->n_full[lt & 1], (lt >> 1) & 1, 15);
        inv_n_next = nrm[(lt & 1) * 128 + row];
        __syncwarp();
        if (elect_one()) mbar_arrive(&bars->n_empty[lt & 1]);
      }
      const float inv_n = px_ok ? inv_n_next : 0.f;
      const float zs = inv_n * inv_tau;
      const float zl = zs * kLog2e;
      // forward-only: the tensor pipe runs one pair ahead (two S buffers), so the next tile's row norms are taken
      // first -- its X chunks are already streaming and their ring slots are refilled only after these reads
      if (!RC_PAIR_NORM_WARP && !kBwd && pj + n_clusters < prm.n_pairs) norm_tile();
      const uint32_t sidx = kBwd ? 0u : (lt & 1u);
      const uint32_t trow = trow0 + sidx * 256;
      ROLE_WAIT_SMX(warp == 4, &bars->s_full[sidx], (kBwd ? lt : (lt >> 1)) & 1u, 8, 8);
      tc_fence_after();
      if (warp == 4) RC_EV(lt, 10);    // S(lt) complete (seen by the softmax warps)
      RC_T0(tsm);
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
      // e = exp(z - m) for this half's columns; P stays in registers as packed bf16 until the P buffer is free.
      // (Measured, not adopted: the scale / sum / e*s arithmetic as packed fp32x2 -- FFMA2 / FADD2, 6 instead of 9
      // issued instructions per two columns -- is 2.7 % slower, 2.57 against 2.50 ms.)
      uint32_t pk[64];
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
      float sy[R];
#pragma unroll
      for (int j = 0; j < R; ++j) sy[j] = 0.f;
      // 16 text columns per step
      auto smx_step = [&](const uint32_t (&r)[16], int c) {
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
      };
      if (RC_SMX_DOUBLE_BUFFER) {
        // the TMEM load of step c+1 is in flight while step c is computed (two 16-column register buffers)
        uint32_t ra[16], rb[16];
        tmem_ld_32x16(trow, ra);
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
          if (c * 16 < Kh) {
            tmem_ld_wait16(ra);
            if ((c + 1) * 16 < Kh) tmem_ld_32x16(trow + (c + 1) * 16, rb);
            smx_step(ra, c);
          }
          if ((c + 1) * 16 < Kh) {
            tmem_ld_wait16(rb);
            if ((c + 2) * 16 < Kh) tmem_ld_32x16(trow + (c + 2) * 16, ra);
            smx_step(rb, c + 1);
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (c * 16 < Kh) {
            uint32_t r[16];
            tmem_ld_32x16(trow + c * 16, r);
            tmem_ld_wait();
            smx_step(r, c);
          }
        }
      }
      tc_fence_before();
      if (warp == 4) RC_EV(lt, 11);    // exp pass done
      if (warp == 11) RC_EV(lt, 15);   // exp pass done (last softmax warp)
      arrive_leader_warp(&bars->s_empty[sidx]);      // S columns are free: the tensor pipe may start the next S in them
      float sum = (s0 + s1) + (s2 + s3);
      float sez = (q0 + q1) + (q2 + q3);
      // targets of this row: y_j, w_j (w_j = 0: ignored); tz = sum_j w_j s[y_j] over the targets in this half's columns
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
      named_bar_sync(2, 256);
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
      RC_TACC(1, tsm);
      if (!kBwd) {
        if (half == 0 && valid && prm.lse) prm.lse[m] = lse;
        continue;
      }
      {
        const float coefb = gscale * inv_wsum;
        const float coef = coefb * wtot;
        const float inv_sum = 1.f / sum;
        // P is stored pre-scaled: G[p][k] = rs_p (e_pk - sum_p [k = y_p]) = d loss / d(xhat_p . that_k) / |x_p|, so the
        // dX accumulators need no row scale and the same tile is the A operand of the dText GEMM (dT = G^T X).
        const float rsv = inv_n * inv_tau * coef * inv_sum;
        if (half == 0) {
          const float cj = coef * sez * zs * inv_sum - coefb * tz * zs;
          const float csv = inv_n * inv_n * cj;
          // the dX epilogue works on packed bf16 pixel pairs: dx = acc + (-cs) * x
          const float cs_n = __shfl_down_sync(0xffffffffu, csv, 1);
          if ((lane & 1) == 0) {
            const uint32_t c2 = pack_bf16x2(-csv, -cs_n);
            const int idx = ((lt % kScaleBufs) * 2 + (int)rank) * 64 + (row >> 1);    // [tile buffer][owner = this CTA][pixel pair]
            sc_s[idx] = c2;
            st_async_remote_b32(sc_peer + idx * 4, c2, (lt & 1) ? sc_peer1 : sc_peer0);   // peer copy: 4 tx bytes on ITS barrier
          }
          dlt_acc -= cj;
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->sc_full[lt & 1]);                // own copy (one arrival per writer warp)
          if (row == 0) mbar_arrive_expect_tx(&bars->sc_full[lt & 1], 64 * 4);   // the peer's 64 st.async land here
        }
        // the dX MMAs of the previous pair have finished reading P: store this pair's P
        ROLE_WAIT_SMX(warp == 4, &bars->p_empty, (lt & 1) ^ 1, 9, 8);
        if (warp == 4) RC_EV(lt, 12);  // P buffer free
        RC_T0(tst);
        const uint32_t rs2 = pack_bf16x2(rsv, rsv);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c * 32 < Kh) {
            const int k0 = cb + c * 32;
            uint8_t* sub = prow + (k0 >> 6) * 16384;
            const int cbase = (k0 & 32) >> 3;
#pragma unroll
            for (int g = 0; g < 4; ++g)
              *reinterpret_cast<uint4*>(sub + (((cbase + g) ^ sw) << 4)) =
                  make_uint4(bf2_mul(pk[c * 16 + g * 4], rs2), bf2_mul(pk[c * 16 + g * 4 + 1], rs2),
                             bf2_mul(pk[c * 16 + g * 4 + 2], rs2), bf2_mul(pk[c * 16 + g * 4 + 3], rs2));
          }
        }
        // G[row][y_j] = rs (e_y - sum * (weight of target y_j) / (weight of the row))  (softmax - onehot), formed before rounding
#pragma unroll
        for (int j = 0; j < R; ++j) {
          if (mine[j]) {
            const float ey = fast_exp2(fmaf(sy[j], zl, -ml));
            float gv;
            if (R == 1) {
              gv = (ey - sum) * rsv;
            } else {
              float wk = 0.f;                 // targets of the block that name the same text row
#pragma unroll
              for (int i = 0; i < R; ++i) wk += (yj[i] == yj[j]) ? wj[i] : 0.f;
              gv = (ey - sum * (wtot > 0.f ? wk / wtot : 0.f)) * rsv;      // rs = rsv carries the block's total weight
            }
            const int kk = yj[j] & 63;
            uint8_t* sub = prow + (yj[j] >> 6) * 16384;
            *reinterpret_cast<__nv_bfloat16*>(sub + (((kk >> 3) ^ sw) << 4) + (kk & 7) * 2) = __float2bfloat16_rn(gv);
          }
        }
        fence_proxy_async_smem();                 // P is read by the tensor cores (async proxy)
        arrive_leader_warp(&bars->p_full);
        if (warp == 4) RC_EV(lt, 13);  // P(lt) stored
        if (warp == 11) RC_EV(lt, 16); // P(lt) stored (last softmax warp)
        RC_TACC(2, tst);
        if (prm.store_g) {                        // dText: the finished G tile goes to global memory as it sits in smem
          named_bar_sync(6, 256);
          if (rtid == 128 && tile_ok) {
            const int gpx0 = px - row;
            for (int j = 0; j < n_kchunks; ++j) tma_store_3d(&map_g, smem + kOffP + j * 16384, j * 64, gpx0, b);
            tma_store_commit();
          }
        }
        if (half == 0 && valid && prm.lse && !(kKB && prm.lse_in != nullptr)) prm.lse[m] = lse;      // after the hand-off: nothing waits behind this store
        uint32_t sxq[2][16];
        int s_n8 = 0;
        int64_t s_off = 0;
        const bool smx_unit = RC_PAIR_SMX_UNIT && !(kKB && prm.acc_dx);
        if (smx_unit) {
          // x of unit 1 (own channel row, pixels [64,128) of the tile this warp half covers): requested now, used after the norms
          const int t1 = 2 * pj + half;
          if (t1 < prm.n_tiles) {
            const int b1 = div_tiles(prm, t1);
            const int px1 = (t1 - b1 * prm.tiles_per_img) * kTilePx + 64;
            const int d1 = (int)rank * 128 + row;
            s_off = ((int64_t)((kKB && prm.kb > 0) ? 0 : b1) * prm.D + d1) * prm.HW + px1;
            const int64_t left = prm.HW - px1;
            s_n8 = left >= 64 ? 8 : (left > 0 ? (int)(left >> 3) : 0);
          }
          const int slb = lane & 1;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int n8 = min(2, max(0, s_n8 - (c * 2 + slb) * 2));
#pragma unroll
            for (int j = 0; j < 2; ++j)
              ldg_px16(prm.x + s_off + (int64_t)(j - slb) * prm.HW + c * 32 + slb * 16, prm.wide != 0, n8, &sxq[c][j * 8], 0);
          }
        }
        if (RC_PAIR_SMX_UNIT != 2 && !RC_PAIR_NORM_WARP && pj + n_clusters < prm.n_pairs) norm_tile();
        if (warp == 4) RC_EV(lt, 14);  // row norms of the next tile done
        if (smx_unit) {
          const int slb = lane & 1;
          const uint32_t upp = 2u * (uint32_t)n_blk;
          const uint32_t uc1 = lt * upp + 1u;                                   // the epilogue's unit counter at (this pair, unit 1)
          RC_WAIT(mbar_wait, &bars->sc_full[lt & 1], (lt >> 1) & 1, 10);       // -cs of both tiles (the peer's arrive by st.async)
          const uint32_t* sc = sc_s + (lt % kScaleBufs) * 128 + half * 64;
          RC_WAIT(mbar_wait, &bars->acc_full[1], (uc1 >> 1) & 1, 11);
          tc_fence_after();
          const uint32_t tacc = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + half * 64 + 128;     // accumulator buffer 1
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t acc[32];
            tmem_ld_32x32(tacc + c * 32, acc);
            tmem_ld_wait();
            if (c == 1) { tc_fence_before(); arrive_leader_warp(&bars->acc_empty[1]); }
            {       // (row 2p+j, piece b) -> own row, pixels [c*32, +32)
#pragma unroll
              for (int r = 0; r < 8; ++r) {
                const uint32_t send = slb ? sxq[c][r] : sxq[c][8 + r];
                const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
                if (slb) sxq[c][r] = recv; else sxq[c][8 + r] = recv;
              }
            }
            const uint4* scp = reinterpret_cast<const uint4*>(sc + 32 + c * 16);          // pxh = 1
            uint32_t o[16];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 sv = scp[g];
              const uint32_t cs4[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int i = 4 * g + j;
                const uint32_t a = pack_bf16x2(__uint_as_float(acc[2 * i]), __uint_as_float(acc[2 * i + 1]));
                o[i] = bf2_fma(cs4[j], sxq[c][i], a);
              }
            }
            {       // own row -> (row 2p+j, piece b): 64 contiguous bytes per lane pair and row, straight to global memory
#pragma unroll
              for (int r = 0; r < 8; ++r) {
                const uint32_t send = slb ? o[r] : o[8 + r];
                const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
                if (slb) o[r] = recv; else o[8 + r] = recv;
              }
            }
            const int n8 = min(2, max(0, s_n8 - (c * 2 + slb) * 2));
#pragma unroll
            for (int j = 0; j < 2; ++j)
              stg_px16(prm.dx + ((kKB && prm.kb > 0) ? (int64_t)div_tiles(prm, 2 * pj + half) * prm.D * prm.HW : 0) + s_off +
                           (int64_t)(j - slb) * prm.HW + c * 32 + slb * 16, prm.wide != 0, n8, &o[j * 8]);
          }
        }
        if (RC_PAIR_SMX_UNIT == 2 && !RC_PAIR_NORM_WARP && pj + n_clusters < prm.n_pairs) norm_tile();
      }
    }
    if (kBwd && prm.store_g && rtid == 128) tma_store_wait_all0();
    if (half == 0) {
      loss_acc = warp_sum(loss_acc); w_acc = warp_sum(w_acc); dlt_acc = warp_sum(dlt_acc);
      if (lane == 0) {
        if (prm.loss_sum) atomicAdd(prm.loss_sum, (double)loss_acc);
        if (prm.w_sum) atomicAdd(prm.w_sum, (double)w_acc);
        if (kBwd && prm.dlogtau) atomicAdd(prm.dlogtau, (double)dlt_acc);
      }
    }
  } else if (kBwd) {
    // ================ dX epilogue warps: own channel rows, 64 contiguous pixels per accumulator ================
    // Warps 12-15 (half 0) take accumulator columns [0,64) = pixels of CTA 0's tile, warps 16-19 (half 1) columns
    // [64,128) = pixels of CTA 1's tile; accumulator unit (blk, pxh) covers pixels [pxh*64, +64) of both tiles.
    const int half = warp >= 16 ? 1 : 0;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t trow_acc = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + half * 64;
    const bool wide = prm.wide != 0;
    const int units_per_pair = n_blk * 2;
    // prefetch cursor: x of the NEXT 32-pixel chunk to be consumed from each of the two register buffers
    int f_pj = cluster_id, f_unit = 0;
    int64_t f_off = 0;        // element offset of (row, first pixel of the unit) in X
    int f_n8 = 0;             // valid 8-pixel groups in the 64-pixel span (0..8)
    int f_b = prm.B, f_px = 0, f_d = 0;   // TMA store coordinates of the unit (image index B = out of bounds: nothing is written)
    auto cursor_set = [&]() {
      f_n8 = 0; f_off = 0; f_b = prm.B; f_px = 0; f_d = 0;
      const int t = 2 * f_pj + half;
      if (f_pj < prm.n_pairs && t < prm.n_tiles) {
        const int b = div_tiles(prm, t);
        const int px0 = (t - b * prm.tiles_per_img) * kTilePx + (f_unit & 1) * 64;
        const int d = (f_unit >> 1) * 256 + (int)rank * 128 + row;
        f_b = b; f_px = px0; f_d = d - lane;
        f_off = ((int64_t)((kKB && prm.kb > 0) ? 0 : b) * prm.D + d) * prm.HW + px0;      // kb mode: X has one image
        const int64_t left = prm.HW - px0;
        f_n8 = left >= 64 ? 8 : (left > 0 ? (int)(left >> 3) : 0);
      }
    };
    const bool smx_unit = RC_PAIR_SMX_UNIT && !(kKB && prm.acc_dx);      // unit 1 of every pair is drained by the softmax warps
    auto cursor_next = [&]() {
      ++f_unit;
      if (smx_unit && f_unit == 1) ++f_unit;
      if (f_unit >= units_per_pair) { f_unit = 0; f_pj += n_clusters; }
      cursor_set();
    };
    // Global accesses are made by lane PAIRS: lanes 2p, 2p+1 own channel rows 2p, 2p+1; access j of a 32-pixel chunk
    // touches row 2p+j, and lane b = lane & 1 moves the b-th 16-pixel (32-byte) piece -- so one warp instruction
    // covers 16 rows x 64 contiguous bytes instead of 32 rows x 32 bytes (half the L1 tag work).  A 2x2 exchange
    // inside the pair (pair_swap) turns "row 2p+j, piece b" into "own row, piece j" and back.
    const int lb = lane & 1;
    const uint64_t pol_x = RC_EPI_X_HINT ? l2_policy_evict_first() : 0;     // last use of these X lines in this kernel
    uint32_t xq[2][16];       // chunk c of the current unit; before pair_swap: [j][8] = (row 2p+j, piece b)
    auto pair_swap = [&](uint32_t* e) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const uint32_t send = lb ? e[r] : e[8 + r];
        const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
        if (lb) e[r] = recv; else e[8 + r] = recv;
      }
    };
    auto fetch = [&](int c) {
      int n8 = min(2, max(0, f_n8 - (c * 2 + lb) * 2));      // valid 8-pixel groups of this lane's piece
      if (prm.ablate & 1) n8 = 0;
#pragma unroll
      for (int j = 0; j < 2; ++j)
        ldg_px16(prm.x + f_off + (int64_t)(j - lb) * prm.HW + c * 32 + lb * 16, wide, n8, &xq[c][j * 8], pol_x);
    };
    uint8_t* stg0 = smem + kOffStg + (warp - 12) * kStgBufs * kStgBytes;
    uint32_t sc_ = 0;                                    // chunks staged so far (buffer = parity)
    const uint64_t pol_first = l2_policy_evict_first();  // dX is write-once: keep it from displacing X in L2
    cursor_set();
    fetch(0); fetch(1);
    uint32_t uc = 0, lt = 0;
    for (int pj = cluster_id; pj < prm.n_pairs; pj += n_clusters, ++lt) {
      ROLE_WAIT_EPI(warp == 12, &bars->sc_full[lt & 1], (lt >> 1) & 1, 10, 7);
      const uint32_t* sc = sc_s + (lt % kScaleBufs) * 128 + half * 64;
      for (int unit = 0; unit < units_per_pair; ++unit, ++uc) {
        if (smx_unit && unit == 1) continue;
        const int ab = uc & 1;
        const int pxh = unit & 1;
        // where this unit's output goes (same cursor arithmetic as the prefetch, one unit behind)
        const int o_b = f_b, o_px = f_px, o_d = f_d;
        ROLE_WAIT_EPI(warp == 12, &bars->acc_full[ab], (uc >> 1) & 1, 11, 7);
        tc_fence_after();
        if (warp == 12) RC_EV(lt, 20 + unit * 2);      // accumulator of this unit complete
        RC_T0(tep);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t acc[32];
          RC_T0(t3);
          tmem_ld_32x32(trow_acc + ab * 128 + c * 32, acc);
          tmem_ld_wait();
          RC_TACC(3, t3);
          RC_T0(t13);
          if (c == 1) { tc_fence_before(); arrive_leader_warp(&bars->acc_empty[ab]); }
          RC_TACC(13, t13);
          RC_T0(t4);
          pair_swap(&xq[c][0]);             // -> own row, pixels [c*32, +32) in order
          RC_TACC(4, t4);                   // first use of the prefetched x: exposed global-load latency shows here
          RC_T0(t7);
          const uint4* scp = reinterpret_cast<const uint4*>(sc + pxh * 32 + c * 16);    // -cs2 of 16 pixel pairs
          uint32_t o[16];
#pragma unroll
          for (int g = 0; g < 4; ++g) {       // four pixel pairs per step, packed bf16x2 arithmetic: dx = acc - cs x
            const uint4 sv = scp[g];
            const uint32_t cs4[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int i = 4 * g + j;
              const uint32_t a = pack_bf16x2(__uint_as_float(acc[2 * i]), __uint_as_float(acc[2 * i + 1]));
              o[i] = bf2_fma(cs4[j], xq[c][i], a);
            }
          }
          RC_TACC(7, t7);
          // x of the NEXT unit's chunk c: requested as soon as this chunk's x registers are free (before the staging wait
          // and the store, not after them) -- the first use of the prefetched x is the hottest stall of the launch
          if (RC_EPI_FETCH_EARLY) {
            RC_T0(t6);
            if (c == 0) cursor_next();        // both chunks of the next unit are fetched relative to the advanced cursor
            fetch(c);
            RC_TACC(6, t6);
          }
          {
            // own row of the warp's [32 d][32 px] staging tile (64-byte swizzle), then one TMA store per warp
            RC_T0(t5);
            uint8_t* stg = stg0 + (kStgBufs == 2 ? (sc_ & 1u) * kStgBytes : 0u);
            ++sc_;
            if (lane == 0) {                                  // the store that last used THIS buffer has read it
              if (kStgBufs == 2) tma_store_wait_read0_keep1();
              else tma_store_wait_read0();
            }
            __syncwarp();
            RC_TACC(5, t5);
            RC_T0(t8);
            uint8_t* srow = stg + lane * 64;
            const int sw64 = (lane >> 1) & 3;
#pragma unroll
            for (int g = 0; g < ((prm.ablate & 128) ? 0 : 4); ++g)
              *reinterpret_cast<uint4*>(srow + ((g ^ sw64) << 4)) = make_uint4(o[g * 4], o[g * 4 + 1], o[g * 4 + 2], o[g * 4 + 3]);
            RC_TACC(8, t8);
            RC_T0(t9);
            fence_proxy_async_smem();
            __syncwarp();
            RC_TACC(9, t9);
            RC_T0(t12);
            if (lane == 0 && !(prm.ablate & 2)) {
              if (kKB && prm.acc_dx) tma_reduce_add_3d(&map_dx, stg, o_px + c * 32, o_d, o_b);
              else tma_store_3d_hint(&map_dx, stg, o_px + c * 32, o_d, o_b, pol_first);
              tma_store_commit();
            }
            RC_TACC(12, t12);
          }
          if (!RC_EPI_FETCH_EARLY) {
            RC_T0(t6);
            if (c == 0) cursor_next();
            fetch(c);
            RC_TACC(6, t6);
          }
        }
        RC_TACC(2, tep);
        if (warp == 12) RC_EV(lt, 21 + unit * 2);      // unit drained and stored
      }
    }
    if (lane == 0) tma_store_wait_all0();       // shared memory must stay valid until the last bulk store has read it
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
                        const double* w_sum_in, float* lse, double* loss_sum, double* w_sum, double* dlogtau, void* g_out,
                        int rep, int keep_w, const float* lse_in, int kb, int acc_dx, const int32_t* k_dev,
                        const float* log_tau_dev, cudaStream_t s) {
  using namespace pair;
  const bool bwd = dx != nullptr;
  // kb > 0: all kb blocks of 256 candidate rows in one launch -- B counts the blocks (virtual images), X has ONE image,
  // the text maps span all blocks (row offset = block * 256), dX has one [D][HW] slab per block
  const int Kp = kb > 0 ? 256 : (K + 63) / 64 * 64;
  const int Kall = kb > 0 ? kb * 256 : Kp;          // rows of the text matrices
  CUtensorMap m_xs, m_t, m_tt, m_dx, m_g;
  int rcode;
  {
    const uint64_t dims[3] = {(uint64_t)HW, (uint64_t)D, (uint64_t)B};
    const uint64_t xdims[3] = {(uint64_t)HW, (uint64_t)D, (uint64_t)(kb > 0 ? 1 : B)};
    const uint64_t str[3] = {2, (uint64_t)HW * 2, (uint64_t)D * HW * 2};
    const uint32_t box_s[3] = {64, 64, 1};
    if ((rcode = make_tmap_bf16(&m_xs, xsrc, 3, xdims, str, box_s, "pair map_x_s"))) return rcode;
    const uint32_t box_o[3] = {32, 32, 1};
    if ((rcode = make_tmap_bf16(&m_dx, bwd ? dx : xsrc, 3, bwd ? dims : xdims, str, box_o, "pair map_dx"))) return rcode;
    const uint64_t tdims[2] = {(uint64_t)D, (uint64_t)Kall}, tstr[2] = {2, (uint64_t)D * 2};
    const uint32_t tbox[2] = {64, (uint32_t)(Kp / 2)};
    if ((rcode = make_tmap_bf16(&m_t, t_bf16, 2, tdims, tstr, tbox, "pair map_t"))) return rcode;
    const uint64_t ttdims[2] = {(uint64_t)Kall, (uint64_t)D}, ttstr[2] = {2, (uint64_t)Kall * 2};
    const uint32_t ttbox[2] = {64, 128};
    if ((rcode = make_tmap_bf16(&m_tt, bwd ? tt_bf16 : t_bf16, 2, bwd ? ttdims : tdims, bwd ? ttstr : tstr,
                                bwd ? ttbox : tbox, "pair map_tt"))) return rcode;
  }
  {
    // G [B][HW][Kp] bf16: a tile is stored exactly as it sits in shared memory (four K-major 128B-swizzled sub-tiles)
    const uint64_t gdims[3] = {(uint64_t)Kp, (uint64_t)HW, (uint64_t)B};
    const uint64_t gstr[3] = {2, (uint64_t)Kp * 2, (uint64_t)HW * Kp * 2};
    const uint32_t gbox[3] = {64, 128, 1};
    if (g_out != nullptr) {
      if ((rcode = make_tmap_bf16(&m_g, g_out, 3, gdims, gstr, gbox, "pair map_g"))) return rcode;
    } else {
      m_g = m_dx;
    }
  }
  Params prm;
  prm.store_g = g_out != nullptr;
  prm.dbg = debug_timing_buffer();
  prm.B = B; prm.D = D; prm.K = K; prm.Kp = Kp; prm.HW = HW;
  prm.tiles_per_img = (int)((HW + kTilePx - 1) / kTilePx);
  if ((int64_t)B * prm.tiles_per_img > 0x3fffffff) return fail(RC_ERR_UNSUPPORTED, "rc_infonce_bf16: too many tiles");
  prm.n_tiles = B * prm.tiles_per_img;
  prm.tpi_magic = prm.tiles_per_img == 1 ? 0xffffffffu : (uint32_t)(0x100000000ull / (uint64_t)prm.tiles_per_img);
  prm.n_pairs = (prm.n_tiles + 1) / 2;
  prm.x = reinterpret_cast<const __nv_bfloat16*>(xsrc);
  prm.dx = reinterpret_cast<__nv_bfloat16*>(dx);
  prm.ablate = 0;
  prm.split_c = prm.split_b = -1;
#ifdef RC_BRINGUP      // ablation / issue-order switches of the bring-up build; the shipped library has no environment knobs
  { const char* ab = getenv("RANGECLIP_B200_ABLATE"); prm.ablate = ab ? atoi(ab) : 0; }
  if (const char* sp = getenv("RANGECLIP_B200_SPLIT")) sscanf(sp, "%d,%d", &prm.split_c, &prm.split_b);
#endif
  prm.wide = (HW % 16 == 0) && (reinterpret_cast<uintptr_t>(xsrc) % 32 == 0) && (reinterpret_cast<uintptr_t>(dx) % 32 == 0);
  prm.inv_norm = inv_norm; prm.y = y; prm.w = w; prm.inv_tau = inv_tau; prm.grad_scale = grad_scale;
  prm.w_sum_in = w_sum_in; prm.lse = lse; prm.loss_sum = loss_sum; prm.w_sum = w_sum; prm.dlogtau = dlogtau;
  prm.keep_w = keep_w; prm.lse_in = lse_in; prm.kb = kb; prm.acc_dx = acc_dx;
  prm.k_dev = k_dev; prm.log_tau_dev = log_tau_dev;
  if (kb > 0 && (prm.tiles_per_img & 1)) return fail(RC_ERR_UNSUPPORTED, "rc_infonce_bf16_kblocks: HW must be a multiple of 256");
  int n_clusters = num_sms() / 2;
  if (n_clusters > prm.n_pairs) n_clusters = prm.n_pairs;
  const int grid = 2 * n_clusters;
  auto launch = [&](auto kernel) -> int {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return fail(RC_ERR_CUDA, "rc_infonce_bf16(pair): smem opt-in: %s", cudaGetErrorString(e));
    kernel<<<grid, kThreads, kSmemBytes, s>>>(m_xs, m_t, m_tt, m_dx, m_g, prm);
    return check_launch("rc_infonce_bf16(pair)");
  };
  if (rep == 4) return bwd ? launch(infonce_umma_pair_kernel<true, 4, false>) : launch(infonce_umma_pair_kernel<false, 4, false>);
  if (keep_w || lse_in != nullptr || kb > 0 || acc_dx) return bwd ? launch(infonce_umma_pair_kernel<true, 1, true>) : launch(infonce_umma_pair_kernel<false, 1, true>);
  return bwd ? launch(infonce_umma_pair_kernel<true, 1, false>) : launch(infonce_umma_pair_kernel<false, 1, false>);
}

}  // namespace rc
