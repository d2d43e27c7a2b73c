// Device-side contrast-set builder: model.py:234-268 without a host round trip.
//
// The reference builds the contrast set of the pixel-text InfoNCE on the host: torch.unique of the sampled labels ->
// .tolist() -> Python set logic over the similarity tables -> np.random.choice / torch.randperm -> torch.unique of the
// concatenation.  Every step of that is a device->host synchronisation in front of the fused loss kernel.  This kernel does
// the same construction from the label histogram of the sampled pixels (rc_sample_label_counts) in ONE launch of one block:
//   present   = labels >= 1 with a non-zero count                                  (model.py:226,233; include_label0: >= 0,
//               the candidate set of predict, model.py:147-156, keeps the background label)
//   candidate = union of the similarity lists of the present labels, minus present  (model.py:240-252)
//   chosen    = n_curriculum candidates drawn without replacement (all of them if there are fewer)   (model.py:254-259)
//   random    = n_rand labels drawn without replacement from everything not present / chosen         (model.py:261-266)
//   contrast  = sorted union (torch.unique, model.py:268); label_map[c] = position of c in it or -1   (model.py:276-277)
// Draws are "the n smallest of independent 64-bit keys", key(c) = splitmix64(seed, phase, c): a counter-based stream, so the
// result depends only on (seed, inputs) -- not the reference's NumPy / CPU-torch streams (parity is statistical there, bit-exact
// against oracle/rangeclip_oracle.py:contrast_build_device which restates this file).  The set is capped at k_cap rows
// (distractors are trimmed first; present labels beyond the cap are dropped from the map and flagged).
#include "common.cuh"

namespace rc {
namespace contrast {

constexpr int kThreads = 1024;
constexpr int kBins = 2048;          // 11 bits per radix pass
constexpr int kMaxC = 16384;
enum : uint8_t { kPresent = 1, kCand = 2, kChosen = 4, kRand = 8 };

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ uint64_t draw_key(uint64_t seed, uint32_t phase, uint32_t c) {
  return splitmix64(splitmix64(seed ^ ((uint64_t)phase << 56)) + (uint64_t)c);
}

struct Shared {
  int hist[kBins];
  int warp_tot[32];
  int result[4];
};

// block-wide sum of one int per thread
__device__ __forceinline__ int block_sum(int v, Shared& sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh.warp_tot[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) t += sh.warp_tot[i];
  return t;
}

// exclusive prefix over the block of one int per thread (thread order); total in *total
__device__ __forceinline__ int block_excl_scan(int v, Shared& sh, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  __syncthreads();
  if (lane == 31) sh.warp_tot[wid] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) {
    const int w = sh.warp_tot[i];
    if (i < wid) base += w;
    tot += w;
  }
  *total = tot;
  return base + inc - v;
}

// Marks (state |= out_flag) the `n_take` elements with the smallest keys among those with (state & any_of) != 0 and
// (state & none_of) == 0.  n_take must not exceed the number of eligible elements (the caller clamps); n_take == all: no draw.
__device__ void select_smallest(uint8_t* state, int C, uint8_t any_of, uint8_t none_of, bool all_states_eligible,
                                int n_eligible, int n_take, uint64_t seed, uint32_t phase, uint8_t out_flag, Shared& sh) {
  auto eligible = [&](int c) -> bool {
    const uint8_t s = state[c];
    return (all_states_eligible || (s & any_of)) && !(s & none_of);
  };
  if (n_take <= 0) return;
  if (n_take >= n_eligible) {
    for (int c = threadIdx.x; c < C; c += kThreads) if (eligible(c)) state[c] |= out_flag;
    __syncthreads();
    return;
  }
  // radix select on the top 44 key bits: find the prefix T with #(key44 < T) < n_take <= #(key44 <= T)
  uint64_t prefix = 0;        // the digits fixed so far (high bits)
  int need = n_take;          // rank wanted inside the current prefix class
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 53 - 11 * pass;
    for (int i = threadIdx.x; i < kBins; i += kThreads) sh.hist[i] = 0;
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kThreads) {
      if (!eligible(c)) continue;
      const uint64_t key = draw_key(seed, phase, (uint32_t)c);
      if (pass == 0 || (key >> (shift + 11)) == prefix) atomicAdd(&sh.hist[(int)((key >> shift) & (kBins - 1))], 1);
    }
    __syncthreads();
    // bins 2t, 2t+1 per thread: the bin holding the need-th smallest
    const int h0 = sh.hist[2 * threadIdx.x], h1 = sh.hist[2 * threadIdx.x + 1];
    int tot;
    const int before = block_excl_scan(h0 + h1, sh, &tot);
    if (need > before && need <= before + h0) { sh.result[0] = 2 * threadIdx.x; sh.result[1] = before; }
    else if (need > before + h0 && need <= before + h0 + h1) { sh.result[0] = 2 * threadIdx.x + 1; sh.result[1] = before + h0; }
    __syncthreads();
    prefix = (prefix << 11) | (uint64_t)sh.result[0];
    need -= sh.result[1];
    __syncthreads();
  }
  // everything below the 44-bit prefix is taken; inside the prefix class (one element unless 44 key bits collide) the
  // first `need` by label order
  if (threadIdx.x == 0) sh.result[2] = 0;
  __syncthreads();
  for (int c0 = 0; c0 < C; c0 += kThreads) {        // label order: block-ordered claim of the tie class
    const int c = c0 + threadIdx.x;
    bool tie = false;
    if (c < C && eligible(c)) {
      const uint64_t k44 = draw_key(seed, phase, (uint32_t)c) >> 20;
      if (k44 < prefix) state[c] |= out_flag;
      tie = k44 == prefix;
    }
    int tot;
    const int rank = block_excl_scan(tie ? 1 : 0, sh, &tot);
    const int taken = sh.result[2];
    if (tie && taken + rank < need) state[c] |= out_flag;
    __syncthreads();
    if (threadIdx.x == 0) sh.result[2] = taken + tot;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads, 1)
contrast_build_kernel(const int32_t* __restrict__ counts, int C, const int32_t* __restrict__ sim_off,
                      const int32_t* __restrict__ sim_items, int n_curriculum, int n_rand, int k_cap, uint64_t seed,
                      const int64_t* __restrict__ seed_dev, int first_label, int32_t* __restrict__ label_map,
                      int64_t* __restrict__ contrast_out, int32_t* __restrict__ k_out) {
  extern __shared__ uint8_t state[];        // [C] flags
  __shared__ Shared sh;
  if (seed_dev != nullptr) seed ^= (uint64_t)seed_dev[0];      // a seed that lives on the device (CUDA-graph replays: a new draw each time)
  for (int c = threadIdx.x; c < C; c += kThreads) state[c] = (c >= first_label && counts[c] > 0) ? kPresent : 0;
  __syncthreads();
  if (sim_off != nullptr && n_curriculum > 0) {
    // a warp per present label, its lanes along the similarity list (coalesced; one thread per label walks a list of a
    // hundred entries as a chain of dependent loads)
    const int lane = threadIdx.x & 31;
    for (int c = threadIdx.x >> 5; c < C; c += kThreads / 32) {
      if (!(state[c] & kPresent)) continue;
      const int end = sim_off[c + 1];
      for (int j = sim_off[c] + lane; j < end; j += 32) {
        const int d = sim_items[j];
        if (d >= 0 && d < C && !(state[d] & kPresent)) state[d] = kCand;      // benign race: every writer stores kCand
      }
    }
    __syncthreads();
  }
  int np = 0, nc = 0;
  for (int c = threadIdx.x; c < C; c += kThreads) { np += (state[c] & kPresent) ? 1 : 0; nc += (state[c] & kCand) ? 1 : 0; }
  const int n_present = block_sum(np, sh);
  const int n_cand = block_sum(nc, sh);
  // how many distractors fit: the reference takes min(wanted, available); the cap trims the random part first
  int flags = 0;
  int n_keep_present = n_present;
  if (n_present > k_cap) { n_keep_present = k_cap; flags |= 1; }
  int take_cur = n_cand >= n_curriculum ? n_curriculum : n_cand;        // model.py:254-259
  int room = k_cap - n_keep_present;
  if (take_cur > room) { take_cur = room; flags |= 2; }
  room -= take_cur;
  int n_free = C - n_present - take_cur;                                // model.py:261-264 (label 0 included)
  int take_rand = n_rand < n_free ? n_rand : n_free;
  if (take_rand < 0) take_rand = 0;
  if (take_rand > room) { take_rand = room; flags |= 2; }
  select_smallest(state, C, kCand, 0, false, n_cand, take_cur, seed, 1u, kChosen, sh);
  select_smallest(state, C, 0, kPresent | kChosen, true, n_free, take_rand, seed, 2u, kRand, sh);
  // sorted union + map: contiguous chunk of labels per thread, block prefix of the member counts
  const int per = (C + kThreads - 1) / kThreads;
  const int c_begin = threadIdx.x * per, c_end = min(C, c_begin + per);
  int mine = 0;
  for (int c = c_begin; c < c_end; ++c) mine += (state[c] & (kPresent | kChosen | kRand)) ? 1 : 0;
  int total;
  int rank = block_excl_scan(mine, sh, &total);
  for (int c = c_begin; c < c_end; ++c) {
    const bool member = (state[c] & (kPresent | kChosen | kRand)) != 0;
    int32_t pos = -1;
    if (member) {
      if (rank < k_cap) { pos = rank; contrast_out[rank] = c; }
      ++rank;
    }
    label_map[c] = pos;
  }
  const int K = total < k_cap ? total : k_cap;
  for (int i = K + threadIdx.x; i < k_cap; i += kThreads) contrast_out[i] = -1;      // pad rows (rc_text_prepare: zero rows)
  if (threadIdx.x == 0) { k_out[0] = K; k_out[1] = flags; k_out[2] = n_present; k_out[3] = take_cur + take_rand; }
}

}  // namespace contrast
}  // namespace rc

extern "C" int rc_contrast_build(const int32_t* counts, int C, const int32_t* sim_off, const int32_t* sim_items,
                                 int n_curriculum, int n_rand, int k_cap, uint64_t seed, const int64_t* seed_dev,
                                 int include_label0, int32_t* label_map, int64_t* contrast, int32_t* k_out, void* stream) {
  using namespace rc::contrast;
  RC_REQUIRE(counts && label_map && contrast && k_out, "rc_contrast_build: null pointer");
  RC_REQUIRE(C >= 1 && k_cap >= 1 && n_curriculum >= 0 && n_rand >= 0, "rc_contrast_build: bad argument");
  RC_REQUIRE((sim_off == nullptr) == (sim_items == nullptr), "rc_contrast_build: sim_off and sim_items go together");
  if (C > kMaxC) return rc::fail(RC_ERR_UNSUPPORTED, "rc_contrast_build: C=%d labels exceed %d", C, kMaxC);
  contrast_build_kernel<<<1, kThreads, (size_t)C, (cudaStream_t)stream>>>(counts, C, sim_off, sim_items, n_curriculum, n_rand,
                                                                         k_cap, seed, seed_dev, include_label0 ? 0 : 1, label_map, contrast, k_out);
  return rc::check_launch("rc_contrast_build");
}
