// K5/K6: smoothness (TV-L1) forward sums and backward, shared-memory tiled.
// Replaces RangeCLIP/src/depth_segmentation_model/model.py:332-334 (two F.l1_loss over four
// strided slices) and its autograd (sgn, mul, two slice-scatter adds) with one read of X for the
// forward and one read of X (+ one read-modify-write of dX) for the backward.  sign(0) = 0 as
// torch.sgn (the nearest-upsampled decoder makes half of all differences exactly 0).
#include "common.cuh"

namespace rc {

constexpr int kTvThreads = 256;
constexpr int kTvSmemFloats = 12288;  // 48 KB tile budget (rows incl. halo) x W

template <typename T>
__device__ __forceinline__ void tv_load_rows(const T* __restrict__ plane, int H, int W, int h_first, int n_rows,
                                             float* __restrict__ tile, bool vec_ok) {
  // rows h_first .. h_first+n_rows-1 (clamped to the plane; rows outside are left untouched)
  const int lo = h_first < 0 ? 0 : h_first;
  const int hi = (h_first + n_rows) > H ? H : (h_first + n_rows);
  if (hi <= lo) return;
  const T* src = plane + (int64_t)lo * W;
  float* dst = tile + (int64_t)(lo - h_first) * W;
  const int total = (hi - lo) * W;
  if (vec_ok) {
    for (int i = threadIdx.x * 8; i < total; i += kTvThreads * 8) {
      float v[8];
      load8(src + i, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[i + j] = v[j];
    }
  } else {
    for (int i = threadIdx.x; i < total; i += kTvThreads) dst[i] = ElemIO<T>::ld(src + i);
  }
}

template <typename T>
__global__ void __launch_bounds__(kTvThreads)
tv_fwd_kernel(const T* __restrict__ x, int64_t planes, int H, int W, int TH, double* __restrict__ sums) {
  extern __shared__ float tile[];  // [(TH+1)][W]
  const int tiles_per_plane = (H + TH - 1) / TH;
  const int64_t n_tiles = planes * tiles_per_plane;
  const bool vec_ok = (W % 8 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  double acc_h = 0.0, acc_v = 0.0;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t pl = t / tiles_per_plane;
    const int h0 = (int)(t - pl * tiles_per_plane) * TH;
    const int rows = min(TH, H - h0);
    __syncthreads();
    tv_load_rows(x + pl * (int64_t)H * W, H, W, h0, rows + 1, tile, vec_ok);
    __syncthreads();
    const bool has_below = (h0 + rows) < H;
    float sh = 0.f, sv = 0.f;
    for (int i = threadIdx.x; i < rows * W; i += kTvThreads) {
      const int r = i / W, c = i - r * W;
      const float a = tile[i];
      if (c + 1 < W) sh += fabsf(a - tile[i + 1]);
      if (r + 1 < rows || has_below) sv += fabsf(a - tile[i + W]);
    }
    acc_h += (double)sh;
    acc_v += (double)sv;
  }
  acc_h = warp_sum(acc_h);
  acc_v = warp_sum(acc_v);
  __shared__ double red[2][kTvThreads / 32];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][wid] = acc_h; red[1][wid] = acc_v; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < kTvThreads / 32; ++i) { a += red[0][i]; b += red[1][i]; }
    atomicAdd(&sums[0], a);
    atomicAdd(&sums[1], b);
  }
}

__device__ __forceinline__ float sgnf(float v) { return (float)((v > 0.f) - (v < 0.f)); }

template <typename T>
__global__ void __launch_bounds__(kTvThreads)
tv_bwd_kernel(const T* __restrict__ x, int64_t planes, int H, int W, int TH, const float* __restrict__ scale,
              T* __restrict__ dx, int accumulate, const float* __restrict__ dx_scale) {
  extern __shared__ float tile[];  // [(TH+2)][W], row 0 = h0-1
  const int tiles_per_plane = (H + TH - 1) / TH;
  const int64_t n_tiles = planes * tiles_per_plane;
  const bool vec_ok = (W % 8 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  const float sh = scale[0], sv = scale[1];
  const float ds = (dx_scale != nullptr) ? dx_scale[0] : 1.f;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t pl = t / tiles_per_plane;
    const int h0 = (int)(t - pl * tiles_per_plane) * TH;
    const int rows = min(TH, H - h0);
    __syncthreads();
    tv_load_rows(x + pl * (int64_t)H * W, H, W, h0 - 1, rows + 2, tile, vec_ok);
    __syncthreads();
    T* out = dx + pl * (int64_t)H * W + (int64_t)h0 * W;
    const float* ctr = tile + W;
    for (int i = threadIdx.x; i < rows * W; i += kTvThreads) {
      const int r = i / W, c = i - r * W;
      const int h = h0 + r;
      const float a = ctr[i];
      float g = 0.f;
      if (c + 1 < W) g += sh * sgnf(a - ctr[i + 1]);
      if (c >= 1) g -= sh * sgnf(ctr[i - 1] - a);
      if (h + 1 < H) g += sv * sgnf(a - ctr[i + W]);
      if (h >= 1) g -= sv * sgnf(ctr[i - W] - a);
      if (accumulate) g += ds * ElemIO<T>::ld(out + i);
      ElemIO<T>::st(out + i, g);
    }
  }
}

static int tv_tile_rows(int H, int W, int halo) {
  int r = kTvSmemFloats / W - halo;
  if (r > 32) r = 32;
  if (r > H) r = H;
  return r;
}

}  // namespace rc

extern "C" int rc_tv_fwd(const void* x, rc_dtype x_dtype, int64_t planes, int H, int W, double* sums, void* stream) {
  RC_REQUIRE(x && sums, "rc_tv_fwd: null pointer");
  RC_REQUIRE(planes >= 0 && H >= 1 && W >= 1, "rc_tv_fwd: bad shape");
  if (planes == 0) return RC_OK;
  const int TH = rc::tv_tile_rows(H, W, 1);
  if (TH < 1) return rc::fail(RC_ERR_UNSUPPORTED, "rc_tv_fwd: W=%d too wide for the 48 KB tile", W);
  const int64_t n_tiles = planes * ((H + TH - 1) / TH);
  const int64_t cap = (int64_t)rc::num_sms() * 8;
  const int grid = (int)(n_tiles < cap ? n_tiles : cap);
  const size_t smem = (size_t)(TH + 1) * W * sizeof(float);
  cudaStream_t s = (cudaStream_t)stream;
  if (x_dtype == RC_F32)
    rc::tv_fwd_kernel<float><<<grid, rc::kTvThreads, smem, s>>>((const float*)x, planes, H, W, TH, sums);
  else
    rc::tv_fwd_kernel<__nv_bfloat16><<<grid, rc::kTvThreads, smem, s>>>((const __nv_bfloat16*)x, planes, H, W, TH, sums);
  return rc::check_launch("rc_tv_fwd");
}

extern "C" int rc_tv_bwd(const void* x, rc_dtype x_dtype, int64_t planes, int H, int W, const float* scale,
                         void* dx, int accumulate, const float* dx_scale, void* stream) {
  RC_REQUIRE(x && dx && scale, "rc_tv_bwd: null pointer");
  RC_REQUIRE(planes >= 0 && H >= 1 && W >= 1, "rc_tv_bwd: bad shape");
  if (planes == 0) return RC_OK;
  const int TH = rc::tv_tile_rows(H, W, 2);
  if (TH < 1) return rc::fail(RC_ERR_UNSUPPORTED, "rc_tv_bwd: W=%d too wide for the 48 KB tile", W);
  const int64_t n_tiles = planes * ((H + TH - 1) / TH);
  const int64_t cap = (int64_t)rc::num_sms() * 8;
  const int grid = (int)(n_tiles < cap ? n_tiles : cap);
  const size_t smem = (size_t)(TH + 2) * W * sizeof(float);
  cudaStream_t s = (cudaStream_t)stream;
  if (x_dtype == RC_F32)
    rc::tv_bwd_kernel<float><<<grid, rc::kTvThreads, smem, s>>>((const float*)x, planes, H, W, TH, scale, (float*)dx, accumulate, dx_scale);
  else
    rc::tv_bwd_kernel<__nv_bfloat16><<<grid, rc::kTvThreads, smem, s>>>((const __nv_bfloat16*)x, planes, H, W, TH, scale,
                                                                         (__nv_bfloat16*)dx, accumulate, dx_scale);
  return rc::check_launch("rc_tv_bwd");
}
