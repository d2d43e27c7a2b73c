// L2-normalised pixel rows as an operator of its own: x_hat[b][:,p] = x[b][:,p] / max(|x[b][:,p]|, 1e-12) over the channel
// dimension of an NCHW tensor (F.normalize(x, p=2, dim=1): the decoder tail, utils/src/decoder.py:114) and its backward
//   dx = (g - x_hat (x_hat . g)) / max(|x|, 1e-12).
// compute_loss_shared2x2 takes the decoder's output_conv result BEFORE that tail; the smoothness term and the
// area pooling need the normalised rows, and eager PyTorch spends ~10 full-tensor passes on normalize + its autograd.
// HBM-bound column walk: a thread owns 8 consecutive pixels and walks the D channels twice (reduce, then write).
#include "common.cuh"

namespace rc {

template <typename T>
__global__ void __launch_bounds__(256)
normalize_rows_fwd_kernel(const T* __restrict__ x, int B, int D, int64_t HW, float* __restrict__ out, float* __restrict__ inv_norm) {
  const int64_t gpi = HW / 8;
  const int64_t n = (int64_t)B * gpi;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = g / gpi;
    const int64_t p0 = (g - b * gpi) * 8;
    const T* src = x + b * (int64_t)D * HW + p0;
    float* dst = out + b * (int64_t)D * HW + p0;
    float ss[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ss[j] = 0.f;
#pragma unroll 4
    for (int d = 0; d < D; ++d) {
      float v[8];
      load8(src + (int64_t)d * HW, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss[j] = fmaf(v[j], v[j], ss[j]);
    }
    float inv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) inv[j] = 1.f / fmaxf(sqrtf(ss[j]), 1e-12f);
    if (inv_norm) store8(inv_norm + b * HW + p0, inv);
#pragma unroll 4
    for (int d = 0; d < D; ++d) {
      float v[8];
      load8(src + (int64_t)d * HW, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= inv[j];
      store8(dst + (int64_t)d * HW, v);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
normalize_rows_bwd_kernel(const float* __restrict__ xhat, const float* __restrict__ gin, const float* __restrict__ inv_norm, int B, int D,
                          int64_t HW, T* __restrict__ dx) {
  const int64_t gpi = HW / 8;
  const int64_t n = (int64_t)B * gpi;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = g / gpi;
    const int64_t p0 = (g - b * gpi) * 8;
    const int64_t off = b * (int64_t)D * HW + p0;
    float dot[8], inv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dot[j] = 0.f;
#pragma unroll 4
    for (int d = 0; d < D; ++d) {
      float a[8], c[8];
      load8(xhat + off + (int64_t)d * HW, a);
      load8(gin + off + (int64_t)d * HW, c);
#pragma unroll
      for (int j = 0; j < 8; ++j) dot[j] = fmaf(a[j], c[j], dot[j]);
    }
    load8(inv_norm + b * HW + p0, inv);
#pragma unroll 4
    for (int d = 0; d < D; ++d) {
      float a[8], c[8];
      load8(xhat + off + (int64_t)d * HW, a);
      load8(gin + off + (int64_t)d * HW, c);
#pragma unroll
      for (int j = 0; j < 8; ++j) c[j] = (c[j] - a[j] * dot[j]) * inv[j];
      store8(dx + off + (int64_t)d * HW, c);
    }
  }
}

// Backward of the smoothness term THROUGH the normalisation, in one kernel:  d/dx [ s_h * sum |xh[.,w] - xh[.,w+1]| + s_v * sum |xh[h,.] -
// xh[h+1,.]| ] with xh = x / max(|x|, 1e-12) over the channels (model.py:332-334 on the decoder tail's output, decoder.py:114).
// rc_tv_bwd would materialise g = d(TV)/d(xh) (the size of xh) and rc_normalize_rows_bwd would read xh and g twice; here a thread
// owns 8 consecutive pixels of one image row, forms g for a channel from the saved xh of its own row and the rows above / below
// (neighbouring threads' rows: L1 hits -- the kernel uses no shared memory), walks the channels once for x_hat . g and once more
// for dx = (g - x_hat (x_hat . g)) / |x|.  sign(0) = 0 as in rc_tv_bwd (the 2x2-shared pixels of the upsampled map, quirk Q8).
__device__ __forceinline__ float sgnf(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

template <typename T>
__global__ void __launch_bounds__(256)
tv_normalize_bwd_kernel(const float* __restrict__ xhat, const float* __restrict__ inv_norm, const float* __restrict__ scale, int B, int D,
                        int H, int W, T* __restrict__ dx) {
  const int gw = W / 8;                       // 8-pixel groups per image row
  const int64_t HW = (int64_t)H * W;
  const int64_t n = (int64_t)B * H * gw;
  const float sh = scale[0], sv = scale[1];
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (int64_t)gridDim.x * blockDim.x) {
    const int gx = (int)(g % gw);
    const int64_t rest = g / gw;
    const int h = (int)(rest % H);
    const int64_t b = rest / H;
    const int w0 = gx * 8;
    const int64_t off = b * (int64_t)D * HW + (int64_t)h * W + w0;
    const bool has_up = h > 0, has_dn = h + 1 < H, has_l = w0 > 0, has_r = w0 + 8 < W;
    // g of the lane's 8 pixels for the channel plane at `p`
    auto grad8 = [&](const float* p, float (&c)[8], float (&gv)[8]) {
      load8(p, c);
      float up[8], dn[8];
      if (has_up) load8(p - W, up);
      if (has_dn) load8(p + W, dn);
      const float left = has_l ? __ldg(p - 1) : 0.f, right = has_r ? __ldg(p + 8) : 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a = 0.f;
        const bool hr = j < 7 || has_r, hl = j > 0 || has_l;
        if (hr) a += sh * sgnf(c[j] - (j < 7 ? c[j + 1] : right));           // d |x[w] - x[w+1]| / d x[w]
        if (hl) a -= sh * sgnf((j > 0 ? c[j - 1] : left) - c[j]);            // d |x[w-1] - x[w]| / d x[w]
        if (has_dn) a += sv * sgnf(c[j] - dn[j]);
        if (has_up) a -= sv * sgnf(up[j] - c[j]);
        gv[j] = a;
      }
    };
    float dot[8], inv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dot[j] = 0.f;
#pragma unroll 2
    for (int d = 0; d < D; ++d) {
      float c[8], gv[8];
      grad8(xhat + off + (int64_t)d * HW, c, gv);
#pragma unroll
      for (int j = 0; j < 8; ++j) dot[j] = fmaf(c[j], gv[j], dot[j]);
    }
    load8(inv_norm + b * HW + (int64_t)h * W + w0, inv);
#pragma unroll 2
    for (int d = 0; d < D; ++d) {
      float c[8], gv[8];
      grad8(xhat + off + (int64_t)d * HW, c, gv);
#pragma unroll
      for (int j = 0; j < 8; ++j) gv[j] = (gv[j] - c[j] * dot[j]) * inv[j];
      store8(dx + off + (int64_t)d * HW, gv);
    }
  }
}

static int rows_grid(int B, int64_t HW) {
  const int64_t blocks = ((int64_t)B * (HW / 8) + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace rc

extern "C" int rc_normalize_rows_fwd(const void* x, rc_dtype dtype, int B, int D, int64_t HW, float* out, float* inv_norm, void* stream) {
  using namespace rc;
  RC_REQUIRE(B >= 0 && D >= 1 && HW >= 0, "rc_normalize_rows_fwd: bad shape");
  if (B == 0 || HW == 0) return RC_OK;
  RC_REQUIRE(x && out, "rc_normalize_rows_fwd: null pointer");
  if (HW % 8 != 0) return fail(RC_ERR_UNSUPPORTED, "rc_normalize_rows_fwd: HW=%lld must be a multiple of 8", (long long)HW);
  RC_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(inv_norm)) & 31) == 0,
             "rc_normalize_rows_fwd: pointers must be 32-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == RC_F32) normalize_rows_fwd_kernel<float><<<rows_grid(B, HW), 256, 0, s>>>((const float*)x, B, D, HW, out, inv_norm);
  else normalize_rows_fwd_kernel<__nv_bfloat16><<<rows_grid(B, HW), 256, 0, s>>>((const __nv_bfloat16*)x, B, D, HW, out, inv_norm);
  return check_launch("rc_normalize_rows_fwd");
}

extern "C" int rc_normalize_rows_bwd(const float* xhat, const float* g, const float* inv_norm, rc_dtype dtype, int B, int D, int64_t HW,
                                     void* dx, void* stream) {
  using namespace rc;
  RC_REQUIRE(B >= 0 && D >= 1 && HW >= 0, "rc_normalize_rows_bwd: bad shape");
  if (B == 0 || HW == 0) return RC_OK;
  RC_REQUIRE(xhat && g && inv_norm && dx, "rc_normalize_rows_bwd: null pointer");
  if (HW % 8 != 0) return fail(RC_ERR_UNSUPPORTED, "rc_normalize_rows_bwd: HW=%lld must be a multiple of 8", (long long)HW);
  RC_REQUIRE(((reinterpret_cast<uintptr_t>(xhat) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(dx) |
               reinterpret_cast<uintptr_t>(inv_norm)) & 31) == 0, "rc_normalize_rows_bwd: pointers must be 32-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == RC_F32)
    normalize_rows_bwd_kernel<float><<<rows_grid(B, HW), 256, 0, s>>>(xhat, g, inv_norm, B, D, HW, (float*)dx);
  else
    normalize_rows_bwd_kernel<__nv_bfloat16><<<rows_grid(B, HW), 256, 0, s>>>(xhat, g, inv_norm, B, D, HW, (__nv_bfloat16*)dx);
  return check_launch("rc_normalize_rows_bwd");
}

extern "C" int rc_tv_normalize_bwd(const float* xhat, const float* inv_norm, const float* scale, rc_dtype dtype, int B, int D, int H, int W,
                                   void* dx, void* stream) {
  using namespace rc;
  RC_REQUIRE(B >= 0 && D >= 1 && H >= 0 && W >= 0, "rc_tv_normalize_bwd: bad shape");
  if (B == 0 || H == 0 || W == 0) return RC_OK;
  RC_REQUIRE(xhat && inv_norm && scale && dx, "rc_tv_normalize_bwd: null pointer");
  if (W % 8 != 0) return fail(RC_ERR_UNSUPPORTED, "rc_tv_normalize_bwd: W=%d must be a multiple of 8", W);
  RC_REQUIRE(((reinterpret_cast<uintptr_t>(xhat) | reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(inv_norm)) & 31) == 0,
             "rc_tv_normalize_bwd: pointers must be 32-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = rows_grid(B, (int64_t)H * W);
  if (dtype == RC_F32) tv_normalize_bwd_kernel<float><<<grid, 256, 0, s>>>(xhat, inv_norm, scale, B, D, H, W, (float*)dx);
  else tv_normalize_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(xhat, inv_norm, scale, B, D, H, W, (__nv_bfloat16*)dx);
  return check_launch("rc_tv_normalize_bwd");
}
