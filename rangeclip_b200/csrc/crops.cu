// Front half of prepare_image_contrast_data (dataloader.py:238-282): the object crops of a batch, resized /
// centre-cropped / normalised for the CLIP image encoder, in ONE launch.
//
// The reference slices `image[:, ymin:ymax, xmin:xmax]` per object and hands the list to `clip_processor(images=crops,
// return_tensors="pt", padding=True, do_rescale=False)` (dataloader.py:254,276).  With the torchvision backend of
// transformers (5.5.0: image_processing_backends.py, TorchvisionBackend._preprocess) that is, per crop:
//   resize     shortest edge -> S, longest -> int(S * long / short)       (get_resize_output_image_size, default_to_square=False)
//              bicubic, antialias=True, align_corners=False               (tvF.resize -> ATen upsample_bicubic2d_aa)
//   centre crop to Sc x Sc at top = int((rh - Sc) / 2.0), left = int((rw - Sc) / 2.0)   (TorchvisionBackend.center_crop)
//   normalise  (v - mean[c]) / std[c] in fp32                              (tvF.normalize; rescaling is off: do_rescale=False)
// -- one resize launch pair per distinct crop shape, a crop, a normalisation and a stack.  Here a thread owns one output
// value: it evaluates the separable antialiased bicubic filter (ATen's arithmetic: fp32 scale / centre / weights, the cubic
// convolution kernel with a = -0.5, support 2 * max(scale, 1), weights normalised by their sum; horizontal pass first)
// straight from the un-cropped image, so neither the crops nor the resized intermediates exist in memory.
#include "common.cuh"

namespace rc {
namespace crops {

__device__ __forceinline__ float cubic_aa(float x) {        // ATen upsample_bicubic2d_aa filter (a = -0.5)
  const float a = -0.5f;
  x = fabsf(x);
  if (x < 1.0f) return ((a + 2.0f) * x - (a + 3.0f)) * x * x + 1.0f;
  if (x < 2.0f) return (((x - 5.0f) * x + 8.0f) * x - 4.0f) * a;
  return 0.0f;
}

struct Axis {          // one output coordinate of one axis: source window and the pieces the weights are made of
  int lo, n;
  float center, invscale, total;
};

__device__ __forceinline__ Axis make_axis(int in_size, int out_size, int o) {
  Axis ax;
  const float scale = (float)in_size / (float)out_size;       // area_pixel_compute_scale, align_corners = false
  const float support = scale >= 1.0f ? 2.0f * scale : 2.0f;
  ax.invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
  ax.center = scale * ((float)o + 0.5f);
  ax.lo = max((int)(ax.center - support + 0.5f), 0);
  ax.n = min((int)(ax.center + support + 0.5f), in_size) - ax.lo;
  float total = 0.f;
  for (int j = 0; j < ax.n; ++j) total += cubic_aa(((float)(j + ax.lo) - ax.center + 0.5f) * ax.invscale);
  ax.total = total;
  return ax;
}
__device__ __forceinline__ float axis_weight(const Axis& ax, int j) {
  return cubic_aa(((float)(j + ax.lo) - ax.center + 0.5f) * ax.invscale) / ax.total;
}

template <typename T>
__global__ void __launch_bounds__(256)
clip_crops_kernel(const T* __restrict__ img, int B, int Cc, int H, int W, const int32_t* __restrict__ boxes,
                  const int32_t* __restrict__ image_index, int n, int S, int Sc, const float* __restrict__ mean,
                  const float* __restrict__ stdv, float* __restrict__ out) {
  const int per = Cc * Sc * Sc;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < (int64_t)n * per; t += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / per);
    int r = (int)(t - (int64_t)i * per);
    const int c = r / (Sc * Sc);
    r -= c * Sc * Sc;
    const int oy = r / Sc, ox = r - oy * Sc;
    const int x0 = boxes[4 * i], y0 = boxes[4 * i + 1], x1 = boxes[4 * i + 2], y1 = boxes[4 * i + 3];
    const int b = image_index[i];
    const int ch = y1 - y0, cw = x1 - x0;
    float v = 0.f;
    if (b >= 0 && b < B && ch > 0 && cw > 0 && x0 >= 0 && y0 >= 0 && x1 <= W && y1 <= H) {
      // get_resize_output_image_size: (short, long) -> (S, int(S * long / short)) -- Python: int product, true division
      int rh, rw;
      if (cw <= ch) { rw = S; rh = (int)((double)((int64_t)S * ch) / (double)cw); }
      else { rh = S; rw = (int)((double)((int64_t)S * cw) / (double)ch); }
      const int top = (int)((double)(rh - Sc) / 2.0), left = (int)((double)(rw - Sc) / 2.0);
      const int ry = top + oy, rx = left + ox;
      if (ry >= 0 && ry < rh && rx >= 0 && rx < rw) {           // (centre crop of a smaller image pads with zeros)
        const Axis ay = make_axis(ch, rh, ry), ax = make_axis(cw, rw, rx);
        const T* src = img + (((int64_t)b * Cc + c) * H + y0) * W + x0;
        for (int j = 0; j < ay.n; ++j) {
          const T* row = src + (int64_t)(ay.lo + j) * W + ax.lo;
          float h = 0.f;
          for (int k = 0; k < ax.n; ++k) h += axis_weight(ax, k) * ElemIO<T>::ld(row + k);      // horizontal pass first (ATen order)
          v += axis_weight(ay, j) * h;
        }
      }
      v = (v - mean[c]) / stdv[c];
    }
    out[t] = v;
  }
}

}  // namespace crops
}  // namespace rc

extern "C" int rc_clip_crops(const void* images, rc_dtype dtype, int B, int C, int H, int W, const int32_t* boxes,
                             const int32_t* image_index, int n, int shortest_edge, int crop_size, const float* mean,
                             const float* stdv, float* out, void* stream) {
  RC_REQUIRE(images && boxes && image_index && mean && stdv && out, "rc_clip_crops: null pointer");
  RC_REQUIRE(B >= 1 && C >= 1 && H >= 1 && W >= 1 && n >= 0 && shortest_edge >= 1 && crop_size >= 1, "rc_clip_crops: bad shape");
  if (n == 0) return RC_OK;
  const int64_t total = (int64_t)n * C * crop_size * crop_size;
  const int64_t blocks = (total + 255) / 256;
  const int cap = rc::num_sms() * 16;
  const int grid = (int)(blocks < cap ? blocks : cap);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == RC_F32)
    rc::crops::clip_crops_kernel<float><<<grid, 256, 0, s>>>((const float*)images, B, C, H, W, boxes, image_index, n, shortest_edge,
                                                             crop_size, mean, stdv, out);
  else
    rc::crops::clip_crops_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)images, B, C, H, W, boxes, image_index, n,
                                                                     shortest_edge, crop_size, mean, stdv, out);
  return rc::check_launch("rc_clip_crops");
}
