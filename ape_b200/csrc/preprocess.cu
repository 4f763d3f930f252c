// preprocess.cu — the image side of the predictor's input pipeline.
//
// The reference's DefaultPredictor.__call__ (ape/engine/defaults.py:203-230) resizes every image on the host with
//   self.aug.get_transform(img).apply_image(img)          ResizeShortestEdge -> ResizeTransform (detectron2 @ 017abbf)
// which for uint8 images is PIL's Image.resize(..., BILINEAR) (Pillow, libImaging/Resample.c: separable convolution
// with a triangle filter whose support grows with the down-scaling factor, 8-bit fixed-point coefficients with
// PRECISION_BITS = 22, a uint8 intermediate image between the horizontal and the vertical pass), then converts to a
// float32 CHW tensor on the host and uploads 12.6 MB per 1024^2 image.  Here the raw uint8 HWC image is uploaded
// (a quarter of the bytes at equal size, 1/16 for the 512^2 -> 1024^2 case of BASELINE configs[0]) and both passes run
// on the device, BIT-EXACT with Pillow: same coefficient tables (built on the host in double precision by
// ape_resample_coeffs_u8, the arithmetic of precompute_coeffs / normalize_coeffs_8bpc), same rounding, same clipping, same
// uint8 intermediate.  The vertical pass writes the float32 CHW tensor the model's input dict carries (optionally with the
// BGR -> RGB flip of defaults.py:218-220 folded into the channel index).
#include <math.h>

#include "common.cuh"

namespace ape {
namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;  // Resample.c

__device__ __forceinline__ int clip8(int v) {  // clip8(): in >> PRECISION_BITS clamped to [0, 255]
  v >>= PRECISION_BITS;
  return min(max(v, 0), 255);
}

// horizontal pass: src [H, W, C] uint8 (row pitch src_pitch bytes) -> tmp [H, new_w, C] uint8
__global__ void __launch_bounds__(256) resample_h_kernel(const uint8_t *__restrict__ src, long long src_pitch,
                                                         uint8_t *__restrict__ tmp, const int *__restrict__ bounds,
                                                         const int *__restrict__ kk, int ksize, int W, int C, int new_w) {
  pdl_prologue();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;  // byte inside the output row
  if (t >= new_w * C) return;
  const int y = blockIdx.y;
  const int xo = t / C, c = t - xo * C;
  const int xmin = __ldg(bounds + 2 * xo), xmax = __ldg(bounds + 2 * xo + 1);
  const int *k = kk + (size_t)xo * ksize;
  const uint8_t *row = src + (size_t)y * src_pitch + (size_t)xmin * C + c;
  int ss = 1 << (PRECISION_BITS - 1);
  for (int x = 0; x < xmax; ++x) ss += (int)__ldg(row + (size_t)x * C) * __ldg(k + x);
  tmp[((size_t)y * new_w) * C + t] = (uint8_t)clip8(ss);
}

// vertical pass: tmp [H, new_w, C] uint8 -> out [C, new_h, new_w] float32 (plane / row strides in elements), channel c of
// the source written to plane (flip ? C-1-c : c)
__global__ void __launch_bounds__(256) resample_v_kernel(const uint8_t *__restrict__ tmp, float *__restrict__ out,
                                                         long long plane_stride, long long row_stride,
                                                         const int *__restrict__ bounds, const int *__restrict__ kk, int ksize,
                                                         int C, int new_w, int flip) {
  pdl_prologue();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= new_w * C) return;
  const int yo = blockIdx.y;
  const int xo = t / C, c = t - xo * C;
  const int ymin = __ldg(bounds + 2 * yo), ymax = __ldg(bounds + 2 * yo + 1);
  const int *k = kk + (size_t)yo * ksize;
  const size_t pitch = (size_t)new_w * C;
  const uint8_t *col = tmp + (size_t)ymin * pitch + t;
  int ss = 1 << (PRECISION_BITS - 1);
  for (int y = 0; y < ymax; ++y) ss += (int)__ldg(col + (size_t)y * pitch) * __ldg(k + y);
  const int cc = flip ? C - 1 - c : c;
  out[(size_t)cc * plane_stride + (size_t)yo * row_stride + xo] = (float)clip8(ss);
}

inline double bilinear_filter(double x) {  // Resample.c: bilinear_filter, support 1.0
  if (x < 0.0) x = -x;
  if (x < 1.0) return 1.0 - x;
  return 0.0;
}

}  // namespace
}  // namespace ape

// Number of taps per output sample (precompute_coeffs: ksize = ceil(support) * 2 + 1, support = max(scale, 1)).
extern "C" int ape_resample_ksize(int in_size, int out_size) {
  if (in_size <= 0 || out_size <= 0) return ape::fail(APE_ERR_INVALID_ARG, "resample: sizes must be positive");
  double filterscale = (double)((float)in_size - 0.0f) / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 1.0 * filterscale;
  return (int)ceil(support) * 2 + 1;
}

// HOST arrays: bounds [out_size, 2] (first source sample, tap count), kk [out_size, ksize] fixed-point taps.
extern "C" int ape_resample_coeffs_u8(int in_size, int out_size, int *bounds, int *kk) {
  using namespace ape;
  if (bounds == nullptr || kk == nullptr) return fail(APE_ERR_NULL_PTR, "resample: null coefficient buffer");
  const int ksize = ape_resample_ksize(in_size, out_size);
  if (ksize < 0) return ksize;
  const float in0 = 0.0f, in1 = (float)in_size;  // the box of Image.resize without `box=`: the whole image
  double filterscale, scale;
  filterscale = scale = (double)(in1 - in0) / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 1.0 * filterscale;
  double *k = (double *)malloc(sizeof(double) * (size_t)ksize);
  if (k == nullptr) return fail(APE_ERR_INVALID_ARG, "resample: out of host memory");
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = in0 + (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    int x;
    for (x = 0; x < xmax; ++x) {
      const double w = bilinear_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (; x < ksize; ++x) k[x] = 0;
    for (x = 0; x < ksize; ++x)  // normalize_coeffs_8bpc
      kk[(size_t)xx * ksize + x] = k[x] < 0 ? (int)(-0.5 + k[x] * (1 << PRECISION_BITS)) : (int)(0.5 + k[x] * (1 << PRECISION_BITS));
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  free(k);
  return APE_OK;
}

// src [H, W, C] uint8 on the device (row pitch src_pitch bytes), tmp [H, new_w, C] uint8 workspace, out float32 planes.
// bounds_h / kk_h: [new_w, 2] / [new_w, ksize_h]; bounds_v / kk_v: [new_h, 2] / [new_h, ksize_v] (device copies of the tables above).
extern "C" int ape_resample_u8(const uint8_t *src, int64_t src_pitch, uint8_t *tmp, float *out, int64_t plane_stride,
                               int64_t row_stride, const int *bounds_h, const int *kk_h, int ksize_h, const int *bounds_v,
                               const int *kk_v, int ksize_v, int H, int W, int C, int new_h, int new_w, int flip_channels,
                               void *stream) {
  using namespace ape;
  if (!src || !tmp || !out || !bounds_h || !kk_h || !bounds_v || !kk_v) return fail(APE_ERR_NULL_PTR, "resample: null pointer");
  if (H <= 0 || W <= 0 || new_h <= 0 || new_w <= 0 || C < 1 || C > 4 || ksize_h < 1 || ksize_v < 1 || H > 65535 || new_h > 65535)
    return fail(APE_ERR_INVALID_ARG, "resample: bad geometry H=%d W=%d C=%d -> %dx%d", H, W, C, new_h, new_w);
  if (src_pitch < (int64_t)W * C || row_stride < new_w || plane_stride < (int64_t)new_h * row_stride)
    return fail(APE_ERR_INVALID_ARG, "resample: pitch / strides smaller than a row / plane");
  cudaStream_t st = (cudaStream_t)stream;
  const int bx = (new_w * C + 255) / 256;
  APE_LAUNCH(resample_h_kernel, dim3(bx, H), 256, 0, st, src, (long long)src_pitch, tmp, bounds_h, kk_h, ksize_h, W, C, new_w);
  int rc = check_launch("resample_h_kernel");
  if (rc) return rc;
  APE_LAUNCH(resample_v_kernel, dim3(bx, new_h), 256, 0, st, (const uint8_t *)tmp, out, (long long)plane_stride,
             (long long)row_stride, bounds_v, kk_v, ksize_v, C, new_w, flip_channels);
  return check_launch("resample_v_kernel");
}
