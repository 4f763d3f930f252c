/*
 * oracle/msda_ref.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's multi-scale deformable attention forward.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this; nothing under ape_b200/ does.
 *
 * Follows, statement by statement, the reference CUDA kernel
 *   ape/layers/csrc/MsDeformAttn/ms_deform_im2col_cuda.cuh:237-299  (per-output loop, in-range test)
 *   ape/layers/csrc/MsDeformAttn/ms_deform_im2col_cuda.cuh:33-84    (bilinear corner fetch / weights)
 * whose result the reference's portable path
 *   ape/layers/multi_scale_deform_attn.py:84-124 (multi_scale_deformable_attn_pytorch: grid_sample,
 *   bilinear, zeros padding, align_corners=False)
 * reproduces up to fp32 summation order.
 *
 * Pinning: tests/test_oracle_golden.py checks this file against golden vectors produced by
 * importing the reference's multi_scale_deformable_attn_pytorch from /root/reference
 * (tests/golden/gen_msda_golden.py, committed with its outputs).
 *
 * Arithmetic: inputs fp32, accumulation in double or float (acc_double flag): the double
 * variant is the "exact" answer the CUDA kernels are compared to; the float variant mirrors
 * the reference kernel's scalar_t=float accumulation order (level-major, point-minor).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ms_deform_im2col_cuda.cuh:33-84 */
static double bilinear_d(const float *bottom, int height, int width, int nheads, int channels,
                         double h, double w, int m, int c) {
  const int h_low = (int)floor(h), w_low = (int)floor(w);
  const int h_high = h_low + 1, w_high = w_low + 1;
  const double lh = h - h_low, lw = w - w_low, hh = 1 - lh, hw = 1 - lw;
  const int w_stride = nheads * channels, h_stride = width * w_stride;
  const int h_low_off = h_low * h_stride, h_high_off = h_low_off + h_stride;
  const int w_low_off = w_low * w_stride, w_high_off = w_low_off + w_stride;
  const int base = m * channels + c;
  double v1 = 0, v2 = 0, v3 = 0, v4 = 0;
  if (h_low >= 0 && w_low >= 0) v1 = bottom[h_low_off + w_low_off + base];
  if (h_low >= 0 && w_high <= width - 1) v2 = bottom[h_low_off + w_high_off + base];
  if (h_high <= height - 1 && w_low >= 0) v3 = bottom[h_high_off + w_low_off + base];
  if (h_high <= height - 1 && w_high <= width - 1) v4 = bottom[h_high_off + w_high_off + base];
  const double w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
  return w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;
}

static float bilinear_f(const float *bottom, int height, int width, int nheads, int channels,
                        float h, float w, int m, int c) {
  const int h_low = (int)floorf(h), w_low = (int)floorf(w);
  const int h_high = h_low + 1, w_high = w_low + 1;
  const float lh = h - h_low, lw = w - w_low, hh = 1 - lh, hw = 1 - lw;
  const int w_stride = nheads * channels, h_stride = width * w_stride;
  const int h_low_off = h_low * h_stride, h_high_off = h_low_off + h_stride;
  const int w_low_off = w_low * w_stride, w_high_off = w_low_off + w_stride;
  const int base = m * channels + c;
  float v1 = 0, v2 = 0, v3 = 0, v4 = 0;
  if (h_low >= 0 && w_low >= 0) v1 = bottom[h_low_off + w_low_off + base];
  if (h_low >= 0 && w_high <= width - 1) v2 = bottom[h_low_off + w_high_off + base];
  if (h_high <= height - 1 && w_low >= 0) v3 = bottom[h_high_off + w_low_off + base];
  if (h_high <= height - 1 && w_high <= width - 1) v4 = bottom[h_high_off + w_high_off + base];
  const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
  return w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;
}

/*
 * value [B,S,H,D] f32; shapes [L,2] i64 (H_l,W_l); starts [L] i64; loc [B,Q,H,L,P,2] f32 (x,y);
 * attn [B,Q,H,L,P] f32; out [B,Q,H*D] f32.  Returns 0.
 * ms_deform_im2col_cuda.cuh:253-298: one output scalar per (b,q,m,c); levels outer, points inner.
 */
int msda_ref_forward(const float *value, const int64_t *shapes, const int64_t *starts,
                     const float *loc, const float *attn, float *out, int B, int S, int H, int D,
                     int L, int Q, int P, int acc_double, int nthreads) {
  const long long rows = (long long)B * Q * H;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
#endif
  for (long long row = 0; row < rows; ++row) {
    const int m = (int)(row % H);
    const int b = (int)(row / ((long long)H * Q));
    const long long wbase = row * L * P; /* data_weight_ptr, :267 */
    const int qid_stride = H * D;
    const float *vb = value + (size_t)b * S * qid_stride;
    for (int c = 0; c < D; ++c) {
      double col_d = 0;
      float col_f = 0;
      long long wp = wbase, lp = wbase << 1;
      for (int l = 0; l < L; ++l) {
        const int level_start = (int)starts[l];
        const int sh = (int)shapes[2 * l], sw = (int)shapes[2 * l + 1];
        const float *vl = vb + (size_t)level_start * qid_stride;
        for (int p = 0; p < P; ++p) {
          const float loc_w = loc[lp], loc_h = loc[lp + 1], weight = attn[wp];
          if (acc_double) {
            const double h_im = (double)loc_h * sh - 0.5, w_im = (double)loc_w * sw - 0.5;
            if (h_im > -1 && w_im > -1 && h_im < sh && w_im < sw)
              col_d += bilinear_d(vl, sh, sw, H, D, h_im, w_im, m, c) * weight;
          } else {
            const float h_im = loc_h * sh - 0.5f, w_im = loc_w * sw - 0.5f;
            if (h_im > -1 && w_im > -1 && h_im < sh && w_im < sw)
              col_f += bilinear_f(vl, sh, sw, H, D, h_im, w_im, m, c) * weight;
          }
          wp += 1;
          lp += 2;
        }
      }
      out[row * D + c] = acc_double ? (float)col_d : col_f;
    }
  }
  return 0;
}

int msda_ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
