// vlf_pool.cu — language-side attention pooling of VisionLanguageFusion for ONE language token
// (the "name" prompt case; BiMultiHeadAttention.forward, ape/layers/fuse_helper.py:67-166, restructured as
// in ape_b200/layers/vision_language_fusion.py:single_token).  For every head h:
//     t[s]   = v_s . qa[h] + qc[h]                       (scores of all S vision tokens against the one key)
//     w[s]   = clamp(t[s] - max_all(t), -5e4, 5e4)        (stable_softmax_2d + clamps, fuse_helper.py:88-97)
//     wl[s]  = clamp(w[s] - max_s(w), -5e4, 5e4)          (:99-108)
//     p      = softmax_s(wl);   pooled[h,:] = sum_s p[s] * v_s
// Three launches, v is read twice (16-bit), nothing of size S x 2048 is ever formed:
//   (1) scores + per-CTA maxima, one warp per token;  (2) fold maxima (per head and global);
//   (3) per-CTA partial sums of exp() and exp()*v over a strip of tokens, thread = channel.
// The caller folds the [strips] partials (deterministic order) and applies the tiny projections.
#include "common.cuh"

namespace ape {
namespace {

constexpr int kMaxHeads = 8;
constexpr int kStrip = 256;  // tokens per CTA in the pooling pass

template <typename T>
__global__ void __launch_bounds__(256)
vlf_scores_kernel(const T *__restrict__ v, const float *__restrict__ qa, const float *__restrict__ qc,
                  float *__restrict__ scores, float *__restrict__ blockmax, int S, int C, int NH) {
  // grid (ceil(S/8), B); 8 warps = 8 tokens per CTA; lane owns channels [8*lane + 256*i, +8)
  extern __shared__ float s_qa[];  // NH * C
  __shared__ float s_max[8][kMaxHeads];
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < NH * C; i += 256) s_qa[i] = qa[(size_t)b * NH * C + i];
  __syncthreads();
  const int s = blockIdx.x * 8 + warp;
  float acc[kMaxHeads];
#pragma unroll
  for (int h = 0; h < kMaxHeads; ++h) acc[h] = 0.f;
  if (s < S) {
    const T *row = v + ((size_t)b * S + s) * C;
    for (int c0 = lane * 8; c0 < C; c0 += 256) {
      float f[8];
      Elem<T>::unpack(*reinterpret_cast<const uint4 *>(row + c0), f);
#pragma unroll
      for (int h = 0; h < kMaxHeads; ++h)
        if (h < NH) {
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[h] = fmaf(f[k], s_qa[h * C + c0 + k], acc[h]);
        }
    }
  }
#pragma unroll
  for (int h = 0; h < kMaxHeads; ++h) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[h] += __shfl_xor_sync(0xffffffffu, acc[h], o);
  }
  if (lane == 0) {
    for (int h = 0; h < NH; ++h) {
      float t = -INFINITY;
      if (s < S) {
        t = acc[h] + qc[b * NH + h];
        scores[((size_t)b * NH + h) * S + s] = t;
      }
      s_max[warp][h] = t;
    }
  }
  __syncthreads();
  if (threadIdx.x < NH) {
    float m = s_max[0][threadIdx.x];
    for (int w = 1; w < 8; ++w) m = fmaxf(m, s_max[w][threadIdx.x]);
    blockmax[((size_t)b * gridDim.x + blockIdx.x) * NH + threadIdx.x] = m;
  }
}
// fp32 rows are 4 elements per 16 bytes: dedicated unpack-free variant
template <>
__global__ void __launch_bounds__(256)
vlf_scores_kernel<float>(const float *__restrict__ v, const float *__restrict__ qa, const float *__restrict__ qc,
                         float *__restrict__ scores, float *__restrict__ blockmax, int S, int C, int NH) {
  extern __shared__ float s_qa[];
  __shared__ float s_max[8][kMaxHeads];
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < NH * C; i += 256) s_qa[i] = qa[(size_t)b * NH * C + i];
  __syncthreads();
  const int s = blockIdx.x * 8 + warp;
  float acc[kMaxHeads];
#pragma unroll
  for (int h = 0; h < kMaxHeads; ++h) acc[h] = 0.f;
  if (s < S) {
    const float *row = v + ((size_t)b * S + s) * C;
    for (int c = lane; c < C; c += 32) {
      const float f = row[c];
#pragma unroll
      for (int h = 0; h < kMaxHeads; ++h)
        if (h < NH) acc[h] = fmaf(f, s_qa[h * C + c], acc[h]);
    }
  }
#pragma unroll
  for (int h = 0; h < kMaxHeads; ++h) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[h] += __shfl_xor_sync(0xffffffffu, acc[h], o);
  }
  if (lane == 0) {
    for (int h = 0; h < NH; ++h) {
      float t = -INFINITY;
      if (s < S) {
        t = acc[h] + qc[b * NH + h];
        scores[((size_t)b * NH + h) * S + s] = t;
      }
      s_max[warp][h] = t;
    }
  }
  __syncthreads();
  if (threadIdx.x < NH) {
    float m = s_max[0][threadIdx.x];
    for (int w = 1; w < 8; ++w) m = fmaxf(m, s_max[w][threadIdx.x]);
    blockmax[((size_t)b * gridDim.x + blockIdx.x) * NH + threadIdx.x] = m;
  }
}

// maxes[0] = global max over everything; maxes[1 + b*NH + h] = max over s of scores[b,h,:]
__global__ void vlf_max_kernel(const float *__restrict__ blockmax, int B, int nblk, int NH, float *__restrict__ maxes) {
  __shared__ float s_row[64];
  const int t = threadIdx.x;
  if (t < B * NH) {
    const int b = t / NH, h = t % NH;
    float m = -INFINITY;
    for (int i = 0; i < nblk; ++i) m = fmaxf(m, blockmax[((size_t)b * nblk + i) * NH + h]);
    maxes[1 + t] = m;
    s_row[t] = m;
  }
  __syncthreads();
  if (t == 0) {
    float g = -INFINITY;
    for (int i = 0; i < B * NH; ++i) g = fmaxf(g, s_row[i]);
    maxes[0] = g;
  }
}

// grid (strips, B), blockDim = C (thread = channel).  partial[b, strip, h, 0..C-1] = sum_s e*v, [.., C] = sum_s e
template <typename T>
__global__ void __launch_bounds__(1024)
vlf_pool_kernel(const T *__restrict__ v, const float *__restrict__ scores, const float *__restrict__ maxes,
                float *__restrict__ partial, int S, int C, int NH, int stable_2d) {
  __shared__ float s_e[kStrip][kMaxHeads];
  const int b = blockIdx.y, strip = blockIdx.x, c = threadIdx.x;
  const int s0 = strip * kStrip, n = min(kStrip, S - s0);
  const float gmax = stable_2d ? maxes[0] : 0.f;
  for (int i = threadIdx.x; i < n * NH; i += blockDim.x) {
    const int r = i / NH, h = i % NH;
    const float t = scores[((size_t)b * NH + h) * S + s0 + r];
    const float w = fminf(fmaxf(t - gmax, -50000.f), 50000.f);
    const float rmax = fminf(fmaxf(maxes[1 + b * NH + h] - gmax, -50000.f), 50000.f);
    s_e[r][h] = expf(fminf(fmaxf(w - rmax, -50000.f), 50000.f));
  }
  __syncthreads();
  float acc[kMaxHeads], se[kMaxHeads];
#pragma unroll
  for (int h = 0; h < kMaxHeads; ++h) acc[h] = se[h] = 0.f;
  const T *col = v + ((size_t)b * S + s0) * C + c;
  for (int r = 0; r < n; ++r) {
    const float x = Elem<T>::to_f(col[(size_t)r * C]);
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h)
      if (h < NH) {
        const float e = s_e[r][h];
        acc[h] = fmaf(e, x, acc[h]);
        se[h] += e;
      }
  }
  float *dst = partial + ((size_t)b * gridDim.x + strip) * NH * (C + 1);
  for (int h = 0; h < NH; ++h) {
    dst[h * (C + 1) + c] = acc[h];
    if (c == 0) dst[h * (C + 1) + C] = se[h];
  }
}

}  // namespace
}  // namespace ape

using namespace ape;

extern "C" int64_t ape_vlf_pool_workspace_bytes(int B, int S, int C, int NH) {
  const int64_t nblk = (S + 7) / 8, strips = (S + kStrip - 1) / kStrip;
  return ((int64_t)B * NH * S + (int64_t)B * nblk * NH + 1 + (int64_t)B * NH + (int64_t)B * strips * NH * (C + 1)) * 4;
}

extern "C" int ape_vlf_pool(const void *v, const float *qa, const float *qc, void *workspace, float **partial_out,
                            int *strips_out, int B, int S, int C, int NH, int stable_softmax_2d, int dtype, void *stream) {
  if (B <= 0 || S <= 0 || C <= 0 || NH <= 0 || NH > kMaxHeads || C > 1024 || C % 32 != 0 || B * NH > 64 ||
      (dtype != APE_DTYPE_F32 && C % 256 != 0))
    return fail(APE_ERR_UNSUPPORTED, "vlf_pool: B=%d S=%d C=%d NH=%d not supported", B, S, C, NH);
  if (!v || !qa || !qc || !workspace) return fail(APE_ERR_NULL_PTR, "vlf_pool: null pointer argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nblk = (S + 7) / 8, strips = (S + kStrip - 1) / kStrip;
  float *scores = reinterpret_cast<float *>(workspace);
  float *blockmax = scores + (size_t)B * NH * S;
  float *maxes = blockmax + (size_t)B * nblk * NH;
  float *partial = maxes + 1 + (size_t)B * NH;
  const size_t smem = (size_t)NH * C * 4;
  if (dtype == APE_DTYPE_F32) vlf_scores_kernel<float><<<dim3(nblk, B), 256, smem, st>>>((const float *)v, qa, qc, scores, blockmax, S, C, NH);
  else if (dtype == APE_DTYPE_F16) vlf_scores_kernel<__half><<<dim3(nblk, B), 256, smem, st>>>((const __half *)v, qa, qc, scores, blockmax, S, C, NH);
  else vlf_scores_kernel<__nv_bfloat16><<<dim3(nblk, B), 256, smem, st>>>((const __nv_bfloat16 *)v, qa, qc, scores, blockmax, S, C, NH);
  if (int rc = check_launch("vlf_scores_kernel")) return rc;
  vlf_max_kernel<<<1, 64, 0, st>>>(blockmax, B, nblk, NH, maxes);
  if (int rc = check_launch("vlf_max_kernel")) return rc;
  if (dtype == APE_DTYPE_F32) vlf_pool_kernel<float><<<dim3(strips, B), C, 0, st>>>((const float *)v, scores, maxes, partial, S, C, NH, stable_softmax_2d);
  else if (dtype == APE_DTYPE_F16) vlf_pool_kernel<__half><<<dim3(strips, B), C, 0, st>>>((const __half *)v, scores, maxes, partial, S, C, NH, stable_softmax_2d);
  else vlf_pool_kernel<__nv_bfloat16><<<dim3(strips, B), C, 0, st>>>((const __nv_bfloat16 *)v, scores, maxes, partial, S, C, NH, stable_softmax_2d);
  if (partial_out) *partial_out = partial;
  if (strips_out) *strips_out = strips;
  return check_launch("vlf_pool_kernel");
}
