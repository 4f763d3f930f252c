// msda_pair.cu — second-generation fused multi-scale deformable attention for the encoder (many queries), 16-bit value.
//
// Same function as msda_fused_fwd_kernel (msda_fwd.cu): the tail of MultiScaleDeformableAttention.forward
// (ape/layers/multi_scale_deform_attn.py:283-348: softmax over L*P, sampling-location arithmetic) plus the bilinear
// gather of ms_deform_im2col_cuda.cuh:237-299, but attacking the two limits ncu showed for the first kernel (round 1:
// L1 wavefronts 81 % AND issue slots 81 % busy, 248 M warp instructions per launch):
//
//  * PAIR layout of `value`: [B][S][H][2][D] — entry (s, h) holds the D channels of token s AND of token s+1 (its right
//    neighbour inside a level row) in one aligned 128-byte line.  A bilinear sample is then 2 lines (top row, bottom row)
//    instead of 4 half-used ones: 8 lanes x 16 B read one line, so every L1 wavefront carries 128 useful bytes.  The pair
//    tensor is written by ape_msda_pair_values (or directly by the value-projection GEMM's epilogue).  At the left / right
//    border of a row the pair is shifted so that it never leaves the row (weights move with it): slot 1 of a row's last
//    entry is never read.
//  * 16-bit blend: lane (x-half, 16-byte chunk) multiplies its 8 channels of the top and bottom line by the packed
//    (top, bottom) weights of its x-half with HFMA2 and accumulates the P points of ONE level in 16 bit (8 terms); the
//    level sums are added in fp32, and the two x-halves are combined in fp32 by one shuffle per channel at the end.  The
//    reference's own half kernel accumulates all L*P*4 terms in half (ms_deform_im2col_cuda.cuh:270,290).
//  * records are produced by the 8 lanes that consume them (lane j < L handles level j: 4 logits + 8 offsets with vector
//    loads, softmax statistics by 3 shuffles inside the 8-lane group), so the kernel has no block-wide barrier after the
//    level table is loaded and no staging of logits in shared memory.
//
// Requirements (host-checked; everything else takes msda_fused_fwd_kernel): fp16 / bf16 value with D = 32, P = 4, L <= 8,
// every level at least 2 pixels wide, 16-byte aligned offset rows.
#include "common.cuh"

namespace ape {
namespace {

constexpr int kPairMaxLevels = 8;

struct PairParams {
  const void *value2;   // [B][S][H][2][32] 16-bit
  const int64_t *shapes, *starts;
  const void *offs, *logits;
  const float *ref;
  void *out;            // [B][Q][H*32]
  int64_t offs_row_stride, logit_row_stride;
  int B, S, H, L, Q;
  int ht_log2, ref_dim;
  // self-attention over the pyramid (Q == S): a CTA's queries form a tile_w x (QT / tile_w) pixel tile of one level, so
  // the sampled texels overlap vertically as well as horizontally (tile_w = 0: QT consecutive queries)
  int tile_w, n_tiles, head_major;
  int tiles_cum[kPairMaxLevels + 1];  // first tile of every level
  int tiles_x[kPairMaxLevels];        // tiles per row of every level
};

struct __align__(16) PairRec {
  unsigned off;     // byte offset of pair entry (y0, x0) of this head inside the batch image
  unsigned wl, wr;  // packed 16-bit (top, bottom) weights of slot 0 / slot 1 (bilinear x attention)
  unsigned dy;      // bytes to the bottom row's entry (0 when the top row is used twice)
};

template <typename T> struct H2;
template <> struct H2<__half> {
  using V = __half2;
  static __device__ __forceinline__ unsigned pack(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<unsigned *>(&h); }
  static __device__ __forceinline__ float2 tof(V v) { return __half22float2(v); }
  static __device__ __forceinline__ V lo2(V v) { return __low2half2(v); }
  static __device__ __forceinline__ V hi2(V v) { return __high2half2(v); }
};
template <> struct H2<__nv_bfloat16> {
  using V = __nv_bfloat162;
  static __device__ __forceinline__ unsigned pack(float a, float b) { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<unsigned *>(&h); }
  static __device__ __forceinline__ float2 tof(V v) { return __bfloat1622float2(v); }
  static __device__ __forceinline__ V lo2(V v) { return __low2bfloat162(v); }
  static __device__ __forceinline__ V hi2(V v) { return __high2bfloat162(v); }
};

template <typename TO>
__device__ __forceinline__ void ld4(const TO *p, float *f) {
  if constexpr (sizeof(TO) == 4) {
    const float4 v = *reinterpret_cast<const float4 *>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  } else {
    const uint2 v = *reinterpret_cast<const uint2 *>(p);
    float g[8];
    Elem<TO>::unpack(make_uint4(v.x, v.y, 0u, 0u), g);
    f[0] = g[0]; f[1] = g[1]; f[2] = g[2]; f[3] = g[3];
  }
}
template <typename TO>
__device__ __forceinline__ void ld8(const TO *p, float *f) {
  if constexpr (sizeof(TO) == 4) {
    const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    Elem<TO>::unpack(*reinterpret_cast<const uint4 *>(p), f);
  }
}

// One sampling point -> pair record.  In-range test and corner validity as ms_deform_im2col_cuda.cuh:279-291 / :36-80.
template <typename T>
__device__ __forceinline__ PairRec make_pair_rec(float x, float y, float a, int Hl, int Wl, int start, int h, int H, unsigned dyb) {
  const float h_im = y * (float)Hl - 0.5f;
  const float w_im = x * (float)Wl - 0.5f;
  const bool in_range = h_im > -1.f && w_im > -1.f && h_im < (float)Hl && w_im < (float)Wl;
  const float lh = h_im - floorf(h_im), lw = w_im - floorf(w_im);
  const int h_low = __float2int_rd(h_im), w_low = __float2int_rd(w_im);  // saturating; NaN -> 0 (in_range false then)
  // vertical: top = row h_low, bottom = row h_low + 1
  const float wt = (in_range && h_low >= 0) ? (1.f - lh) * a : 0.f;
  const float wb = (in_range && h_low < Hl - 1) ? lh * a : 0.f;
  const int y0 = min(max(h_low, 0), Hl - 1);
  const bool two_rows = h_low >= 0 && h_low < Hl - 1;
  // h_low == -1: the bottom corner is row 0 = the clamped y0 and dy = 0 reads it as the "bottom" line (top weight is 0);
  // h_low == Hl-1: bottom weight is 0 and the top line is read twice.
  // horizontal: left = column w_low, right = w_low + 1; the pair entry is [x0, x0 + 1] with x0 kept inside [0, Wl-2]
  float cl, cr;  // weights of slot 0 / slot 1 (a NaN / Inf location is out of range: every weight must be exactly 0)
  int x0;
  if (!in_range) {
    x0 = min(max(w_low, 0), Wl - 2); cl = 0.f; cr = 0.f;
  } else if (w_low < 0) {            // w_low == -1 when in range: only the right corner (column 0) exists -> slot 0
    x0 = 0; cl = lw; cr = 0.f;
  } else if (w_low >= Wl - 1) {  // w_low == Wl-1 when in range: only the left corner (column Wl-1) exists -> slot 1 of Wl-2
    x0 = Wl - 2; cl = 0.f; cr = 1.f - lw;
  } else {
    x0 = w_low; cl = 1.f - lw; cr = lw;
  }
  PairRec r;
  r.off = (unsigned)(((start + y0 * Wl + x0) * H + h) * 128);
  r.wl = H2<T>::pack(wt * cl, wb * cl);
  r.wr = H2<T>::pack(wt * cr, wb * cr);
  r.dy = two_rows ? dyb : 0u;
  return r;
}

// grid = (q_tiles * head_tiles, B), 256 threads = 32 rows (b, q, h) x 8 lanes; dynamic smem = 32 * 4L * 16 bytes.
template <typename T, typename TO>
__global__ void __launch_bounds__(256) msda_pair_fused_kernel(const PairParams p) {
  pdl_prologue();
  using V = typename H2<T>::V;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_lvl[kPairMaxLevels * 3];
  const int tid = threadIdx.x;
  if (tid < p.L) {
    s_lvl[tid * 3 + 0] = (int)p.shapes[tid * 2 + 0];
    s_lvl[tid * 3 + 1] = (int)p.shapes[tid * 2 + 1];
    s_lvl[tid * 3 + 2] = (int)p.starts[tid];
  }
  __syncthreads();
  const int LP = p.L * 4;
  const int r = tid >> 3, j = tid & 7;  // row inside the CTA, lane inside the row group
  const int HT = 1 << p.ht_log2, head_tiles = p.H >> p.ht_log2, QT = 32 >> p.ht_log2;
  const int b = blockIdx.y;
  const int tile = p.head_major ? (int)(blockIdx.x % (unsigned)p.n_tiles) : (int)(blockIdx.x / (unsigned)head_tiles);
  const int htile = p.head_major ? (int)(blockIdx.x / (unsigned)p.n_tiles) : (int)(blockIdx.x % (unsigned)head_tiles);
  const int h = htile * HT + (r & (HT - 1));
  const int qi = r >> p.ht_log2;  // query inside the tile, 0 .. QT-1
  int q;
  bool active;
  if (p.tile_w > 0) {
    int l = 0;
    while (l + 1 < p.L && tile >= p.tiles_cum[l + 1]) ++l;
    const int t = tile - p.tiles_cum[l];
    const int ty = t / p.tiles_x[l], tx = t - ty * p.tiles_x[l];
    const int th = QT / p.tile_w;
    const int x = tx * p.tile_w + qi % p.tile_w, y = ty * th + qi / p.tile_w;
    active = x < s_lvl[l * 3 + 1] && y < s_lvl[l * 3];
    q = s_lvl[l * 3 + 2] + y * s_lvl[l * 3 + 1] + x;
  } else {
    q = tile * QT + qi;
    active = q < p.Q;
  }
  PairRec *rec = reinterpret_cast<PairRec *>(smem_raw) + r * LP;

  // ---- records: lane j < L owns level j of this row ----------------------------------------------------------------
  {
    const size_t bq = (size_t)b * p.Q + (active ? q : 0);
    const bool mine = active && j < p.L;
    float lg[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    if (mine) ld4<TO>(reinterpret_cast<const TO *>(p.logits) + bq * p.logit_row_stride + h * LP + j * 4, lg);
    float m = fmaxf(fmaxf(lg[0], lg[1]), fmaxf(lg[2], lg[3]));
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float e[4], sum = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      e[k] = mine ? __expf(lg[k] - m) : 0.f;
      sum += e[k];
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (mine) {
      const float ri = 1.f / sum;
      const int Hl = s_lvl[j * 3], Wl = s_lvl[j * 3 + 1], start = s_lvl[j * 3 + 2];
      float o8[8];
      ld8<TO>(reinterpret_cast<const TO *>(p.offs) + bq * p.offs_row_stride + (h * LP + j * 4) * 2, o8);
      const float *rp = p.ref + (bq * p.L + j) * p.ref_dim;
      float sx, sy;
      if (p.ref_dim == 2) {  // multi_scale_deform_attn.py:298-303: ref + off / (W_l, H_l)
        sx = 1.f / (float)Wl;
        sy = 1.f / (float)Hl;
      } else {               // :304-311: ref_xy + off / P * ref_wh * 0.5
        sx = rp[2] * 0.125f;
        sy = rp[3] * 0.125f;
      }
      const float rx = rp[0], ry = rp[1];
      const unsigned dyb = (unsigned)(Wl * p.H * 128);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        rec[j * 4 + k] = make_pair_rec<T>(fmaf(o8[2 * k], sx, rx), fmaf(o8[2 * k + 1], sy, ry), e[k] * ri, Hl, Wl, start, h, p.H, dyb);
    }
  }
  __syncwarp();  // a row's records are written and read by the same 8 lanes of one warp
  if (!active) return;

  // ---- gather ---------------------------------------------------------------------------------------------------------
  const int xh = j >> 2;  // which texel of the pair this lane multiplies
  const char *vb = reinterpret_cast<const char *>(p.value2) + (size_t)b * p.S * p.H * 128 + xh * 64 + (j & 3) * 16;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 1
  for (int l = 0; l < p.L; ++l) {
    PairRec rc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) rc[k] = rec[l * 4 + k];
    uint4 top[4], bot[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      top[k] = ldg_nc_v4(reinterpret_cast<const uint4 *>(vb + rc[k].off));
      bot[k] = ldg_nc_v4(reinterpret_cast<const uint4 *>(vb + rc[k].off + rc[k].dy));
    }
    V hacc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned w = xh ? rc[k].wr : rc[k].wl;
      const V w2 = *reinterpret_cast<const V *>(&w);
      const V wt = H2<T>::lo2(w2), wb = H2<T>::hi2(w2);  // (top, top) / (bottom, bottom); folds into HFMA2 operand selects
      const V *tv = reinterpret_cast<const V *>(&top[k]);
      const V *bv = reinterpret_cast<const V *>(&bot[k]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        hacc[i] = (k == 0) ? __hmul2(tv[i], wt) : __hfma2(tv[i], wt, hacc[i]);
        hacc[i] = __hfma2(bv[i], wb, hacc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = H2<T>::tof(hacc[i]);
      acc[2 * i] += f.x;
      acc[2 * i + 1] += f.y;
    }
  }
  // the two x-halves of a row sit 4 lanes apart
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 4);
  if (xh == 0) {
    uint4 *ob = reinterpret_cast<uint4 *>(p.out) + (((size_t)b * p.Q + q) * p.H + h) * 4 + (j & 3);
    stg_stream_v4(ob, Elem<T>::pack(acc));
  }
}

// value [B][S][C] (C = H*32, any 16-bit) -> pair layout [B][S][H][2][32]; optional per-token zero mask (key_padding_mask,
// multi_scale_deform_attn.py:286-287).  One thread per 16 bytes of output.
__global__ void __launch_bounds__(256) msda_pair_values_kernel(const uint4 *__restrict__ value, long long ld16, uint4 *__restrict__ out,
                                                               const unsigned char *__restrict__ mask, int S, int H, long long total) {
  pdl_prologue();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  // idx = (((b*S + s)*H + h)*2 + slot)*4 + c
  const int c = (int)(idx & 3), slot = (int)((idx >> 2) & 1);
  const long long e = idx >> 3;  // (b*S + s)*H + h
  const int h = (int)(e % H);
  const long long bs = e / H;
  const int s = (int)(bs % S);
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (s + slot < S) {
    const long long src = bs + slot;  // same image: s + slot < S
    if (!mask || !mask[src]) v = value[src * ld16 + h * 4 + c];
  }
  out[idx] = v;
}

}  // namespace
}  // namespace ape

using namespace ape;

extern "C" int ape_msda_pair_values(const void *value, int64_t ld, void *value2, const uint8_t *token_mask, int B, int S, int H,
                                    int D, int dtype, void *stream) {
  if (dtype != APE_DTYPE_F16 && dtype != APE_DTYPE_BF16) return fail(APE_ERR_UNSUPPORTED, "msda_pair_values: 16-bit value only");
  if (D != 32) return fail(APE_ERR_UNSUPPORTED, "msda_pair_values: head dim %d (only 32)", D);
  if (B < 0 || S < 0 || H <= 0) return fail(APE_ERR_INVALID_ARG, "msda_pair_values: bad sizes");
  if (B == 0 || S == 0) return APE_OK;
  if (!value || !value2) return fail(APE_ERR_NULL_PTR, "msda_pair_values: null pointer argument");
  if (ld < (int64_t)H * D || (ld * 2) % 16 || (reinterpret_cast<uintptr_t>(value) & 15) || (reinterpret_cast<uintptr_t>(value2) & 127))
    return fail(APE_ERR_INVALID_ARG, "msda_pair_values: value rows must be 16-byte aligned, value2 128-byte aligned");
  const long long total = (long long)B * S * H * 8;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  APE_LAUNCH((msda_pair_values_kernel), (unsigned)((total + 255) / 256), 256, 0, st, reinterpret_cast<const uint4 *>(value), ld * 2 / 16,
                                                                        reinterpret_cast<uint4 *>(value2), token_mask, S, H, total);
  return check_launch("msda_pair_values_kernel");
}

extern "C" int ape_msda_pair_supported(const int *host_shapes, int L, int H, int D, int P, int dtype) {
  if (dtype != APE_DTYPE_F16 && dtype != APE_DTYPE_BF16) return 0;
  if (D != 32 || P != 4 || L <= 0 || L > kPairMaxLevels || H <= 0 || H > 32 || (H & (H - 1))) return 0;
  long long S = 0;
  for (int l = 0; l < L; ++l) {
    if (host_shapes[2 * l] < 1 || host_shapes[2 * l + 1] < 2) return 0;
    S += (long long)host_shapes[2 * l] * host_shapes[2 * l + 1];
  }
  return S * H * 128 < (1LL << 32) ? 1 : 0;  // 32-bit byte offsets inside one image
}

extern "C" int ape_msda_pair_fused_fwd(const void *value2, const int64_t *shapes, const int64_t *starts, const int *host_shapes,
                                       const void *offsets, int64_t offs_row_stride, const void *logits, int64_t logit_row_stride,
                                       const float *ref, int ref_dim, void *out, int B, int S, int H, int D, int L, int Q, int P,
                                       int dtype, int offs_dtype, int heads_per_cta, int tile_w, int head_major, void *stream) {
  if (!host_shapes) return fail(APE_ERR_NULL_PTR, "msda_pair: null host_shapes");
  if (!ape_msda_pair_supported(host_shapes, L, H, D, P, dtype))
    return fail(APE_ERR_UNSUPPORTED, "msda_pair: needs a 16-bit value with D=32, P=4, L<=8, power-of-two H, levels >= 2 wide");
  if (offs_dtype != dtype && offs_dtype != APE_DTYPE_F32) return fail(APE_ERR_UNSUPPORTED, "msda_pair: offsets in the value dtype or fp32");
  if (ref_dim != 2 && ref_dim != 4) return fail(APE_ERR_INVALID_ARG, "msda_pair: ref_dim must be 2 or 4");
  if (B < 0 || Q < 0 || B > 65535) return fail(APE_ERR_INVALID_ARG, "msda_pair: bad B / Q");
  long long total = 0;
  for (int l = 0; l < L; ++l) total += (long long)host_shapes[2 * l] * host_shapes[2 * l + 1];
  if (total != S) return fail(APE_ERR_INVALID_ARG, "msda_pair: level shapes do not sum to S=%d", S);
  if (B == 0 || Q == 0) return APE_OK;
  if (!value2 || !shapes || !starts || !offsets || !logits || !ref || !out) return fail(APE_ERR_NULL_PTR, "msda_pair: null pointer argument");
  const int eo = dtype_size(offs_dtype);
  if ((reinterpret_cast<uintptr_t>(offsets) & 15) || (offs_row_stride * eo) % 16 || (reinterpret_cast<uintptr_t>(logits) & (4 * eo - 1)) ||
      (logit_row_stride * eo) % (4 * eo) || (reinterpret_cast<uintptr_t>(value2) & 127) || (reinterpret_cast<uintptr_t>(out) & 15))
    return fail(APE_ERR_INVALID_ARG, "msda_pair: offsets rows 16-byte aligned, logits rows 4-element aligned, value2 128-byte aligned");
  if (offs_row_stride < (int64_t)H * L * P * 2 || logit_row_stride < (int64_t)H * L * P)
    return fail(APE_ERR_INVALID_ARG, "msda_pair: row strides smaller than a row");
  int ht = heads_per_cta > 0 ? heads_per_cta : (Q >= 128 ? 1 : (H < 32 ? H : 32));
  if (ht > H || (ht & (ht - 1)) || ht > 32) return fail(APE_ERR_INVALID_ARG, "msda_pair: heads_per_cta=%d", ht);
  PairParams p{};
  p.value2 = value2; p.shapes = shapes; p.starts = starts; p.offs = offsets; p.logits = logits; p.ref = ref; p.out = out;
  p.offs_row_stride = offs_row_stride; p.logit_row_stride = logit_row_stride;
  p.B = B; p.S = S; p.H = H; p.L = L; p.Q = Q; p.ref_dim = ref_dim;
  int lg = 0;
  while ((1 << lg) < ht) ++lg;
  p.ht_log2 = lg;
  const int QT = 32 >> lg;
  if (tile_w < 0) tile_w = (Q == S && QT >= 8) ? 8 : 0;  // default for self-attention: 8 x (QT/8) pixel tiles
  while (tile_w > QT && !(tile_w & (tile_w - 1))) tile_w >>= 1;  // small calls run with fewer queries per CTA than the tile asked for
  if (tile_w > 0 && (Q != S || QT % tile_w != 0 || (tile_w & (tile_w - 1))))
    return fail(APE_ERR_INVALID_ARG, "msda_pair: tile_w=%d needs Q == S and a power of two dividing %d", tile_w, QT);
  p.tile_w = tile_w;
  p.head_major = head_major ? 1 : 0;
  if (tile_w > 0) {
    const int th = QT / tile_w;
    int cum = 0;
    for (int l = 0; l < L; ++l) {
      p.tiles_cum[l] = cum;
      p.tiles_x[l] = (host_shapes[2 * l + 1] + tile_w - 1) / tile_w;
      cum += p.tiles_x[l] * ((host_shapes[2 * l] + th - 1) / th);
    }
    p.tiles_cum[L] = cum;
    p.n_tiles = cum;
  } else {
    p.n_tiles = (Q + QT - 1) / QT;
  }
  dim3 grid((unsigned)(p.n_tiles * (H >> lg)), (unsigned)B);
  const size_t smem = (size_t)32 * L * 4 * sizeof(PairRec);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == APE_DTYPE_F16) {
    if (offs_dtype == APE_DTYPE_F16) APE_LAUNCH((msda_pair_fused_kernel<__half, __half>), grid, 256, smem, st, p);
    else APE_LAUNCH((msda_pair_fused_kernel<__half, float>), grid, 256, smem, st, p);
  } else {
    if (offs_dtype == APE_DTYPE_BF16) APE_LAUNCH((msda_pair_fused_kernel<__nv_bfloat16, __nv_bfloat16>), grid, 256, smem, st, p);
    else APE_LAUNCH((msda_pair_fused_kernel<__nv_bfloat16, float>), grid, 256, smem, st, p);
  }
  return check_launch("msda_pair_fused_kernel");
}
