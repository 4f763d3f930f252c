// nms.cu — greedy hard-NMS over score-sorted boxes (torchvision.ops.nms semantics), used by
//   the two-stage proposal selection   ape/modeling/ape_deta/deformable_transformer_vl.py:591-596
//   the final class-aware NMS           ape/modeling/ape_deta/fast_rcnn.py:192 (detectron2 batched_nms)
// Two launches: (1) upper-triangular IoU>thr bit matrix, 64x64 boxes per CTA; (2) one CTA scans the
// sorted list in chunks of 64: one warp resolves the 64x64 diagonal block in registers, then all threads OR
// the rows of the boxes kept in that chunk into the suppression bitset out of shared memory (the rows of the
// next chunk are prefetched with cp.async meanwhile) — the serial dependency is 64 register steps per chunk.
// IoU arithmetic is written with explicit round-to-nearest intrinsics (no FMA contraction):
//   inter / (area_a + area_b - inter) > thr, widths/heights clamped at 0, as torchvision's devIoU.
#include "common.cuh"

namespace ape {
namespace {

__device__ __forceinline__ bool iou_gt(const float4 a, const float4 b, float thr) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float w = fmaxf(__fsub_rn(right, left), 0.f), h = fmaxf(__fsub_rn(bottom, top), 0.f);
  const float inter = __fmul_rn(w, h);
  const float sa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  const float sb = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter)) > thr;
}

// grid (col_blocks, row_blocks), 64 threads; mask[row * col_blocks + cb] bit j = IoU(row, cb*64+j) > thr, j > row
__global__ void __launch_bounds__(64) nms_mask_kernel(const float4 *__restrict__ boxes, int n, const int *__restrict__ n_dev,
                                                      float thr, unsigned long long *__restrict__ mask, int col_blocks) {
  const int rb = blockIdx.y, cb = blockIdx.x;
  if (cb < rb) return;  // lower triangle never read
  if (n_dev) n = min(n, max(*n_dev, 0));  // only the first *n_dev boxes are real (static-shape callers)
  if (cb * 64 >= n) return;
  __shared__ float4 cols[64];
  const int t = threadIdx.x;
  const int col = cb * 64 + t;
  if (col < n) cols[t] = boxes[col];
  __syncthreads();
  const int row = rb * 64 + t;
  if (row >= n) return;
  const float4 a = boxes[row];
  const int ncols = min(64, n - cb * 64);
  unsigned long long bits = 0;
  for (int j = (rb == cb) ? t + 1 : 0; j < ncols; ++j)
    if (iou_gt(a, cols[j], thr)) bits |= 1ull << j;
  mask[(size_t)row * col_blocks + cb] = bits;
}

// single CTA, 1024 threads; keep[i] = 1 iff sorted box i survives; *count = number kept.
// Per chunk of 64 boxes: (a) warp 0 resolves the 64x64 diagonal block with the rows in registers (lane i holds
// rows i and i+32; the row broadcasts are shuffles that do not depend on the running bitset, so only a test +
// predicated OR per box sits on the serial chain); (b) all threads OR the rows of the kept boxes into the
// suppression bitset from SHARED memory — the chunk's rows were fetched with cp.async while the previous chunk
// was being resolved (double buffer), so no global-memory latency sits between chunks.
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(1024) nms_scan_kernel(const unsigned long long *__restrict__ mask, int n, const int *__restrict__ n_dev,
                                                        int pitch, unsigned char *__restrict__ keep, int *__restrict__ count) {
  extern __shared__ unsigned long long s_dyn[];
  if (n_dev) n = min(n, max(*n_dev, 0));
  const int col_blocks = (n + 63) / 64;  // active chunks; rows of `mask` are `pitch` words apart
  unsigned long long *removed = s_dyn;                      // col_blocks words
  unsigned long long *rows[2] = {s_dyn + pitch, s_dyn + pitch + 64 * (size_t)pitch};  // [64][pitch] each
  __shared__ unsigned long long s_keepbits;
  const int t = threadIdx.x, nt = blockDim.x;
  for (int w = t; w < col_blocks; w += nt) removed[w] = 0;
  auto prefetch = [&](int c, int buf) {
    // rows base..base+63, words c..col_blocks-1 (word c = the diagonal block)
    const int base = c * 64, nb = min(64, n - base), nw = col_blocks - c;
    for (int i = t; i < nb * nw; i += nt) {
      const int r = i / nw, w = c + (i - r * nw);
      cp_async8(&rows[buf][r * pitch + w], mask + (size_t)(base + r) * pitch + w);
    }
    cp_async_commit();
  };
  if (col_blocks > 0) prefetch(0, 0);
  int total = 0;
  for (int c = 0; c < col_blocks; ++c) {
    const int buf = c & 1;
    const int base = c * 64, nb = min(64, n - base);
    cp_async_wait_all();
    __syncthreads();  // chunk c's rows are in rows[buf]; removed[] of the previous chunk is complete
    if (c + 1 < col_blocks) prefetch(c + 1, buf ^ 1);
    if (t < 32) {
      const unsigned long long r0 = (t < nb) ? rows[buf][t * pitch + c] : 0ull;
      const unsigned long long r1 = (t + 32 < nb) ? rows[buf][(t + 32) * pitch + c] : 0ull;
      unsigned long long rem = removed[c], kept = 0;
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const unsigned long long row = __shfl_sync(0xffffffffu, i < 32 ? r0 : r1, i & 31);
        if (i < nb && !((rem >> i) & 1ull)) {
          kept |= 1ull << i;
          rem |= row;  // diagonal block: bits j > i only
        }
      }
      if (t == 0) s_keepbits = kept;
      total += __popcll(kept);
    }
    __syncthreads();
    const unsigned long long kept = s_keepbits;
    if (t < nb) keep[base + t] = (unsigned char)((kept >> t) & 1ull);
    // suppress later chunks: thread (g, wl) ORs rows g, g+8, .. of the chunk for word c+1+wl (+128, ..)
    const int g = t >> 7, wl = t & 127;
    for (int w = c + 1 + wl; w < col_blocks; w += 128) {
      unsigned long long acc = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = g + 8 * j;
        if ((kept >> i) & 1ull) acc |= rows[buf][i * pitch + w];
      }
      if (acc) atomicOr(&removed[w], acc);
    }
  }
  if (t == 0) *count = total;
}

// Fallback for very long lists (the double-buffered rows do not fit shared memory): same algorithm, rows read
// straight from global memory.
__global__ void __launch_bounds__(256) nms_scan_big_kernel(const unsigned long long *__restrict__ mask, int n, const int *__restrict__ n_dev,
                                                           int pitch, unsigned char *__restrict__ keep, int *__restrict__ count) {
  extern __shared__ unsigned long long removed[];  // col_blocks words
  if (n_dev) n = min(n, max(*n_dev, 0));
  const int col_blocks = (n + 63) / 64;
  __shared__ unsigned long long s_diag[64];
  __shared__ unsigned long long s_keepbits;
  __shared__ int s_total;
  const int t = threadIdx.x;
  for (int w = t; w < col_blocks; w += blockDim.x) removed[w] = 0;
  if (t == 0) s_total = 0;
  __syncthreads();
  for (int c = 0; c < col_blocks; ++c) {
    const int base = c * 64;
    const int nb = min(64, n - base);
    if (t < nb) s_diag[t] = mask[(size_t)(base + t) * pitch + c];
    __syncthreads();
    if (t == 0) {
      unsigned long long rem = removed[c], kept = 0;
      for (int i = 0; i < nb; ++i) {
        if (!((rem >> i) & 1ull)) {
          kept |= 1ull << i;
          rem |= s_diag[i];
        }
      }
      s_keepbits = kept;
      s_total += __popcll(kept);
    }
    __syncthreads();
    const unsigned long long kept = s_keepbits;
    if (t < nb) keep[base + t] = (unsigned char)((kept >> t) & 1ull);
    for (int w = c + 1 + t; w < col_blocks; w += blockDim.x) {
      unsigned long long acc = removed[w], k = kept;
      while (k) {
        const int i = __ffsll((long long)k) - 1;
        k &= k - 1;
        acc |= mask[(size_t)(base + i) * pitch + w];
      }
      removed[w] = acc;
    }
    __syncthreads();
  }
  if (t == 0) *count = s_total;
}

}  // namespace
}  // namespace ape

using namespace ape;

extern "C" int64_t ape_nms_workspace_bytes(int n) {
  const int64_t cb = (n + 63) / 64;
  return (int64_t)n * cb * 8;
}

static int nms_launch(const float *boxes_sorted, int n, const int *n_dev, float iou_threshold, void *workspace, uint8_t *keep,
                      int *count, void *stream) {
  if (n < 0) return fail(APE_ERR_INVALID_ARG, "nms: n=%d", n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n == 0) {
    if (count) cudaMemsetAsync(count, 0, sizeof(int), st);
    return APE_OK;
  }
  if (!boxes_sorted || !workspace || !keep || !count) return fail(APE_ERR_NULL_PTR, "nms: null pointer argument");
  if (reinterpret_cast<uintptr_t>(boxes_sorted) & 15) return fail(APE_ERR_INVALID_ARG, "nms: boxes must be 16-byte aligned");
  const int cb = (n + 63) / 64;
  if ((size_t)cb * 8 > 200 * 1024) return fail(APE_ERR_UNSUPPORTED, "nms: n=%d too large for the single-CTA scan", n);
  if (n_dev) cudaMemsetAsync(keep, 0, (size_t)n, st);  // entries past *n_dev are never visited
  nms_mask_kernel<<<dim3(cb, cb), 64, 0, st>>>(reinterpret_cast<const float4 *>(boxes_sorted), n, n_dev, iou_threshold,
                                               reinterpret_cast<unsigned long long *>(workspace), cb);
  if (int rc = check_launch("nms_mask_kernel")) return rc;
  const size_t smem = (size_t)cb * 8 * (1 + 2 * 64);
  if (smem <= 200 * 1024) {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return fail((int)e, "nms: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    nms_scan_kernel<<<1, 1024, smem, st>>>(reinterpret_cast<const unsigned long long *>(workspace), n, n_dev, cb, keep, count);
    return check_launch("nms_scan_kernel");
  }
  const size_t smem_big = (size_t)cb * 8;
  if (smem_big > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(nms_scan_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_big);
    if (e != cudaSuccess) return fail((int)e, "nms: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  nms_scan_big_kernel<<<1, 256, smem_big, st>>>(reinterpret_cast<const unsigned long long *>(workspace), n, n_dev, cb, keep, count);
  return check_launch("nms_scan_kernel");
}

extern "C" int ape_nms_sorted(const float *boxes_sorted, int n, float iou_threshold, void *workspace, uint8_t *keep,
                              int *count, void *stream) {
  return nms_launch(boxes_sorted, n, nullptr, iou_threshold, workspace, keep, count, stream);
}

extern "C" int ape_nms_sorted_dev(const float *boxes_sorted, int n_max, const int *n_dev, float iou_threshold, void *workspace,
                                  uint8_t *keep, int *count, void *stream) {
  if (!n_dev) return fail(APE_ERR_NULL_PTR, "nms: null n_dev");
  return nms_launch(boxes_sorted, n_max, n_dev, iou_threshold, workspace, keep, count, stream);
}
