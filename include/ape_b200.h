/*
 * ape_b200.h — C-ABI of libape_b200.so, the sm_100a (B200) kernels behind APE's
 * detection forward pass.
 *
 * Every entry point is `extern "C"`, takes plain device pointers and sizes, a
 * `cudaStream_t` passed as `void*`, never allocates, never synchronises, and is
 * safe to call under CUDA-graph capture.  Return value: 0 = ok, <0 = invalid
 * argument (text in ape_last_error()), >0 = a cudaError_t from the launch.
 *
 * Reference interfaces replaced (paths relative to the APE repository):
 *   ape_msda_fwd            <- torch.ops.ape.ms_deform_attn_forward
 *                              ape/layers/csrc/vision.cpp:76-79 (registration)
 *                              ape/layers/csrc/MsDeformAttn/ms_deform_attn.h:21-40 (dispatch)
 *                              ape/layers/csrc/MsDeformAttn/ms_deform_attn_cuda.cu:21-81 (host)
 *                              ape/layers/csrc/MsDeformAttn/ms_deform_im2col_cuda.cuh:237-299 (kernel)
 *   ape_msda_fused_fwd      <- MultiScaleDeformableAttention.forward tail
 *                              ape/layers/multi_scale_deform_attn.py:283-348
 *                              (softmax + sampling-location arithmetic + gather in one launch)
 * The Python side that binds these (ctypes) and registers the reference's
 * operator names lives in ape_b200/_lib.py and ape_b200/ops.py; the binding a
 * reference maintainer would add is shown in INTEGRATION.md.
 */
#ifndef APE_B200_H_
#define APE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APE_ABI_VERSION 3

/* element types accepted by the kernels (value of the `dtype` argument) */
#define APE_DTYPE_F32 0
#define APE_DTYPE_F16 1
#define APE_DTYPE_BF16 2

/* status codes (<0); positive return values are cudaError_t */
#define APE_OK 0
#define APE_ERR_INVALID_ARG (-1)
#define APE_ERR_UNSUPPORTED (-2)
#define APE_ERR_NULL_PTR (-3)

int ape_abi_version(void);

/* Text of the last non-zero status returned on the calling thread ("" if none). */
const char *ape_last_error(void);

/* Number of kernels this library has launched since load (all threads). */
uint64_t ape_launch_count(void);

/*
 * Multi-scale deformable attention, forward.
 *
 *   out[b,q,h,:] = sum_{l<L} sum_{p<P} attn[b,q,h,l,p] *
 *                  bilinear_zero_pad(value_l[b,:,h,:], loc[b,q,h,l,p] * (W_l,H_l) - 0.5)
 *
 * value   [B,S,H,D]      contiguous, dtype
 * spatial_shapes [L,2]   int64 (H_l, W_l), DEVICE memory (as in the reference op)
 * level_start    [L]     int64, DEVICE memory
 * loc     [B,Q,H,L,P,2]  contiguous, dtype, (x,y) normalised to [0,1]
 * attn    [B,Q,H,L,P]    contiguous, dtype
 * out     [B,Q,H*D]      contiguous, dtype; fully overwritten (no pre-zeroing needed)
 *
 * Accumulation is always fp32 (the reference accumulates in `dtype`,
 * ms_deform_im2col_cuda.cuh:270).  Sample kept iff h_im>-1 && w_im>-1 &&
 * h_im<H_l && w_im<W_l (…cuh:285-291), per-corner validity as …cuh:56-78.
 */
int ape_msda_fwd(const void *value, const int64_t *spatial_shapes, const int64_t *level_start,
                 const void *loc, const void *attn, void *out, int B, int S, int H, int D, int L,
                 int Q, int P, int dtype, void *stream);

/*
 * Same contract as ape_msda_fwd but selects a kernel variant explicitly
 * (used by bench.py / tests to sweep mappings; `variant` < 0 = default).
 *   variant = heads_per_cta (1,2,4,8) | unroll<<8 ; 0x1000 = scalar fallback kernel
 */
int ape_msda_fwd_variant(const void *value, const int64_t *spatial_shapes, const int64_t *level_start,
                         const void *loc, const void *attn, void *out, int B, int S, int H, int D,
                         int L, int Q, int P, int dtype, int variant, void *stream);

/*
 * Fused tail of MultiScaleDeformableAttention.forward (multi_scale_deform_attn.py:283-348):
 * takes the raw outputs of the sampling_offsets / attention_weights linears and the
 * reference points, performs softmax over L*P, the sampling-location arithmetic and
 * the gather in one launch (sampling locations and attention weights never touch HBM).
 *
 * value     [B,S,H,D]        dtype
 * offsets   [B,Q,H,L,P,2]    offs_dtype (raw linear output), row stride `offs_row_stride` elements per (b,q)
 * logits    [B,Q,H,L*P]      offs_dtype (raw linear output), row stride `logit_row_stride` elements per (b,q)
 * ref       [B,Q,L,ref_dim]  fp32, ref_dim 2 (points) or 4 (boxes cx,cy,w,h)
 *   ref_dim 2: loc = ref + off / (W_l,H_l)                      (…py:298-303)
 *   ref_dim 4: loc = ref_xy + off / P * ref_wh * 0.5            (…py:304-311)
 * out       [B,Q,H*D]        dtype
 */
int ape_msda_fused_fwd(const void *value, const int64_t *spatial_shapes, const int64_t *level_start,
                       const void *offsets, int64_t offs_row_stride, const void *logits,
                       int64_t logit_row_stride, const float *ref, int ref_dim, void *out, int B,
                       int S, int H, int D, int L, int Q, int P, int dtype, int offs_dtype,
                       void *stream);

/*
 * Second-generation fused kernel for calls with many queries (the encoder: Q = S = 87 296 at 1024^2), 16-bit value, D = 32,
 * P = 4.  `value2` is the PAIR layout [B][S][H][2][D]: entry (s, h) = channels of token s followed by those of token s+1, one
 * aligned 128-byte line, so a bilinear sample costs two full L1 lines instead of four half-used ones; the four corners of
 * the P points of one level are blended with packed 16-bit FMAs, levels are summed in fp32 (the reference's half kernel
 * accumulates everything in half, ms_deform_im2col_cuda.cuh:270,290).  Same sampling semantics as ape_msda_fused_fwd.
 *   ape_msda_pair_values     value [B,S,H*D] (row pitch ld elements) -> value2; token_mask [B*S] bytes or NULL zeroes masked
 *                            tokens (key_padding_mask, multi_scale_deform_attn.py:286-287)
 *   ape_msda_pair_supported  1 if the geometry is covered (host_shapes int32 [L,2] on the HOST: every level >= 2 wide)
 *   ape_msda_pair_fused_fwd  heads_per_cta: 0 = auto (1 for Q >= 128: a CTA's 32 rows are queries of one head);
 *                            tile_w: for Q == S the CTA's queries form a tile_w x (32 / heads_per_cta / tile_w) PIXEL tile of one
 *                            level (texel re-use in both directions); 0 = consecutive queries, -1 = auto (8);
 *                            head_major: CTA order (0: heads of one tile adjacent, 1: tiles of one head adjacent)
 */
int ape_msda_pair_values(const void *value, int64_t ld, void *value2, const uint8_t *token_mask, int B, int S, int H, int D,
                         int dtype, void *stream);
int ape_msda_pair_supported(const int *host_shapes, int L, int H, int D, int P, int dtype);
int ape_msda_pair_fused_fwd(const void *value2, const int64_t *spatial_shapes, const int64_t *level_start,
                            const int *host_shapes, const void *offsets, int64_t offs_row_stride, const void *logits,
                            int64_t logit_row_stride, const float *ref, int ref_dim, void *out, int B, int S, int H, int D,
                            int L, int Q, int P, int dtype, int offs_dtype, int heads_per_cta, int tile_w, int head_major,
                            void *stream);

/*
 * Multi-scale deformable attention, backward  <- torch.ops.ape.ms_deform_attn_backward
 *   (ape/layers/csrc/vision.cpp:78, ms_deform_attn.h:42-61, ms_deform_attn_cuda.cu:84-160,
 *    ms_deform_im2col_cuda.cuh:86-146 per-sample arithmetic, :301-920 kernels).
 * Inputs as ape_msda_fwd plus grad_out [B,Q,H*D] (dtype).  Outputs: grad_value_f32 [B,S,H,D] ALWAYS fp32 and pre-zeroed by
 * the caller (vector atomics accumulate into it; the caller casts to dtype), grad_loc [B,Q,H,L,P,2] and grad_attn
 * [B,Q,H,L,P] in dtype, fully overwritten.  fp32 arithmetic for every dtype.
 */
int ape_msda_bwd(const void *value, const int64_t *spatial_shapes, const int64_t *level_start, const void *loc,
                 const void *attn, const void *grad_out, float *grad_value_f32, void *grad_loc, void *grad_attn, int B,
                 int S, int H, int D, int L, int Q, int P, int dtype, void *stream);

/*
 * Tensor-core linear layer: C[M,N] = act(A[M,K] * W[N,K]^T + bias) (+ residual), tcgen05 / TMA / TMEM.
 * Replaces the nn.Linear (cuBLAS) calls of the detection path (vit_eva_clip.py:225-232,266-267,125-132;
 * deformable_transformer_vl.py:36-54; multi_scale_deform_attn.py:278-295,353; vision_language_align.py:36-48).
 *
 * A [M,K] and W [N,K] (nn.Linear weight layout) are fp16 or bf16 (in_dtype), K contiguous, row pitches
 * lda / ldw in elements (16-byte aligned rows).  fp32 accumulation in tensor memory.
 * bias: fp32 [N] or NULL.  residual: [M,N] of out_dtype with pitch ldr, or NULL.
 * act: 0 none, 1 ReLU, 2 GELU(erf), 3 SwiGLU over interleaved (gate, up) column pairs -> C is [M, N/2],
 *      4 clamp to +-50000 (VisionLanguageAlign, vision_language_align.py:49-51).
 * out_dtype: APE_DTYPE_* of C (pitch ldc elements).  tile_n: 0 = auto, or 128 / 256 (| 0x1000 single CTA, 0x4000 cluster of
 * two CTAs sharing the weight tile by TMA multicast, 0x2000 CTA-pair MMA; 0x4000 forces the multicast cluster).
 */
int ape_gemm_tn(const void *A, int64_t lda, const void *W, int64_t ldw, void *C, int64_t ldc, const float *bias,
                const void *residual, int64_t ldr, int M, int N, int K, int in_dtype, int out_dtype, int act,
                int tile_n, void *stream);

/* ape_gemm_tn with the residual in its own element type (res_dtype): the engine keeps the residual stream / pre-LayerNorm
 * sums in fp32 (out_dtype F32) while GEMM operands and LayerNorm outputs are 16-bit, so e.g. a 16-bit residual is added
 * into an fp32 output (encoder: query + output_proj(...)) or an fp32 residual into an fp32 output (ViT: x + proj(...)). */
int ape_gemm_tn_ex(const void *A, int64_t lda, const void *W, int64_t ldw, void *C, int64_t ldc, const float *bias,
                   const void *residual, int64_t ldr, int res_dtype, int M, int N, int K, int in_dtype, int out_dtype,
                   int act, int tile_n, void *stream);

/*
 * ape_gemm_tn_ex with a LayerNorm folded around it (the sub-LayerNorms of the EVA-02 block, vit_eva_clip.py:266,130):
 *   consume: ln_part [M, ln_nparts, 2] per-row partial (sum, sum of squares) of the RAW 16-bit A written by its producer;
 *            W must hold gamma .* W, ln_colsum [N] its row sums (of the 16-bit values), bias = beta W^T + b; the epilogue forms
 *            rstd * (A W'^T - mean * colsum) + bias (+ residual); fp32 output only.  ln_inv_c = 1 / C, ln_eps as the LayerNorm.
 *   produce: stats_out [M, ceil(N/2/64), 2] with act = SwiGLU: (sum, sum of squares) of every 64-column slab of the output.
 * Either half may be disabled with NULL.
 */
int ape_gemm_tn_fused(const void *A, int64_t lda, const void *W, int64_t ldw, void *C, int64_t ldc, const float *bias,
                      const void *residual, int64_t ldr, int res_dtype, int M, int N, int K, int in_dtype, int out_dtype,
                      int act, int tile_n, const float *ln_part, int ln_nparts, const float *ln_colsum, float ln_inv_c,
                      float ln_eps, float *stats_out, int stats_nslab, void *stream);

/*
 * 3x3 convolution (stride 1, zero padding 1) over NHWC activations as an implicit GEMM on the tcgen05 kernel: the 3x3
 * convolutions of SimpleFeaturePyramid (vit_eva_clip.py:804-847) and of the mask head (deformable_detr_segm_vl.py:741-747).
 * x [B,H,W,Cin], w [Cout,3,3,Cin] (= the Conv2d weight permuted (0,2,3,1)), y [B,H,W,Cout], bias fp32 [Cout] or NULL;
 * fp16 / bf16, fp32 accumulation; act 0 / 1 (ReLU) / 2 (GELU).  Cin % 64 == 0; the image must be a whole number of
 * tw x (128/tw) pixel tiles, tw the largest power of two <= 128 dividing W.
 */
int ape_conv3x3_nhwc(const void *x, const void *w, void *y, const float *bias, int B, int H, int W, int Cin, int Cout,
                     int dtype, int act, void *stream);

/*
 * ape_gemm_tn with the 2-D rotary embedding of the ViT (VisionRotaryEmbeddingFast, utils_eva02.py:248-252,346) fused into
 * the epilogue: C = A W^T + bias, then t' = t*cos + rotate_half(t)*sin on output columns [0, rope_cols) — the q and k
 * thirds of the fused qkv projection (vit_eva_clip.py:225-262) — in fp32 before the single rounding to the 16-bit
 * output.  cos/sin fp32 [npos, 64]; row m uses position pos_map[m] (int32, device) or m % npos when NULL; head_dim 64.
 */
int ape_gemm_tn_rope(const void *A, int64_t lda, const void *W, int64_t ldw, void *C, int64_t ldc, const float *bias,
                     int M, int N, int K, int in_dtype, int out_dtype, const float *cos_table, const float *sin_table,
                     const int *pos_map, int npos, int head_dim, int rope_cols, int tile_n, void *stream);

/*
 * Development aid (no reference counterpart): when device_buffer is not NULL, every CTA of the following ape_gemm_tn*
 * launches (single-CTA / multicast variants) writes 8 clock64() stamps to device_buffer[8 * blockIdx.x ..]:
 * entry, set-up done, first operands landed, last MMA issued, first / last accumulator complete, epilogue drained, exit.
 * NULL switches it off.  Process-wide, not for concurrent use.
 */
void ape_gemm_set_trace(long long *device_buffer);

/*
 * LayerNorm over the last dimension (nn.LayerNorm / inner_attn_ln / ffn_ln of vit_eva_clip.py:505-523,266,130;
 * norms of the detrex transformer layers).  fp32 statistics; x [rows, C] pitch ldx (in_dtype), y pitch ldy
 * (out_dtype); weight/bias fp32 [C].  Pitches must cover C rounded up to 8 elements; padding elements
 * are written as 0.  row_map (int32 [rows], device) or NULL: output row of input row r (window partition).
 */
int ape_layernorm(const void *x, int64_t ldx, void *y, int64_t ldy, const float *weight, const float *bias,
                  const int *row_map, int rows, int C, float eps, int in_dtype, int out_dtype, void *stream);

/*
 * Extended LayerNorm for the deformable encoder (C % 8 == 0, C <= 1024); one pass over the activations for
 *   t = LN(x; weight, bias, eps)                 norms[1] of a detrex BaseTransformerLayer (deformable_transformer_vl.py:36-54)
 *   t = LN(t; weight2, bias2, eps2)  if weight2  layer_norm_v of the next VisionLanguageFusion (fuse_helper.py:224)
 *   y = t + col_add[image, :]        if col_add  gamma_v * delta_v of that fusion (one fp32 [C] vector per image for
 *                                                "name" prompts; col_add_stride elements between images, rows_per_image rows each)
 *   y2 = y + row_add[row, :]         if y2       query + query_pos (multi_scale_deform_attn.py:262-263); row_add / y2 in out_dtype
 */
int ape_layernorm_ex(const void *x, int64_t ldx, void *y, int64_t ldy, const float *weight, const float *bias, float eps,
                     const float *weight2, const float *bias2, float eps2, const float *col_add, int64_t col_add_stride,
                     int rows_per_image, const void *row_add, int64_t ld_add, void *y2, int64_t ldy2, int rows, int C,
                     int in_dtype, int out_dtype, void *stream);

/*
 * GroupNorm over token-major activations x [B, rows_per_image, C] (the neck's GroupNorm(32, 256) after each
 * 1x1 conv, configs/…1080k.py:42-55): statistics per (image, group) over all rows x C/groups channels, fp32,
 * deterministic.  y_batch_stride: elements between images of y (0 = rows_per_image * ldy), so the result can be
 * written straight into its slice of the flattened multi-level feature tensor.
 * workspace: ape_groupnorm_workspace_bytes(B, rows_per_image, C) bytes.
 */
int64_t ape_groupnorm_workspace_bytes(int B, int rows_per_image, int C);
int ape_groupnorm_nhwc(const void *x, int64_t ldx, void *y, int64_t ldy, int64_t y_batch_stride, const float *weight,
                       const float *bias, void *workspace, int B, int rows_per_image, int C, int groups, float eps,
                       int in_dtype, int out_dtype, void *stream);

/*
 * In-place 2-D rotary embedding on the q and k thirds of a fused qkv buffer [M, 3*C] (pitch ld):
 * t' = t*cos + rotate_half(t)*sin (VisionRotaryEmbeddingFast, utils_eva02.py:248-252,346).
 * cos/sin fp32 [npos, head_dim]; token m uses row pos_map[m] (int32, device) or m % npos when NULL.
 */
int ape_rope_qk(void *qkv, int64_t ld, const float *cos_table, const float *sin_table, const int *pos_map, int M,
                int C, int head_dim, int npos, int dtype, void *stream);

/*
 * Softmax attention core of the ViT blocks (Attention.forward, ape/modeling/backbone/vit_eva_clip.py:218-319, the
 * part xformers / F.scaled_dot_product_attention computes there): out = softmax(q k^T * scale) v per (sequence, head),
 * flash-attention style on tcgen05 tensor cores.  qkv [num_seq * n, >= 3*heads*64] (pitch ld elements): columns
 * [0,C) = q, [C,2C) = k, [2C,3C) = v with C = heads*64, head h at columns h*64 (the layout the fused qkv GEMM writes,
 * RoPE already applied); sequence s owns rows [s*n, (s+1)*n).  out [num_seq * n, >= C] (pitch ldo).  fp16 or bf16,
 * head_dim 64, n a multiple of 128; fp32 softmax statistics and accumulation.
 */
int ape_attn_fwd(const void *qkv, int64_t ld, void *out, int64_t ldo, int num_seq, int n, int heads, int head_dim,
                 float scale, int dtype, void *stream);
/* Same with sequences padded to n rows: only the first n_valid keys of every sequence take part in the softmax (the rows
 * beyond must hold finite values, e.g. zeros).  Used for the self-attention over the 900 decoder queries
 * (deformable_transformer_vl.py:142-147: nn.MultiheadAttention, 8 heads x 32 — heads zero-padded to 64 channels). */
/* stats_out (or NULL): fp32 [rows, heads, 2] — per row and head the (sum, sum of squares) of the 64 output values as stored,
 * consumed by ape_gemm_tn_fused to fold the LayerNorm that follows (inner_attn_ln, vit_eva_clip.py:266) into the projection.
 * seq_stride (0 = n): rows between consecutive sequences; sequences may be packed tighter than the 128-row tile (the text
 * tower packs 77-token prompts at a stride of 80 rows): rows of a tile past n_valid are then neither attended nor written.
 * causal: key t only sees keys <= t (eva02_clip/transformer.py:714-720).  total_rows (0 = derived): rows of the qkv buffer. */
int ape_attn_fwd_ex(const void *qkv, int64_t ld, void *out, int64_t ldo, int num_seq, int n, int n_valid, int heads,
                    int head_dim, float scale, int dtype, float *stats_out, int seq_stride, int causal, int64_t total_rows,
                    void *stream);

/* Kernel structure behind ape_attn_fwd*: 0 = P through shared memory (4 CTAs / SM), 1 = P in tensor memory as the A operand of
 * tcgen05.mma, double-buffered scores (2 CTAs / SM).  set >= 0 selects it for the process; returns the value in force
 * (initially APE_ATTN_VARIANT or the built-in default). */
int ape_attn_variant(int set);

/*
 * Cross attention with separate Q / K / V tensors and 64- or 256-channel heads: the two softmax attentions of
 * VisionLanguageFusion for phrase / text prompts (BiMultiHeadAttention.forward, ape/layers/fuse_helper.py:67-166: 8 heads x 256;
 * vision <- language: queries = S vision tokens, keys = N_t phrases; language <- vision: the transposed roles), one
 * flash-attention pass per direction — the S x N_t score matrix (31 GB in fp32 at 1536^2 / 5 000 phrases) is never formed.
 * q [num_seq * nq, >= heads*head_dim], k / v [num_seq * nkv, ...], out like q; nq a multiple of 128, nkv a multiple of 64 (rows
 * padded by the caller with finite values), only the first n_valid keys of every sequence count.  fp16 / bf16, fp32 softmax
 * statistics and accumulation.  The reference's global-max shift and +-5e4 clamps are softmax-invariant for LayerNormed
 * inputs (see csrc/attn_xfwd.cu).
 */
int ape_attn_cross_fwd(const void *q, int64_t ldq, const void *k, int64_t ldk, const void *v, int64_t ldv, void *out, int64_t ldo,
                       int num_seq, int nq, int nkv, int n_valid, int heads, int head_dim, float scale, int dtype, void *stream);

/*
 * Language-side attention pooling of VisionLanguageFusion for a single language token ("name" prompts):
 * softmax over the S vision tokens of scores t[s,h] = v_s . qa[h] + qc[h] (with the reference's global-max shift
 * and +-5e4 clamps, fuse_helper.py:88-110) and the p-weighted sum of v.  v [B,S,C] dtype; qa [B,NH,C], qc [B,NH] fp32.
 * Results are left as per-strip partials in the workspace: *partial_out -> [B, *strips_out, NH, C+1] fp32
 * (last column = sum of exp).  workspace: ape_vlf_pool_workspace_bytes(B,S,C,NH) bytes.
 */
int64_t ape_vlf_pool_workspace_bytes(int B, int S, int C, int NH);
int ape_vlf_pool(const void *v, const float *qa, const float *qc, void *workspace, float **partial_out, int *strips_out,
                 int B, int S, int C, int NH, int stable_softmax_2d, int dtype, void *stream);

/*
 * Greedy hard NMS over boxes already sorted by descending score (torchvision.ops.nms semantics; replaces the
 * nms call inside batched_nms of deformable_transformer_vl.py:591-596 and fast_rcnn.py:192).
 * boxes_sorted [n,4] fp32 xyxy (16-byte aligned); keep [n] bytes (1 = survives); *count = survivors (device int).
 * workspace: ape_nms_workspace_bytes(n) bytes of device memory.
 */
int64_t ape_nms_workspace_bytes(int n);
int ape_nms_sorted(const float *boxes_sorted, int n, float iou_threshold, void *workspace, uint8_t *keep, int *count,
                   void *stream);
/* Same, for static-shape callers (CUDA graphs): buffers are sized for n_max boxes, only the first min(n_max, *n_dev)
 * (device int) are real; keep[i] = 0 for the rest.  workspace: ape_nms_workspace_bytes(n_max). */
int ape_nms_sorted_dev(const float *boxes_sorted, int n_max, const int *n_dev, float iou_threshold, void *workspace,
                       uint8_t *keep, int *count, void *stream);

/*
 * Class-aware NMS when (almost) every (query, class) pair is a candidate (test_score_thresh 0.0: 900 queries x 1203 names;
 * fast_rcnn.py:129-192 -> batched_nms, class by class as torchvision's _batched_nms_vanilla).  All classes share the Q
 * per-query boxes, so one Q x Q IoU bit matrix serves every class; one warp sorts and scans one class.
 * boxes [Q,4] fp32 xyxy (16-byte aligned), scores [Q,N] fp32 (row pitch ld_scores), row_valid [Q] bytes or NULL,
 * candidates = valid rows with score > score_thresh.  out [N,Q] fp32 CLASS-major: the score where (q, c) survives, -inf
 * elsewhere (a top-k over it yields the detections in descending score order).  Q <= 1024.
 * workspace: ape_nms_classwise_workspace_bytes(Q).
 */
int64_t ape_nms_classwise_workspace_bytes(int Q);
int ape_nms_classwise(const float *boxes, const float *scores, int64_t ld_scores, const uint8_t *row_valid, int Q, int N,
                      float score_thresh, float iou_threshold, void *workspace, float *out, void *stream);

/*
 * Input pipeline of the predictor: the ResizeShortestEdge of DefaultPredictor.__call__
 * (ape/engine/defaults.py:203-230 -> detectron2 ResizeTransform.apply_image -> PIL Image.resize(BILINEAR) for uint8 images)
 * on the device, BIT-EXACT with Pillow's libImaging/Resample.c (triangle filter widened by the down-scaling factor, 22-bit
 * fixed-point taps, uint8 intermediate between the horizontal and the vertical pass).
 *   ape_resample_ksize      taps per output sample for one axis (host)
 *   ape_resample_coeffs_u8  HOST tables of one axis: bounds [out,2] int32 (first source sample, tap count), kk [out,ksize] int32
 *   ape_resample_u8         src [H,W,C] uint8 (row pitch src_pitch bytes) -> out float32 planes [C][new_h][new_w] with element
 *                           strides (plane_stride, row_stride); tmp [H,new_w,C] uint8 workspace; tables as DEVICE copies;
 *                           flip_channels: source channel c lands in plane C-1-c (BGR -> RGB, defaults.py:218-220)
 */
int ape_resample_ksize(int in_size, int out_size);
int ape_resample_coeffs_u8(int in_size, int out_size, int *bounds_host, int *kk_host);
int ape_resample_u8(const uint8_t *src, int64_t src_pitch, uint8_t *tmp, float *out, int64_t plane_stride, int64_t row_stride,
                    const int *bounds_h, const int *kk_h, int ksize_h, const int *bounds_v, const int *kk_v, int ksize_v, int H,
                    int W, int C, int new_h, int new_w, int flip_channels, void *stream);

/* Decoder reference-point update (deformable_transformer_vl.py:268-300, 4-d reference points): new_ref [B,Q,4] =
 * sigmoid(delta + inverse_sigmoid(ref, eps)) and ref_in [B,Q,L,4] = new_ref[:,:,None] * cat(valid_ratios, valid_ratios)[:,None]
 * (valid_ratios [B,L,2]); all fp32, same operation order as the PyTorch sequence. */
int ape_ref_update(const float *delta, const float *ref, const float *valid_ratios, float *new_ref, float *ref_in, int B, int Q,
                   int L, float eps, void *stream);

/* y[b,n] = W[n,:] . x[b,:] + bias[n] in fp32 for 1..4 input rows (the folded language-side maps of VisionLanguageFusion for one
 * language token, ape/layers/fuse_helper.py:67-166 restructured): W [N,K] row-major, K a multiple of 4, W / x 16-byte aligned. */
int ape_gemv_f32(const float *W, const float *x, const float *bias, float *y, int B, int N, int K, void *stream);

/*
 * Instance-mask post-processing for the detections that survive the final selection (deformable_detr_segm_vl.py:569-603 and
 * detectron2 detector_postprocess / paste_masks_in_image), without the full-resolution fp32 maps:
 *   ape_mask_crop   logits [Q,h,w] dtype (one image), index [K] int64 kept queries, boxes [K,4] fp32 xyxy in padded-image
 *                   pixels -> out [K,S,S] bytes = (ROIAlign(S, scale 1, sampling_ratio 0, aligned) of (bilinear upsample to
 *                   Hp x Wp, align_corners=False, > 0) >= 0.5), i.e. BitMasks(sigmoid(mask) > 0.5).crop_and_resize(boxes, S).
 *                   workspace: ape_mask_crop_workspace_bytes(K, Hp, Wp) (one bit per upsampled pixel).
 *   ape_mask_paste  masks [N,S,S] bytes (0 / 1), boxes [N,4] fp32 in OUTPUT-image pixels -> out [N,img_h,img_w] bytes (bool):
 *                   bilinear grid_sample (zeros padding, align_corners=False) of the mask inside its box >= threshold.
 * boxes 16-byte aligned.
 */
int64_t ape_mask_crop_workspace_bytes(int K, int Hp, int Wp);
int ape_mask_crop(const void *logits, const int64_t *index, const float *boxes, void *workspace, uint8_t *out, int K, int h, int w,
                  int Hp, int Wp, int S, int dtype, void *stream);
int ape_mask_paste(const uint8_t *masks, const float *boxes, uint8_t *out, int N, int S, int img_h, int img_w, float threshold,
                   void *stream);
/*
 * The pasted masks as COCO run-length codes without ever writing the dense masks (the evaluators encode every pasted mask with
 * cocoapi's mask_util.encode(np.array(mask[:, :, None], order="F")) right away: ape/evaluation/d3_evaluation.py:466-468,
 * refcoco_evaluation.py:450-452, detectron2 instances_to_coco_json): runs of equal pixels in COLUMN-major order.
 *   ape_mask_paste_rle  pass 1 (positions == NULL): col_count [N,img_w] int32 <- run boundaries per (mask, column);
 *                       pass 2: positions <- the boundary positions x*img_h + y, in order, at col_offset [N,img_w] int64 =
 *                       exclusive scan of col_count over the whole array.  Same pixel arithmetic as ape_mask_paste.
 *   ape_rle_to_string   HOST: run lengths (first run = zeros) -> cocoapi's compressed "counts" string (rleToString); returns
 *                       the number of characters (out holds up to 7 per count).
 */
int ape_mask_paste_rle(const uint8_t *masks, const float *boxes, int N, int S, int img_h, int img_w, float threshold,
                       int *col_count, const int64_t *col_offset, int *positions, void *stream);
int ape_rle_to_string(const uint32_t *counts, int m, char *out);

#ifdef __cplusplus
}
#endif
#endif /* APE_B200_H_ */
