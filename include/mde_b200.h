/* mde_b200.h -- C ABI of the B200-native AdaBins head / loss / external-info path.
 *
 * One entry point per kernel group.  Plain pointers and sizes only (no torch types); every pointer is DEVICE
 * memory owned by the caller (torch allocates; kernels never allocate or free); every call enqueues work on
 * the given cudaStream_t and returns without synchronising (CUDA-graph capturable), unless noted.
 * Return value: MDE_OK (0) or a negative MDE_ERR_* code; mde_error_string() names it.  There is no CPU
 * fallback anywhere behind this header.
 *
 * Process model: ONE process drives ONE device (the one-process-per-GPU layout of the reference's mp.spawn / DDP,
 * train.py:576-640).  The library keeps a little process-global state that is neither per-device nor thread-safe: the launch
 * counter, "shared-memory attribute already set" flags per kernel instantiation, the profiling buffer of
 * mde_tc_debug_profile and the SyncBatchNorm wait bound.  Calls from several host threads must be serialised by the caller.
 *
 * Each declaration cites the reference interface it replaces (paths relative to the reference tree
 * DylanAuty/MDE-biological-vision-systems).
 */
#ifndef MDE_B200_H
#define MDE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mde_stream_t; /* cudaStream_t */

enum {
  MDE_OK = 0,
  MDE_ERR_BAD_SHAPE = -1,   /* a size / stride / divisibility precondition failed */
  MDE_ERR_BAD_POINTER = -2, /* null or misaligned pointer */
  MDE_ERR_BAD_ARCH = -3,    /* device is not sm_100 */
  MDE_ERR_LAUNCH = -4,      /* cudaGetLastError() != cudaSuccess after the launch */
  MDE_ERR_UNSUPPORTED = -5, /* mode not implemented */
  MDE_ERR_DRIVER = -6       /* driver entry point (cuTensorMapEncodeTiled) unavailable */
};

enum { MDE_F32 = 0, MDE_F64 = 1 };
enum { MDE_I64 = 0, MDE_I32 = 1, MDE_U8 = 2 }; /* label element types */
enum { MDE_NORM_LINEAR = 0, MDE_NORM_SOFTMAX = 1, MDE_NORM_SIGMOID = 2 };

int mde_version(void);
const char* mde_error_string(int code);
/* 0 if the current device is compute capability 10.x, MDE_ERR_BAD_ARCH otherwise. Synchronous, no launch. */
int mde_check_device(void);
/* Number of kernels launched through this library since load (the bench's "gpu_launches" evidence). */
int64_t mde_launch_count(void);
/* Programmatic dependent launch (csrc/common.cuh: launch_pdl / pdl_sync): kernels of the inference step can be launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization; they wait for their predecessor on the device (griddepcontrol.wait) before
 * their first global-memory access, so a kernel's launch latency and prologue overlap the previous kernel's execution.
 * mask: bit 0 = chains of small kernels (transformer GEMMs / attention / LayerNorm, regressor, query fold), bit 1 = persistent
 * tcgen05 kernels (conv3x3, point-wise GEMM, patch embedding, fused chain), bit 2 = streaming kernels (bias / SiLU / pooling,
 * squeeze-excite gate, resize + concat, stem); 0 = off; mask < 0 queries.  Returns the mask in effect (default: environment
 * MDE_PDL, else 0 -- measured on B200: a gain for eagerly launched steps only, none under CUDA-graph replay).
 * Results are identical either way.  Host-side state of the process. */
int mde_set_pdl(int mask);

/* ---- K3: label -> embedding gather -------------------------------------------------------------------
 * Replaces SemanticsLoader.get_semantics (ExternalInfoLoaders/SemanticsLoader.py:102-145) and
 * InstanceSegmentationLoader.get_instance_segmentation (ExternalInfoLoaders/InstanceSegmentationLoader.py:89-121):
 * clamp (labels outside [0, rows-1] -> background), table.index_select + permute to planar [B,D,H,W].
 *   labels      int64 [B*HW]           (read)
 *   labels_out  int64 [B*HW] or NULL   (clamped labels written back; may alias labels == in-place clamp)
 *   table       [rows, D]  of out_dtype, row-major (or [B, rows, D] when table_image_stride != 0)
 *   out         [B, D, HW] of out_dtype
 *   background  >= 0: clamp target; < 0: no clamp, out-of-range labels raise *oob_flag (int32, may be NULL)
 *               and produce zeros (the reference's index_select would raise IndexError).
 */
int mde_gather_embed(const int64_t* labels, int64_t* labels_out, const void* table, void* out, int B, int64_t HW,
                     int rows, int D, int background, int out_dtype, int64_t table_image_stride,
                     int32_t* oob_flag, mde_stream_t stream);

/* Same gather on the labels' wire formats ("next" row (f)3): label_dtype MDE_I64 (batch tensors, dataloader.py:200-205),
 * MDE_I32 (the .npz instance label / area maps, dataloader.py:136-150) or MDE_U8 (semantic maps after
 * astype(np.ubyte), dataloader.py:121-133 -- a -1 "no prediction" label arrives as 255 and is clamped like any other
 * out-of-range value).  labels_out (optional) receives the clamped labels as int64. */
int mde_gather_embed_labels(const void* labels, int label_dtype, int64_t* labels_out, const void* table, void* out, int B,
                            int64_t HW, int rows, int D, int background, int out_dtype, int64_t table_image_stride,
                            int32_t* oob_flag, mde_stream_t stream);

/* The same clamp + gather written straight into channels [c0, c0+D) of a (spatially padded) channels_last tensor
 * out_nhwc [B, Ho, Wo, pitch] at pixel (y + pad_top, x + pad_left): the embedding planes land where the encoder reads them
 * (input insertion, models/unet_adaptive_bins.py:194-211), without the planar [B,D,H,W] tensor and its transpose.  fp32
 * table [rows, D] (<= 48 KB), clamping mode only (0 <= background < rows); labels_out int64 [B*H*W] or NULL.
 * image_nchw (optional, fp32 [B, c0, H, W], 1 <= c0 <= 3): the leading channels [0, c0) -- the RGB planes the reference
 * concatenates in front (x = cat(x, semantics)) -- are written by the same pass, so the whole input row is one store stream. */
int mde_gather_embed_nhwc(const void* labels, int label_dtype, int64_t* labels_out, const float* table, const float* image_nchw,
                          float* out_nhwc, int B, int H, int W, int rows, int D, int background, int pitch, int c0, int Ho,
                          int Wo, int pad_top, int pad_left, mde_stream_t stream);

/* Per-image class histogram -> per-image table of area fractions count/HW (float64), the gather table of
 * SemanticsLoader.get_semantics_inst_areas (SemanticsLoader.py:88-99).
 *   counts int32 [B, rows] workspace (zeroed by the call); frac float64 [B, rows] output. */
int mde_class_area_table(const int64_t* labels, int B, int64_t HW, int rows, int32_t* counts, double* frac,
                         mde_stream_t stream);

/* int64 -> float32 cast of instance areas (InstanceSegmentationLoader.py:112) */
int mde_cast_i64_f32(const int64_t* in, float* out, int64_t n, mde_stream_t stream);

/* ---- A3: per-pixel 1x1-conv MLP  C_in -> 10 -> 10 with ReLU (unet_adaptive_bins.py:146-174, used :196-228) ----
 *   x [B, C_in, HW] float32 (C_in = 1 or 3), each input is divided by in_scale first (the /(H*W) of :219,:226; pass 1.0f for none)
 *   w0 [H1, C_in], b0 [H1], w1 [H2, H1], b1 [H2]   (H1 = H2 = 10 in the reference; <= 16 supported)
 *   out: written at channel offset of a [B, out_channels_total, HW] tensor (so the caller can write straight
 *        into the concatenated encoder input):  out + (b*out_batch_stride + c*HW + p)                      */
int mde_aux_mlp_fwd(const float* x, int64_t x_batch_stride, const float* w0, const float* b0, const float* w1,
                    const float* b1, float* out, int64_t out_batch_stride, int B, int C_in, int H1, int H2,
                    int64_t HW, float in_scale, mde_stream_t stream);
/* backward: grad wrt x (may be NULL) and wrt the four parameter tensors (accumulated with atomics; caller zeroes) */
int mde_aux_mlp_bwd(const float* x, int64_t x_batch_stride, const float* w0, const float* b0, const float* w1,
                    const float* b1, const float* gout, int64_t gout_batch_stride, float* gx, float* gw0, float* gb0,
                    float* gw1, float* gb1, int B, int C_in, int H1, int H2, int64_t HW, float in_scale,
                    mde_stream_t stream);

/* y[p,c] = act(x[p,c] + bias[c]) (+ residual[p,c]) on channels_last activations [pixels, C] (C % 4 == 0); act 0 none,
 * 1 SiLU; y may alias x.  Epilogue of a convolution whose eval-mode BatchNorm is folded into the filter. */
int mde_bias_act_nhwc(const float* x, const float* bias, const float* residual, float* y, int64_t pixels, int C, int act,
                      mde_stream_t stream);

/* Same epilogue written into a zero-padded tensor y [B, H+pad_top+pad_bottom, W+pad_left+pad_right, C]: folds the F.pad of a
 * following TensorFlow-"SAME" stride-2 convolution into this pass. */
int mde_bias_act_pad_nhwc(const float* x, const float* bias, float* y, int B, int H, int W, int C, int pad_top,
                          int pad_bottom, int pad_left, int pad_right, int act, mde_stream_t stream);

/* ---- K1b/K2 prologue: bin-width regressor + normalisation + cumsum (miniViT.py:17-21,35-45;
 * unet_adaptive_bins.py:292-296).  t0 [B, E] rows at stride t0_stride (token 0 of the transformer output).
 *   w1 [H,E] b1 [H] w2 [H,H] b2 [H] w3 [n_bins,H] b3 [n_bins]   (E = 128, H = 256 in the reference)
 *   y_raw [B,n_bins] (pre-normalisation regressor output, kept for backward), widths_normed [B,n_bins],
 *   edges [B,n_bins+1], centers [B,n_bins]                                                            */
int mde_regressor_bins_fwd(const float* t0, int64_t t0_stride, const float* w1, const float* b1, const float* w2,
                           const float* b2, const float* w3, const float* b3, int B, int E, int H, int n_bins,
                           int norm_mode, float min_val, float max_val, float* y_raw, float* widths_normed,
                           float* edges, float* centers, mde_stream_t stream);

/* ---- split-bf16 pairs ("bf16x3"): the operand format of the tensor-core kernels below.  An fp32 tensor of n elements
 * travels as uint16 planes[2][n]: plane 0 = bf16_rn(v), plane 1 = bf16_rn(v - plane 0); products are formed as
 * hi*hi + mid*hi + hi*mid with fp32 accumulation (~2^-17 relative error per product, against ~2^-13 for TF32): this is what
 * keeps depth maps within north_star's 1e-3 of the fp32 reference on every pixel.  n % 8 == 0, 16-byte aligned pointers.
 *   mde_split_bf16       fp32 [n] -> planes [2][n]            mde_merge_bf16   planes -> fp32 (hi + mid)
 *   mde_split_bf16_nchw  fp32 NCHW [B,C,P] -> planes [2][B,P,C] (NHWC element order; C even) */
int mde_split_bf16(const float* x, uint16_t* planes, int64_t n, mde_stream_t stream);
int mde_merge_bf16(const uint16_t* planes, float* out, int64_t n, mde_stream_t stream);
int mde_split_bf16_nchw(const float* x_nchw, uint16_t* planes_nhwc, int B, int C, int64_t P, mde_stream_t stream);

/* ---- K1a: patch-embedding conv (kernel = stride = patch) + positional rows (models/layers.py:11-12,17-19) as a
 * TMA-fed tcgen05 split-K GEMM on split-bf16 pairs (three bf16 products per K step, see below).
 * x_pair: channels_last activations as a pair, planes[2][B,h,w,C]; w_pair: the conv filter in channels_last order
 * [E,patch,patch,C] as a pair, planes[2][E*patch*patch*C] (mde_split_bf16 of the permuted filter);
 * bias [E]; pos [>=S, E] (positional_encodings); tokens [S,B,E] fp32 with S = (h/patch)*(w/patch); E must be 128,
 * C % 8 == 0, (patch * C) % 64 == 0.  ws: mde_patch_embed_ws_floats(...) floats of scratch (split-K partials). */
int64_t mde_patch_embed_ws_floats(int B, int h, int w, int patch, int C);
int mde_patch_embed_fwd(const uint16_t* x_pair, const uint16_t* w_pair, const float* bias, const float* pos, float* tokens,
                        float* ws, int B, int h, int w, int C, int patch, int E, mde_stream_t stream);

/* ---- K1c: 3x3 / stride 1 / pad 1 convolution on channels_last activations as a TMA-fed tcgen05 implicit GEMM (fp32
 * accumulation) with a fused per-channel affine + LeakyReLU epilogue: mViT.conv3x3 (models/miniViT.py:16,27)
 * and the DecoderBN blocks Conv3x3 -> BatchNorm(eval) -> LeakyReLU / conv3 (models/unet_adaptive_bins.py:39-49,73).
 *   y = lrelu(conv(x) * scale[co] + shift[co]); scale / shift may be NULL (1 / 0; pass the conv bias as shift);
 *   lrelu_slope 1.0f = no activation.
 * mde_conv3x3_nhwc_x3_fwd (the model's default): x_pair planes[2][B,H,W,C] and w_pair planes[2][dx][dy][Cout][C]
 *   (mde_conv3x3_prep_weight_x3) are split-bf16 pairs, three bf16 products per K step; y is fp32 [B,H,W,Cout]
 *   (y_is_pair == 0) or a pair planes[2][B,H,W,Cout] for the next tensor-core consumer (y_is_pair != 0).
 *   Requires C % 8 == 0, Cout % 4 == 0 (% 8 for pair output).  products = 3: hi*hi + mid*hi + hi*mid (fp32-grade, the 1e-3
 *   contract); products = 1: the hi planes only, ONE bf16 product per K step -- the bf16 mode of the north star (depth within
 *   2e-2 of the fp32 reference): a third of the tensor work and half the operand traffic.
 * mde_conv3x3_nhwc_fwd (single-pass TF32): x_nhwc fp32 whose values are already TF32-representable (rounded by their
 *   producer -- the tensor core would otherwise truncate them), w_prep = [dx][dy][Cout][C] TF32-rounded
 *   (mde_conv3x3_prep_weight, operand_scale normally 1.0f); round_tf32 != 0 rounds the outputs to TF32.  C % 4 == 0.
 * Both: any such Cout (N tiles of <= 256 channels; the last tile may be ragged).
 * mde_conv3x3_small_nhwc_fwd: exact-fp32 direct kernel for Cout <= 4 (the noAdaBins decoder's conv3, :78-80);
 *   w_oihw is the torch-layout filter [Cout,C,3,3], bias may be NULL. */
int mde_conv3x3_prep_weight(const float* w_oihw, float* w_prep, int Cout, int C, float operand_scale,
                            mde_stream_t stream);
int mde_conv3x3_prep_weight_x3(const float* w_oihw, uint16_t* w_pair, int Cout, int C, mde_stream_t stream);
int mde_conv3x3_nhwc_fwd(const float* x_nhwc, const float* w_prep, const float* scale, const float* shift, float* y_nhwc,
                         int B, int H, int W, int C, int Cout, float lrelu_slope, int round_tf32, mde_stream_t stream);
int mde_conv3x3_nhwc_x3_fwd(const uint16_t* x_pair, const uint16_t* w_pair, const float* scale, const float* shift, void* y,
                            int y_is_pair, int B, int H, int W, int C, int Cout, float lrelu_slope, int products,
                            mde_stream_t stream);
int mde_conv3x3_small_nhwc_fwd(const float* x_nhwc, const float* w_oihw, const float* bias, float* y_nhwc, int B, int H, int W,
                               int C, int Cout, mde_stream_t stream);

/* ---- 1x1 convolution on channels_last activations as a tcgen05 GEMM with fp32-grade accuracy (DecoderBN.conv2,
 * models/unet_adaptive_bins.py:61; the point-wise convolutions of the EfficientNet passthrough body):
 *   y[m][n] = act(sum_k x[m][k] * w[n][k] + bias[n]) (+ residual[m][n])
 * x fp32 [M][K] (M = B*H*W pixels, K = C_in contiguous, K % 8 == 0); w_pair = split-bf16 pair of the [N][K] filter, planes
 * [2][N][K] (mde_split_bf16); bias [N] or NULL; act 0 none / 1 SiLU; residual fp32 [M][ldr] or NULL; y fp32 [M][ldc].
 * The activations are split into (hi, mid) bf16 inside the kernel (in shared memory), so callers pass plain fp32.
 * gate (optional, with rows_per_image = H*W): fp32 [M / rows_per_image][K] per-image channel scale multiplied into x (in fp32)
 * before the product -- the squeeze-excite gate of the MBConv block the 1x1 projection closes (geffnet SqueezeExcite.forward:
 * x * sigmoid(...)), saving the pass that would materialise x * gate.
 * out_pad (optional, host array {H, W, pad_top, pad_bottom, pad_left, pad_right}; residual must be NULL): row m = (b, y, x) is
 * written at (b, y + pad_top, x + pad_left) of a [B][H + pt + pb][W + pl + pr][ldc] map (the TensorFlow-SAME padding of the
 * stride-2 depthwise convolution that consumes it); the border is NOT written -- the caller zeroes it. */
int mde_pointwise_x3_fwd(const float* x, const float* gate, int64_t rows_per_image, const uint16_t* w_pair, const float* bias,
                         int act, const float* residual, float* y, int64_t M, int N, int K, int64_t ldc, int64_t ldr,
                         const int* out_pad, mde_stream_t stream);

/* ---- squeeze-excite helpers of the EfficientNet passthrough body (inference):
 * mde_bias_act_pool_nhwc: y[p][c] = act(x[p][c] + bias[c]) on [B][HW][C] (in place allowed) and, in the same pass, the
 *   per-(image, slab, channel) sums partial[B][slabs][C] of y (slabs = mde_pool_slabs(B, HW)); fixed summation order.
 * mde_se_gate: gate[b][c] = sigmoid(w2[c][:] . silu(w1 . mean_b + b1) + b2[c]) with mean_b = inv_hw * sum_s partial[b][s][:];
 *   w1 [R][C], w2 [C][R] (the 1x1 conv_reduce / conv_expand filters of geffnet's SqueezeExcite). */
/* mde_stem_conv3x3s2_nhwc: the encoder stem in exact fp32 -- 3x3 / stride 2 convolution (geffnet conv_stem, TensorFlow-SAME
 * padding as pad_top / pad_left + bounds), folded-BatchNorm bias and SiLU (act = 1): x [B][Hi][Wi][Cin] fp32 NHWC, Cin % 4 == 0;
 * w_tcc [9][Cin][Cout] (tap = dy*3 + dx, C_out innermost); y [B][Ho][Wo][Cout]; Cout in {32, 48}. */
int mde_stem_conv3x3s2_nhwc(const float* x, const float* w_tcc, const float* bias, float* y, int B, int Hi, int Wi, int Cin, int Cout,
                            int pad_top, int pad_left, int Ho, int Wo, int act, mde_stream_t stream);
int mde_pool_slabs(int B, int64_t HW);
/* mde_depthwise_bias_act_pool_nhwc: the depthwise k x k convolution (k = 3 / 5, stride 1 / 2, zero padding pad_top / pad_left and
 * whatever the output size implies at the far edges -- TensorFlow-SAME included), its folded-BatchNorm bias, SiLU (act = 1) and
 * the slab sums of the result in ONE pass: x [B][Hi][Wi][C] fp32 NHWC, w [k][k][C] (channel innermost), y [B][Ho][Wo][C],
 * partial [B][mde_pool_slabs(B, Ho*Wo)][C].  Plain fp32 FMAs (a depthwise convolution has no channel reduction). */
int mde_depthwise_bias_act_pool_nhwc(const float* x, const float* w, const float* bias, float* y, float* partial, int B, int Hi,
                                     int Wi, int C, int k, int stride, int pad_top, int pad_left, int Ho, int Wo, int act,
                                     mde_stream_t stream);
int mde_bias_act_pool_nhwc(const float* x, const float* bias, float* y, float* partial, int B, int64_t HW, int C, int act,
                           mde_stream_t stream);
int mde_se_gate(const float* partial, int slabs, float inv_hw, const float* w1, const float* b1, const float* w2,
                const float* b2, float* gate, int B, int C, int R, mde_stream_t stream);

/* Batched NT GEMM on tcgen05 (TF32 inputs, fp32 accumulate):  C[b][m][n] (+)= alpha * sum_k A[b][m][k] * B[b][n][k].
 * A [batch][M][K] with row pitch lda and batch stride a_batch (floats; multiples of 4), B [batch][N][K] likewise,
 * C [batch][M][N] with row pitch ldc / batch stride c_batch.  splits > 1 splits the K range over CTAs and accumulates
 * into C with atomicAdd (the caller zeroes C).  The building block of the head's backward pass (autograd of
 * models/layers.py:31-36 + unet_adaptive_bins.py:286): d feat = W'^T gl, d W' = gl^T feat. */
int mde_gemm_nt_tf32(const float* A, int64_t lda, int64_t a_batch, const float* B, int64_t ldb, int64_t b_batch, float* C,
                     int64_t ldc, int64_t c_batch, int batch, int M, int N, int K, int splits, float alpha,
                     mde_stream_t stream);
/* Same with an nn.Linear-style epilogue: + bias[N] (added once), act 0 none / 1 ReLU, and split_out != 0 writing each
 * output row as [v | v - trunc_tf32(v) | v] (ldc >= 3N) -- the A operand of a following 3xTF32 product
 * (A' = [a | a_lo | a] against B' = [b_hi | b_hi | b_lo] along K gives fp32-grade accuracy on the TF32 tensor cores). */
int mde_gemm_nt_tf32_ex(const float* A, int64_t lda, int64_t a_batch, const float* B, int64_t ldb, int64_t b_batch, float* C,
                        int64_t ldc, int64_t c_batch, int batch, int M, int N, int K, int splits, float alpha,
                        const float* bias, int act, int split_out, mde_stream_t stream);
/* Deterministic split-K: plane_stride > 0 makes K split s store its partial product at C + s * plane_stride floats (the
 * consumer sums the planes; the actual number of planes is min(splits, ceil(K / 32))). */
int mde_gemm_nt_tf32_planes(const float* A, int64_t lda, int64_t a_batch, const float* B, int64_t ldb, int64_t b_batch,
                            float* C, int64_t ldc, int64_t c_batch, int batch, int M, int N, int K, int splits, float alpha,
                            const float* bias, int act, int split_out, int64_t plane_stride, mde_stream_t stream);

/* C[M,N] = act(A[M,K] W[N,K]^T + bias[N]); act: 0 none, 1 ReLU, 2 LeakyReLU(0.01).  fp32 SIMT, row-major with leading
 * dimensions lda/ldw/ldc (the nn.Linear building block of the regressor and the encoder layers). */
int mde_linear_fwd(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc, int M, int N,
                   int K, int act, mde_stream_t stream);
/* normalisation + widths + cumsum edges + centres from a ready regressor output y_raw [B,n_bins]
 * (the tail of mde_regressor_bins_fwd; miniViT.py:36-44, unet_adaptive_bins.py:292-296) */
int mde_bins_finalize_fwd(float* y_raw, int B, int n_bins, int norm_mode, float min_val, float max_val,
                          float* widths_normed, float* edges, float* centers, mde_stream_t stream);

/* ---- K1b: one post-LN transformer encoder layer (models/layers.py:8-9,23: nn.TransformerEncoderLayer(128, 4, 1024),
 * ReLU, eps 1e-5, eval semantics).  Tokens x, y are [S, NB, E] row-major (row = s*NB + n), E = 128, E/heads = 32.
 * Parameter tensors keep torch's layouts: in_w [3E,E], out_w [E,E], l1_w [FF,E], l2_w [E,FF].
 * ws: mde_encoder_layer_ws_floats(S,NB,E,FF) floats of scratch.  y may not alias x. */
int64_t mde_encoder_layer_ws_floats(int S, int NB, int E, int FF);
int mde_encoder_layer_fwd(const float* x, float* y, const float* in_w, const float* in_b, const float* out_w,
                          const float* out_b, const float* ln1_w, const float* ln1_b, const float* l1_w, const float* l1_b,
                          const float* l2_w, const float* l2_b, const float* ln2_w, const float* ln2_b, float* ws, int S,
                          int NB, int E, int heads, int FF, float eps, mde_stream_t stream);

/* Tensor-core form of the same layer: the four nn.Linear products run on mde_gemm_nt_tf32_ex in 3xTF32.  x3: tokens in
 * split form [S*NB][3E] = [v | v - trunc_tf32(v) | v] (mde_split3_tf32 makes it from plain rows); the *_w3 weights are
 * [w_hi | w_hi | w_lo] along K (w_hi = round_tf32(w), w_lo = round_tf32(w - w_hi)): in_w3 [3E][3E], out_w3 [E][3E],
 * l1_w3 [FF][3E], l2_w3 [E][3FF].  y: next layer's split-form input (y_split != 0, [S*NB][3E]) or plain [S,NB,E].
 * ws: mde_encoder_layer_tc_ws_floats(S,NB,E,FF) floats. */
int64_t mde_encoder_layer_tc_ws_floats(int S, int NB, int E, int FF);
int mde_split3_tf32(const float* in, float* out, int64_t rows, int E, mde_stream_t stream);
int mde_encoder_layer_tc_fwd(const float* x3, float* y, int y_split, const float* in_w3, const float* in_b,
                             const float* out_w3, const float* out_b, const float* ln1_w, const float* ln1_b,
                             const float* l1_w3, const float* l1_b, const float* l2_w3, const float* l2_b, const float* ln2_w,
                             const float* ln2_b, float* ws, int S, int NB, int E, int heads, int FF, float eps,
                             mde_stream_t stream);

/* ---- K1d: range-attention contraction  y[b,n,p] = sum_k x[b,k,p] * q[b,n,k]  (layers.py:31-36) ----------
 * x [B,K,P] float32 (NCHW with P = h*w), q [B,N,K] float32, y [B,N,P] float32.
 * impl 0 = SIMT fp32 (exact fp32 FMA); the tensor-core form is mde_range_attention_tc (split-bf16 pair operands). */
int mde_range_attention(const float* x, const float* q, float* y, int B, int K, int N, int64_t P, int impl,
                        mde_stream_t stream);

/* ---- training: weight gradient of the 3x3 convolutions (autograd of models/miniViT.py:16 and the DecoderBN convs,
 * models/unet_adaptive_bins.py:39-49,73) on the NT GEMM above, ONE launch for the nine taps:
 *   dW9[ky*3+kx][co][ci] = sum_k dyT3[kx][co][k] * xT[ci][k + (ky-1)*Wp]
 * with k over the zero-padded pixel axis (b, y+1, x+1) of pitch Wp (>= W+2, % 4 == 0), Kp = ld = B*(H+2)*Wp.
 * mde_nhwc_to_cpad_tf32 builds the operands from NHWC fp32 tensors (channel-major, padded, TF32-rounded RNA so that the
 * tensor cores' operand read is exact): xT [Cin][ld] with shift3 == 0, and dyT3 [3][Cout][ld] with shift3 != 0 -- three
 * copies shifted by kx-1 along k, which bakes the horizontal tap into the data (the TMA cannot start a box at an
 * unaligned element of the contiguous axis; the vertical tap is an aligned shift of its K coordinate).
 * splits > 1: split-K with atomic accumulation (dW9 zeroed by the call).
 * The input gradient (dgrad) is mde_conv3x3_nhwc_x3_fwd on the spatially flipped, channel-transposed filter. */
int mde_nhwc_to_cpad_tf32(const float* x_nhwc, float* out, int B, int H, int W, int C, int Wp, int shift3, mde_stream_t stream);
int mde_conv3x3_wgrad_tf32(const float* dyT3, const float* xT, float* dW9, int Cout, int Cin, int64_t Kp, int64_t ld, int Wp,
                           int splits, mde_stream_t stream);

/* ---- K2: streaming bin pipeline  pred[b,p] = sum_j softmax_j(logits[b,:,p]) * centers[b,j]
 * (unet_adaptive_bins.py:286 Softmax(dim=1) and :298-300).  logits [B,n_bins,P], centers [B,n_bins], pred [B,P]. */
int mde_bins_pred_fwd(const float* logits, const float* centers, float* pred, int B, int n_bins, int64_t P,
                      mde_stream_t stream);
/* 1x1 conv 128 -> n_bins (+bias) on the range-attention maps (unet_adaptive_bins.py:190,286), SIMT fp32.
 * ram [B,K,P], w [n_bins,K], bias [n_bins] -> logits [B,n_bins,P] */
int mde_conv1x1_fwd(const float* ram, const float* w, const float* bias, float* logits, int B, int K, int n_bins,
                    int64_t P, mde_stream_t stream);

/* ---- K1d+K1e+K2 fused: range attention -> conv_out -> softmax -> centre-weighted sum, nothing but pred is
 * written (miniViT.py:33 + unet_adaptive_bins.py:286-300).  TMA-fed tcgen05 (three bf16 products per K step on split-bf16
 * pairs, fp32 accumulators in TMEM).
 *   x_pair   activations (the conv3x3 output) as a split-bf16 pair, planes[2][B,P,128] (NHWC element order)
 *   w_pair   planes[2][B,n_bins,128]: the pair of wf = (conv_out.weight @ queries[b]) * log2(e)   (mde_fold_queries
 *            + mde_split_bf16)
 *   biasf    [B,n_bins]     float32 = log2(e) * (conv_out.bias + wf-fold of the producer's bias)  (mde_fold_queries)
 *   centers  [B,n_bins], pred [B,P].   Requires P % 128 == 0, n_bins == 256. */
int mde_head_chain_fwd(const uint16_t* x_pair, const uint16_t* w_pair, const float* biasf, const float* centers, float* pred,
                       int B, int n_bins, int64_t P, mde_stream_t stream);
/* the same fused head with ONE bf16 product per K step (hi planes only; the mid planes are neither read nor multiplied): the
 * bf16 mode -- depth within 2e-2 of the fp32 reference (north star), half the operand traffic and a third of the MMAs */
int mde_head_chain_bf16_fwd(const uint16_t* x_pair, const uint16_t* w_pair, const float* biasf, const float* centers,
                            float* pred, int B, int n_bins, int64_t P, mde_stream_t stream);
/* Training forms of the fused chain (autograd of layers.py:31-36 + unet_adaptive_bins.py:286-300 in hand-written form):
 *  - mde_head_chain_fwd_train: the forward, additionally storing the per-pixel softmax state stats [B,P,2]
 *    (max logit in log2 units, sum_j 2^(z_j - max));
 *  - mde_head_chain_bwd_logits: recomputes the logits on the tensor cores and writes d loss / d logit (natural-log
 *    units, TF32-rounded) as gl [B,P,n_bins] and glT [B,n_bins,P], plus gc [B,n_bins] = d loss / d centres and
 *    gb [B,n_bins] = sum_p gl (both zeroed by the call).  gpred [B,P] is the upstream gradient of pred.
 *  The remaining products (d feat = gl W', d W' = gl^T feat) are mde_gemm_nt_tf32 calls. */
int mde_head_chain_fwd_train(const uint16_t* x_pair, const uint16_t* w_pair, const float* biasf, const float* centers,
                             float* pred, float* stats, int B, int n_bins, int64_t P, mde_stream_t stream);
int mde_head_chain_bwd_logits(const uint16_t* x_pair, const uint16_t* w_pair, const float* biasf, const float* centers,
                              const float* pred, const float* stats, const float* gpred, float* gl, float* glT, float* gc,
                              float* gb, int B, int n_bins, int64_t P, mde_stream_t stream);
/* Stand-alone PixelWiseDotProduct (layers.py:31-36) on the same kernel: y[b,n,p] = sum_k x[b,p,k] q[b,n,k] with x_pair
 * planes[2][B,P,128] and q_pair planes[2][B,128,128]; y fp32 [B,128,P] (NCHW).  K == N == 128, P % 128 == 0. */
int mde_range_attention_tc(const uint16_t* x_pair, const uint16_t* q_pair, float* y, int B, int K, int N, int64_t P,
                           mde_stream_t stream);
/* wf[b] = (w_out [n_bins,N] @ q[b] [N,K]) * log2e * operand_scale   (fp32 FMA, tiled; round_tf32 != 0 additionally rounds
 *         to TF32, RNA)
 * biasf[b,j] = log2e * ( bias[j] + sum_k (w_out @ q[b])[j,k] * feat_bias[k] )            (feat_bias may be NULL)
 * q[b] must be a dense [N,K] block (q_batch_stride == N*K). */
int mde_fold_queries(const float* w_out, const float* bias, const float* q, int64_t q_batch_stride,
                     const float* feat_bias, float* wf, float* biasf, int B, int n_bins, int N, int K,
                     float operand_scale, int round_tf32, mde_stream_t stream);
/* out[i] = round-to-nearest TF32 of in[i]*scale (so the tensor cores' operand truncation is exact for this tensor) */
int mde_round_tf32(const float* in, float* out, int64_t n, float scale, mde_stream_t stream);
/* last barrier-timeout code of the chain kernels (0 = none; synchronises the device).  Test/debug only. */
int mde_tc_last_error(void);
/* tuning aid: device buffer [148][8] int64 of per-role wait cycles filled by the following mde_head_chain_fwd launches
 * (NULL = off; the instrumented kernel is a separate instantiation -- the default one carries no clock reads) */
int mde_tc_debug_profile(long long* buf);

/* ---- A8': noAdaBins epilogue relu(x) + 1e-4 (unet_adaptive_bins.py:240-242) */
int mde_relu_eps_fwd(const float* x, float* y, int64_t n, float eps, mde_stream_t stream);

/* ---- "next" row (f)1: DecoderBN up-sampling step (models/unet_adaptive_bins.py:51-54): bilinear align_corners=True
 * resize of x [B,C1,h,w] to HxW fused with torch.cat((up_x, skip [B,C2,H,W]), 1) -> out [B,C1+C2,H,W]. */
int mde_upsample_concat_fwd(const float* x, const float* skip, float* out, int B, int C1, int C2, int h, int w, int H,
                            int W, mde_stream_t stream);
/* gradient of the resize part, deterministic gather: gout [B,Ctot,H,W] -> gx [B,C1,h,w] (channels_last == 0) or
 * gout [B,H,W,Ctot] -> gx [B,h,w,C1] (channels_last != 0); only the first C1 channels of gout are read.
 * ws: mde_upsample_bwd_ws_bytes(h, w) bytes of scratch (per-axis inverse tap tables, rebuilt by every call). */
int64_t mde_upsample_bwd_ws_bytes(int h, int w);
int mde_upsample_bwd(const float* gout, float* gx, int channels_last, int B, int C1, int Ctot, int h, int w, int H, int W,
                     void* ws, mde_stream_t stream);

/* channels_last variant feeding mde_conv3x3_nhwc_fwd: x_nhwc [B,h,w,C1]; skip [B,H,W,C2] (skip_channels_last != 0) or
 * [B,C2,H,W]; out_nhwc [B,H,W,C1+C2].  C1 % 4 == 0 and C2 % 4 == 0. */
int mde_upsample_concat_nhwc_fwd(const float* x_nhwc, const float* skip, int skip_channels_last, float* out_nhwc, int B,
                                 int C1, int C2, int h, int w, int H, int W, mde_stream_t stream);
/* the same step writing its result as a split-bf16 pair planes[2][B,H,W,Cpitch] (feeds mde_conv3x3_nhwc_x3_fwd);
 * skip must be NHWC.  Cpitch >= C1 + C2, Cpitch % 8 == 0; channels [C1 + C2, Cpitch) are written as zeros -- a pitch rounded up
 * to 32 channels keeps every 64-byte TMA box row of the convolution sector-aligned (the conv then runs with C = Cpitch and a
 * filter zero-padded to match: same K chunks, same result). */
int mde_upsample_concat_nhwc_pair_fwd(const float* x_nhwc, const float* skip_nhwc, uint16_t* out_pair, int B, int C1, int C2,
                                      int Cpitch, int h, int w, int H, int W, mde_stream_t stream);

/* NCHW [B,C,P] -> NHWC [B,P,C] transpose (feeds the head's cuDNN convs and the K-major chain operand) */
int mde_nchw_to_nhwc(const float* in, float* out, int B, int C, int64_t P, mde_stream_t stream);

/* Same transpose into a channel slice of a wider channels_last tensor (row pitch out_pitch channels; `out` points at the
 * slice's first channel): concatenates planar sources (image planes, embedding planes) into the NHWC encoder input. */
int mde_nchw_to_nhwc_slice(const float* in, float* out, int B, int C, int64_t P, int out_pitch, mde_stream_t stream);

/* ... and into a zero-padded channels_last image [B, H+pad_top+pad_bottom, W+pad_left+pad_right, out_pitch]; the border is
 * not written (the caller zeroes it).  Folds the F.pad of a TensorFlow-"SAME" stem convolution into the concatenation. */
int mde_nchw_to_nhwc_slice_padded(const float* in, float* out, int B, int C, int H, int W, int out_pitch, int pad_top,
                                  int pad_bottom, int pad_left, int pad_right, mde_stream_t stream);

/* ---- K4: SILog loss (loss.py:12-25).  pred [B,1,h,w] float32; target [B,1,H,W] float32; mask uint8/bool
 * [B,1,H,W] or NULL (all pixels); interpolate != 0: bilinear align_corners=True resampling of pred to HxW is
 * fused (never materialised).  ws: >= mde_silog_ws_bytes() bytes of scratch (holds {sum g, sum g^2, n} as float64
 * afterwards, which mde_silog_bwd reads, followed by per-block partial sums).  loss: float32 scalar.  One launch, no host
 * synchronisation, bit-reproducible (fixed summation order). */
int64_t mde_silog_ws_bytes(void);
int mde_silog_fwd(const float* pred, const float* target, const uint8_t* mask, int B, int h, int w, int H, int W,
                  int interpolate, void* ws, float* loss, mde_stream_t stream);
/* grad_pred [B,1,h,w] (zeroed by the call, accumulated with atomics) given upstream scalar grad (device ptr) */
int mde_silog_bwd(const float* pred, const float* target, const uint8_t* mask, int B, int h, int w, int H, int W,
                  int interpolate, const void* ws, const float* grad_loss, float* grad_pred, mde_stream_t stream);

/* ---- K5: bin-centre chamfer loss (loss.py:33-46 -> pytorch3d.loss.chamfer_distance, K=1 squared L2 both ways,
 * point mean, batch mean; targets < min_target dropped).  edges [B,n_bins+1] float32; target [B,HW] float32.
 * The centres must be ascending (they are for any edges the model produces: bin widths are positive); the kernel checks
 * and returns NaN for an unsorted centre vector rather than a silently wrong value.
 * ws: >= mde_chamfer_ws_bytes(B,n_bins) bytes scratch (keeps per-image/per-centre statistics for mde_chamfer_bwd;
 * those per-centre sums are only collected when want_grad != 0).  loss: float32 scalar (NaN if an image has no valid
 * target, like the reference's 0/0). */
int64_t mde_chamfer_ws_bytes(int B, int n_bins);
int mde_chamfer_fwd(const float* edges, const float* target, int B, int n_bins, int64_t HW, float min_target, int want_grad,
                    void* ws, float* loss, mde_stream_t stream);
int mde_chamfer_bwd(const float* edges, int B, int n_bins, const void* ws, const float* grad_loss,
                    float* grad_edges, mde_stream_t stream);

/* ---- K4 + K5 fused: both losses of train.py:414-419 in ONE pass over the target depth (the depth
 * read is shared, 1.13 MB/img): SILog with mask = target > silog_min_depth (train.py:414) and chamfer over targets >=
 * chamfer_min_target (loss.py:40).  Same scratch buffers and backward entry points as the separate forms; the SILog
 * backward for the derived mask is mde_silog_bwd_thr. */
int mde_depth_losses_fwd(const float* pred, const float* edges, const float* target, int B, int h, int w, int H, int W,
                         int n_bins, int interpolate, float silog_min_depth, float chamfer_min_target, int want_grad,
                         void* silog_ws, void* chamfer_ws, float* silog_loss, float* chamfer_loss, mde_stream_t stream);
int mde_silog_bwd_thr(const float* pred, const float* target, float min_depth, int B, int h, int w, int H, int W,
                      int interpolate, const void* ws, const float* grad_loss, float* grad_pred, mde_stream_t stream);

/* ---- "next" row (f)2: evaluation epilogue + metrics (evaluate.py:50-71,128-152; train.py:543-568; utils.py:119-139).
 * pred [B,1,h,w] (bilinearly up-sampled to HxW with align_corners=True inside the kernel when h,w != H,W), gt [B,1,H,W];
 * pred is clipped to [min_depth_eval, max_depth_eval] (nan -> min, inf -> max); valid = gt in (min, max) inside the crop
 * box rows [crop_y0, crop_y1) x cols [crop_x0, crop_x1) (pass 0,H,0,W for no crop).  out [B,10] float32 per image:
 * a1 a2 a3 abs_rel rmse log_10 rmse_log silog sq_rel n_valid (NaN metrics when n_valid == 0, like numpy's empty mean).
 * ws: mde_eval_metrics_ws_bytes(B) bytes of scratch (zeroed by the call). */
int64_t mde_eval_metrics_ws_bytes(int B);
int mde_eval_metrics_fwd(const float* pred, const float* gt, int B, int h, int w, int H, int W, float min_depth_eval,
                         float max_depth_eval, int crop_y0, int crop_y1, int crop_x0, int crop_x1, void* ws, float* out,
                         mde_stream_t stream);
/* mirror test-time augmentation of infer.py:108-118: out = 0.5 * (clip(a, lo, hi) + clip(flip_w(b_flipped), lo, hi));
 * a, b_flipped, out are [rows, w] float32 (rows = B*h). */
int mde_flip_average(const float* a, const float* b_flipped, float* out, int64_t rows, int w, float lo, float hi,
                     mde_stream_t stream);

/* ---- section 8(e): SyncBatchNorm building blocks (train.py:296, nn.SyncBatchNorm semantics: statistics over the global
 * batch) on channels_last activations x [N pixels, C] (C % 4 == 0).  A layer is
 *   forward : mde_bn_stats_nhwc -> all-reduce(SUM) of stats (+ the pixel count) over ranks -> mde_bn_apply_nhwc
 *   backward: mde_bn_bwd_reduce_nhwc -> all-reduce(SUM) of sums -> mde_bn_bwd_apply_nhwc
 * stats / sums: float64 [2C] = per-channel (sum x, sum x^2) / (sum dy, sum dy*xhat).  zero_next == NULL: the producing
 * call zeroes stats / sums itself (memset); zero_next != NULL: stats / sums must already be zero and the kernel zeroes
 * zero_next [2C] for the NEXT call (double buffering: no memset launches in the training loop);
 * count = global number of pixels per channel; save_mean / save_invstd [C] are written by apply and read by the backward;
 * running_mean / running_var (may be NULL) get torch's momentum update with the unbiased variance.
 * The local (pre-all-reduce) sums are d bias and d weight. */
int mde_bn_stats_nhwc(const float* x, int64_t N, int C, double* stats, double* zero_next, mde_stream_t stream);
int mde_bn_apply_nhwc(const float* x, float* y, int64_t N, int C, const double* stats, double count, const float* weight,
                      const float* bias, float eps, float* save_mean, float* save_invstd, float* running_mean,
                      float* running_var, float momentum, mde_stream_t stream);
int mde_bn_bwd_reduce_nhwc(const float* x, const float* dy, int64_t N, int C, const float* mean, const float* invstd,
                           double* sums, double* zero_next, mde_stream_t stream);
int mde_bn_bwd_apply_nhwc(const float* x, const float* dy, float* dx, int64_t N, int C, const float* mean,
                          const float* invstd, const float* weight, const double* sums, double count, mde_stream_t stream);

/* Fused statistics exchange over NVLink peer memory instead of an all-reduce call (one process per GPU, <= 8 ranks): every
 * rank owns a symmetric-memory arena mapped by all peers; peer_bases[r] (HOST array of `world` device addresses) is rank r's
 * arena in this process.  The *_p2p statistics kernels publish the rank's [2C] float64 partial sums into slot [rank] of every
 * peer's arena at slot_off (bytes; room for world * 2C doubles) and then raise flag [rank] = epoch at flag_off (world uint64
 * per arena); the *_p2p consumers spin on the `world` flags of their own arena (my_base) until all reached `epoch`, sum the
 * partials and continue like the non-p2p kernels.  The caller alternates two (slot_off, flag_off) pairs by epoch parity and
 * increases epoch by one per call; `local` is the rank's own float64 [2C+1] scratch row (sums + a ticket), zero on entry,
 * `zero_next` the row of the next call (zeroed by this one). */
int mde_bn_stats_p2p_nhwc(const float* x, int64_t N, int C, double* local, double* zero_next, const uint64_t* peer_bases,
                          int world, int rank, int64_t slot_off, int64_t flag_off, uint64_t epoch, mde_stream_t stream);
int mde_bn_apply_p2p_nhwc(const float* x, float* y, int64_t N, int C, uint64_t my_base, int64_t slot_off, int64_t flag_off,
                          int world, uint64_t epoch, double count, const float* weight, const float* bias, float eps,
                          float* save_mean, float* save_invstd, float* running_mean, float* running_var, float momentum,
                          mde_stream_t stream);
int mde_bn_bwd_reduce_p2p_nhwc(const float* x, const float* dy, int64_t N, int C, const float* mean, const float* invstd,
                               double* local, double* zero_next, const uint64_t* peer_bases, int world, int rank,
                               int64_t slot_off, int64_t flag_off, uint64_t epoch, mde_stream_t stream);
int mde_bn_bwd_apply_p2p_nhwc(const float* x, const float* dy, float* dx, int64_t N, int C, const float* mean,
                              const float* invstd, const float* weight, uint64_t my_base, int64_t slot_off, int64_t flag_off,
                              int world, uint64_t epoch, double count, mde_stream_t stream);
/* CUDA-graph replay of a training step (training.GraphedTrainStep): kernel arguments are frozen at capture, but SyncBatchNorm's
 * epoch must advance by one per step.  While a replay counter is set (device address of a uint64 that the graph's first node
 * increments once per replay), the *_p2p launches take `epoch` as a BASE that the kernels add to the counter's current value,
 * and slot_off / flag_off as the parity-0 offsets of the direction (the kernels add (epoch & 1) * world * 2C * 8 resp.
 * world * 8 bytes themselves).  NULL restores the eager meaning.  Host-side state of the process (one process per GPU). */
int mde_bn_p2p_set_epoch_counter(const uint64_t* counter);
/* Bound of the peer-flag wait inside the *_p2p kernels (seconds, default 600: all ranks must enter every BatchNorm layer
 * within it -- the role NCCL's collective timeout plays for the reference's nn.SyncBatchNorm, train.py:296).  The wait backs
 * off with __nanosleep; on expiry the rank counts the event (mde_bn_peer_timeouts) and traps.  Both calls synchronise. */
int mde_bn_set_peer_timeout_seconds(double seconds);
int mde_bn_peer_timeouts(void);
/* diagnostic: time (ns) block (0,0) of the consuming kernels spent waiting for peer flags, and the number of waits, since the
 * last reset -- the rank skew + NVLink latency SyncBatchNorm adds to a training step (bench.py reports it per step) */
int mde_bn_wait_stats(uint64_t* wait_ns, uint64_t* waits, int reset);

#ifdef __cplusplus
}
#endif
#endif /* MDE_B200_H */
