/*
 * instantrestore_b200 — C ABI of the B200 (sm_100a) kernels behind the InstantRestore hot path.
 *
 * The reference (snap-research/InstantRestore) is pure Python: it has no FFI. Every entry point below
 * replaces a chain of torch/diffusers library calls reached from the cited reference lines; the
 * reference-side binding is the ctypes stub in INTEGRATION.md (and instantrestore_b200/_lib.py here).
 *
 * Conventions
 *  - Plain pointers + sizes only; the caller owns every buffer (torch caching allocator on the Python side).
 *    Kernels never allocate or free device memory and never synchronise; they launch on `stream`.
 *  - Activations are fp16, channel-last ("NHWC" == token-major [B, H*W, C]); accumulation is fp32.
 *  - Return 0 on success, a negative IR_ERR_* otherwise; ir_last_error_string() describes the last failure
 *    on the calling thread. Nothing throws across the ABI.
 *  - All entry points are re-entrant and CUDA-graph capturable (TMA descriptors travel as kernel parameters).
 *  - Device pointers must be 16-byte aligned; row strides are in ELEMENTS and must be multiples of 8.
 */
#ifndef INSTANTRESTORE_B200_H_
#define INSTANTRESTORE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IR_OK 0
#define IR_ERR_SHAPE (-1)       /* unsupported / inconsistent shape */
#define IR_ERR_ALIGN (-2)       /* misaligned pointer or stride */
#define IR_ERR_ARCH (-3)        /* device is not sm_100 */
#define IR_ERR_CUDA (-4)        /* CUDA runtime / driver error */
#define IR_ERR_ARG (-5)         /* NULL or invalid argument */

#define IR_ACT_NONE 0
#define IR_ACT_GEGLU 1          /* out[:, j] = v[:, j] * gelu_erf(g[:, j]); weights interleaved, see ir_conv_gemm */
#define IR_ACT_SILU 2

typedef void* ir_stream_t;      /* cudaStream_t */

const char* ir_last_error_string(void);
int ir_version(void);
/* 0 when the current device is a B200-class (sm_100) GPU, IR_ERR_ARCH / IR_ERR_CUDA otherwise. */
int ir_check_device(void);
/* Programmatic dependent launch (cudaLaunchAttributeProgrammaticStreamSerialization) for every kernel launched or
 * captured after the call: each kernel runs its set-up before griddepcontrol.wait, so consecutive kernels of one
 * stream overlap tail and prologue. Shortens a single request's critical path; off by default because blocked
 * dependents cost throughput when several requests are in flight. Returns the previous setting. Env: IR_PDL=1. */
int ir_set_pdl(int enabled);
/* Number of kernel launches issued (or captured into a CUDA graph) through this library since load. */
unsigned long long ir_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * ir_conv_gemm — im2col-free implicit GEMM on tcgen05 tensor cores (TMA-staged, TMEM accumulators).
 * Replaces nn.Conv2d 3x3 / 3x3 stride-2 / 1x1 and nn.Linear as reached from
 *   reference face_replace/models/unet_2d_condition/block.py:1061,1106,2239,2282,2397,2414 (ResnetBlock2D,
 *   Downsample2D, Upsample2D convs), unet.py:289,612 (conv_in/conv_out) and the q/k/v/out/proj/ff Linears
 *   called from face_replace/models/attn_processors.py:222,229-230,267.
 *
 *   out[m, n] = epilogue( sum_{tap, c} A[pixel(m) + tap, c] * W[n, tap * c_in + c] )
 *
 * A: fp16 [batch, h_in, w_in, >= c_in] (pixel stride a_row_stride elements). ksize 1 (stride 1) or 3
 *    (stride 1 or 2, zero padding 1). A linear layer is ksize=1 with batch=1, h_in=1, w_in=tokens.
 * W: fp16 [c_out, ksize*ksize*c_in], K contiguous, tap-major then channel.  c_in % 64 == 0.
 * Epilogue: + bias[n] (fp32, optional); round to fp16; + residual[m, n] (fp16, optional);
 *    act: IR_ACT_GEGLU expects W rows interleaved in blocks of 64 (64 value rows, then their 64 gate rows)
 *    and writes c_out/2 columns; IR_ACT_SILU applies x*sigmoid(x).
 * out: fp16 [M, c_out or c_out/2] with row stride out_row_stride.
 * Small-M layers (few output tiles, large K) are split along K across a thread-block cluster and reduced through
 * distributed shared memory in rank order: results are deterministic (no atomics); the split factor depends on the
 * launch shape only.
 */
typedef struct {
  const void* a;
  int batch, h_in, w_in, c_in;
  int a_row_stride;
  int ksize, stride;
  const void* w;
  int c_out;
  const float* bias;
  const void* residual;
  int res_row_stride;
  int act;
  void* out;
  int out_row_stride;
  int tile_n;  /* 0 = auto; else 64, 128, 160 or 256 */
  int split_k; /* 0 = auto; 1 = off; 2, 4, 8 = K split over a thread-block cluster of that size (DSMEM reduce) */
  int m_sub; /* A-B measurement: 0 = auto, 1 = never stack two M tiles per CTA */
  int no_persistent; /* A-B measurement: 0 = auto, 1 = one tile per CTA, 2 = always the persistent kernel (needs split_k <= 1) */
  int pad_hi_only; /* 3x3 stride 2 only: 0 = zero padding 1 on every side; 1 = one row/column of zeros at the
                      bottom/right only (diffusers Downsample2D(padding=0) of the VAE encoder) */
  int cta_pair; /* A-B measurement: 0 = auto, 1 = never, 2 = always the CTA-pair (tcgen05 cta_group::2, 256 x 256
                   tiles over two SMs) kernel; needs c_out % 256 == 0, no K split, >= 2 M tiles */
  void* gn_partial; /* optional fp32 [batch, h_out*w_out/32, gn_groups, 2]: (mean, M2) of every 32-pixel slab of the
                       stored outputs per GroupNorm group — pass A of the NEXT GroupNorm, computed in the epilogue (or by
                       a separate pass when the kernel chosen for the shape has no fused statistics). Needs
                       c_out / gn_groups in {4, 8, 16} and h_out*w_out % 128 == 0. Hand it to ir_groupnorm.partial_in */
  int gn_groups;
  int halo; /* A-B measurement: 0 = auto, 1 = never, 2 = always the halo kernel (3x3 stride 1, w_in % 128 == 0, even
               h_in, c_out % 128 == 0): each 64-channel input slice is staged once per tile as a (rows+2) x 130 pixel box
               and the nine taps read shifted views of it. On the halo kernel cta_pair = 1 keeps 128-wide outputs
               (c_out % 256 != 0) on the single-CTA kernel; the default is the 128-wide CTA pair, whose epilogue moves the
               residual and the outputs with TMA (out / residual 16-byte aligned, res_row_stride % 8 == 0; otherwise the
               single-CTA kernel runs) */
  int wide_io; /* A-B measurement: 0 = auto (256-bit residual loads / output stores in the epilogue whenever out and
                  residual rows are 32-byte aligned; results are bit-identical), 1 = always 128-bit */
  void* col_partial; /* optional fp32 [M / 32, c_out - col_begin, 2]: per (32-row slab, output column >= col_begin) the (mean, M2)
                        of the stored fp16 outputs, written by the GEMM epilogue (AdaIN statistics of the V third of a fused
                        QKV projection: reference attn_processors.py:7-10, 244-245 `x.mean / x.std over tokens`). Needs a plain
                        epilogue (no residual / activation), M % 128 == 0, c_out % 32 == 0, col_begin % 32 == 0. Hand it to
                        ir_adain_coeffs.own_partial / ref_partial */
  int col_begin;
  int upsample2x; /* != 0: nearest-neighbour 2x upsampling followed by the 3x3 stride-1 convolution (diffusers Upsample2D:
                     F.interpolate(scale_factor=2.0, mode="nearest") + conv; reference block.py:2366,2476 and the VAE decoder's
                     upsamplers) WITHOUT materialising the upsampled tensor: each of the four output sub-pixel phases
                     (py, px) is a 2x2 convolution on the low-resolution input whose taps are sums of the 3x3 taps that land
                     on the same input pixel (4/9 of the multiply-adds). h_in, w_in are the LOW-resolution sizes; out is
                     [batch, 2*h_in, 2*w_in, c_out]; w is the phase-folded matrix [4 phases (py*2+px)][c_out][2x2 taps (a*2+b)][c_in]
                     with tap (a, b) of phase (py, px) reading input pixel (y + py - 1 + a, x + px - 1 + b). Needs ksize 3,
                     stride 1, no residual / GEGLU / col_partial; never runs on the halo kernel */
  int tma_store; /* A-B measurement: 0 = auto, 1 = never, 2 = always the TMA-store epilogue of the persistent kernel (output
                    tile packed into a swizzled shared-memory box, one cp.async.bulk.tensor store per 128 x 64 half tile
                    instead of per-thread row stores; results are bit-identical). Needs 128-wide full N tiles
                    (c_out % 128 == 0), no residual / GEGLU / upsample2x */
} ir_conv_gemm_params;
int ir_conv_gemm(const ir_conv_gemm_params* p, ir_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * ir_shared_attn_fwd — fused shared-image attention: softmax(Q K^T * scale) V with K/V streamed from
 * [own tokens (optional)] ++ [n_ref reference images], online softmax, AdaIN on reference values applied
 * as a per-(batch, ref, channel) affine  V' = adain_scale * V + adain_shift  inside the kernel.
 * Replaces reference face_replace/models/attn_processors.py:232-264 (SharedAttnProcessor.forward: head split,
 * adain(), torch.cat, get_attention_scores (baddbmm + softmax), torch.bmm, batch_to_head_dim) and the same
 * lines of AttnProcessor.forward (:76-82).  head_dim is 64.
 *
 * q:      fp16 [batch, s_q, *]   head h occupies columns q_col_off + 64*h .. +63 (row stride q_row_stride)
 * k_own/v_own: fp16 [batch or 1, s_own, *] (NULL: no own chunk); s_own need not be a multiple of the KV tile
 *         (tail keys are masked). own_shared != 0: the same K/V serves every batch entry (constant caption).
 * k_ref/v_ref: fp16 [batch, n_ref, s_ref, *], head h at columns ref_col_off + 64*h (NULL when n_ref == 0).
 *         Padded reference slots are expected to be ZERO-FILLED by the caller (reference quirk,
 *         pix2pix_turbo.py:269-273): they still receive softmax mass.
 * adain_scale/shift: fp32 [batch, n_ref, heads*64] or NULL.
 * out:    fp16 [batch, s_q, heads*64] (row stride out_row_stride).
 * chunk_mass: optional fp32 [batch, heads, n_chunks] — attention probability mass per KV chunk ([own], ref 0, ...)
 *         averaged over queries (what gradio_demo.py:118-133 derives from the dense matrix). NULL to skip. Needs
 *         n_ref > 0 and `workspace` of batch*heads*ceil(s_q/256)*256*(n_chunks+1)*8 bytes; disables KV splitting.
 * When batch * heads * ceil(s_q / 256) cannot fill the 148 SMs (a single identity), the KV sequence is split into
 * ranges handled by separate CTAs; partial (O, max, sum) land in `workspace` and are merged in split order.
 */
typedef struct {
  const void* q;
  int q_row_stride, q_col_off;
  const void* k_own;
  const void* v_own;
  int own_row_stride, k_own_col_off, v_own_col_off;
  int s_own;
  int own_shared;
  const void* k_ref;
  const void* v_ref;
  int ref_row_stride, ref_col_off;
  int n_ref, s_ref;
  const float* adain_scale;
  const float* adain_shift;
  int batch, heads, s_q;
  float scale;
  void* out;
  int out_row_stride;
  float* chunk_mass;
  int kv_splits;          /* 0 = auto; 1 = off; 2..16 = KV ranges processed by separate CTAs and merged (needs workspace) */
  void* workspace;        /* device scratch for split-KV partials (may be NULL: splitting is then disabled) */
  size_t workspace_bytes;
} ir_shared_attn_params;
/* Upper bound of the split-KV scratch for (batch, heads, s_q); 0 when the shape never splits. */
size_t ir_shared_attn_workspace_bytes(int batch, int heads, int s_q);
int ir_shared_attn_fwd(const ir_shared_attn_params* p, ir_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * ir_groupnorm — GroupNorm(+optional SiLU) on channel-last fp16, statistics in fp32 (slab-parallel partial moments
 * merged with Chan's formula, then one affine(+SiLU) pass).
 * Replaces diffusers ResnetBlock2D.norm1/norm2 + nonlinearity, Transformer2DModel.norm and
 * unet.py:1167-1169 (conv_norm_out + conv_act).
 * x: fp16 [batch, hw, channels] (row stride x_row_stride); gamma/beta fp32 [channels];
 * out fp16 [batch, hw, channels]; workspace: ir_groupnorm_workspace_bytes(batch, groups) bytes of device memory
 * (per-(batch, slab, group) partial moments), caller-owned.
 */
typedef struct {
  const void* x;
  int x_row_stride;
  int batch, hw, channels, groups;
  float eps;
  const float* gamma;
  const float* beta;
  int silu;
  void* out;
  int out_row_stride;
  void* workspace;
  const void* partial_in; /* optional: slab moments written by ir_conv_gemm (gn_partial) for exactly this x; pass A is skipped */
  int fused; /* 0 = auto: ONE launch with ONE read of x (thread-block cluster per (image, channel slice), tile in
                registers, per-CTA moments merged through distributed shared memory) whenever
                ir_groupnorm_fused_supported() — tensors up to 12 MB whose rows split into whole groups x whole 16-byte
                vectors (the UNet norms of a single-identity step); partial_in and workspace are then unused. 1 = never (three-kernel path), 2 = require it (A-B measurement, tests) */
} ir_groupnorm_params;
size_t ir_groupnorm_workspace_bytes(int batch, int groups);
int ir_groupnorm_fused_supported(int batch, int hw, int channels, int groups); /* 1 / 0 */
int ir_groupnorm(const ir_groupnorm_params* p, ir_stream_t stream);

/* ir_layernorm — LayerNorm over the last dim of fp16 [rows, channels]; fp32 statistics.
 * Replaces diffusers BasicTransformerBlock.norm1/2/3 (reached from block.py:2355-2362). */
typedef struct {
  const void* x;
  int x_row_stride;
  int rows, channels;
  float eps;
  const float* gamma;
  const float* beta;
  void* out;
  int out_row_stride;
} ir_layernorm_params;
int ir_layernorm(const ir_layernorm_params* p, ir_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * ir_adain_coeffs — AdaIN statistics folded to an affine (attn_processors.py:7-18, 244-245):
 *   style_mean/std over tokens of the own V;  content mean/std over tokens of each reference V;
 *   scale = style_std / content_std,  shift = style_mean - content_mean * scale
 *   (std unbiased, +1e-5 added to the std as in the reference).
 * v_own: fp16 [batch, s_own, *] columns v_col_off .. +channels; v_ref: fp16 [batch, n_ref, s_ref, *].
 * scale/shift: fp32 [batch, n_ref, channels].
 * workspace: ir_adain_workspace_bytes(batch, n_ref, channels) bytes (per-slab partial moments), caller-owned.
 * With own_partial / ref_partial the mean / var reduction has already happened in the epilogue of the GEMM that produced V
 * (north_star: "AdaIN mean/var ... fused into the preceding conv epilogue"); what is left is one merge launch.
 */
typedef struct {
  const void* v_own;
  int own_row_stride, v_col_off, s_own;
  const void* v_ref;
  int ref_row_stride, ref_col_off, n_ref, s_ref;
  int batch, channels;
  float eps;
  float* scale;
  float* shift;
  void* workspace;
  const void* own_partial; /* optional, both or neither: (mean, M2) per (32-token slab, channel) of the own V [batch, s_own/32, channels] */
  const void* ref_partial; /* and of the reference V [batch, n_ref, s_ref/32, channels], as written by ir_conv_gemm.col_partial in the
                              epilogues of the QKV projections (padded reference slots: zero-filled like their V). Then v_own /
                              v_ref / workspace are not read and the statistics pass over V does not run. */
} ir_adain_coeffs_params;
size_t ir_adain_workspace_bytes(int batch, int n_ref, int channels);
int ir_adain_coeffs(const ir_adain_coeffs_params* p, ir_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * ir_concat_freeu — builds the up-block resnet input cat([hidden, skip], channels) and applies FreeU
 * (block.py:2314-2325, 2442-2453, 3495-3520): hidden[:, :C_h/2] *= b and
 * skip <- fourier_filter(skip, threshold=1, scale=s) evaluated in closed form (4 DFT coefficients per plane).
 * backbone_scale == 1 and skip_scale == 1 give a plain concat.
 * hidden fp16 [batch, hw, c_hidden]; skip fp16 [batch, hw, c_skip]; out fp16 [batch, hw, c_hidden + c_skip].
 * (The reference scales `hidden` in place; that tensor has no other reader, so only `out` carries the result.)
 */
typedef struct {
  const void* hidden;
  const void* skip;
  int batch, h, w, c_hidden, c_skip;
  float backbone_scale;
  float skip_scale;
  void* out;
  int two_pass; /* A-B measurement: 0 = auto (8x8 and 16x16 FreeU stages run as ONE launch with the skip planes in registers),
                   1 = always the concat kernel followed by the two-pass filter kernel (same sums, results within an fp16 ulp) */
} ir_concat_freeu_params;
int ir_concat_freeu(const ir_concat_freeu_params* p, ir_stream_t stream);

/* ir_upsample_nearest2x — F.interpolate(scale_factor=2, mode="nearest") on channel-last fp16
 * (diffusers Upsample2D, reached from block.py:2366,2476). x [batch,h,w,c] -> out [batch,2h,2w,c]. */
int ir_upsample_nearest2x(const void* x, void* out, int batch, int h, int w, int c, ir_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * ir_latent_in — scheduler add_noise + layout change (pix2pix_turbo.py:250-251, 310-311):
 *   out[b, hw, c] = fp16( a * x[b, c, hw] + s * noise[b, c, hw] ), channels zero-padded to c_pad.
 * x, noise: fp32 NCHW [batch, c, hw] (noise may be NULL). out: fp16 [batch, hw, c_pad].
 * ir_latent_out — DDPM pred_original_sample + layout change (pix2pix_turbo.py:277,331):
 *   xt = a * x + s * noise;  out[b, c, hw] = (xt[b, c, hw] - s * eps[b, hw, c]) / a     (fp32 NCHW)
 * eps: fp16 [batch, hw, c] channel-last (row stride eps_row_stride); x, noise: fp32 NCHW (noise may be NULL).
 */
int ir_latent_in(const float* x, const float* noise, float a, float s, void* out, int batch, int c, int hw,
                 int c_pad, ir_stream_t stream);
int ir_latent_out(const void* eps, int eps_row_stride, const float* x, const float* noise, float a, float s, float* out,
                  int batch, int c, int hw, ir_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * VAE side of the pipeline (reference pix2pix_turbo.py:245,291,333; models/model.py:15-63).
 * ir_softmax_rows — in-place row softmax of an fp16 [rows, cols] score matrix, fp32 math: p = softmax(x * scale).
 *   (the VAE mid-block attention is single-head with head_dim 512: scores are materialised by ir_conv_gemm exactly as
 *   the reference's baddbmm does, diffusers Attention.get_attention_scores.)
 * ir_image_in  — fp16 or fp32 NCHW image [batch, c, hw] -> fp16 channel-last [batch, hw, c_pad] (zero-padded).
 * ir_image_in_patches3x3 — the same image as 3x3 patches (zero padding 1) in ONE 64-wide K block per pixel:
 *   out[b, y, x, (ky*3+kx)*c + ch] = image[b, ch, y+ky-1, x+kx-1], zero for k >= 9c (needs 9c <= 64). The 3 -> 128 channel
 *   conv_in of the VAE encoder (diffusers Encoder.conv_in as patched by reference models/model.py:15-30) then is ir_conv_gemm
 *   with ksize 1, c_in 64 and the weight laid out [c_out, (ky, kx, ch) padded to 64]: the 27 products per output
 *   instead of 576 (a 3x3 convolution over 64 zero-padded channels).
 * ir_image_out — fp16 channel-last [batch, hw, >= c] (row stride y_row_stride) -> NCHW clamp(lo, hi), fp16 or fp32.
 * ir_vae_sample — DiagonalGaussianDistribution.sample() * scaling_factor with the normal draw injected:
 *   moments fp16 channel-last [batch, hw, >= 2c] (mean | logvar); out[b,c,hw] = (mean + exp(0.5*clamp(logvar,-30,20))
 *   * eps[b,c,hw]) * scale, fp32 NCHW; eps may be NULL (mode).
 */
int ir_softmax_rows(void* x, int rows, int cols, int row_stride, float scale, ir_stream_t stream);
int ir_image_in(const void* x, int x_is_fp32, void* out, int batch, int c, int hw, int c_pad, ir_stream_t stream);
int ir_image_in_patches3x3(const void* x, int x_is_fp32, void* out, int batch, int c, int h, int w, ir_stream_t stream);
int ir_image_out(const void* y, int y_row_stride, float lo, float hi, void* out, int out_is_fp32, int batch, int c,
                 int hw, ir_stream_t stream);
int ir_vae_sample(const void* moments, int m_row_stride, const float* eps, float scale, float* out, int batch, int c,
                  int hw, ir_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Host pre / post-processing of the reference entry on the GPU (byte / integer work, bit-exact with the reference):
 * ir_resample_u8_pass — one separable pass of Pillow's 8-bit resampling (Image.resize(..., LANCZOS), which
 *   torchvision's Resize(512, LANCZOS) of face_replace/inference/test.py:54-56 calls), restricted to the CenterCrop(512)
 *   window (:57):  out[o, j, c] = clip8((2^21 + sum_t kk[first + o, t] * in[bounds[first + o].lo + t, j, c]) >> 22), c = 0..2.
 *   in: uint8, element (a, j, c) at in + a * in_stride_axis + j * in_stride_other + c (bytes); bounds: int32 [n, 2] =
 *   (first tap, tap count), kk: int32 [n, ksize] fixed-point weights, both computed on the host as Pillow's
 *   precompute_coeffs / normalize_coeffs_8bpc do. out_f16_norm == 0: uint8 out (the 8-bit intermediate image of the
 *   first pass); != 0: ToTensor + Normalize(0.5, 0.5) (:58-59) and the fp16 cast of :92 applied, fp16 out. Strides of
 *   `out` are in elements: (o, j, c) at o * out_stride_axis + j * out_stride_other + c * out_stride_c.
 * ir_u8_to_f16 — crop + normalise when no resize is needed: uint8 window [h, w, 3] (byte strides) -> fp16 NCHW [3, h, w].
 * ir_image_out_u8 — vis_utils.tensor2im(unnorm=True) (face_replace/training/utils/vis_utils.py:14-23) on the fp16 NCHW
 *   prediction [batch, 3, hw]: * 0.5, + 0.5 (each rounded to fp16), clamp [0, 1], * 255 (fp16), truncate -> uint8 [batch, hw, 3].
 */
int ir_resample_u8_pass(const void* in, long in_stride_axis, long in_stride_other, int n_out, int n_other, const int* bounds,
                        const int* kk, int ksize, int first, void* out, long out_stride_axis, long out_stride_other,
                        long out_stride_c, int out_f16_norm, ir_stream_t stream);
int ir_u8_to_f16(const void* in, long in_stride_y, long in_stride_x, int h, int w, void* out, ir_stream_t stream);
int ir_image_out_u8(const void* pred, void* out, int batch, int hw, ir_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* INSTANTRESTORE_B200_H_ */
