/* wiw.h -- C ABI of libwiw_b200.so, the sm_100a kernels behind the World-in-World SVD-dagger
 * world-model worker.
 *
 * Boundary (last row): the reference has no native code; what its worker
 * calls inside task_fn is Python -> torch library kernels.  This ABI is what a Python FFI
 * (ctypes) binds instead: raw device pointers + shapes + a cudaStream_t, int status
 * (0 = ok) and wiw_last_error().  No torch types cross it.  The caller (Python) owns every
 * buffer; the library owns nothing but a cached SM count and the TMA-encode entry point.
 *
 * Each entry point names the reference call site it replaces (paths relative to
 * /root/reference; DP = FTsvd/diffusers-private/diffusers).
 *
 * Layouts: activations are fp16, channels-last: a "[n][y][x][c]" tensor for 2-D ops and
 * "[b][f][s][c]" for temporal ops (s = y*W + x), i.e. the reference's NCHW/"(b f) c h w"
 * tensors permuted once at the model boundary.  Weights are fp16, output-major with the
 * reduction axis contiguous ("K-major"): conv filters [cout][ky][kx][cin].  Bias / lookup
 * tables / norm parameters are fp32.
 */
#ifndef WIW_H
#define WIW_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* wiw_stream_t; /* cudaStream_t */

/* Last error message of the calling process (empty string if none). */
const char* wiw_last_error(void);
/* ABI version; bumped on any signature change. */
int wiw_abi_version(void);
/* 1 if the current device is sm_100 (B200) and the kernels can run. */
int wiw_device_ok(void);

/* ------------------------------------------------------------------------------------------
 * Implicit-GEMM contraction on tcgen05 tensor cores (TMA-fed, TMEM accumulators).
 *
 *   out[n, y, x, :] = s_acc * ( sum_taps sum_c A[c + cbase_t, x + ox_t, p_t, y + oy_t, n] * W[:, t*Cin + c]
 *                               + bias[:] + tab[((n*Y + y)*X + x) / tab_div % tab_mod, :] )      (act applied here)
 *                     + s_res * res[n, y, x, :] + s_aux * aux[n, y, x, :]
 *
 * A is a 5-D strided view (c, x, p, y, n) of an fp16 tensor (c contiguous); out-of-range
 * x / y / c coordinates read as zero (that is the conv zero padding).  One call covers:
 *   nn.Linear            (attention_processor.py:2358-2391, attention.py FeedForward, proj_in/out
 *                         transformer_temporal.py:238,275): 1 tap, X = rows, Y = N = 1
 *   nn.Conv2d 3x3 s1 p1  (resnet.py:269,285; unet conv_in/conv_out): 9 taps
 *   nn.Conv2d 3x3 s2 p1  (downsampling.py:69-146): 9 taps over a space-to-depth view (p = row parity)
 *   nn.Conv3d (3,1,1)    (resnet.py:570,586; time_conv_out): 3 taps along y = frame
 * geglu != 0: W rows are interleaved per N-tile as [bn/2 value rows | bn/2 gate rows]; the
 *   epilogue writes value * gelu_erf(gate) (activations.py:107-123), out has n_total/2 channels.
 * Returns 0 on success.
 */
typedef struct {
    const void* a;          /* fp16 */
    int64_t a_dim[5];       /* extents of (c, x, p, y, n) */
    int64_t a_stride[5];    /* element strides; a_stride[0] must be 1 */
    const void* w;          /* fp16 [n_total][k_total] */
    int64_t k_total;        /* = num_taps * chunks_per_tap * 64 */
    int64_t n_total;        /* GEMM N (multiple of bn) */
    void* out;              /* fp16 (c, x, y, n) view */
    int64_t out_dim[4];
    int64_t out_stride[4];
    const void* res;        /* optional fp16, same extents as out */
    int64_t res_stride[4];
    const void* aux;        /* optional fp16, same extents as out */
    int64_t aux_stride[4];
    const float* bias;      /* optional fp32 [n_total] */
    const void* tab;        /* optional fp16 [tab_rows][tab_stride] */
    int32_t tab_div;
    int32_t tab_mod;
    float s_acc, s_res, s_aux;
    int32_t bx, by;         /* M-tile = bz images x by rows x bx pixels, bx*by*bz == 128 (bz: last field) */
    int32_t bn;             /* N-tile, multiple of 32 (64 if geglu), <= 256 */
    int32_t geglu;
    int32_t num_taps;       /* <= 9 */
    int32_t chunks_per_tap; /* Cin / 64 */
    int32_t tap_cbase[9];
    int32_t tap_ox[9];
    int32_t tap_p[9];
    int32_t tap_oy[9];
    int32_t tab_stride;     /* row stride of tab in elements (0 = n_total) */
    int32_t act;            /* applied to (acc + bias + tab): 0 none, 1 SiLU, 2 GELU (erf), 3 quick GELU x*sigmoid(1.702x) */
    int32_t bz;             /* images per M-tile (0 or 1 = one image): lets 9-row / 18-row feature maps fill their tiles */
} wiw_gemm_desc;

int wiw_conv_gemm(const wiw_gemm_desc* d, wiw_stream_t stream);


/* ------------------------------------------------------------------------------------------
 * Normalisation (HBM-bound, fp32 statistics).
 * wiw_groupnorm: nn.GroupNorm(groups) [+ SiLU] over x[samples][rows_per_sample][C] channels-last
 *   (resnet.py:267,283 spatial: sample = one frame; resnet.py:569,583 temporal: sample = one batch
 *   element, rows = F*h*w; transformer_temporal.py:237; unet conv_norm_out :252).
 *   workspace: >= wiw_groupnorm_workspace_floats(samples, groups) floats.
 * wiw_layernorm: nn.LayerNorm(C) over rows of x[M][C] (attention.py:371-432,659-694); optional
 *   pre-add of tab[(row / tab_div) % tab_mod][C] in fp16 (the frame-position embedding,
 *   transformer_temporal.py:360).
 */
int wiw_groupnorm_workspace_floats(int samples, int groups);
int wiw_groupnorm(const void* x, void* y, int samples, long long rows_per_sample, int C, int G,
                  const float* gamma, const float* beta, float eps, int silu, float* workspace, wiw_stream_t stream);
int wiw_layernorm(const void* x, void* y, long long M, int C, const float* gamma, const float* beta, float eps,
                  const void* tab /* fp16 [tab_mod][C] */, int tab_div, int tab_mod, wiw_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * EDM / Euler sampler glue (bit-exact fp32 arithmetic, no FMA contraction).
 * wiw_sampler_prepare: pipeline :589-593 + scheduler.scale_model_input :313-314.  latents fp16
 *   [F][hw][4], image_latents fp16 [hw][4] -> unet_in fp16 [2][F][hw][64] channels 0..7
 *   (row 0 = unconditional: image channels zero).  div = sqrt(sigma_i^2 + 1) computed by the host in fp32.
 * wiw_sampler_step: CFG combine (pipeline :606-608, per-frame scale gs fp16 [F]) + scheduler.step
 *   :635-673 (v-prediction Euler); v fp16 [2][F][hw][vstride] (channels 0..3), latents updated in place.
 *   c_out = -sigma/sqrt(sigma^2+1), sig2p1 = sigma^2+1, dt = sigma_next - sigma, all fp32 from the host.
 */
int wiw_sampler_prepare(const void* latents, const void* image_latents, void* unet_in, int F, int hw, float div,
                        wiw_stream_t stream);
int wiw_sampler_step(const void* v, void* latents, const void* gs, int F, int hw, int vstride, float c_out,
                     float sig2p1, float sigma, float dt, wiw_stream_t stream);

/* emb[r] = silu(fp16(fp16(te + act[r]) + noise[r])) -- unet :482-487 + resnet.py:345; fp16 [R][C]. */
int wiw_emb_silu(const void* te, const void* act, const void* noise, void* out, int R, int C, wiw_stream_t stream);
/* torch.cat([a, b], dim=channel) on channels-last rows (unet_3d_blocks.py:1489,1612). */
int wiw_concat_channels(const void* a, const void* b, void* out, long long M, int C1, int C2, wiw_stream_t stream);
/* F.interpolate(scale_factor=2, mode="nearest") (upsampling.py:169). */
int wiw_upsample_nearest2x(const void* x, void* y, int N, int H, int W, int C, wiw_stream_t stream);
/* Temporal self-attention over F frames per (batch element, pixel, head), head_dim 64
 * (attention.py:736): qkv fp16 [B][F][S][3C] -> out [B][F][S][C]. */
int wiw_attention_temporal(const void* qkv, void* out, int B, int F, int S, int C, wiw_stream_t stream);
/* Spatial self-attention softmax(q k^T / 8) v per (image, head), head_dim 64, tcgen05 flash kernel
 * (F.scaled_dot_product_attention at attention_processor.py:2383 <- attention.py:507):
 * qkv fp16 [images][S][3C] (q | k | v, head h at columns h*64), out fp16 [images][S][C]. */
int wiw_attention_spatial(const void* qkv, void* out, int images, int S, int C, wiw_stream_t stream);
/* In-place softmax(scale * row) on fp16 [rows][S] (VAE mid-block attention, unet_3d_blocks.py:958-969). */
int wiw_softmax_rows(void* s, long long rows, int S, float scale, wiw_stream_t stream);
/* Batched transpose fp16 [batch][R][C] -> [batch][C][R]. */
int wiw_transpose(const void* x, void* y, int batch, int R, int C, wiw_stream_t stream);
/* (x/2+0.5).clamp(0,1) -> round(*255) uint8, channels-last fp16 -> [F][3][HW] (image_processor.py:239,147). */
int wiw_frames_to_u8(const void* x, void* out, long long F, long long HW, int cs, wiw_stream_t stream);
/* TemporalDecoder tail: time_conv_out Conv3d(3,3,(3,1,1)) + quantisation
 * (autoencoder_kl_temporal_decoder.py:154-160; pipeline :307; image_processor.py:239,147).
 * x fp16 [F][HW][cs]; w fp32 [3][3][3] (co,ci,kt); out_u8 [F][3][HW] and/or out_f32 [F][3][HW] (pre-quantisation). */
int wiw_vae_finish(const void* x, const float* w, const float* bias, void* out_u8, float* out_f32, int F,
                   long long HW, int cs, wiw_stream_t stream);
/* ------------------------------------------------------------------------------------------
 * Conditioning front-end (once per request): CLIP ViT image encoder of StableVideoDiffusionPipeline._encode_image
 * (DP/pipelines/stable_video_diffusion/pipeline_stable_video_diffusion.py:183-229; the encoder itself is
 * transformers.CLIPVisionModelWithProjection).  The Linear / LayerNorm layers go through wiw_conv_gemm / wiw_layernorm.
 */
/* _resize_with_antialiasing (pipeline :643-747: Gaussian blur, then bicubic with align_corners=True) as one separable linear
 * operator per axis + the CLIP normalisation (x - mean[c]) / std[c]: uint8 planes [planes][H][W] (read as value / 255) ->
 * fp16 [planes][oh][ow].  bounds_* int32 [out][2] = (first source index, taps), w_* fp32 [out][ksize_*] (host-built,
 * wiw_b200/clip.py); tmp fp32 [planes][H][ow]; channel of plane p = p % channels. */
int wiw_resize_separable_f32(const void* src, void* tmp, void* dst, long long planes, int channels, int H, int W, int oh, int ow,
                             const int* bounds_h, const float* w_h, int ksize_h, const int* bounds_v, const float* w_v, int ksize_v,
                             const float* mean, const float* inv_std, wiw_stream_t stream);
/* im2col rows of a stride == kernel patch convolution (CLIPVisionEmbeddings.patch_embedding): x fp16 [B][C][G*ps][G*ps] ->
 * fp16 [B*G*G][kpad], column (c*ps + ky)*ps + kx, zero padded to kpad (multiple of 64). */
int wiw_patchify(const void* x, void* out, int B, int C, int G, int ps, int kpad, wiw_stream_t stream);
/* Multi-head self-attention for short sequences, any even head_dim <= 128 (CLIPAttention: 257 tokens, 16 heads of 80):
 * qkv fp16 [batch*S][3*heads*d] (q | k | v, head-major) -> out fp16 [batch*S][heads*d]; softmax(q k^T * scale) v in fp32. */
int wiw_attention_small(const void* qkv, void* out, int batch, int S, int heads, int d, float scale, wiw_stream_t stream);

/* Pillow-exact BICUBIC resize of uint8 planes [planes][H][W] -> [planes][oh][ow] (22-bit fixed-point two-pass resampler,
 * 8-bit intermediate in `tmp` [planes][H][ow]) followed by a 256-entry requantisation LUT: replaces the CPU
 * `img.resize(save_size, Image.BICUBIC)` + to_tensor + (clip*255).astype(uint8) of downstream/api_models/__init__.py:159-166,229.
 * bounds_* int32 [out][2] = (first source index, taps), kk_* int32 [out][ksize] built by wiw_b200/resize.py. */
int wiw_resize_bicubic_u8(const void* src, void* tmp, void* dst, long long planes, int H, int W, int oh, int ow,
                          const int* bounds_h, const int* kk_h, int ksize_h, const int* bounds_v, const int* kk_v, int ksize_v,
                          const void* lut, wiw_stream_t stream);
/* Boundary layout converters: fp16 [N][C][P] <-> channels-last [N][P][cs] (cs >= C, padding zero). */
int wiw_nchw_to_nhwc(const void* src, void* dst, long long N, int C, long long P, int cs, wiw_stream_t stream);
int wiw_nhwc_to_nchw(const void* src, void* dst, long long N, int C, long long P, int cs, wiw_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* WIW_H */
