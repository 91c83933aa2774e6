/* aether_b200 -- C ABI of the B200-native Aether denoising hot path.
 *
 * The reference (InternRobotics/Aether) has no FFI: its seam is constructor injection of three module
 * objects into a diffusers pipeline (aether/pipelines/aetherv1_pipeline_cogvideox.py:274-288).  Each entry
 * point below replaces the device work behind one of those objects' methods and is what a binding for that
 * method calls (INTEGRATION.md shows the ctypes stubs):
 *
 *   aether_dit_forward            <- transformer(hidden_states, encoder_hidden_states, timestep, ofs,
 *                                    image_rotary_emb, ...)[0]            pipeline :865-875
 *   aether_cfg_dpm_step           <- CFG combine :895-899 + scheduler.step(...) :907-915 + bf16 cast :916
 *   aether_vae_*                  <- vae.encode(x).latent_dist / vae.decode(z).sample   :557-620, :931, :936
 *   aether_scale_reduce / aether_blend_crossfade
 *                                 <- compute_scale (aether/utils/postprocess_utils.py:847-864) and the
 *                                    linear cross-fade of evaluation/video_depth/launch_aether.py:166-285
 *   aether_gemm_bf16, aether_attention_bf16, aether_ln_modulate, ...   the per-kernel entry points the
 *                                    above are composed of; exported so each kernel can be parity-tested and
 *                                    profiled in isolation.
 *
 * Conventions: every function returns an int status (AETHER_OK = 0) and never throws; all data pointers are
 * DEVICE pointers owned by the caller; shapes are explicit; `stream` is a cudaStream_t passed as void*;
 * no function allocates device memory or synchronises the device; handles are re-entrant per handle and
 * one handle belongs to one GPU.  bf16 tensors are `uint16_t`-sized elements (no torch types anywhere).
 */
#ifndef AETHER_B200_H_
#define AETHER_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AETHER_OK 0
#define AETHER_ERR_INVALID 1   /* bad argument (shape, alignment, null pointer) */
#define AETHER_ERR_CUDA 2      /* CUDA runtime / driver error (message on stderr) */
#define AETHER_ERR_WORKSPACE 3 /* workspace too small */

#define AETHER_ABI_VERSION 2
int32_t aether_abi_version(void);
/* 1 when a CUDA device of compute capability 10.x is present, else 0 (never throws). */
int32_t aether_device_ok(void);

/* ---------------------------------------------------------------- dense contractions (tcgen05) */

/* C[M,N] = epi(A[M,K] . W[N,K]^T), all bf16 row-major, fp32 accumulate in TMEM.
 *   epilogue 0: + bias        1: gelu_tanh(+ bias)       2: C = C + gate[b, n] * (acc + bias)   (in place)
 * For epilogue 2 row r belongs to batch b = r / S, token s = r % S; tokens s < St use gate_txt, others
 * gate_vid (both fp32 [B, N] with batch stride gate_bstride elements).
 * Output columns >= f16_from_col (a multiple of 8; < 0 = none) are written as fp16 instead of bf16 (used for
 * the V third of the fused QKV projection when the fp16-PV attention mode is on).
 * M >= 1024 runs on CTA pairs (tcgen05 cta_group::2, cluster of two CTAs per 256 x 256 tile), smaller M on single
 * CTAs; results are identical.
 * Replaces nn.Linear inside CogVideoXBlock / CogVideoXPatchEmbed / proj_out (pipeline :865). */
int aether_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int32_t M,
                     int32_t N, int32_t K, const float* bias, int32_t epilogue, const float* gate_vid,
                     const float* gate_txt, int64_t gate_bstride, int32_t S, int32_t St, int32_t f16_from_col,
                     void* stream);

/* Fused QKV projection: qkv[rows, 3*H*64] = A[rows,K] . W[3*H*64,K]^T + bias, followed INSIDE the GEMM epilogue by
 * QK-LayerNorm(64) of every q and k head (gamma/beta fp32 [64]) and the rotary embedding of the video tokens
 * (token s = row % S >= St; rope_cos / rope_sin fp32 [S - St, 64], both NULL = no rotary embedding).  Same values as
 * aether_gemm_bf16 followed by aether_qk_norm_rope.  Replaces to_q/to_k/to_v + norm_q/norm_k + apply_rotary_emb of
 * CogVideoXAttnProcessor2_0 (pipeline :865). */
int aether_gemm_qkv_norm_rope_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* qkv, int32_t rows,
                                   int32_t K, const float* bias, int32_t S, int32_t St, int32_t H, const float* qn_g,
                                   const float* qn_b, const float* kn_g, const float* kn_b, float eps,
                                   const float* rope_cos, const float* rope_sin, int32_t f16_from_col, void* stream);

/* out[B,S,H*64] (bf16) = softmax(Q K^T * softmax_scale) V over qkv[B,S,3,H,64], non-causal, head_dim 64.
 * `v_fp16` is the kernel-variant id (all variants compute the same function to the tolerance of the tests):
 *   5 decoupled S / P TMEM buffers, skewed MMA schedule (product default)   0 baseline (P aliases S)
 *   2 fp16 P/V with 40 % of the exponentials as an FMA-pipe polynomial (the V third of qkv must then hold fp16 bit
 *   patterns).  Any other id returns AETHER_ERR_INVALID.  (Twelve further variants explored in round 1 -- none faster
 *   inside the power-capped step -- are kept, unbuilt, under tools/experiments/attention/.)
 * Replaces F.scaled_dot_product_attention in CogVideoXAttnProcessor2_0 (pipeline :865). */
int aether_attention_bf16(const void* qkv, void* out, int32_t B, int32_t S, int32_t H, float softmax_scale,
                          int32_t v_fp16, void* stream);
/* Same with caller-provided scratch (aether_attention_workspace_bytes; 0 = none needed): variant 5 then cuts the LAST,
 * partially filled wave of its grid -- (query blocks x heads x batch) % #SM work items, 20 of 148 SMs busy for a whole
 * item time at S = 15076, H = 48 -- along the keys so that it fills every SM, and merges the partial (O, max, sum)
 * triples with a small kernel.  Results agree with the unsplit call to fp32 merge round-off. */
int64_t aether_attention_workspace_bytes(int32_t B, int32_t S, int32_t H, int32_t v_fp16);
int aether_attention_bf16_ws(const void* qkv, void* out, int32_t B, int32_t S, int32_t H, float softmax_scale,
                             int32_t v_fp16, void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- bandwidth-bound DiT kernels */

/* y[r,:] = LN(x[r,:]; gamma, beta, eps) * (1 + scale[b,:]) + shift[b,:]   (CogVideoXLayerNormZero / AdaLayerNorm)
 * x, y bf16 [B*S, D]; gamma/beta fp32 [D] (nullable = no affine); shift/scale fp32 with batch stride
 * mod_bstride; text rows (s < St) use the *_txt vectors.  If gamma2 != NULL a second LayerNorm(gamma2, beta2)
 * is applied to the result of the first before modulation (norm_final followed by norm_out.norm). */
int aether_ln_modulate(const void* x, void* y, int32_t B, int32_t S, int32_t St, int32_t D, const float* gamma,
                       const float* beta, float eps, const float* gamma2, const float* beta2,
                       const float* shift_vid, const float* scale_vid, const float* shift_txt,
                       const float* scale_txt, int64_t mod_bstride, void* stream);

/* In place on qkv[B,S,3,H,64]: q,k <- LayerNorm_64(q,k; eps) (affine gq/bq, gk/bk fp32[64]); then for tokens
 * s >= St the interleaved-pair RoPE with cos/sin fp32 [S-St, 64].  (attn.norm_q/norm_k + apply_rotary_emb) */
int aether_qk_norm_rope(void* qkv, int32_t B, int32_t S, int32_t St, int32_t H, const float* gq, const float* bq,
                        const float* gk, const float* bk, float eps, const float* cos, const float* sin,
                        void* stream);

/* y[b,n] = sum_k act(x[b,k]) * W[n,k] + bias[n];  x fp32 [B,K] (B <= 8), W bf16 [N,K], y fp32 [B,N].
 * act: 0 identity, 1 SiLU.  Used for the timestep MLP and all AdaLN-Zero linears of a step in one launch. */
int aether_small_m_linear(const float* x, const void* W, const float* bias, float* y, int32_t B, int32_t N,
                          int32_t K, int32_t act, void* stream);

/* emb[b, :] = [cos(t_b * f_i) | sin(t_b * f_i)] (flip_sin_to_cos) or [sin | cos];  fp32 [B, dim]. */
int aether_timestep_sinusoid(const int64_t* timesteps, float* emb, int32_t B, int32_t dim, int32_t flip_sin_to_cos,
                             float freq_shift, void* stream);

/* patches[(b,f,y,x), c*p*p + dy*p + dx] = in[b,f,c,p*y+dy,p*x+dx]   (bf16; p == 2) */
int aether_patchify(const void* in, void* patches, int32_t B, int32_t F, int32_t C, int32_t H, int32_t W,
                    void* stream);
/* out[b,f,c,2y+dy,2x+dx] = tok[(b,f,y,x), c*4 + dy*2 + dx]   (bf16; tok row stride ld_tok elements) */
int aether_unpatchify(const void* tok, int64_t ld_tok, void* out, int32_t B, int32_t F, int32_t C, int32_t H,
                      int32_t W, void* stream);
/* x[b, s, :] += pos[s, :]   (bf16 [B,S,D] += bf16 [S,D]; learned / sincos positional table) */
int aether_add_pos_embed(void* x, const void* pos, int32_t B, int32_t S, int32_t D, void* stream);

/* ---------------------------------------------------------------- the DiT forward (handle API) */

typedef struct AetherDitConfig {
  int32_t num_heads, head_dim, num_layers;
  int32_t in_channels, out_channels, patch_size;
  int32_t time_embed_dim, text_embed_dim;
  int32_t flip_sin_to_cos;
  float freq_shift, norm_eps;
  int32_t ff_mult;
  int32_t attention_fp16_pv;   /* attention variant id passed to aether_attention_bf16 (5 = product default); for
                                  id 2 the QKV GEMM emits the V third as fp16 */
  int32_t fused_qkv_epilogue;  /* 1: QK-LayerNorm + RoPE run inside the QKV GEMM epilogue (one launch less per layer;
                                  measured slower, default 0).  Per handle -- the library reads no environment. */
  int32_t attention_split_tail; /* 1 (default): reserve scratch so attention variant 5 cuts its partially filled tail wave
                                  along the keys (aether_attention_bf16_ws); 0: single launch -- results then do not
                                  depend on how many (batch x head x query-block) items there are */
} AetherDitConfig;

typedef struct AetherDitLayerWeights {
  const void* w_qkv;  const float* b_qkv;    /* [3D, D] bf16 (to_q | to_k | to_v), [3D] */
  const void* w_out;  const float* b_out;    /* [D, D], [D] */
  const void* w_ff1;  const float* b_ff1;    /* [ff_mult*D, D] */
  const void* w_ff2;  const float* b_ff2;    /* [D, ff_mult*D] */
  const float* norm1_g; const float* norm1_b;  /* [D] (nullable) */
  const float* norm2_g; const float* norm2_b;
  const float* qn_g; const float* qn_b; const float* kn_g; const float* kn_b;   /* [head_dim] */
} AetherDitLayerWeights;

typedef struct AetherDitWeights {
  const void* w_time1; const float* b_time1;   /* [T, D] bf16, [T] */
  const void* w_time2; const float* b_time2;   /* [T, T] */
  const void* w_text;  const float* b_text;    /* [D, text_embed_dim] */
  const void* w_patch; const float* b_patch;   /* [D, in_channels*p*p] */
  /* all AdaLN linears concatenated row-wise: per layer [norm1.linear (6D) | norm2.linear (6D)], then
   * norm_out.linear (2D):  [(12*L + 2) * D, T] bf16 and its bias */
  const void* w_adaln; const float* b_adaln;
  const float* normf_g; const float* normf_b;  /* norm_final */
  const float* normo_g; const float* normo_b;  /* norm_out.norm */
  const void* w_proj;  const float* b_proj;    /* [p*p*out_channels, D] */
  const void* pos_embedding;                   /* optional bf16 [St + Sv, D] added after patch embed, or NULL */
  const AetherDitLayerWeights* layers;         /* [num_layers] host array */
} AetherDitWeights;

typedef struct AetherDit AetherDit;

int aether_dit_create(const AetherDitConfig* cfg, const AetherDitWeights* w, AetherDit** out);
void aether_dit_destroy(AetherDit* h);
/* attach / detach (NULL) the additive positional table bf16 [St+Sv, D] used by the next forwards
 * (CogVideoXPatchEmbed learned / sin-cos branch); the buffer stays owned by the caller. */
int aether_dit_set_pos_embedding(AetherDit* h, const void* pos_bf16);
/* Measurement hooks (bench.py roofline): when enabled every attention launch of a forward is bracketed by a
 * CUDA event pair recorded on the launching stream; after the caller synchronised that stream,
 * aether_dit_read_timing returns the summed duration and the number of launches of the LAST forward. */
int aether_dit_enable_timing(AetherDit* h, int32_t enable);
int aether_dit_read_timing(AetherDit* h, float* attention_ms_total, int32_t* attention_launches);
/* bytes of scratch the forward needs for batch B, F latent frames of HxW (latent pixels), St text tokens */
int64_t aether_dit_workspace_bytes(const AetherDit* h, int32_t B, int32_t F, int32_t H, int32_t W, int32_t St);
/* hidden [B,F,Cin,H,W] bf16; text [B,St,text_dim] bf16; timesteps int64 [B]; cos/sin fp32 [F*(H/p)*(W/p), hd]
 * (nullable = no RoPE); out [B,F,Cout,H,W] bf16.  n_layers < 0 runs all layers (>= 0: first n, for profiling). */
int aether_dit_forward(AetherDit* h, const void* hidden, const void* text, const int64_t* timesteps,
                       const float* rope_cos, const float* rope_sin, void* out, int32_t B, int32_t F, int32_t H,
                       int32_t W, int32_t St, void* workspace, int64_t workspace_bytes, int32_t n_layers,
                       void* stream);

/* Same forward with the loop's concatenations folded in (pipeline :832-869): the model input is the channel concat
 * of latents[latents_batch, F, latents_channels, H, W] (latents_batch == 1 broadcasts over the CFG batch =
 * torch.cat([latents] * 2)) and cond[B, F, in_channels - latents_channels, H, W]; text [text_batch, St, dim] and
 * timesteps [timesteps_batch] broadcast when their batch is 1 (prompt_embeds.repeat, t.expand). */
int aether_dit_forward_split(AetherDit* h, const void* latents, int32_t latents_batch, int32_t latents_channels,
                             const void* cond, const void* text, int32_t text_batch, const int64_t* timesteps,
                             int32_t timesteps_batch, const float* rope_cos, const float* rope_sin, void* out,
                             int32_t B, int32_t F, int32_t H, int32_t W, int32_t St, void* workspace,
                             int64_t workspace_bytes, int32_t n_layers, void* stream);

/* ---------------------------------------------------------------- 3-D causal VAE (K9)
 * Device work behind AutoencoderKLCogVideoX.encode / .decode (pipeline :557-620, :931, :936).  Activations are
 * channels-last x[T, H, W, C] bf16 (one batch item; the reference enables slicing).  The tiling / frame-batching
 * / conv-cache control flow lives in aether_b200/vae.py and mirrors diffusers. */

/* y[T_out,H_out,W_out,Cout] = conv(x[T_in,H_in,W_in,Cin], w) + bias (+ resid), implicit GEMM on tcgen05.
 * x is time-padded by the caller (T_in = T_out + kt - 1); spatial zero padding of pad_h/pad_w leading rows/cols
 * (trailing padding is implicit); stride 1 or 2 spatially.  w_packed: bf16 [Cout, kt*kh*kw*ceil64(Cin)], tap-major
 * then channel, zero-padded channels.  Cin % 8 == 0, Cout % 8 == 0. */
int aether_conv3d_bf16(const void* x, int32_t T_in, int32_t H_in, int32_t W_in, int32_t Cin, const void* w_packed,
                       const float* bias, const void* resid, void* y, int32_t T_out, int32_t H_out, int32_t W_out,
                       int32_t Cout, int32_t kt, int32_t kh, int32_t kw, int32_t stride, int32_t pad_h, int32_t pad_w,
                       void* stream);
/* Same operation, restricted to the one-CTA kernel.  aether_conv3d_bf16 runs Cout >= 128 on CTA pairs (tcgen05
 * cta_group::2: two neighbouring output tiles share every weight tile); this entry is the pair kernel's bit-exact
 * reference in the parity tests and the baseline of its A/B timing. */
int aether_conv3d_bf16_1cta(const void* x, int32_t T_in, int32_t H_in, int32_t W_in, int32_t Cin, const void* w_packed,
                            const float* bias, const void* resid, void* y, int32_t T_out, int32_t H_out, int32_t W_out,
                            int32_t Cout, int32_t kt, int32_t kh, int32_t kw, int32_t stride, int32_t pad_h,
                            int32_t pad_w, void* stream);
/* GroupNorm statistics of x[N, C] (N = T*H*W): mean_rstd[2*G] = {mean_g, 1/sqrt(var_g + eps)}; deterministic.
 * workspace: aether_gn_workspace_floats(C) floats. */
int64_t aether_gn_workspace_floats(int32_t C);
int aether_gn_stats(const void* x, int64_t N, int32_t C, int32_t G, float eps, float* workspace, float* mean_rstd,
                    void* stream);
/* y = SiLU?( GN(x) [* zy[map] + zb[map]] ): GroupNorm or CogVideoXSpatialNorm3D (zy/zb = conv_y/conv_b of the latent
 * at latent resolution [Tz, hz, wz, .] with row stride zld elements (0 = C); tmap[t] = latent frame of frame t;
 * rows/cols map by floor(y*hz/H)). */
int aether_gn_apply(const void* x, void* y, int64_t N, int32_t C, int32_t G, const float* mean_rstd,
                    const float* gamma, const float* beta, const void* zy, const void* zb, int32_t zld,
                    const int32_t* tmap, int32_t H, int32_t W, int32_t hz, int32_t wz, int32_t silu, void* stream);
/* out[t', y', x', :] = in[tmap[t'], y'/sy, x'/sx, :]   (F.interpolate nearest; first-frame rule encoded in tmap) */
int aether_upsample_nearest(const void* in, void* out, const int32_t* tmap, int32_t To, int32_t Ho, int32_t Wo,
                            int32_t Hi, int32_t Wi, int32_t sy, int32_t sx, int32_t C, void* stream);
/* out[t'] = in[ia[t']] (ib[t'] < 0) or the mean of frames ia[t'], ib[t']   (temporal avg_pool1d keeping frame 0) */
int aether_avgpool_time(const void* in, void* out, const int32_t* ia, const int32_t* ib, int32_t To,
                        int64_t frame_elems, void* stream);
/* layout conversions between the pipeline's NCTHW tensors and channels-last (Cp = padded channel count) */
int aether_ncthw_to_thwc(const void* in, void* out, int32_t C, int32_t Cp, int64_t thw, void* stream);
int aether_thwc_to_ncthw(const void* in, void* out, int32_t C, int32_t Cp, int64_t thw, void* stream);
/* z (NCTHW [L, P]) = mean + exp(0.5*clamp(logvar,-30,20)) * noise from channels-last moments [P, Cp]
 * (DiagonalGaussianDistribution.sample; noise NULL = mode()). */
int aether_posterior_sample(const void* moments, int32_t Cp, int32_t L, const void* noise, void* z, int64_t P,
                            void* stream);
/* blend_v (axis 1) / blend_h (axis 2) of the tiled VAE, in place on b, bf16 roundings of the torch expression */
int aether_tile_blend(const void* a, void* b, int32_t T, int32_t Ha, int32_t Wa, int32_t Hb, int32_t Wb, int32_t C,
                      int32_t axis, int32_t extent, void* stream);

/* ---------------------------------------------------------------- 3-D causal VAE, handle level
 * aether_vae_encode / aether_vae_decode run the whole AutoencoderKLCogVideoX.encode / .decode schedule of ONE batch item
 * (3 x 3 spatial tiling + tile blends, frame batching with conv caches, every resnet / norm / resampling stage) as one
 * call that only enqueues kernels on `stream`: the `vae_encode_tile` / `vae_decode_tile` entry points of SURVEY.md 8(b).
 * Device work behind vae.encode(x).latent_dist (pipeline :557-620) and vae.decode(z).sample (:931, :936). */
typedef struct AetherVaeConfig {
  int32_t in_channels, out_channels, latent_channels;
  int32_t num_blocks;                /* len(block_out_channels); channel counts come with the parameters */
  int32_t layers_per_block, norm_num_groups;
  float norm_eps;
  int32_t temporal_compression_ratio;
  int32_t use_tiling;
  int32_t num_latent_frames_batch_size, num_sample_frames_batch_size;   /* diffusers: 2 and 8 */
  int32_t tile_sample_min_height, tile_sample_min_width, tile_latent_min_height, tile_latent_min_width;
  /* tiling arithmetic of diffusers tiled_encode / tiled_decode, evaluated by the caller (int(...) of float expressions):
   * stride between tiles, blend extent, kept extent -- in input pixels / latent pixels as appropriate */
  int32_t enc_overlap_h, enc_overlap_w, enc_blend_h, enc_blend_w, enc_limit_h, enc_limit_w;
  int32_t dec_overlap_h, dec_overlap_w, dec_blend_h, dec_blend_w, dec_limit_h, dec_limit_w;
} AetherVaeConfig;

/* One packed parameter, addressed by its diffusers state-dict prefix (e.g. "decoder.up_blocks.0.resnets.1.conv1"):
 *   kind 0  convolution: data = bf16 [cout, kt*kh*kw*ceil64(cin)] (aether_conv3d_bf16 layout), bias fp32 [cout]
 *   kind 1  GroupNorm affine ("...norm1" for the encoder, "...norm1.norm_layer" for SpatialNorm3D): data = gamma fp32,
 *           bias = beta fp32
 *   kind 2  SpatialNorm3D conv_y | conv_b fused ("...norm1.conv_yb"): data = bf16 [2C, latent_channels], bias fp32 [2C];
 *           cin = index of its first row inside the kind-3 concatenation (0 if there is none), cout = 2C
 *   kind 3  (optional, name free) ALL kind-2 matrices row-concatenated: data = bf16 [cout, latent_channels], bias fp32
 *           [cout]; lets the decoder evaluate every SpatialNorm's 1x1x1 convs of a frame batch with one GEMM */
typedef struct AetherVaeParam {
  const char* name;
  int32_t kind;
  const void* data;
  const float* bias;
  int32_t kt, kh, kw, cin, cout;     /* kind 0 only */
} AetherVaeParam;

typedef struct AetherVae AetherVae;
int aether_vae_create(const AetherVaeConfig* cfg, const AetherVaeParam* params, int32_t n_params, AetherVae** out);
void aether_vae_destroy(AetherVae* h);
/* op: 0 = encode of x[C, T, H, W] pixels, 1 = decode of z[L, T, H, W] latents.  Scratch bytes / number of kernel launches
 * and copies / result dims {T', H', W', C'} (channels-last moments for op 0, decoded frames for op 1) of one call. */
int64_t aether_vae_workspace_bytes(const AetherVae* h, int32_t op, int32_t T, int32_t H, int32_t W);
int64_t aether_vae_launch_count(const AetherVae* h, int32_t op, int32_t T, int32_t H, int32_t W);
int aether_vae_output_shape(const AetherVae* h, int32_t op, int32_t T, int32_t H, int32_t W, int32_t* dims4);
/* x: bf16 [in_channels, T, H, W] with element strides (stride_c, stride_t, stride_h, 1) -> moments: channels-last bf16
 * [T', H/8, W/8, ceil8(2 * latent_channels)] = (mean | logvar) of the posterior (input of aether_posterior_sample). */
int aether_vae_encode(const AetherVae* h, const void* x, int64_t stride_c, int64_t stride_t, int64_t stride_h, int32_t T,
                      int32_t H, int32_t W, void* moments, void* workspace, int64_t workspace_bytes, void* stream);
/* z: bf16 [latent_channels, T, H, W] (strides as above) -> sample: contiguous bf16 [out_channels, T', 8H, 8W]. */
int aether_vae_decode(const AetherVae* h, const void* z, int64_t stride_c, int64_t stride_t, int64_t stride_h, int32_t T,
                      int32_t H, int32_t W, void* sample, void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- scheduler step (K8) */

typedef struct AetherDpmCoeffs {
  float sqrt_alpha, sqrt_one_minus_alpha;   /* v-prediction -> x0 */
  float m1, m2, m3, m4, m_noise;            /* DPM-Solver++(2M) SDE multipliers (scheduling_dpm_cogvideox.get_mult) */
  int32_t second_order;                     /* 0: first-order return (old_x0 NULL or last step) */
  int32_t prediction_type;                  /* 0 = v_prediction, 1 = epsilon */
} AetherDpmCoeffs;

/* model_out: bf16 [n_cfg, N] (n_cfg 1 or 2: uncond first, cond second) or fp32 if model_out_fp32;
 * guidance: u + g (c - u) in fp32.  sample bf16 [N]; old_x0 fp32 [N] or NULL; noise1/noise2 bf16 [N].
 * Outputs: prev_bf16 [N] (nullable), prev_f32 [N] (nullable), x0_f32 [N].  Rounding follows torch type
 * promotion in the reference (coef * bf16 tensor rounds to bf16 before the fp32 combine). */
int aether_cfg_dpm_step(const void* model_out, int32_t model_out_fp32, int32_t n_cfg, float guidance,
                        const void* sample, const float* old_x0, const void* noise1, const void* noise2,
                        const AetherDpmCoeffs* c, void* prev_bf16, float* prev_f32, float* x0_f32, int64_t N,
                        void* stream);

/* ---------------------------------------------------------------- sliding-window blend (K10) */

/* All of these work on a 3-D region [n0, n1, n2] (frames x rows x cols; strides s0, s1 in ELEMENTS, unit stride
 * on the last axis) of disparity buffers that are fp32 (a raw pipeline window) or fp64 (an already blended
 * accumulation; the reference's np.ones(float64) result buffers).
 *
 * aether_scale_reduce: work[0] = sum(f32(p)*f32(t)), work[1] = sum(f32(p)*f32(p)) as fp64 (deterministic two-stage
 *   reduction); `work` is a device buffer of aether_scale_reduce_work_bytes() bytes, 8-byte aligned, ZEROED ONCE by
 *   the caller when it is allocated (the kernel leaves its block counter at zero again).  scale = f32(work[0]) /
 *   f32(work[1]) (0 when the denominator is 0) is compute_scale (postprocess_utils.py:847-864).
 * aether_blend_crossfade: dst = acc * w + (scale*win) * (1-w), w = np.linspace(1, 0, n_axis)[index along axis]
 *   (launch_aether.py:217-250 spatial, axis 2 or 1;  :281-284 temporal, axis 0)
 * aether_scale_copy: dst = apply_scale ? scale*src : src, widened to fp64   (launch_aether.py:220-227, :277-280)
 *   Both take the scale either from the host (`scale`, scale_sums == NULL) or, without any host round trip, from
 *   the device buffer a preceding aether_scale_reduce on the same stream filled (`scale_sums` = its `work`).
 * aether_disparity_to_depth: dst(fp64) = clip(1 / src, 0, 100)   (launch_aether.py:347) */
int64_t aether_scale_reduce_work_bytes(void);
int aether_scale_reduce(const void* pred, int32_t pred_is_f64, int64_t pred_s0, int64_t pred_s1, const void* target,
                        int32_t target_is_f64, int64_t target_s0, int64_t target_s1, int64_t n0, int64_t n1,
                        int64_t n2, void* work, void* stream);
int aether_blend_crossfade(double* dst, int64_t dst_s0, int64_t dst_s1, const void* acc, int32_t acc_is_f64,
                           int64_t acc_s0, int64_t acc_s1, const void* win, int32_t win_is_f64, int64_t win_s0,
                           int64_t win_s1, double scale, const double* scale_sums, int64_t n0, int64_t n1, int64_t n2,
                           int32_t axis, void* stream);
int aether_scale_copy(double* dst, int64_t dst_s0, int64_t dst_s1, const void* src, int32_t src_is_f64, int64_t src_s0,
                      int64_t src_s1, double scale, const double* scale_sums, int32_t apply_scale, int64_t n0,
                      int64_t n1, int64_t n2, void* stream);
/* One link of the blend chain in one call (SURVEY 8(b) `blend_tiles`): append window `win` (extents e0 x e1 x e2, fp32 or
 * fp64) to the fp64 accumulation `buf` along `axis`; the window starts at index `start`, the accumulation so far ends at
 * `prev_end` (overlap = prev_end - start >= 1).  Enqueues scale_reduce + blend_crossfade (in place on the overlap) +
 * scale_copy (the new part); `work` as for aether_scale_reduce. */
int aether_blend_link(double* buf, int64_t buf_s0, int64_t buf_s1, const void* win, int32_t win_is_f64, int64_t win_s0,
                      int64_t win_s1, int64_t e0, int64_t e1, int64_t e2, int32_t axis, int64_t start, int64_t prev_end,
                      void* work, void* stream);
int aether_disparity_to_depth(double* dst, int64_t dst_s0, int64_t dst_s1, const void* src, int32_t src_is_f64,
                              int64_t src_s0, int64_t src_s1, int64_t n0, int64_t n1, int64_t n2, void* stream);

/* ---------------------------------------------------------------- input side on the device
 * aether_resize_bilinear_u8: dst[T, H, W, 3] = cv2.resize(src[t] (h x w x 3 uint8), (W, H)) with OpenCV's default
 *   INTER_LINEAR, bit for bit (prepare_input, evaluation/video_depth/launch_aether.py:388-403, before its / 255.0).
 * aether_u8_frames_to_model_input: dst bf16 [F, 3, H, W] = 2 * (src / 255) - 1 of a uint8 crop src[F, H, W, 3] (element
 *   strides stride_t, stride_h; pixels contiguous): `/ 255.0`, the centre crop at the target size, the layout change,
 *   the [-1, 1] map and the bf16 cast of pipeline :451-512 in one pass. */
/* points[T, H, W, 3] (fp64) = camera-to-world un-projection of depth = 1 / clip(disparity, 1e-8, 1e8) (disparity fp32 or
 * fp64 [T, H, W]); cam[t] = {1/f, -cx/f, 1/f, -cy/f, pose rows 0..2 (3 x 4, row-major)} as 16 doubles.  `project` of
 * aether/utils/postprocess_utils.py:381-403 for every frame of scripts/demo.py:404-420. */
int aether_project_points(const void* disparity, int32_t disparity_is_f64, const double* cam, double* points, int32_t T,
                          int32_t H, int32_t W, void* stream);
int aether_resize_bilinear_u8(const void* src, void* dst, int32_t T, int32_t h, int32_t w, int32_t H, int32_t W,
                              void* stream);
int aether_u8_frames_to_model_input(const void* src, int64_t stride_t, int64_t stride_h, void* dst, int32_t F, int32_t H,
                                    int32_t W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AETHER_B200_H_ */
