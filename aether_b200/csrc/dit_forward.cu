// aether_dit_forward: the whole CogVideoXTransformer3DModel.forward (rotary / "5B" branch) as a fixed
// sequence of the kernels in this directory, on one stream, no host synchronisation, no allocation.
// Replaces the call at aether/pipelines/aetherv1_pipeline_cogvideox.py:865-875.
//
// Per forward (L layers):  1 sinusoid + 2 timestep GEMV + 1 AdaLN GEMV (all 2L+1 modulation vectors of the
// step at once) + text_proj GEMM + patchify + patch GEMM, then per layer
//   ln_modulate -> QKV GEMM -> qk_norm_rope -> attention -> to_out GEMM (gate*x + residual, in place)
//   ln_modulate -> FF1 GEMM (GELU-tanh) -> FF2 GEMM (gate*x + residual, in place)
// and the tail  double-LayerNorm+modulate -> proj_out GEMM -> unpatchify.
// The residual stream is one bf16 [B, St+Sv, D] buffer with the text tokens first, so diffusers'
// torch.cat([encoder_hidden_states, hidden_states]) is free.
#include <new>

#include "host_util.h"

namespace aether {
int gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N, int K,
              const float* bias, int epilogue, const float* gate_vid, const float* gate_txt, int64_t gate_bstride,
              int S, int St, int f16_from_col, cudaStream_t stream);
int gemm_qkv_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* qkv, int rows, int K, const float* bias,
                  int S, int St, int H, const float* gq, const float* bq, const float* gk, const float* bk, float eps,
                  const float* cosb, const float* sinb, int f16_from_col, cudaStream_t stream);
int attention_bf16_ws(const void* qkv, void* out, int B, int S, int H, float softmax_scale, int v_fp16, void* workspace,
                      int64_t workspace_bytes, cudaStream_t stream);
int64_t attention_v3_workspace_bytes(int B, int S, int H);
int ln_modulate(const void* x, void* y, int B, int S, int St, int D, const float* gamma, const float* beta, float eps,
                const float* gamma2, const float* beta2, const float* shift_vid, const float* scale_vid,
                const float* shift_txt, const float* scale_txt, int64_t mod_bstride, cudaStream_t stream);
int qk_norm_rope(void* qkv, int B, int S, int St, int H, const float* gq, const float* bq, const float* gk,
                 const float* bk, float eps, const float* cosb, const float* sinb, cudaStream_t stream);
int small_m_linear(const float* x, const void* W, const float* bias, float* y, int B, int N, int K, int act,
                   cudaStream_t stream);
int timestep_sinusoid(const int64_t* t, int t_stride, float* emb, int B, int dim, int flip, float shift,
                      cudaStream_t stream);
int patchify2(const void* in0, int C0, int B0, const void* in1, int C1, void* patches, int B, int F, int H, int W,
              cudaStream_t stream);
int unpatchify(const void* tok, int64_t ld_tok, void* out, int B, int F, int C, int H, int W, cudaStream_t stream);
int add_pos_embed(void* x, const void* pos, int B, int S, int D, cudaStream_t stream);
}  // namespace aether

struct AetherDit {
  AetherDitConfig cfg;
  AetherDitWeights w;
  AetherDitLayerWeights* layers;
  // optional live timing of the dominant kernel (attention): one CUDA event pair per layer, recorded on the
  // launching stream, read back by aether_dit_read_timing after the caller synchronised.
  bool timing = false;
  cudaEvent_t* ev = nullptr;     // [2 * num_layers]
  int ev_used = 0;               // launches recorded by the last forward
};

namespace {
inline int64_t align256(int64_t x) { return (x + 255) & ~int64_t(255); }

struct Workspace {
  char* hidden;    // bf16 [B*S, D]      residual stream (text | video)
  char* xn;        // bf16 [B*S, D]      normalised / modulated activations
  char* qkv;       // bf16 [B*S, 3D]
  char* attn;      // bf16 [B*S, D]
  char* ffh;       // bf16 [B*S, ff*D]   (also hosts the patch matrix [B*Sv, Cin*4] and proj tokens before layer 0 / after the last)
  float* temb_sin; // fp32 [B, D]
  float* temb_h;   // fp32 [B, T]
  float* temb;     // fp32 [B, T]
  float* mod;      // fp32 [B, (12L+2) D]
  char* attn_ws;   // scratch of the attention kernel's split tail wave (may be empty)
  int64_t attn_ws_bytes;
  int64_t total;
};

Workspace carve(const AetherDitConfig& c, int B, int S, char* base) {
  const int64_t D = int64_t(c.num_heads) * c.head_dim;
  const int64_t rows = int64_t(B) * S;
  Workspace ws;
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align256(bytes);
    return p;
  };
  ws.hidden = take(rows * D * 2);
  ws.xn = take(rows * D * 2);
  ws.qkv = take(rows * 3 * D * 2);
  ws.attn = take(rows * D * 2);
  ws.ffh = take(rows * c.ff_mult * D * 2);
  ws.temb_sin = reinterpret_cast<float*>(take(int64_t(B) * D * 4));
  ws.temb_h = reinterpret_cast<float*>(take(int64_t(B) * c.time_embed_dim * 4));
  ws.temb = reinterpret_cast<float*>(take(int64_t(B) * c.time_embed_dim * 4));
  ws.mod = reinterpret_cast<float*>(take(int64_t(B) * (12 * int64_t(c.num_layers) + 2) * D * 4));
  ws.attn_ws_bytes = (c.attention_fp16_pv == 5 && c.attention_split_tail != 0)
                         ? aether::attention_v3_workspace_bytes(B, S, c.num_heads) : 0;
  ws.attn_ws = take(ws.attn_ws_bytes);
  ws.total = off;
  return ws;
}
}  // namespace

extern "C" int aether_dit_create(const AetherDitConfig* cfg, const AetherDitWeights* w, AetherDit** out) {
  if (!cfg || !w || !out || !w->layers) return AETHER_ERR_INVALID;
  if (cfg->head_dim != 64 || cfg->patch_size != 2 || cfg->num_layers <= 0 || cfg->num_heads <= 0 ||
      (cfg->num_heads * cfg->head_dim) % 256 != 0 || cfg->num_heads % 2 != 0 || cfg->ff_mult <= 0 ||
      cfg->time_embed_dim % 8 != 0 || cfg->text_embed_dim % 8 != 0 || cfg->in_channels % 2 != 0) {
    fprintf(stderr, "[aether_b200] dit_create: unsupported geometry (need head_dim 64, patch 2, D %% 256 == 0)\n");
    return AETHER_ERR_INVALID;
  }
  AetherDit* h = new (std::nothrow) AetherDit;
  if (!h) return AETHER_ERR_INVALID;
  h->cfg = *cfg;
  h->w = *w;
  h->layers = new (std::nothrow) AetherDitLayerWeights[cfg->num_layers];
  if (!h->layers) {
    delete h;
    return AETHER_ERR_INVALID;
  }
  for (int i = 0; i < cfg->num_layers; ++i) h->layers[i] = w->layers[i];
  h->w.layers = h->layers;
  *out = h;
  return AETHER_OK;
}

extern "C" void aether_dit_destroy(AetherDit* h) {
  if (!h) return;
  if (h->ev) {
    for (int i = 0; i < 2 * h->cfg.num_layers; ++i) cudaEventDestroy(h->ev[i]);
    delete[] h->ev;
  }
  delete[] h->layers;
  delete h;
}

extern "C" int aether_dit_enable_timing(AetherDit* h, int32_t enable) {
  if (!h) return AETHER_ERR_INVALID;
  if (enable && !h->ev) {
    h->ev = new (std::nothrow) cudaEvent_t[2 * h->cfg.num_layers];
    if (!h->ev) return AETHER_ERR_INVALID;
    for (int i = 0; i < 2 * h->cfg.num_layers; ++i) AETHER_CUDA_OK(cudaEventCreate(&h->ev[i]));
  }
  h->timing = enable != 0;
  h->ev_used = 0;
  return AETHER_OK;
}

extern "C" int aether_dit_read_timing(AetherDit* h, float* attention_ms_total, int32_t* attention_launches) {
  if (!h || !attention_ms_total || !attention_launches || !h->ev) return AETHER_ERR_INVALID;
  float total = 0.f;
  for (int i = 0; i < h->ev_used; ++i) {
    float ms = 0.f;
    AETHER_CUDA_OK(cudaEventElapsedTime(&ms, h->ev[2 * i], h->ev[2 * i + 1]));   // fails if not yet complete
    total += ms;
  }
  *attention_ms_total = total;
  *attention_launches = h->ev_used;
  return AETHER_OK;
}

extern "C" int aether_dit_set_pos_embedding(AetherDit* h, const void* pos_bf16) {
  if (!h) return AETHER_ERR_INVALID;
  h->w.pos_embedding = pos_bf16;
  return AETHER_OK;
}

extern "C" int64_t aether_dit_workspace_bytes(const AetherDit* h, int32_t B, int32_t F, int32_t H, int32_t W,
                                              int32_t St) {
  if (!h || B <= 0 || F <= 0 || H <= 0 || W <= 0 || St < 0) return -1;
  const int S = St + F * (H / 2) * (W / 2);
  return carve(h->cfg, B, S, nullptr).total + 256;
}

// Model input = channel concat of in0[B0, F, C0, H, W] (B0 == 1: broadcast over the batch, the reference's
// torch.cat([latents] * 2)) and in1[B, F, C1, H, W] (C1 may be 0); text [text_batch, St, dim] and timesteps
// [timesteps_batch] broadcast when their batch is 1 (prompt_embeds.repeat / t.expand in the reference :862-869).
static int dit_forward_impl(AetherDit* h, const void* in0, int C0, int B0, const void* in1, int C1, const void* text,
                            int text_batch, const int64_t* timesteps, int timesteps_batch, const float* rope_cos,
                            const float* rope_sin, void* out, int32_t B, int32_t F, int32_t H, int32_t W, int32_t St,
                            void* workspace, int64_t workspace_bytes, int32_t n_layers, void* stream_) {
  using namespace aether;
  if (!h || !in0 || !text || !timesteps || !out || !workspace) return AETHER_ERR_INVALID;
  const AetherDitConfig& c = h->cfg;
  AETHER_CHECK_ARG(B > 0 && B <= 8 && F > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && St >= 0);
  AETHER_CHECK_ARG((rope_cos == nullptr) == (rope_sin == nullptr));
  AETHER_CHECK_ARG(C0 + C1 == c.in_channels && (B0 == 1 || B0 == B) && (text_batch == 1 || text_batch == B) &&
                   (timesteps_batch == 1 || timesteps_batch == B));
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int D = c.num_heads * c.head_dim;
  const int T = c.time_embed_dim;
  const int L = c.num_layers;
  const int Sv = F * (H / 2) * (W / 2);
  const int S = St + Sv;
  const int rows = B * S;
  const int Kp = c.in_channels * 4;
  const int Np = c.out_channels * 4;
  AETHER_CHECK_ARG(int64_t(Kp) <= int64_t(c.ff_mult) * D && Np % 8 == 0);

  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  Workspace ws = carve(c, B, S, base);
  if (ws.total + (base - reinterpret_cast<char*>(workspace)) > workspace_bytes) return AETHER_ERR_WORKSPACE;
  const int64_t mod_stride = (12 * int64_t(L) + 2) * D;   // per-batch stride of ws.mod
  int rc;
#define RUN(expr)        \
  do {                   \
    rc = (expr);         \
    if (rc) return rc;   \
  } while (0)

  // ---- timestep embedding + every AdaLN modulation vector of this step
  RUN(timestep_sinusoid(timesteps, timesteps_batch == 1 ? 0 : 1, ws.temb_sin, B, D, c.flip_sin_to_cos, c.freq_shift,
                        stream));
  RUN(small_m_linear(ws.temb_sin, h->w.w_time1, h->w.b_time1, ws.temb_h, B, T, D, 0, stream));
  RUN(small_m_linear(ws.temb_h, h->w.w_time2, h->w.b_time2, ws.temb, B, T, T, 1, stream));
  RUN(small_m_linear(ws.temb, h->w.w_adaln, h->w.b_adaln, ws.mod, B, (int)mod_stride, T, 1, stream));

  // ---- patch embed: text projection into rows [0, St), video patches into rows [St, S) of each batch
  char* patches = ws.ffh;   // [B*Sv, Kp] bf16 (ffh is free until the first FF1)
  RUN(patchify2(in0, C0, B0, in1, C1, patches, B, F, H, W, stream));
  for (int b = 0; b < B; ++b) {
    char* hb = ws.hidden + int64_t(b) * S * D * 2;
    if (St > 0)
      RUN(gemm_bf16(reinterpret_cast<const char*>(text) + int64_t(text_batch == 1 ? 0 : b) * St * c.text_embed_dim * 2,
                    c.text_embed_dim,
                    h->w.w_text, c.text_embed_dim, hb, D, St, D, c.text_embed_dim, h->w.b_text, 0, nullptr, nullptr,
                    0, 0, 0, -1, stream));
    RUN(gemm_bf16(patches + int64_t(b) * Sv * Kp * 2, Kp, h->w.w_patch, Kp, hb + int64_t(St) * D * 2, D, Sv, D, Kp,
                  h->w.b_patch, 0, nullptr, nullptr, 0, 0, 0, -1, stream));
  }
  if (h->w.pos_embedding) RUN(add_pos_embed(ws.hidden, h->w.pos_embedding, B, S, D, stream));

  // ---- transformer blocks
  const int run_layers = (n_layers < 0 || n_layers > L) ? L : n_layers;
  const float softmax_scale = 1.0f / sqrtf(float(c.head_dim));
  // cfg.fused_qkv_epilogue = 1 runs QK-LayerNorm + RoPE inside the QKV GEMM epilogue (aether_gemm_qkv_norm_rope_bf16).
  // Measured on B200 at S = 15076: 267.7 ms/step fused vs 264.5 ms with the separate qk_norm_rope launch -- four
  // epilogue warps doing 64-wide LayerNorms with row-private cos/sin loads take longer than the 24.5k-clk main loop
  // that should hide them, so the separate HBM-bound kernel (0.12 ms) stays the default.
  const bool fused_qk = c.fused_qkv_epilogue != 0;
  for (int l = 0; l < run_layers; ++l) {
    const AetherDitLayerWeights& lw = h->layers[l];
    // CogVideoXLayerNormZero chunk order: shift, scale, gate, enc_shift, enc_scale, enc_gate
    const float* m1 = ws.mod + int64_t(l) * 12 * D;
    const float* m2 = m1 + 6 * int64_t(D);
    RUN(ln_modulate(ws.hidden, ws.xn, B, S, St, D, lw.norm1_g, lw.norm1_b, c.norm_eps, nullptr, nullptr, m1, m1 + D,
                    m1 + 3 * D, m1 + 4 * D, mod_stride, stream));
    const int f16_from = (c.attention_fp16_pv == 1 || c.attention_fp16_pv == 2) ? 2 * D : -1;
    if (fused_qk && c.head_dim == 64) {
      // to_q/k/v + norm_q/norm_k + rotary embedding in ONE launch: LayerNorm(64) and RoPE run in the GEMM epilogue
      RUN(gemm_qkv_bf16(ws.xn, D, lw.w_qkv, D, ws.qkv, rows, D, lw.b_qkv, S, St, c.num_heads, lw.qn_g, lw.qn_b, lw.kn_g,
                        lw.kn_b, 1e-6f, rope_cos, rope_sin, f16_from, stream));
    } else {
      RUN(gemm_bf16(ws.xn, D, lw.w_qkv, D, ws.qkv, 3 * D, rows, 3 * D, D, lw.b_qkv, 0, nullptr, nullptr, 0, 0, 0, f16_from,
                    stream));
      RUN(qk_norm_rope(ws.qkv, B, S, St, c.num_heads, lw.qn_g, lw.qn_b, lw.kn_g, lw.kn_b, 1e-6f, rope_cos, rope_sin,
                       stream));
    }
    if (h->timing) AETHER_CUDA_OK(cudaEventRecord(h->ev[2 * l], stream));
    RUN(attention_bf16_ws(ws.qkv, ws.attn, B, S, c.num_heads, softmax_scale, c.attention_fp16_pv,
                          ws.attn_ws_bytes > 0 ? ws.attn_ws : nullptr, ws.attn_ws_bytes, stream));
    if (h->timing) {
      AETHER_CUDA_OK(cudaEventRecord(h->ev[2 * l + 1], stream));
      h->ev_used = l + 1;
    }
    RUN(gemm_bf16(ws.attn, D, lw.w_out, D, ws.hidden, D, rows, D, D, lw.b_out, 2, m1 + 2 * D, m1 + 5 * D, mod_stride,
                  S, St, -1, stream));
    RUN(ln_modulate(ws.hidden, ws.xn, B, S, St, D, lw.norm2_g, lw.norm2_b, c.norm_eps, nullptr, nullptr, m2, m2 + D,
                    m2 + 3 * D, m2 + 4 * D, mod_stride, stream));
    RUN(gemm_bf16(ws.xn, D, lw.w_ff1, D, ws.ffh, int64_t(c.ff_mult) * D, rows, c.ff_mult * D, D, lw.b_ff1, 1, nullptr,
                  nullptr, 0, 0, 0, -1, stream));
    RUN(gemm_bf16(ws.ffh, int64_t(c.ff_mult) * D, lw.w_ff2, int64_t(c.ff_mult) * D, ws.hidden, D, rows, D,
                  c.ff_mult * D, lw.b_ff2, 2, m2 + 2 * D, m2 + 5 * D, mod_stride, S, St, -1, stream));
  }

  // ---- tail: norm_final -> norm_out (AdaLayerNorm, chunk order shift, scale) -> proj_out -> unpatchify
  const float* mo = ws.mod + int64_t(L) * 12 * D;
  RUN(ln_modulate(ws.hidden, ws.xn, B, S, St, D, h->w.normf_g, h->w.normf_b, c.norm_eps, h->w.normo_g, h->w.normo_b,
                  mo, mo + D, nullptr, nullptr, mod_stride, stream));
  char* proj = ws.ffh;   // [B*Sv, Np]
  for (int b = 0; b < B; ++b)
    RUN(gemm_bf16(ws.xn + (int64_t(b) * S + St) * D * 2, D, h->w.w_proj, D, proj + int64_t(b) * Sv * Np * 2, Np, Sv, Np,
                  D, h->w.b_proj, 0, nullptr, nullptr, 0, 0, 0, -1, stream));
  RUN(unpatchify(proj, Np, out, B, F, c.out_channels, H, W, stream));
#undef RUN
  return AETHER_OK;
}

extern "C" int aether_dit_forward(AetherDit* h, const void* hidden_in, const void* text, const int64_t* timesteps,
                                  const float* rope_cos, const float* rope_sin, void* out, int32_t B, int32_t F,
                                  int32_t H, int32_t W, int32_t St, void* workspace, int64_t workspace_bytes,
                                  int32_t n_layers, void* stream_) {
  if (!h) return AETHER_ERR_INVALID;
  return dit_forward_impl(h, hidden_in, h->cfg.in_channels, B, nullptr, 0, text, B, timesteps, B, rope_cos, rope_sin,
                          out, B, F, H, W, St, workspace, workspace_bytes, n_layers, stream_);
}

extern "C" int aether_dit_forward_split(AetherDit* h, const void* latents, int32_t latents_batch,
                                        int32_t latents_channels, const void* cond, const void* text,
                                        int32_t text_batch, const int64_t* timesteps, int32_t timesteps_batch,
                                        const float* rope_cos, const float* rope_sin, void* out, int32_t B, int32_t F,
                                        int32_t H, int32_t W, int32_t St, void* workspace, int64_t workspace_bytes,
                                        int32_t n_layers, void* stream_) {
  if (!h || !cond) return AETHER_ERR_INVALID;
  return dit_forward_impl(h, latents, latents_channels, latents_batch, cond, h->cfg.in_channels - latents_channels, text,
                          text_batch, timesteps, timesteps_batch, rope_cos, rope_sin, out, B, F, H, W, St, workspace,
                          workspace_bytes, n_layers, stream_);
}
