// Bandwidth-bound kernels of the DiT forward (K2, K5, K6, K7 and the QK-norm/RoPE pre-pass
// of K1).  All are single-pass over their tensors with 16-byte vector accesses; algorithmic bytes per
// launch are given at each kernel.  They replace the LayerNorm / modulation / rotary / conv-patchify /
// unpatchify pieces of diffusers' CogVideoXTransformer3DModel.forward, which the reference calls at
// aether/pipelines/aetherv1_pipeline_cogvideox.py:865-875.
#include "host_util.h"
#include "ptx.cuh"

namespace aether {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}
// 8 consecutive fp32 parameters as two 16-byte loads (p must be 32-byte aligned: all callers index at
// multiples of 8 floats from cudaMalloc'ed / D-strided vectors).
__device__ __forceinline__ void load8(const float* __restrict__ p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// ------------------------------------------------------------------------------------------------
// K2: LayerNorm (+ optional second LayerNorm) + AdaLN modulation.  One warp per row, the row lives in
// registers (VPL 16-byte vectors per lane), two-pass statistics (mean, then centred variance) like torch.
// Algorithmic bytes: 2 * B*S*D*2 (read x, write y) + O(D) parameters.
// ------------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(256)
ln_modulate_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int rows, int S, int St,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                   const float* __restrict__ gamma2, const float* __restrict__ beta2,
                   const float* __restrict__ shift_vid, const float* __restrict__ scale_vid,
                   const float* __restrict__ shift_txt, const float* __restrict__ scale_txt, int64_t mod_bstride) {
  constexpr int D = VPL * 256;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int b = row / S, s = row - b * S;
  const float* shift = (s < St ? shift_txt : shift_vid) + b * mod_bstride;
  const float* scale = (s < St ? scale_txt : scale_vid) + b * mod_bstride;
  const uint4* xr = reinterpret_cast<const uint4*>(x + int64_t(row) * D);
  float v[VPL][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    unpack8(xr[i * 32 + lane], v[i]);
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += v[i][j];
  }
  float mean = warp_sum(sum) * (1.0f / D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = v[i][j] - mean;
      sq += d * d;
    }
  float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (gamma != nullptr) {
      float gm[8], bt[8];
      load8(gamma + c, gm);
      load8(beta + c, bt);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = (v[i][j] - mean) * rstd * gm[j] + bt[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = (v[i][j] - mean) * rstd;
    }
  }
  if (gamma2 != nullptr) {   // second LayerNorm over the (fp32) result of the first
    sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[i][j];
    mean = warp_sum(sum) * (1.0f / D);
    sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        sq += d * d;
      }
    rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      float gm[8], bt[8];
      load8(gamma2 + c, gm);
      load8(beta2 + c, bt);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = (v[i][j] - mean) * rstd * gm[j] + bt[j];
    }
  }
  uint4* yr = reinterpret_cast<uint4*>(y + int64_t(row) * D);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 8;
    float o[8], sc[8], sh[8];
    load8(scale + c, sc);
    load8(shift + c, sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = v[i][j] * (1.0f + sc[j]) + sh[j];
    yr[i * 32 + lane] = pack8(o);
  }
}

// Hot-path variant (single LayerNorm): persistent blocks; the combined parameters
//   A[c] = gamma[c] * (1 + scale[c]),  Bv[c] = beta[c] * (1 + scale[c]) + shift[c]
// of the current (batch, text|video) segment are staged ONCE per block in shared memory (2 * D fp32), so a row
// costs 12 KB of HBM traffic and no parameter re-reads from L1/L2 (the generic kernel re-reads 4 vectors per row).
template <int VPL>
__global__ void __launch_bounds__(256, 2)
ln_modulate_smem_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int S, int St,
                        const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                        const float* __restrict__ shift_vid, const float* __restrict__ scale_vid,
                        const float* __restrict__ shift_txt, const float* __restrict__ scale_txt, int64_t mod_bstride) {
  constexpr int D = VPL * 256;
  extern __shared__ float prm[];            // [2][D]
  float* sA = prm;
  float* sB = prm + D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // segments: (b, text) then (b, video) for every batch item
  for (int seg = 0; seg < 2 * B; ++seg) {
    const int b = seg >> 1;
    const bool is_txt = (seg & 1) == 0;
    const int r0 = b * S + (is_txt ? 0 : St);
    const int nrows = is_txt ? St : S - St;
    if (nrows <= 0) continue;
    const float* shift = (is_txt ? shift_txt : shift_vid) + b * mod_bstride;
    const float* scale = (is_txt ? scale_txt : scale_vid) + b * mod_bstride;
    __syncthreads();                        // previous segment's readers are done with sA / sB
    for (int c = threadIdx.x; c < D; c += 256) {
      const float g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f, sc = 1.f + scale[c];
      sA[c] = g * sc;
      sB[c] = bt * sc + shift[c];
    }
    __syncthreads();
    for (int r = blockIdx.x * 8 + warp; r < nrows; r += gridDim.x * 8) {
      const int64_t row = int64_t(r0 + r);
      const uint4* xr = reinterpret_cast<const uint4*>(x + row * D);
      // The row stays PACKED in registers (VPL x 16 bytes) and is widened three times (sum, squares, output): the
      // fp32 copy cost 214 registers at D = 3072 and left ONE 8-warp block per SM; packed (and with the launch bound
      // that stops ptxas from caching the widened values) two blocks are resident.
      uint4 raw[VPL];
#pragma unroll
      for (int i = 0; i < VPL; ++i) raw[i] = xr[i * 32 + lane];
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        float v[8];
        unpack8(raw[i], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += v[j];
      }
      const float mean = warp_sum(sum) * (1.0f / D);
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        float v[8];
        unpack8(raw[i], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[j] - mean;
          sq += d * d;
        }
      }
      const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
      uint4* yr = reinterpret_cast<uint4*>(y + row * D);
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int c = (i * 32 + lane) * 8;
        const float4 a0 = *reinterpret_cast<const float4*>(sA + c), a1 = *reinterpret_cast<const float4*>(sA + c + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(sB + c), b1 = *reinterpret_cast<const float4*>(sB + c + 4);
        float v[8], o[8];
        unpack8(raw[i], v);
        o[0] = (v[0] - mean) * rstd * a0.x + b0.x; o[1] = (v[1] - mean) * rstd * a0.y + b0.y;
        o[2] = (v[2] - mean) * rstd * a0.z + b0.z; o[3] = (v[3] - mean) * rstd * a0.w + b0.w;
        o[4] = (v[4] - mean) * rstd * a1.x + b1.x; o[5] = (v[5] - mean) * rstd * a1.y + b1.y;
        o[6] = (v[6] - mean) * rstd * a1.z + b1.z; o[7] = (v[7] - mean) * rstd * a1.w + b1.w;
        yr[i * 32 + lane] = pack8(o);
      }
    }
  }
}

int ln_modulate(const void* x, void* y, int B, int S, int St, int D, const float* gamma, const float* beta, float eps,
                const float* gamma2, const float* beta2, const float* shift_vid, const float* scale_vid,
                const float* shift_txt, const float* scale_txt, int64_t mod_bstride, cudaStream_t stream) {
  AETHER_CHECK_ARG(B > 0 && S > 0 && D > 0 && D % 256 == 0 && D <= 4096);
  AETHER_CHECK_ARG(shift_vid && scale_vid);
  AETHER_CHECK_ARG((gamma == nullptr) == (beta == nullptr) && (gamma2 == nullptr) == (beta2 == nullptr));
  if (shift_txt == nullptr) { shift_txt = shift_vid; scale_txt = scale_vid; }
  const int rows = B * S;
  const int grid = (int)ceil_div(rows, 8);
  auto xb = reinterpret_cast<const __nv_bfloat16*>(x);
  auto yb = reinterpret_cast<__nv_bfloat16*>(y);
  if (gamma2 == nullptr && rows >= 1024) {   // hot path: parameters staged in shared memory, persistent blocks
    const int pgrid = num_sms() * 2;          // two resident 8-warp blocks per SM (<= 128 registers per thread)
    const size_t smem = size_t(2) * D * sizeof(float);
#define LAUNCH_S(V)                                                                                            \
  ln_modulate_smem_kernel<V><<<pgrid, 256, smem, stream>>>(xb, yb, B, S, St, gamma, beta, eps, shift_vid,      \
                                                            scale_vid, shift_txt, scale_txt, mod_bstride)
    switch (D / 256) {
      case 1: LAUNCH_S(1); break;
      case 2: LAUNCH_S(2); break;
      case 3: LAUNCH_S(3); break;
      case 4: LAUNCH_S(4); break;
      case 8: LAUNCH_S(8); break;
      case 12: LAUNCH_S(12); break;
      case 16: LAUNCH_S(16); break;
      default: goto generic;
    }
#undef LAUNCH_S
    AETHER_CUDA_OK(cudaGetLastError());
    return AETHER_OK;
  }
generic:
#define LAUNCH(V)                                                                                             \
  ln_modulate_kernel<V><<<grid, 256, 0, stream>>>(xb, yb, rows, S, St, gamma, beta, eps, gamma2, beta2,       \
                                                   shift_vid, scale_vid, shift_txt, scale_txt, mod_bstride)
  switch (D / 256) {
    case 1: LAUNCH(1); break;
    case 2: LAUNCH(2); break;
    case 4: LAUNCH(4); break;
    case 8: LAUNCH(8); break;
    case 3: LAUNCH(3); break;
    case 5: LAUNCH(5); break;
    case 6: LAUNCH(6); break;
    case 7: LAUNCH(7); break;
    case 9: LAUNCH(9); break;
    case 10: LAUNCH(10); break;
    case 11: LAUNCH(11); break;
    case 12: LAUNCH(12); break;
    case 13: LAUNCH(13); break;
    case 14: LAUNCH(14); break;
    case 15: LAUNCH(15); break;
    case 16: LAUNCH(16); break;
    default: AETHER_CHECK_ARG(!"unsupported D (need D % 256 == 0 and D <= 4096)");
  }
#undef LAUNCH
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}

// ------------------------------------------------------------------------------------------------
// QK-LayerNorm(64) + 3-D RoPE, in place on the q and k thirds of qkv[B,S,3,H,64].
// 8 threads per (token, head) vector of 64 (one 16-byte vector each).  Every thread handles the q vector AND the k
// vector of the same (token, head, sub-vector): two independent 16-byte loads are in flight before any math and the
// token's cos/sin row is fetched once for both.  grid.y = token, grid.x covers the H head-vectors.
// Algorithmic bytes: 2 * (2/3) * |qkv|  (+ cos/sin, L2-resident).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void qk_norm_vec(float (&f)[8], const float* __restrict__ g, const float* __restrict__ bt,
                                            int sub, float eps) {
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) sum += f[j];
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  sum += __shfl_xor_sync(0xffffffffu, sum, 4);
  const float mean = sum * (1.0f / 64);
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float d = f[j] - mean;
    sq += d * d;
  }
  sq += __shfl_xor_sync(0xffffffffu, sq, 1);
  sq += __shfl_xor_sync(0xffffffffu, sq, 2);
  sq += __shfl_xor_sync(0xffffffffu, sq, 4);
  const float rstd = rsqrtf(sq * (1.0f / 64) + eps);
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = (f[j] - mean) * rstd;
  float gm[8], bb[8];
  load8(g + sub * 8, gm);
  load8(bt + sub * 8, bb);
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = f[j] * gm[j] + bb[j];
}
__device__ __forceinline__ void rope_vec(float (&f)[8], const float (&c)[8], const float (&sn)[8]) {
  // upstream rounds LayerNorm's output to bf16 before apply_rotary_emb upcasts it again
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = __bfloat162float(__float2bfloat16_rn(f[j]));
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    const float xr = f[j], xi = f[j + 1];
    f[j] = xr * c[j] - xi * sn[j];
    f[j + 1] = xi * c[j + 1] + xr * sn[j + 1];
  }
}

// TOK tokens per thread: all 2 * TOK activation vectors AND the tokens' cos / sin rows are requested before the first
// use (round 2: the one-token version had 32 bytes in flight per thread, then a dependent 64-byte cos/sin fetch --
// 56 % of the copy bandwidth); the LayerNorm affine parameters of the thread's 8 channels are fetched once.
template <int TOK>
__global__ void __launch_bounds__(128)
qk_norm_rope_kernel(__nv_bfloat16* __restrict__ qkv, int total, int S, int St, int H, const float* __restrict__ gq,
                    const float* __restrict__ bq, const float* __restrict__ gk, const float* __restrict__ bk,
                    float eps, const float* __restrict__ cosb, const float* __restrict__ sinb) {
  const int vec = blockIdx.x * 128 + threadIdx.x;   // 16-byte vector index inside the q span of a token
  if (vec >= H * 8) return;                         // whole warps only (H even)
  const int sub = vec & 7;                          // which 8 of the 64 elements
  const int tok0 = blockIdx.y * TOK;                // token = b * S + s
  uint4 qraw[TOK], kraw[TOK];
  float c[TOK][8], sn[TOK][8];
  bool rope[TOK];
#pragma unroll
  for (int i = 0; i < TOK; ++i) {
    const int token = tok0 + i;
    if (token < total) {
      const uint4* qp = reinterpret_cast<const uint4*>(qkv + int64_t(token) * (3 * H * 64)) + vec;
      qraw[i] = qp[0];
      kraw[i] = qp[H * 8];
    }
  }
#pragma unroll
  for (int i = 0; i < TOK; ++i) {
    const int token = tok0 + i;
    const int s = token % S;
    rope[i] = cosb != nullptr && token < total && s >= St;
    if (rope[i]) {
      load8(cosb + int64_t(s - St) * 64 + sub * 8, c[i]);
      load8(sinb + int64_t(s - St) * 64 + sub * 8, sn[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < TOK; ++i) {
    const int token = tok0 + i;
    if (token >= total) break;                      // uniform per block.y: shuffles below stay converged
    float fq[8], fk[8];
    unpack8(qraw[i], fq);
    unpack8(kraw[i], fk);
    qk_norm_vec(fq, gq, bq, sub, eps);
    qk_norm_vec(fk, gk, bk, sub, eps);
    if (rope[i]) {
      rope_vec(fq, c[i], sn[i]);
      rope_vec(fk, c[i], sn[i]);
    }
    uint4* qp = reinterpret_cast<uint4*>(qkv + int64_t(token) * (3 * H * 64)) + vec;
    qp[0] = pack8(fq);
    qp[H * 8] = pack8(fk);
  }
}

int qk_norm_rope(void* qkv, int B, int S, int St, int H, const float* gq, const float* bq, const float* gk,
                 const float* bk, float eps, const float* cosb, const float* sinb, cudaStream_t stream) {
  AETHER_CHECK_ARG(B > 0 && S > 0 && H > 0 && gq && bq && gk && bk);
  AETHER_CHECK_ARG((cosb == nullptr) == (sinb == nullptr));
  constexpr int TOK = 2;
  dim3 grid((unsigned)ceil_div(H * 8, 128), (unsigned)ceil_div(B * S, TOK));
  qk_norm_rope_kernel<TOK><<<grid, 128, 0, stream>>>(reinterpret_cast<__nv_bfloat16*>(qkv), B * S, S, St, H, gq, bq, gk,
                                                      bk, eps, cosb, sinb);
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}

// ------------------------------------------------------------------------------------------------
// K6: y[b, n] = sum_k act(x[b, k]) W[n, k] + bias[n] for tiny B (timestep MLP, every AdaLN-Zero linear of
// the step in ONE launch).  One warp per output row n streams the bf16 weight row with 16-byte loads;
// act(x) is staged once per block in shared memory.  HBM-bound: algorithmic bytes = N*K*2.
// ------------------------------------------------------------------------------------------------
constexpr int SML_MAXB = 8;
__global__ void __launch_bounds__(256)
small_m_linear_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ W,
                      const float* __restrict__ bias, float* __restrict__ y, int B, int N, int K, int act) {
  extern __shared__ float xs[];   // [B][K]
  for (int i = threadIdx.x; i < B * K; i += blockDim.x) {
    float v = x[i];
    if (act == 1) v = v / (1.0f + __expf(-v));
    xs[i] = v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int n = blockIdx.x * warps_per_block + (threadIdx.x >> 5); n < N; n += gridDim.x * warps_per_block) {
    float acc[SML_MAXB];
#pragma unroll
    for (int b = 0; b < SML_MAXB; ++b) acc[b] = 0.f;
    const uint4* wr = reinterpret_cast<const uint4*>(W + int64_t(n) * K);
    for (int kv = lane; kv < K / 8; kv += 32) {
      float w[8];
      unpack8(__ldg(wr + kv), w);
#pragma unroll
      for (int b = 0; b < SML_MAXB; ++b) {
        if (b < B) {
          const float* xb = xs + b * K + kv * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[b] = fmaf(w[j], xb[j], acc[b]);
        }
      }
    }
#pragma unroll
    for (int b = 0; b < SML_MAXB; ++b) {
      if (b < B) {
        const float r = warp_sum(acc[b]);
        if (lane == 0) y[int64_t(b) * N + n] = r + (bias ? bias[n] : 0.f);
      }
    }
  }
}

// Register-resident variant for the hot case (K = 256*VPL <= 512, B = NB <= 2): act(x) is loaded once per
// warp and stays in registers while the warp streams weight rows, so shared memory is not on the path.
template <int VPL, int NB>
__global__ void __launch_bounds__(256)
small_m_linear_reg_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ W,
                          const float* __restrict__ bias, float* __restrict__ y, int N, int act) {
  constexpr int K = VPL * 256;
  const int lane = threadIdx.x & 31;
  float xr[NB][VPL][8];
#pragma unroll
  for (int b = 0; b < NB; ++b)
#pragma unroll
    for (int i = 0; i < VPL; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = x[b * K + (i * 32 + lane) * 8 + j];
        if (act == 1) v = v / (1.0f + __expf(-v));
        xr[b][i][j] = v;
      }
  const int warps_per_block = blockDim.x >> 5;
  const int wstride = gridDim.x * warps_per_block;
  for (int n = blockIdx.x * warps_per_block + (threadIdx.x >> 5); n < N; n += wstride) {
    const uint4* wr = reinterpret_cast<const uint4*>(W + int64_t(n) * K);
    uint4 wv[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) wv[i] = __ldg(wr + i * 32 + lane);
    float acc[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float w[8];
      unpack8(wv[i], w);
#pragma unroll
      for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[b] = fmaf(w[j], xr[b][i][j], acc[b]);
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const float r = warp_sum(acc[b]);
      if (lane == 0) y[int64_t(b) * N + n] = r + (bias ? bias[n] : 0.f);
    }
  }
}

int small_m_linear(const float* x, const void* W, const float* bias, float* y, int B, int N, int K, int act,
                   cudaStream_t stream) {
  AETHER_CHECK_ARG(B > 0 && B <= SML_MAXB && N > 0 && K > 0 && K % 8 == 0);
  AETHER_CHECK_ARG(size_t(B) * K * 4 <= 48 * 1024);
  const int rows_per_block = 8;
  int64_t grid = ceil_div(N, rows_per_block);
  const int64_t cap = int64_t(num_sms()) * 32;
  if (grid > cap) grid = cap;
  auto Wb = reinterpret_cast<const __nv_bfloat16*>(W);
  if ((K == 256 || K == 512) && B <= 2) {
    if (K == 256 && B == 1) small_m_linear_reg_kernel<1, 1><<<(unsigned)grid, 256, 0, stream>>>(x, Wb, bias, y, N, act);
    else if (K == 256) small_m_linear_reg_kernel<1, 2><<<(unsigned)grid, 256, 0, stream>>>(x, Wb, bias, y, N, act);
    else if (B == 1) small_m_linear_reg_kernel<2, 1><<<(unsigned)grid, 256, 0, stream>>>(x, Wb, bias, y, N, act);
    else small_m_linear_reg_kernel<2, 2><<<(unsigned)grid, 256, 0, stream>>>(x, Wb, bias, y, N, act);
  } else {
    small_m_linear_kernel<<<(unsigned)grid, 256, size_t(B) * K * 4, stream>>>(x, Wb, bias, y, B, N, K, act);
  }
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}

// ------------------------------------------------------------------------------------------------
// diffusers get_timestep_embedding(scale=1, max_period=10000): emb = t * exp(-ln(1e4) * i / (half - shift)).
// ------------------------------------------------------------------------------------------------
__global__ void timestep_sinusoid_kernel(const int64_t* __restrict__ t, int t_stride, float* __restrict__ emb, int B,
                                         int dim, int flip, float shift) {
  const int half = dim / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * half) return;
  const int b = idx / half, i = idx - b * half;
  const float freq = expf(-9.210340371976184f * float(i) / (float(half) - shift));
  const float arg = float(t[b * t_stride]) * freq;   // t_stride 0: one timestep broadcast over the batch
  const float sv = sinf(arg), cv = cosf(arg);
  float* e = emb + int64_t(b) * dim;
  if (flip) {
    e[i] = cv;
    e[half + i] = sv;
  } else {
    e[i] = sv;
    e[half + i] = cv;
  }
}

int timestep_sinusoid(const int64_t* t, int t_stride, float* emb, int B, int dim, int flip, float shift,
                      cudaStream_t stream) {
  AETHER_CHECK_ARG(B > 0 && dim > 0 && dim % 2 == 0 && (t_stride == 0 || t_stride == 1));
  const int n = B * dim / 2;
  timestep_sinusoid_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(t, t_stride, emb, B, dim, flip, shift);
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}

// ------------------------------------------------------------------------------------------------
// K5 front half: 2x2 patch gather  in[B,F,C,H,W] -> patches[(b,f,y,x), c*4 + dy*2 + dx]  (K index order of
// the flattened Conv2d weight [D, C, 2, 2]); the contraction itself runs on the tcgen05 GEMM.
// One CTA per (b, f, y): the 2C image rows of the patch row are read coalesced (4-byte pairs) into shared
// memory and written back as 16-byte vectors of the token-major layout.  Algorithmic bytes: 2 * |in|.
// ------------------------------------------------------------------------------------------------
// Two sources are concatenated along the channel axis on the fly (reference :832-859: cat([latents]*2) and
// cat([latent_model_input, latent_condition], dim=2)): channels [0, C0) come from in0[B0, F, C0, H, W] (B0 == 1
// broadcasts the latents over the CFG batch), channels [C0, C0+C1) from in1[B, F, C1, H, W].  C1 == 0: single source.
__global__ void __launch_bounds__(256)
patchify_kernel(const __nv_bfloat16* __restrict__ in0, int C0, int B0, const __nv_bfloat16* __restrict__ in1, int C1,
                __nv_bfloat16* __restrict__ out, int F, int H, int W) {
  extern __shared__ uint32_t tile[];   // [2C][W/2] pairs
  const int C = C0 + C1;
  const int Wp = W / 2, Hp = H / 2;
  const int y = blockIdx.x % Hp;
  const int bf = blockIdx.x / Hp;
  const int b = bf / F, f = bf - b * F;
  const uint32_t* src0 = reinterpret_cast<const uint32_t*>(in0 + (int64_t(B0 == 1 ? 0 : b) * F + f) * C0 * H * W);
  const uint32_t* src1 = C1 ? reinterpret_cast<const uint32_t*>(in1 + (int64_t(b) * F + f) * C1 * H * W) : nullptr;
  const int npairs = 2 * C * Wp;
  for (int i = threadIdx.x; i < npairs; i += blockDim.x) {
    const int r = i / Wp, xp = i - r * Wp;           // r = c*2 + dy
    const int c = r >> 1, dy = r & 1;
    tile[i] = c < C0 ? src0[(int64_t(c) * H + 2 * y + dy) * Wp + xp]
                     : src1[(int64_t(c - C0) * H + 2 * y + dy) * Wp + xp];
  }
  __syncthreads();
  uint32_t* dst = reinterpret_cast<uint32_t*>(out + (int64_t(bf) * Hp + y) * Wp * (C * 4));
  for (int i = threadIdx.x; i < npairs; i += blockDim.x) {
    const int xp = i / (2 * C), r = i - xp * (2 * C);   // output pair index: token xp, k-pair r = c*2 + dy
    dst[i] = tile[r * Wp + xp];
  }
}

int patchify2(const void* in0, int C0, int B0, const void* in1, int C1, void* patches, int B, int F, int H, int W,
              cudaStream_t stream) {
  const int C = C0 + C1;
  AETHER_CHECK_ARG(B > 0 && F > 0 && C0 > 0 && C1 >= 0 && H % 2 == 0 && W % 2 == 0 && in0 && (C1 == 0 || in1));
  AETHER_CHECK_ARG(B0 == 1 || B0 == B);
  const size_t smem = size_t(2) * C * (W / 2) * 4;
  AETHER_CHECK_ARG(smem <= 160 * 1024);
  static SmemGrant grant;
  AETHER_CUDA_OK(ensure_dynamic_smem(grant, patchify_kernel, smem));
  patchify_kernel<<<(unsigned)(B * F * (H / 2)), 256, smem, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(in0), C0, B0, reinterpret_cast<const __nv_bfloat16*>(in1), C1,
      reinterpret_cast<__nv_bfloat16*>(patches), F, H, W);
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}

int patchify(const void* in, void* patches, int B, int F, int C, int H, int W, cudaStream_t stream) {
  return patchify2(in, C, B, nullptr, 0, patches, B, F, H, W, stream);
}

// K7 back half: inverse of the above for the proj_out tokens:  out[b,f,c,2y+dy,2x+dx] = tok[(b,f,y,x), c*4+dy*2+dx]
__global__ void __launch_bounds__(256)
unpatchify_kernel(const __nv_bfloat16* __restrict__ tok, int64_t ld_tok, __nv_bfloat16* __restrict__ out, int C,
                  int H, int W) {
  extern __shared__ uint32_t tile[];   // [2C][W/2] pairs
  const int Wp = W / 2, Hp = H / 2;
  const int y = blockIdx.x % Hp;
  const int bf = blockIdx.x / Hp;
  const int npairs = 2 * C * Wp;
  const __nv_bfloat16* src = tok + (int64_t(bf) * Hp + y) * Wp * ld_tok;
  for (int i = threadIdx.x; i < npairs; i += blockDim.x) {
    const int xp = i / (2 * C), r = i - xp * (2 * C);
    tile[r * Wp + xp] = *reinterpret_cast<const uint32_t*>(src + int64_t(xp) * ld_tok + 2 * r);
  }
  __syncthreads();
  uint32_t* dst = reinterpret_cast<uint32_t*>(out + int64_t(bf) * C * H * W);
  for (int i = threadIdx.x; i < npairs; i += blockDim.x) {
    const int r = i / Wp, xp = i - r * Wp;
    const int c = r >> 1, dy = r & 1;
    dst[(int64_t(c) * H + 2 * y + dy) * Wp + xp] = tile[i];
  }
}

int unpatchify(const void* tok, int64_t ld_tok, void* out, int B, int F, int C, int H, int W, cudaStream_t stream) {
  AETHER_CHECK_ARG(B > 0 && F > 0 && C > 0 && H % 2 == 0 && W % 2 == 0 && ld_tok % 2 == 0 && ld_tok >= 4 * C);
  const size_t smem = size_t(2) * C * (W / 2) * 4;
  AETHER_CHECK_ARG(smem <= 160 * 1024);
  static SmemGrant grant;
  AETHER_CUDA_OK(ensure_dynamic_smem(grant, unpatchify_kernel, smem));
  unpatchify_kernel<<<(unsigned)(B * F * (H / 2)), 256, smem, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(tok), ld_tok, reinterpret_cast<__nv_bfloat16*>(out), C, H, W);
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}

// x[b, s, :] += pos[s, :]  (fp32 add, bf16 storage: matches `embeds + pos_embedding` in bf16)
__global__ void add_pos_embed_kernel(uint4* __restrict__ x, const uint4* __restrict__ pos, int64_t per_batch_vecs,
                                     int64_t total_vecs) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total_vecs) return;
  float a[8], b[8];
  unpack8(x[i], a);
  unpack8(__ldg(pos + (i % per_batch_vecs)), b);
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] += b[j];
  x[i] = pack8(a);
}

int add_pos_embed(void* x, const void* pos, int B, int S, int D, cudaStream_t stream) {
  AETHER_CHECK_ARG(B > 0 && S > 0 && D % 8 == 0);
  const int64_t per = int64_t(S) * D / 8, total = per * B;
  add_pos_embed_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(
      reinterpret_cast<uint4*>(x), reinterpret_cast<const uint4*>(pos), per, total);
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}

}  // namespace aether

// ------------------------------------------------------------------------------------------------ C ABI
using namespace aether;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
extern "C" {
int aether_ln_modulate(const void* x, void* y, int32_t B, int32_t S, int32_t St, int32_t D, const float* gamma,
                       const float* beta, float eps, const float* gamma2, const float* beta2, const float* shift_vid,
                       const float* scale_vid, const float* shift_txt, const float* scale_txt, int64_t mod_bstride,
                       void* stream) {
  return ln_modulate(x, y, B, S, St, D, gamma, beta, eps, gamma2, beta2, shift_vid, scale_vid, shift_txt, scale_txt,
                     mod_bstride, ST(stream));
}
int aether_qk_norm_rope(void* qkv, int32_t B, int32_t S, int32_t St, int32_t H, const float* gq, const float* bq,
                        const float* gk, const float* bk, float eps, const float* cos, const float* sin,
                        void* stream) {
  return qk_norm_rope(qkv, B, S, St, H, gq, bq, gk, bk, eps, cos, sin, ST(stream));
}
int aether_small_m_linear(const float* x, const void* W, const float* bias, float* y, int32_t B, int32_t N, int32_t K,
                          int32_t act, void* stream) {
  return small_m_linear(x, W, bias, y, B, N, K, act, ST(stream));
}
int aether_timestep_sinusoid(const int64_t* timesteps, float* emb, int32_t B, int32_t dim, int32_t flip_sin_to_cos,
                             float freq_shift, void* stream) {
  return timestep_sinusoid(timesteps, 1, emb, B, dim, flip_sin_to_cos, freq_shift, ST(stream));
}
int aether_patchify(const void* in, void* patches, int32_t B, int32_t F, int32_t C, int32_t H, int32_t W,
                    void* stream) {
  return patchify(in, patches, B, F, C, H, W, ST(stream));
}
int aether_unpatchify(const void* tok, int64_t ld_tok, void* out, int32_t B, int32_t F, int32_t C, int32_t H,
                      int32_t W, void* stream) {
  return unpatchify(tok, ld_tok, out, B, F, C, H, W, ST(stream));
}
int aether_add_pos_embed(void* x, const void* pos, int32_t B, int32_t S, int32_t D, void* stream) {
  return add_pos_embed(x, pos, B, S, D, ST(stream));
}
}
