// K9, bandwidth-bound half: GroupNorm / SpatialNorm3D + SiLU, nearest up-sampling, temporal
// average pooling, layout conversions, posterior sampling and the tile blends of AutoencoderKLCogVideoX
// (third-party diffusers module behind vae.encode / vae.decode, aetherv1_pipeline_cogvideox.py:557-620, :931, :936).
// All tensors are channels-last  x[T, H, W, C]  bf16 unless stated; every kernel is a single streaming pass with
// 16-byte vectors (8 channels per thread).  Algorithmic bytes are noted per kernel.
#include "host_util.h"
#include "ptx.cuh"
#include "vae_internal.h"

namespace aether {

__device__ __forceinline__ void unpack8v(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8v(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ------------------------------------------------------------------------------------------------
// GroupNorm statistics over x[N, C] (N = T*H*W positions): per-block partial sums of every 4-channel quad in
// a fixed order, then an fp64 finalize per group -> mean[G], rstd[G].  Deterministic (no floating-point atomics).
// Algorithmic bytes: N*C*2 read.
// Shape of the reduction (round 2): at most ONE 1024-thread block per SM, so a group has <= #SM x (C/G/4) partials
// (38 KB .. 150 KB in total) and a single block can finalize all groups in a few microseconds -- which lets the last
// block to finish do it (one launch instead of two; the previous 8 x 256-thread blocks per SM left 0.3 - 1.2 MB of
// partials, a 9 us stand-alone finalize launch per norm = 3 % of a VAE decode).
// ------------------------------------------------------------------------------------------------
constexpr int GN_THREADS = 1024;

// Finalize of ONE group by ONE warp: the group's partials are read with eight independent L2 loads per lane and pass,
// summed in fp64 per lane in index order and combined by a fixed shuffle tree.  Used by the fused tail of
// gn_partial_kernel and by the stand-alone gn_finalize_kernel: same summation order -> identical bits.
__device__ __forceinline__ void gn_finalize_group(const float* __restrict__ partial, int nblocks, int C, int G, int g,
                                                  int lane, double count, float eps, float* __restrict__ mean_rstd) {
  const int quads_per_group = (C / G) / 4;
  const int n = nblocks * quads_per_group;
  double s = 0.0, q = 0.0;
  for (int i0 = lane; i0 < n; i0 += 32 * 8) {
    float2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * 32;
      if (i < n) {
        const int b = i / quads_per_group, k = i - b * quads_per_group;
        // L2 (coherent) loads: in the fused kernel these values were written by other blocks of the same launch
        v[u] = __ldcg(reinterpret_cast<const float2*>(partial + (int64_t(b) * (C / 4) + g * quads_per_group + k) * 2));
      } else {
        v[u] = make_float2(0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      s += v[u].x;
      q += v[u].y;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane == 0) {
    const double mean = s / count;
    const double var = fmax(q / count - mean * mean, 0.0);
    mean_rstd[2 * g] = float(mean);
    mean_rstd[2 * g + 1] = float(1.0 / sqrt(var + double(eps)));
  }
}

// `counter` (nullable): a zeroed device word.  With it the LAST block to finish (ticket from one integer atomicAdd) also
// finalizes all G groups (one warp per group) and resets the counter; without it the caller launches gn_finalize_kernel.
__global__ void __launch_bounds__(GN_THREADS)
gn_partial_kernel(const __nv_bfloat16* __restrict__ x, int64_t N, int C, float* __restrict__ partial, int G, double count,
                  float eps, float* __restrict__ mean_rstd, unsigned* __restrict__ counter) {
  // thread layout: tpr = C/8 threads per row; rows advance by (GN_THREADS / tpr) * gridDim.x
  const int tpr = C / 8;
  const int lane_c = threadIdx.x % tpr;
  const int row_in_blk = threadIdx.x / tpr;
  const int rows_per_blk = GN_THREADS / tpr;
  float s[2] = {0.f, 0.f}, q[2] = {0.f, 0.f};
  const int64_t stride = int64_t(gridDim.x) * rows_per_blk;
  int64_t r = int64_t(blockIdx.x) * rows_per_blk + row_in_blk;
  // 4 independent 16-byte loads in flight per thread (the loop is pure streaming; one dependent load per
  // iteration left the kernel latency-bound)
  for (; r + 3 * stride < N; r += 4 * stride) {
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(x + (r + k * stride) * C + lane_c * 8));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float f[8];
      unpack8v(u[k], f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[0] += f[j]; q[0] += f[j] * f[j];
        s[1] += f[4 + j]; q[1] += f[4 + j] * f[4 + j];
      }
    }
  }
  for (; r < N; r += stride) {
    float f[8];
    unpack8v(__ldg(reinterpret_cast<const uint4*>(x + r * C + lane_c * 8)), f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s[0] += f[j]; q[0] += f[j] * f[j];
      s[1] += f[4 + j]; q[1] += f[4 + j] * f[4 + j];
    }
  }
  extern __shared__ float red[];   // [GN_THREADS][4]
  red[threadIdx.x * 4 + 0] = s[0]; red[threadIdx.x * 4 + 1] = q[0];
  red[threadIdx.x * 4 + 2] = s[1]; red[threadIdx.x * 4 + 3] = q[1];
  __syncthreads();
  if (threadIdx.x < tpr) {         // fixed-order reduction over the rows handled by this block
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int rr = 0; rr < rows_per_blk; ++rr) {
      const float* p = red + (rr * tpr + threadIdx.x) * 4;
      a0 += p[0]; a1 += p[1]; a2 += p[2]; a3 += p[3];
    }
    float* out = partial + (int64_t(blockIdx.x) * (C / 4) + threadIdx.x * 2) * 2;
    out[0] = a0; out[1] = a1; out[2] = a2; out[3] = a3;
  }
  if (counter == nullptr) return;
  // ---- fused finalize: the last block to arrive sees every block's partials (release: fence + atomic; acquire: atomic + fence)
  __shared__ unsigned ticket;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) ticket = atomicAdd(counter, 1u);
  __syncthreads();
  if (ticket != gridDim.x - 1) return;
  __threadfence();
  for (int g = threadIdx.x >> 5; g < G; g += GN_THREADS / 32)
    gn_finalize_group(partial, int(gridDim.x), C, G, g, threadIdx.x & 31, count, eps, mean_rstd);
  if (threadIdx.x == 0) *counter = 0u;     // ready for the next (stream-ordered) launch
}

// Stand-alone finalize (one block, one warp per group) -- the path of callers without a counter word.
__global__ void __launch_bounds__(GN_THREADS)
gn_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, int G, double count, float eps,
                   float* __restrict__ mean_rstd) {
  for (int g = threadIdx.x >> 5; g < G; g += GN_THREADS / 32)
    gn_finalize_group(partial, nblocks, C, G, g, threadIdx.x & 31, count, eps, mean_rstd);
}

// ------------------------------------------------------------------------------------------------
// y = SiLU?( ((x - mean_g) * rstd_g * gamma_c + beta_c) [* zy[map(pos), c] + zb[map(pos), c]] )
//   GroupNorm (encoder) or CogVideoXSpatialNorm3D (decoder; zy/zb = conv_y(zq)/conv_b(zq) evaluated at LATENT
//   resolution -- a 1x1x1 conv commutes with nearest interpolation -- and gathered here by the nearest-neighbour
//   index maps of F.interpolate).  Output rows may land inside a time-padded buffer (y is just a pointer).
// Algorithmic bytes: 2*N*C*2 (+ the small zy/zb tables, L2-resident).
// ------------------------------------------------------------------------------------------------
struct GnApplyArgs {
  const __nv_bfloat16* x;
  __nv_bfloat16* y;
  int64_t N;
  int C, G;
  const float* mean_rstd;
  const float* gamma;
  const float* beta;
  const __nv_bfloat16* zy;     // nullable: [Tz, hz, wz, C]
  const __nv_bfloat16* zb;
  int H, W, hz, wz;
  int zld;                     // row stride (elements) of the zy / zb tables
  int silu;
  // optional second destination for the rows >= y2_from (the last two frames = the causal conv's cache for the next
  // frame batch, written here instead of by a separate device-to-device copy of what was just stored)
  __nv_bfloat16* y2;
  int64_t y2_from;
};

// Work layout: a thread owns ONE 8-channel vector of the row for the whole launch, so its group statistics and
// affine parameters (mean, rstd*gamma, beta: 24 registers) are fetched once; a block covers 256 / (C/8) consecutive
// rows per pass and walks the tensor in a grid-stride loop over row groups with four independent 16-byte loads in
// flight per thread.  Round 1 looked all of that up per element (8 mean/rstd pairs, 64 B of gamma/beta, an int64
// division per vector, three more for the latent gather) and sat at 17 % of the copy bandwidth.
// MapT: `const int*` (device array, the per-kernel C entry point) or `IMap` (by value, the handle-level executor).
template <bool SPATIAL, class MapT>
__global__ void __launch_bounds__(256) gn_apply_kernel(const GnApplyArgs a, const MapT tmap) {
  const int tpr = a.C / 8;
  const int lane_c = threadIdx.x % tpr;
  const int row_in_blk = threadIdx.x / tpr;
  const int rows_per_blk = 256 / tpr;
  const int c0 = lane_c * 8;
  const int gs = a.C / a.G;
  float mu[8], sc[8], bt[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (c0 + j) / gs;
    mu[j] = a.mean_rstd[2 * g];
    sc[j] = a.mean_rstd[2 * g + 1] * __ldg(a.gamma + c0 + j);
    bt[j] = __ldg(a.beta + c0 + j);
  }
  const uint32_t n_rows = (uint32_t)a.N;
  const uint32_t stride = gridDim.x * rows_per_blk;
  const uint32_t hw = SPATIAL ? uint32_t(a.H) * a.W : 1u;
  for (uint32_t r0 = blockIdx.x * rows_per_blk + row_in_blk; r0 < n_rows; r0 += 4 * stride) {
    uint4 xin[4], zyv[4], zbv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t r = r0 + k * stride;
      if (r < n_rows) {
        xin[k] = __ldg(reinterpret_cast<const uint4*>(a.x + int64_t(r) * a.C + c0));
        if (SPATIAL) {
          const uint32_t t = r / hw, rem = r - t * hw;
          const uint32_t yy = rem / uint32_t(a.W), xx = rem - yy * uint32_t(a.W);
          const int64_t zr = (int64_t(tmap[t]) * a.hz + (yy * uint32_t(a.hz)) / uint32_t(a.H)) * a.wz +
                             (xx * uint32_t(a.wz)) / uint32_t(a.W);
          zyv[k] = __ldg(reinterpret_cast<const uint4*>(a.zy + zr * a.zld + c0));
          zbv[k] = __ldg(reinterpret_cast<const uint4*>(a.zb + zr * a.zld + c0));
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t r = r0 + k * stride;
      if (r >= n_rows) break;
      float f[8];
      unpack8v(xin[k], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j] - mu[j], sc[j], bt[j]);
      if (SPATIAL) {
        float zy[8], zb[8];
        unpack8v(zyv[k], zy);
        unpack8v(zbv[k], zb);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaf(bf16r(f[j]), zy[j], zb[j]);   // norm_f is a bf16 tensor upstream
      }
      if (a.silu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = bf16r(f[j]);                                        // activation input is the bf16 norm output
          f[j] = __fdividef(v, 1.0f + __expf(-v));
        }
      }
      const uint4 o = pack8v(f);
      *reinterpret_cast<uint4*>(a.y + int64_t(r) * a.C + c0) = o;
      if (a.y2 != nullptr && int64_t(r) >= a.y2_from) *reinterpret_cast<uint4*>(a.y2 + (int64_t(r) - a.y2_from) * a.C + c0) = o;
    }
  }
}

// SpatialNorm3D when the activation is an integer multiple of the latent resolution (always the case in the decoder:
// x1, x2, x4, x8): a thread keeps its channel vector AND one latent pixel, loads that pixel's (scale | bias) pair once
// and walks the RX output pixels of the row that map to it -- the generic kernel above fetched the pair again for every
// output pixel (two extra 16-byte L2 requests per 16 bytes of activation; round-2 profile: 2.4 TB/s effective).
template <int RX, class MapT>
__global__ void __launch_bounds__(256) gn_apply_spatial_rx_kernel(const GnApplyArgs a, const MapT tmap) {
  const int tpr = a.C / 8;
  const int lane_c = threadIdx.x % tpr;
  const int item_in_blk = threadIdx.x / tpr;
  const int items_per_blk = 256 / tpr;
  const int c0 = lane_c * 8;
  const int gs = a.C / a.G;
  float mu[8], sc[8], bt[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (c0 + j) / gs;
    mu[j] = a.mean_rstd[2 * g];
    sc[j] = a.mean_rstd[2 * g + 1] * __ldg(a.gamma + c0 + j);
    bt[j] = __ldg(a.beta + c0 + j);
  }
  const uint32_t ry = uint32_t(a.H / a.hz);
  const uint32_t items_per_frame = uint32_t(a.H) * a.wz;
  const uint32_t n_items = uint32_t(a.N / RX);                 // N = T * H * W and W = RX * wz
  const uint32_t stride = gridDim.x * items_per_blk;
  for (uint32_t it = blockIdx.x * items_per_blk + item_in_blk; it < n_items; it += stride) {
    const uint32_t t = it / items_per_frame, rem = it - t * items_per_frame;
    const uint32_t y = rem / uint32_t(a.wz), xz = rem - y * uint32_t(a.wz);
    const int64_t zr = (int64_t(tmap[t]) * a.hz + y / ry) * a.wz + xz;
    const int64_t row0 = (int64_t(t) * a.H + y) * a.W + int64_t(xz) * RX;
    const uint4 zyv = __ldg(reinterpret_cast<const uint4*>(a.zy + zr * a.zld + c0));
    const uint4 zbv = __ldg(reinterpret_cast<const uint4*>(a.zb + zr * a.zld + c0));
    uint4 xin[RX];
#pragma unroll
    for (int k = 0; k < RX; ++k) xin[k] = __ldg(reinterpret_cast<const uint4*>(a.x + (row0 + k) * a.C + c0));
    float zy[8], zb[8];
    unpack8v(zyv, zy);
    unpack8v(zbv, zb);
#pragma unroll
    for (int k = 0; k < RX; ++k) {
      float f[8];
      unpack8v(xin[k], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[j] = fmaf(bf16r(fmaf(f[j] - mu[j], sc[j], bt[j])), zy[j], zb[j]);      // norm_f is a bf16 tensor upstream
        if (a.silu) {
          const float v = bf16r(f[j]);
          f[j] = __fdividef(v, 1.0f + __expf(-v));
        }
      }
      const uint4 o = pack8v(f);
      *reinterpret_cast<uint4*>(a.y + (row0 + k) * a.C + c0) = o;
      if (a.y2 != nullptr && row0 + k >= a.y2_from) *reinterpret_cast<uint4*>(a.y2 + (row0 + k - a.y2_from) * a.C + c0) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// out[t', y', x', :] = in[tmap[t'], y' / sy, x' / sx, :]      nearest up-sampling (F.interpolate) incl. the
// "keep the first frame" temporal rule, which the host encodes in tmap.   Bytes: (1 + sy*sx*T'/T) * |in|.
// ------------------------------------------------------------------------------------------------
template <class MapT>
__global__ void __launch_bounds__(256)
upsample_nearest_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, const MapT tmap, int To,
                        int Ho, int Wo, int Hi, int Wi, int sy, int sx, int cvec) {
  const int64_t total = int64_t(To) * Ho * Wo * cvec;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int cv = int(i % cvec);
    int64_t r = i / cvec;
    const int x = int(r % Wo); r /= Wo;
    const int y = int(r % Ho);
    const int t = int(r / Ho);
    out[i] = __ldg(in + ((int64_t(tmap[t]) * Hi + y / sy) * Wi + x / sx) * cvec + cv);
  }
}

// out[t'] = in[a[t']] if b[t'] < 0 else avg(in[a[t']], in[b[t']])   (F.avg_pool1d(k=2,s=2) keeping frame 0; bf16)
template <class MapT>
__global__ void __launch_bounds__(256)
avgpool_time_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, const MapT ia, const MapT ib, int To,
                    int64_t frame_vecs) {
  const int64_t total = int64_t(To) * frame_vecs;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int t = int(i / frame_vecs);
    const int64_t o = i - int64_t(t) * frame_vecs;
    const uint4 va = __ldg(in + int64_t(ia[t]) * frame_vecs + o);
    if (ib[t] < 0) {
      out[i] = va;
    } else {
      float fa[8], fb[8];
      unpack8v(va, fa);
      unpack8v(__ldg(in + int64_t(ib[t]) * frame_vecs + o), fb);
#pragma unroll
      for (int j = 0; j < 8; ++j) fa[j] = (fa[j] + fb[j]) * 0.5f;
      out[i] = pack8v(fa);
    }
  }
}

// NCTHW [C, T, H, W] (one batch item, any of bf16) -> channels-last [T, H, W, Cp] with channels >= C zeroed.
__global__ void __launch_bounds__(256)
ncthw_to_thwc_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int C, int Cp,
                     int64_t thw) {
  const int64_t total = thw * Cp;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % Cp);
    const int64_t p = i / Cp;
    out[i] = c < C ? in[int64_t(c) * thw + p] : __float2bfloat16(0.f);
  }
}
// channels-last [T*H*W, Cp] -> NCTHW [C, T*H*W] (first C channels)
__global__ void __launch_bounds__(256)
thwc_to_ncthw_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int C, int Cp,
                     int64_t thw) {
  const int64_t total = thw * C;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i / thw);
    const int64_t p = i - int64_t(c) * thw;
    out[i] = in[p * Cp + c];
  }
}

// DiagonalGaussianDistribution.sample: z = mean + exp(0.5 * clamp(logvar, -30, 20)) * noise, all bf16 tensors with
// torch's roundings (each op's result is a bf16 tensor).  moments: channels-last [P, 2*L(+pad)], noise / z: NCTHW.
__global__ void __launch_bounds__(256)
posterior_sample_kernel(const __nv_bfloat16* __restrict__ moments, int Cp, int L, const __nv_bfloat16* __restrict__ noise,
                        __nv_bfloat16* __restrict__ z, int64_t P, int use_noise) {
  const int64_t total = P * L;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i / P);
    const int64_t p = i - int64_t(c) * P;
    const float mean = __bfloat162float(moments[p * Cp + c]);
    if (!use_noise) {
      z[i] = __float2bfloat16_rn(mean);
      continue;
    }
    float lv = __bfloat162float(moments[p * Cp + L + c]);
    lv = fminf(fmaxf(lv, -30.f), 20.f);
    const float stdv = bf16r(expf(bf16r(0.5f * lv)));
    z[i] = __float2bfloat16_rn(mean + bf16r(stdv * __bfloat162float(noise[i])));
  }
}

// blend_v / blend_h of the tiled VAE: b[.., k, ..] = a[.., La - extent + k, ..] * (1 - k/extent) + b[.., k, ..] * (k/extent)
// for k < extent along `axis` (1 = rows, 2 = cols) of channels-last tiles a[T, Ha, Wa, C], b[T, Hb, Wb, C]; each product
// and the sum are rounded to bf16 like the torch expression on bf16 tensors.
__global__ void __launch_bounds__(256)
tile_blend_kernel(const __nv_bfloat16* __restrict__ a, __nv_bfloat16* __restrict__ b, int T, int Ha, int Wa, int Hb,
                  int Wb, int C, int axis, int extent) {
  const int nh = axis == 1 ? extent : Hb, nw = axis == 2 ? extent : Wb;
  const int64_t total = int64_t(T) * nh * nw * C;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % C);
    int64_t r = i / C;
    const int x = int(r % nw); r /= nw;
    const int y = int(r % nh);
    const int t = int(r / nh);
    const int k = axis == 1 ? y : x;
    const int ya = axis == 1 ? Ha - extent + y : y, xa = axis == 2 ? Wa - extent + x : x;
    const float wa = 1.0f - float(double(k) / double(extent)), wb = float(double(k) / double(extent));
    const float va = __bfloat162float(a[((int64_t(t) * Ha + ya) * Wa + xa) * C + c]);
    __nv_bfloat16* pb = b + ((int64_t(t) * Hb + y) * Wb + x) * C + c;
    *pb = __float2bfloat16_rn(bf16r(va * wa) + bf16r(__bfloat162float(*pb) * wb));
  }
}

static unsigned sgrid(int64_t n) {
  int64_t g = ceil_div(n, 256);
  const int64_t cap = int64_t(num_sms()) * 16;
  return (unsigned)(g > cap ? cap : (g < 1 ? 1 : g));
}

// crop of an NCTHW tensor (element strides sC, sT, sH; unit stride along W) -> channels-last [T, H, W, Cp], channels
// >= C zeroed: the `x[:, s:e, i:i+th, j:j+tw]` slices of the tiled / frame-batched VAE without a .contiguous() copy.
__global__ void __launch_bounds__(256)
crop_ncthw_to_thwc_kernel(const __nv_bfloat16* __restrict__ in, int64_t sC, int64_t sT, int64_t sH,
                          __nv_bfloat16* __restrict__ out, int C, int Cp, int T, int H, int W) {
  const int64_t total = int64_t(T) * H * W * Cp;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % Cp);
    int64_t p = i / Cp;
    const int x = int(p % W); p /= W;
    const int y = int(p % H);
    const int t = int(p / H);
    out[i] = c < C ? in[c * sC + t * sT + y * sH + x] : __float2bfloat16(0.f);
  }
}

// dst[t, y0 + y, x0 + x, :] = src[t, y, x, :] for y < h, x < w (channels-last, 16-byte vectors): writes the kept part
// of a blended VAE tile into the assembled output (torch.cat of `tile[:, :lim_h, :lim_w]` in diffusers).
__global__ void __launch_bounds__(256)
copy_region_kernel(const uint4* __restrict__ src, int Hs, int Ws, uint4* __restrict__ dst, int Hd, int Wd, int T, int h,
                   int w, int cvec, int y0, int x0) {
  const int64_t total = int64_t(T) * h * w * cvec;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int cv = int(i % cvec);
    int64_t p = i / cvec;
    const int x = int(p % w); p /= w;
    const int y = int(p % h);
    const int t = int(p / h);
    dst[((int64_t(t) * Hd + y0 + y) * Wd + x0 + x) * cvec + cv] = __ldg(src + ((int64_t(t) * Hs + y) * Ws + x) * cvec + cv);
  }
}

// ---- C++ entry points for the handle-level executor (vae_exec.cu): index maps by value, no device-side tables
int gn_apply_imap(const void* x, void* y, int64_t N, int C, int G, const float* mean_rstd, const float* gamma,
                  const float* beta, const void* zy, const void* zb, int zld, const IMap* tmap, int H, int W, int hz,
                  int wz, int silu, void* y2, int64_t y2_from, cudaStream_t stream) {
  AETHER_CHECK_ARG(x && y && mean_rstd && gamma && beta && N > 0 && C % 8 == 0 && C % G == 0);
  AETHER_CHECK_ARG(256 % (C / 8) == 0 && N < (int64_t(1) << 31) && (C / G) % 4 == 0 && y2_from >= 0);
  GnApplyArgs a{reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y), N, C, G, mean_rstd, gamma,
                beta, reinterpret_cast<const __nv_bfloat16*>(zy), reinterpret_cast<const __nv_bfloat16*>(zb), H, W, hz, wz,
                zld > 0 ? zld : C, silu, reinterpret_cast<__nv_bfloat16*>(y2), y2_from};
  const int rows_per_blk = 256 / (C / 8);
  int64_t blocks = ceil_div(N, int64_t(rows_per_blk) * 4);
  const int64_t cap = int64_t(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  if (zy != nullptr) {
    AETHER_CHECK_ARG(zb && tmap);
    const int rx = (H % hz == 0 && W % wz == 0) ? W / wz : 0;
    int64_t iblocks = ceil_div(N / (rx > 0 ? rx : 1), int64_t(rows_per_blk));
    if (iblocks > cap) iblocks = cap;
    if (rx == 8) gn_apply_spatial_rx_kernel<8, IMap><<<(unsigned)iblocks, 256, 0, stream>>>(a, *tmap);
    else if (rx == 4) gn_apply_spatial_rx_kernel<4, IMap><<<(unsigned)iblocks, 256, 0, stream>>>(a, *tmap);
    else if (rx == 2) gn_apply_spatial_rx_kernel<2, IMap><<<(unsigned)iblocks, 256, 0, stream>>>(a, *tmap);
    else gn_apply_kernel<true, IMap><<<(unsigned)blocks, 256, 0, stream>>>(a, *tmap);
  } else {
    gn_apply_kernel<false, const int*><<<(unsigned)blocks, 256, 0, stream>>>(a, nullptr);
  }
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}
int upsample_nearest_imap(const void* in, void* out, const IMap& tmap, int To, int Ho, int Wo, int Hi, int Wi, int sy,
                          int sx, int C, cudaStream_t stream) {
  AETHER_CHECK_ARG(in && out && C % 8 == 0 && Ho == Hi * sy && Wo == Wi * sx && To <= IMap::kMax);
  upsample_nearest_kernel<IMap><<<sgrid(int64_t(To) * Ho * Wo * (C / 8)), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), tmap, To, Ho, Wo, Hi, Wi, sy, sx, C / 8);
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}
int avgpool_time_imap(const void* in, void* out, const IMap& ia, const IMap& ib, int To, int64_t frame_elems,
                      cudaStream_t stream) {
  AETHER_CHECK_ARG(in && out && frame_elems % 8 == 0 && To <= IMap::kMax);
  avgpool_time_kernel<IMap><<<sgrid(int64_t(To) * frame_elems / 8), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), ia, ib, To, frame_elems / 8);
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}
int crop_ncthw_to_thwc(const void* in, int64_t sC, int64_t sT, int64_t sH, void* out, int C, int Cp, int T, int H, int W,
                       cudaStream_t stream) {
  AETHER_CHECK_ARG(in && out && C <= Cp);
  crop_ncthw_to_thwc_kernel<<<sgrid(int64_t(T) * H * W * Cp), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(in), sC, sT, sH, reinterpret_cast<__nv_bfloat16*>(out), C, Cp, T, H, W);
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}
int copy_region_cl(const void* src, int Hs, int Ws, void* dst, int Hd, int Wd, int T, int h, int w, int C, int y0, int x0,
                   cudaStream_t stream) {
  AETHER_CHECK_ARG(src && dst && C % 8 == 0 && h <= Hs && w <= Ws && y0 + h <= Hd && x0 + w <= Wd);
  copy_region_kernel<<<sgrid(int64_t(T) * h * w * (C / 8)), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(src), Hs, Ws, reinterpret_cast<uint4*>(dst), Hd, Wd, T, h, w, C / 8, y0, x0);
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}

int64_t gn_workspace_floats(int C) { return int64_t(num_sms()) * 8 * (C / 4) * 2; }

int gn_stats(const void* x, int64_t N, int C, int G, float eps, float* workspace, float* mean_rstd,
             unsigned* counter, cudaStream_t stream) {
  AETHER_CHECK_ARG(x && workspace && mean_rstd && N > 0 && C % 8 == 0 && C <= 2048 && GN_THREADS % (C / 8) == 0 &&
                   C % G == 0 && (C / G) % 4 == 0);
  const int rows_per_blk = GN_THREADS / (C / 8);
  int nblocks = (int)ceil_div(N, rows_per_blk);
  const int cap = num_sms();           // one resident 1024-thread block per SM
  if (nblocks > cap) nblocks = cap;
  const double count = double(N) * (C / G);
  gn_partial_kernel<<<nblocks, GN_THREADS, GN_THREADS * 4 * sizeof(float), stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), N, C, workspace, G, count, eps, mean_rstd, counter);
  if (counter == nullptr) gn_finalize_kernel<<<1, GN_THREADS, 0, stream>>>(workspace, nblocks, C, G, count, eps, mean_rstd);
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}

int tile_blend(const void* a, void* b, int T, int Ha, int Wa, int Hb, int Wb, int C, int axis, int extent,
               cudaStream_t stream) {
  AETHER_CHECK_ARG(a && b && (axis == 1 || axis == 2) && extent > 0);
  AETHER_CHECK_ARG(!(axis == 1 && (extent > Ha || extent > Hb || Wa != Wb)) &&
                   !(axis == 2 && (extent > Wa || extent > Wb || Ha != Hb)));
  const int64_t n = int64_t(T) * (axis == 1 ? extent : Hb) * (axis == 2 ? extent : Wb) * C;
  tile_blend_kernel<<<sgrid(n), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(a),
                                                  reinterpret_cast<__nv_bfloat16*>(b), T, Ha, Wa, Hb, Wb, C, axis, extent);
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}

int thwc_to_ncthw(const void* in, void* out, int C, int Cp, int64_t thw, cudaStream_t stream) {
  AETHER_CHECK_ARG(in && out && C <= Cp);
  thwc_to_ncthw_kernel<<<sgrid(thw * C), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(in),
                                                           reinterpret_cast<__nv_bfloat16*>(out), C, Cp, thw);
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}

}  // namespace aether

using namespace aether;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)
extern "C" {

int64_t aether_gn_workspace_floats(int32_t C) { return gn_workspace_floats(C); }

int aether_gn_stats(const void* x, int64_t N, int32_t C, int32_t G, float eps, float* workspace, float* mean_rstd,
                    void* stream) {
  return gn_stats(x, N, C, G, eps, workspace, mean_rstd, nullptr, ST(stream));
}

int aether_gn_apply(const void* x, void* y, int64_t N, int32_t C, int32_t G, const float* mean_rstd,
                    const float* gamma, const float* beta, const void* zy, const void* zb, int32_t zld,
                    const int32_t* tmap, int32_t H, int32_t W, int32_t hz, int32_t wz, int32_t silu, void* stream) {
  if (!x || !y || !mean_rstd || !gamma || !beta || N <= 0 || C % 8 != 0 || C % G != 0) return AETHER_ERR_INVALID;
  if ((zy == nullptr) != (zb == nullptr) || (zy && (!tmap || H <= 0 || W <= 0 || hz <= 0 || wz <= 0)))
    return AETHER_ERR_INVALID;
  if (256 % (C / 8) != 0 || N >= (int64_t(1) << 31) || (C / G) % 4 != 0) return AETHER_ERR_INVALID;
  GnApplyArgs a{CBF(x), BF(y), N, C, G, mean_rstd, gamma, beta, CBF(zy), CBF(zb), H, W, hz, wz, zld > 0 ? zld : C, silu,
                nullptr, 0};
  const int rows_per_blk = 256 / (C / 8);
  int64_t blocks = ceil_div(N, int64_t(rows_per_blk) * 4);
  const int64_t cap = int64_t(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  if (zy != nullptr)
    gn_apply_kernel<true, const int*><<<(unsigned)blocks, 256, 0, ST(stream)>>>(a, tmap);
  else
    gn_apply_kernel<false, const int*><<<(unsigned)blocks, 256, 0, ST(stream)>>>(a, nullptr);
  return cudaGetLastError() == cudaSuccess ? AETHER_OK : AETHER_ERR_CUDA;
}

int aether_upsample_nearest(const void* in, void* out, const int32_t* tmap, int32_t To, int32_t Ho, int32_t Wo,
                            int32_t Hi, int32_t Wi, int32_t sy, int32_t sx, int32_t C, void* stream) {
  if (!in || !out || !tmap || C % 8 != 0 || Ho != Hi * sy || Wo != Wi * sx) return AETHER_ERR_INVALID;
  upsample_nearest_kernel<const int*><<<sgrid(int64_t(To) * Ho * Wo * (C / 8)), 256, 0, ST(stream)>>>(
      reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), tmap, To, Ho, Wo, Hi, Wi, sy, sx, C / 8);
  return cudaGetLastError() == cudaSuccess ? AETHER_OK : AETHER_ERR_CUDA;
}

int aether_avgpool_time(const void* in, void* out, const int32_t* ia, const int32_t* ib, int32_t To,
                        int64_t frame_elems, void* stream) {
  if (!in || !out || !ia || !ib || frame_elems % 8 != 0) return AETHER_ERR_INVALID;
  avgpool_time_kernel<const int*><<<sgrid(int64_t(To) * frame_elems / 8), 256, 0, ST(stream)>>>(
      reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), ia, ib, To, frame_elems / 8);
  return cudaGetLastError() == cudaSuccess ? AETHER_OK : AETHER_ERR_CUDA;
}

int aether_ncthw_to_thwc(const void* in, void* out, int32_t C, int32_t Cp, int64_t thw, void* stream) {
  if (!in || !out || C > Cp) return AETHER_ERR_INVALID;
  ncthw_to_thwc_kernel<<<sgrid(thw * Cp), 256, 0, ST(stream)>>>(CBF(in), BF(out), C, Cp, thw);
  return cudaGetLastError() == cudaSuccess ? AETHER_OK : AETHER_ERR_CUDA;
}
int aether_thwc_to_ncthw(const void* in, void* out, int32_t C, int32_t Cp, int64_t thw, void* stream) {
  return thwc_to_ncthw(in, out, C, Cp, thw, ST(stream));
}
int aether_posterior_sample(const void* moments, int32_t Cp, int32_t L, const void* noise, void* z, int64_t P,
                            void* stream) {
  if (!moments || !z || 2 * L > Cp) return AETHER_ERR_INVALID;
  posterior_sample_kernel<<<sgrid(P * L), 256, 0, ST(stream)>>>(CBF(moments), Cp, L, CBF(noise), BF(z), P,
                                                                noise != nullptr);
  return cudaGetLastError() == cudaSuccess ? AETHER_OK : AETHER_ERR_CUDA;
}
int aether_tile_blend(const void* a, void* b, int32_t T, int32_t Ha, int32_t Wa, int32_t Hb, int32_t Wb, int32_t C,
                      int32_t axis, int32_t extent, void* stream) {
  return tile_blend(a, b, T, Ha, Wa, Hb, Wb, C, axis, extent, ST(stream));
}
}
