// K8: classifier-free-guidance combine + v->x0 + SDE-DPM-Solver++(2M) update (+ bf16 cast), one pass.
//   replaces  aether/pipelines/aetherv1_pipeline_cogvideox.py:876 (.float()), :895-899 (CFG), :907-915
//   (CogVideoXDPMScheduler.step, third-party diffusers) and :916 (.to(bf16)).
// HBM-bound streaming kernel (grid = k * #SM); algorithmic bytes per element: n_cfg*2 (model out) + 2 (sample)
// + 2 (noise) [+ 4 old_x0] read, 2 (bf16 prev) + 4 (x0) written.
#include "host_util.h"
#include "ptx.cuh"

namespace aether {

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// Rounding contract (must mirror torch type promotion in the reference, see oracle/scheduler.py):
//   v      = fp32(model_out)                         ; CFG: u + g * (c - u) in fp32, no FMA contraction
//   x0     = bf16(sqrt_a * sample) - sqrt_1ma * v    ; fp32
//   prev   = (bf16(m1 * sample) - m2 * D) + bf16(m_noise * noise)     with D = x0 or (m3*x0 - m4*old_x0)
struct StepArgs {
  const void* model_out;
  int model_out_fp32, n_cfg;
  float guidance;
  const __nv_bfloat16* sample;
  const float* old_x0;
  const __nv_bfloat16* noise1;
  const __nv_bfloat16* noise2;
  AetherDpmCoeffs c;
  __nv_bfloat16* prev_bf16;
  float* prev_f32;
  float* x0_f32;
  int64_t N;
};

// One element of the update (shared by the scalar and the 8-wide kernel so the rounding sequence exists once).
struct StepOut {
  float prev, x0;
};
__device__ __forceinline__ StepOut step_one(const StepArgs& a, float v, float x, float old_x0, float nz) {
  float x0;
  if (a.c.prediction_type == 0) {
    x0 = __fsub_rn(bf16_round(__fmul_rn(a.c.sqrt_alpha, x)), __fmul_rn(a.c.sqrt_one_minus_alpha, v));
  } else {
    // (sample - sqrt(1-a) * eps) / sqrt(a): bf16 - fp32 -> fp32, then / fp64 scalar -> fp32 division
    x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(a.c.sqrt_one_minus_alpha, v)), a.c.sqrt_alpha);
  }
  float d = x0;
  if (a.c.second_order) d = __fsub_rn(__fmul_rn(a.c.m3, x0), __fmul_rn(a.c.m4, old_x0));
  const float t1 = bf16_round(__fmul_rn(a.c.m1, x));
  const float t3 = bf16_round(__fmul_rn(a.c.m_noise, nz));
  return {__fadd_rn(__fsub_rn(t1, __fmul_rn(a.c.m2, d)), t3), x0};
}
__device__ __forceinline__ float cfg_combine(float u, float cnd, float g) {
  return __fadd_rn(u, __fmul_rn(g, __fsub_rn(cnd, u)));
}

// Scalar kernel: any N / alignment (also the reference for the vector kernel in the tests).
__global__ void __launch_bounds__(256) cfg_dpm_step_kernel(const StepArgs a) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < a.N; i += stride) {
    float v;
    if (a.model_out_fp32) {
      const float* m = reinterpret_cast<const float*>(a.model_out);
      v = a.n_cfg == 2 ? cfg_combine(m[i], m[a.N + i], a.guidance) : m[i];
    } else {
      const __nv_bfloat16* m = reinterpret_cast<const __nv_bfloat16*>(a.model_out);
      v = a.n_cfg == 2 ? cfg_combine(__bfloat162float(m[i]), __bfloat162float(m[a.N + i]), a.guidance)
                       : __bfloat162float(m[i]);
    }
    const __nv_bfloat16* nz = a.c.second_order ? a.noise2 : a.noise1;
    const StepOut o = step_one(a, v, __bfloat162float(a.sample[i]), a.c.second_order ? a.old_x0[i] : 0.f,
                               __bfloat162float(nz[i]));
    if (a.prev_f32) a.prev_f32[i] = o.prev;
    if (a.prev_bf16) a.prev_bf16[i] = __float2bfloat16_rn(o.prev);
    a.x0_f32[i] = o.x0;
  }
}

// 8 elements per thread: every bf16 stream is one 16-byte load / store, every fp32 stream two; all loads of a
// thread are issued before the first use (the fused loop step launches ~13 k threads x 5-7 independent 16-byte
// requests per SM).  Requires N % 8 == 0 and 16-byte aligned pointers (checked by the launcher); same arithmetic.
__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = bf16_lo(w[i]);
    f[2 * i + 1] = bf16_hi(w[i]);
  }
}
template <bool MODEL_F32, int NCFG, bool SECOND>
__global__ void __launch_bounds__(256) cfg_dpm_step_vec8_kernel(const StepArgs a) {
  const int64_t n8 = a.N >> 3;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  const __nv_bfloat16* nzp = SECOND ? a.noise2 : a.noise1;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float u[8], cnd[8], x[8], nz[8], old[8];
    const uint4 qs = __ldg(reinterpret_cast<const uint4*>(a.sample) + i);
    const uint4 qn = __ldg(reinterpret_cast<const uint4*>(nzp) + i);
    if (MODEL_F32) {
      const float4* m = reinterpret_cast<const float4*>(a.model_out);
      const float4 u0 = __ldg(m + 2 * i), u1 = __ldg(m + 2 * i + 1);
      u[0] = u0.x; u[1] = u0.y; u[2] = u0.z; u[3] = u0.w; u[4] = u1.x; u[5] = u1.y; u[6] = u1.z; u[7] = u1.w;
      if (NCFG == 2) {
        const float4 c0 = __ldg(m + 2 * (n8 + i)), c1 = __ldg(m + 2 * (n8 + i) + 1);
        cnd[0] = c0.x; cnd[1] = c0.y; cnd[2] = c0.z; cnd[3] = c0.w; cnd[4] = c1.x; cnd[5] = c1.y; cnd[6] = c1.z; cnd[7] = c1.w;
      }
    } else {
      const uint4* m = reinterpret_cast<const uint4*>(a.model_out);
      const uint4 qu = __ldg(m + i);
      if (NCFG == 2) {
        const uint4 qc = __ldg(m + n8 + i);
        unpack8(qc, cnd);
      }
      unpack8(qu, u);
    }
    if (SECOND) {
      const float4 o0 = __ldg(reinterpret_cast<const float4*>(a.old_x0) + 2 * i);
      const float4 o1 = __ldg(reinterpret_cast<const float4*>(a.old_x0) + 2 * i + 1);
      old[0] = o0.x; old[1] = o0.y; old[2] = o0.z; old[3] = o0.w; old[4] = o1.x; old[5] = o1.y; old[6] = o1.z; old[7] = o1.w;
    }
    unpack8(qs, x);
    unpack8(qn, nz);
    float prev[8], x0[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float v = NCFG == 2 ? cfg_combine(u[e], cnd[e], a.guidance) : u[e];
      const StepOut o = step_one(a, v, x[e], SECOND ? old[e] : 0.f, nz[e]);
      prev[e] = o.prev;
      x0[e] = o.x0;
    }
    float4* xo = reinterpret_cast<float4*>(a.x0_f32) + 2 * i;
    xo[0] = make_float4(x0[0], x0[1], x0[2], x0[3]);
    xo[1] = make_float4(x0[4], x0[5], x0[6], x0[7]);
    if (a.prev_f32) {
      float4* po = reinterpret_cast<float4*>(a.prev_f32) + 2 * i;
      po[0] = make_float4(prev[0], prev[1], prev[2], prev[3]);
      po[1] = make_float4(prev[4], prev[5], prev[6], prev[7]);
    }
    if (a.prev_bf16) {
      uint4 w;
      w.x = pack_bf16x2(prev[0], prev[1]);
      w.y = pack_bf16x2(prev[2], prev[3]);
      w.z = pack_bf16x2(prev[4], prev[5]);
      w.w = pack_bf16x2(prev[6], prev[7]);
      reinterpret_cast<uint4*>(a.prev_bf16)[i] = w;
    }
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int cfg_dpm_step(const StepArgs& a, cudaStream_t stream) {
  AETHER_CHECK_ARG(a.N > 0 && a.model_out && a.sample && a.noise1 && a.x0_f32);
  AETHER_CHECK_ARG(a.n_cfg == 1 || a.n_cfg == 2);
  AETHER_CHECK_ARG(!a.c.second_order || (a.old_x0 && a.noise2));
  const bool vec = a.N % 8 == 0 && aligned16(a.model_out) && aligned16(a.sample) && aligned16(a.noise1) &&
                   aligned16(a.noise2) && aligned16(a.old_x0) && aligned16(a.prev_bf16) && aligned16(a.prev_f32) &&
                   aligned16(a.x0_f32);
  const int64_t items = vec ? a.N / 8 : a.N;
  int64_t grid = ceil_div(items, 256);
  const int64_t cap = int64_t(num_sms()) * 8;
  if (grid > cap) grid = cap;
  if (!vec) {
    cfg_dpm_step_kernel<<<(unsigned)grid, 256, 0, stream>>>(a);
  } else {
    const int sel = (a.model_out_fp32 ? 4 : 0) | (a.n_cfg == 2 ? 2 : 0) | (a.c.second_order ? 1 : 0);
#define AETHER_STEP_CASE(id, F32, NC, SO) \
  case id: cfg_dpm_step_vec8_kernel<F32, NC, SO><<<(unsigned)grid, 256, 0, stream>>>(a); break;
    switch (sel) {
      AETHER_STEP_CASE(0, false, 1, false)
      AETHER_STEP_CASE(1, false, 1, true)
      AETHER_STEP_CASE(2, false, 2, false)
      AETHER_STEP_CASE(3, false, 2, true)
      AETHER_STEP_CASE(4, true, 1, false)
      AETHER_STEP_CASE(5, true, 1, true)
      AETHER_STEP_CASE(6, true, 2, false)
      AETHER_STEP_CASE(7, true, 2, true)
    }
#undef AETHER_STEP_CASE
  }
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}

}  // namespace aether

using namespace aether;
extern "C" {
int aether_cfg_dpm_step(const void* model_out, int32_t model_out_fp32, int32_t n_cfg, float guidance,
                        const void* sample, const float* old_x0, const void* noise1, const void* noise2,
                        const AetherDpmCoeffs* c, void* prev_bf16, float* prev_f32, float* x0_f32, int64_t N,
                        void* stream) {
  if (c == nullptr) return AETHER_ERR_INVALID;
  StepArgs a;
  a.model_out = model_out; a.model_out_fp32 = model_out_fp32; a.n_cfg = n_cfg; a.guidance = guidance;
  a.sample = reinterpret_cast<const __nv_bfloat16*>(sample);
  a.old_x0 = old_x0;
  a.noise1 = reinterpret_cast<const __nv_bfloat16*>(noise1);
  a.noise2 = reinterpret_cast<const __nv_bfloat16*>(noise2);
  a.c = *c;
  a.prev_bf16 = reinterpret_cast<__nv_bfloat16*>(prev_bf16);
  a.prev_f32 = prev_f32; a.x0_f32 = x0_f32; a.N = N;
  return cfg_dpm_step(a, reinterpret_cast<cudaStream_t>(stream));
}

int32_t aether_abi_version(void) { return AETHER_ABI_VERSION; }
int32_t aether_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return 0;
  }
  int dev = 0, major = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  return major == 10 ? 1 : 0;
}
}
