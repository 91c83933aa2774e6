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

__global__ void __launch_bounds__(256) cfg_dpm_step_kernel(const StepArgs a) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < a.N; i += stride) {
    float v;
    if (a.model_out_fp32) {
      const float* m = reinterpret_cast<const float*>(a.model_out);
      if (a.n_cfg == 2) {
        const float u = m[i], cnd = m[a.N + i];
        v = __fadd_rn(u, __fmul_rn(a.guidance, __fsub_rn(cnd, u)));
      } else {
        v = m[i];
      }
    } else {
      const __nv_bfloat16* m = reinterpret_cast<const __nv_bfloat16*>(a.model_out);
      if (a.n_cfg == 2) {
        const float u = __bfloat162float(m[i]), cnd = __bfloat162float(m[a.N + i]);
        v = __fadd_rn(u, __fmul_rn(a.guidance, __fsub_rn(cnd, u)));
      } else {
        v = __bfloat162float(m[i]);
      }
    }
    const float x = __bfloat162float(a.sample[i]);
    float x0;
    if (a.c.prediction_type == 0) {
      x0 = __fsub_rn(bf16_round(__fmul_rn(a.c.sqrt_alpha, x)), __fmul_rn(a.c.sqrt_one_minus_alpha, v));
    } else {
      // (sample - sqrt(1-a) * eps) / sqrt(a): bf16 - fp32 -> fp32, then / fp64 scalar -> fp32 division
      x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(a.c.sqrt_one_minus_alpha, v)), a.c.sqrt_alpha);
    }
    float d = x0;
    const __nv_bfloat16* nz = a.noise1;
    if (a.c.second_order) {
      d = __fsub_rn(__fmul_rn(a.c.m3, x0), __fmul_rn(a.c.m4, a.old_x0[i]));
      nz = a.noise2;
    }
    const float t1 = bf16_round(__fmul_rn(a.c.m1, x));
    const float t3 = bf16_round(__fmul_rn(a.c.m_noise, __bfloat162float(nz[i])));
    const float prev = __fadd_rn(__fsub_rn(t1, __fmul_rn(a.c.m2, d)), t3);
    if (a.prev_f32) a.prev_f32[i] = prev;
    if (a.prev_bf16) a.prev_bf16[i] = __float2bfloat16_rn(prev);
    a.x0_f32[i] = x0;
  }
}

int cfg_dpm_step(const StepArgs& a, cudaStream_t stream) {
  AETHER_CHECK_ARG(a.N > 0 && a.model_out && a.sample && a.noise1 && a.x0_f32);
  AETHER_CHECK_ARG(a.n_cfg == 1 || a.n_cfg == 2);
  AETHER_CHECK_ARG(!a.c.second_order || (a.old_x0 && a.noise2));
  int64_t grid = ceil_div(a.N, 256);
  const int64_t cap = int64_t(num_sms()) * 8;
  if (grid > cap) grid = cap;
  cfg_dpm_step_kernel<<<(unsigned)grid, 256, 0, stream>>>(a);
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
