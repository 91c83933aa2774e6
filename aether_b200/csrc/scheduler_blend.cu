// K8: classifier-free-guidance combine + v->x0 + SDE-DPM-Solver++(2M) update (+ bf16 cast), one pass.
//   replaces  aether/pipelines/aetherv1_pipeline_cogvideox.py:876 (.float()), :895-899 (CFG), :907-915
//   (CogVideoXDPMScheduler.step, third-party diffusers) and :916 (.to(bf16)).
// K10: masked-LSQ scale reduction + linear cross-fade of the sliding-window blend
//   replaces  aether/utils/postprocess_utils.py:847-864 (compute_scale) and
//             evaluation/video_depth/launch_aether.py:166-285 (spatial / temporal blend chain).
// Both are HBM-bound streaming kernels (16-byte vectors where alignment allows, grid = k * #SM).
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

// ------------------------------------------------------------------------------------------------ K10
// compute_scale: numerator = sum(mask*p*t), denominator = sum(mask*p*p) with mask == 1 (the reference passes
// np.ones_like, launch_aether.py:195, :272).  Inputs may be fp32 (a raw pipeline window) or fp64 (an already
// blended np.ones(float64) buffer, launch_aether.py:211/:276); both are cast to fp32 first exactly like
// `torch.from_numpy(x).float()` (postprocess_utils.py:848-851) and the products are formed in fp32.
// The reference then reduces in fp32 (torch.sum); we accumulate the fp32 products in fp64 (deterministic
// grid-stride order + fp64 atomics), which is at least as accurate; the tolerance on `scale` is stated in
// tests/test_blend.py (rel 1e-6).
__device__ __forceinline__ float load_as_f32(const void* p, int is_f64, int64_t idx) {
  return is_f64 ? static_cast<float>(reinterpret_cast<const double*>(p)[idx]) : reinterpret_cast<const float*>(p)[idx];
}

__global__ void __launch_bounds__(256)
scale_reduce_kernel(const void* __restrict__ pred, int pred_f64, int64_t pred_rs, const void* __restrict__ target,
                    int target_f64, int64_t t_rs, int64_t rows, int64_t cols, double* out) {
  double num = 0.0, den = 0.0;
  const int64_t n = rows * cols;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t r = i / cols, c = i - r * cols;
    const float p = load_as_f32(pred, pred_f64, r * pred_rs + c);
    const float t = load_as_f32(target, target_f64, r * t_rs + c);
    num += double(__fmul_rn(p, t));
    den += double(__fmul_rn(p, p));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    num += __shfl_xor_sync(0xffffffffu, num, o);
    den += __shfl_xor_sync(0xffffffffu, den, o);
  }
  __shared__ double snum[8], sden[8];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { snum[w] = num; sden[w] = den; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < 8; ++i) { a += snum[i]; b += sden[i]; }
    atomicAdd(out, a);
    atomicAdd(out + 1, b);
  }
}

// `scale * window` follows numpy promotion: python-float * float32 array -> float32 (NEP 50 weak scalar);
// python-float * float64 array -> float64.
__device__ __forceinline__ double aligned_value(const void* win, int win_f64, int64_t idx, double scale) {
  if (win_f64) return scale * reinterpret_cast<const double*>(win)[idx];
  return double(__fmul_rn(static_cast<float>(scale), reinterpret_cast<const float*>(win)[idx]));
}

// dst = acc * w + aligned * (1 - w),  w = linspace(1, 0, n)[k]  (numpy: start + k * step with step = -1/(n-1),
// last element forced to `stop`), in fp64 like the reference's float64 `result` buffer.  `acc` is the
// previous accumulation (fp32 for the very first window, fp64 afterwards); dst is always fp64.
__global__ void __launch_bounds__(256)
blend_crossfade_kernel(double* __restrict__ dst, int64_t dst_rs, const void* __restrict__ acc, int acc_f64,
                       int64_t acc_rs, const void* __restrict__ win, int win_f64, int64_t win_rs, double scale,
                       int64_t rows, int64_t cols, int64_t inner, int axis_outer, int64_t nw) {
  // logical index space: rows x cols where the weight index k = (axis_outer ? r / inner : c)
  const int64_t n = rows * cols;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  const double step = nw > 1 ? (0.0 - 1.0) / double(nw - 1) : 0.0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t r = i / cols, c = i - r * cols;
    const int64_t k = axis_outer ? r / inner : c;
    const double w = (nw > 1 && k == nw - 1) ? 0.0 : 1.0 + double(k) * step;
    const double a = acc_f64 ? reinterpret_cast<const double*>(acc)[r * acc_rs + c]
                             : double(reinterpret_cast<const float*>(acc)[r * acc_rs + c]);
    dst[r * dst_rs + c] = a * w + aligned_value(win, win_f64, r * win_rs + c, scale) * (1.0 - w);
  }
}

// dst (fp64) = scale * src   (scale == 1 and fp32/fp64 src gives the plain widening copy of the untouched part)
__global__ void __launch_bounds__(256)
scale_copy_kernel(double* __restrict__ dst, int64_t dst_rs, const void* __restrict__ src, int src_f64,
                  int64_t src_rs, double scale, int apply_scale, int64_t rows, int64_t cols) {
  const int64_t n = rows * cols;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t r = i / cols, c = i - r * cols;
    const int64_t si = r * src_rs + c;
    double v;
    if (apply_scale) v = aligned_value(src, src_f64, si, scale);
    else v = src_f64 ? reinterpret_cast<const double*>(src)[si] : double(reinterpret_cast<const float*>(src)[si]);
    dst[r * dst_rs + c] = v;
  }
}

static unsigned stream_grid(int64_t n) {
  int64_t g = ceil_div(n, 256);
  const int64_t cap = int64_t(num_sms()) * 8;
  return (unsigned)(g > cap ? cap : g);
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

int aether_scale_reduce(const void* pred, int32_t pred_is_f64, int64_t pred_rs, const void* target,
                        int32_t target_is_f64, int64_t target_rs, int64_t rows, int64_t cols, double* out,
                        void* stream) {
  if (!pred || !target || !out || rows <= 0 || cols <= 0) return AETHER_ERR_INVALID;
  scale_reduce_kernel<<<stream_grid(rows * cols), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      pred, pred_is_f64, pred_rs, target, target_is_f64, target_rs, rows, cols, out);
  return cudaGetLastError() == cudaSuccess ? AETHER_OK : AETHER_ERR_CUDA;
}
int aether_blend_crossfade(double* dst, int64_t dst_rs, const void* acc, int32_t acc_is_f64, int64_t acc_rs,
                           const void* win, int32_t win_is_f64, int64_t win_rs, double scale, int64_t rows,
                           int64_t cols, int64_t inner, int32_t axis_outer, int64_t n_weights, void* stream) {
  if (!dst || !acc || !win || rows <= 0 || cols <= 0 || inner <= 0) return AETHER_ERR_INVALID;
  if (n_weights != (axis_outer ? rows / inner : cols)) return AETHER_ERR_INVALID;
  blend_crossfade_kernel<<<stream_grid(rows * cols), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      dst, dst_rs, acc, acc_is_f64, acc_rs, win, win_is_f64, win_rs, scale, rows, cols, inner, axis_outer, n_weights);
  return cudaGetLastError() == cudaSuccess ? AETHER_OK : AETHER_ERR_CUDA;
}
int aether_scale_copy(double* dst, int64_t dst_rs, const void* src, int32_t src_is_f64, int64_t src_rs, double scale,
                      int32_t apply_scale, int64_t rows, int64_t cols, void* stream) {
  if (!dst || !src || rows <= 0 || cols <= 0) return AETHER_ERR_INVALID;
  scale_copy_kernel<<<stream_grid(rows * cols), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      dst, dst_rs, src, src_is_f64, src_rs, scale, apply_scale, rows, cols);
  return cudaGetLastError() == cudaSuccess ? AETHER_OK : AETHER_ERR_CUDA;
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
