// K10: the arithmetic of the sliding-window blend chain on the device.
//   replaces  aether/utils/postprocess_utils.py:847-864 (compute_scale)  and
//             evaluation/video_depth/launch_aether.py:166-252 (spatial blend) / :259-285 (temporal blend).
//
// All kernels work on a 3-D region [n0, n1, n2] (frames x rows x cols, unit stride on the last axis) of
// disparity buffers that are fp32 (a raw pipeline window, pipeline :940) or fp64 (an already blended
// accumulation: the reference allocates `np.ones(result_shape)` = float64, launch_aether.py:211/:276).
// HBM-bound streaming kernels: grid-stride, grid = 8 x #SM; algorithmic bytes = region elements x
// (4|8 read per source + 8 written).
#include "host_util.h"

namespace aether {

struct View3 {            // element strides of axis 0 and 1 (axis 2 is contiguous)
  const void* p;
  int is_f64;
  int64_t s0, s1;
};

__device__ __forceinline__ float ld_f32(const View3& v, int64_t off) {
  return v.is_f64 ? static_cast<float>(reinterpret_cast<const double*>(v.p)[off])
                  : reinterpret_cast<const float*>(v.p)[off];
}
__device__ __forceinline__ double ld_f64(const View3& v, int64_t off) {
  return v.is_f64 ? reinterpret_cast<const double*>(v.p)[off] : double(reinterpret_cast<const float*>(v.p)[off]);
}
// `scale * window` follows numpy promotion (NEP 50): python-float * float32 array -> float32 product;
// python-float * float64 array -> float64 product.          (launch_aether.py:204, :274)
__device__ __forceinline__ double aligned(const View3& v, int64_t off, double scale) {
  if (v.is_f64) return scale * reinterpret_cast<const double*>(v.p)[off];
  return double(__fmul_rn(static_cast<float>(scale), reinterpret_cast<const float*>(v.p)[off]));
}

// compute_scale with mask == 1 (np.ones_like, launch_aether.py:195/:272): both operands are cast to fp32
// (`torch.from_numpy(x).float()`, postprocess_utils.py:848-851), the products are fp32; the reference sums
// them in fp32 (torch.sum), we accumulate the same fp32 products in fp64 (grid-stride order + fp64 atomics),
// which is at least as accurate.  Tolerance on the resulting scale: rel 1e-6 (tests/test_blend*.py).
__global__ void __launch_bounds__(256)
scale_reduce_kernel(View3 pred, View3 target, int64_t n0, int64_t n1, int64_t n2, double* out) {
  double num = 0.0, den = 0.0;
  const int64_t n = n0 * n1 * n2;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t i2 = i % n2, r = i / n2, i1 = r % n1, i0 = r / n1;
    const float p = ld_f32(pred, i0 * pred.s0 + i1 * pred.s1 + i2);
    const float t = ld_f32(target, i0 * target.s0 + i1 * target.s1 + i2);
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

// dst = acc * w + (scale*win) * (1 - w),  w = np.linspace(1, 0, n)[k], k = index along `axis`.
// numpy linspace: start + k*step with step = (stop-start)/(n-1), last sample forced to `stop`; n == 1 -> [1.0].
__global__ void __launch_bounds__(256)
blend_crossfade_kernel(double* __restrict__ dst, int64_t d0, int64_t d1, View3 acc, View3 win, double scale,
                       int64_t n0, int64_t n1, int64_t n2, int axis) {
  const int64_t n = n0 * n1 * n2;
  const int64_t nw = axis == 0 ? n0 : (axis == 1 ? n1 : n2);
  const double step = nw > 1 ? (0.0 - 1.0) / double(nw - 1) : 0.0;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t i2 = i % n2, r = i / n2, i1 = r % n1, i0 = r / n1;
    const int64_t k = axis == 0 ? i0 : (axis == 1 ? i1 : i2);
    const double w = (nw > 1 && k == nw - 1) ? 0.0 : 1.0 + double(k) * step;
    const double a = ld_f64(acc, i0 * acc.s0 + i1 * acc.s1 + i2);
    const double b = aligned(win, i0 * win.s0 + i1 * win.s1 + i2, scale);
    dst[i0 * d0 + i1 * d1 + i2] = a * w + b * (1.0 - w);
  }
}

// dst (fp64) = apply_scale ? scale*src (numpy promotion as above) : src
__global__ void __launch_bounds__(256)
scale_copy_kernel(double* __restrict__ dst, int64_t d0, int64_t d1, View3 src, double scale, int apply_scale,
                  int64_t n0, int64_t n1, int64_t n2) {
  const int64_t n = n0 * n1 * n2;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t i2 = i % n2, r = i / n2, i1 = r % n1, i0 = r / n1;
    const int64_t so = i0 * src.s0 + i1 * src.s1 + i2;
    dst[i0 * d0 + i1 * d1 + i2] = apply_scale ? aligned(src, so, scale) : ld_f64(src, so);
  }
}

static unsigned stream_grid(int64_t n) {
  int64_t g = ceil_div(n, 256);
  const int64_t cap = int64_t(num_sms()) * 8;
  return (unsigned)(g > cap ? cap : g);
}

}  // namespace aether

using namespace aether;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
extern "C" {
int aether_scale_reduce(const void* pred, int32_t pred_is_f64, int64_t pred_s0, int64_t pred_s1, const void* target,
                        int32_t target_is_f64, int64_t target_s0, int64_t target_s1, int64_t n0, int64_t n1,
                        int64_t n2, double* out, void* stream) {
  if (!pred || !target || !out || n0 <= 0 || n1 <= 0 || n2 <= 0) return AETHER_ERR_INVALID;
  View3 p{pred, pred_is_f64, pred_s0, pred_s1}, t{target, target_is_f64, target_s0, target_s1};
  scale_reduce_kernel<<<stream_grid(n0 * n1 * n2), 256, 0, ST(stream)>>>(p, t, n0, n1, n2, out);
  return cudaGetLastError() == cudaSuccess ? AETHER_OK : AETHER_ERR_CUDA;
}
int aether_blend_crossfade(double* dst, int64_t dst_s0, int64_t dst_s1, const void* acc, int32_t acc_is_f64,
                           int64_t acc_s0, int64_t acc_s1, const void* win, int32_t win_is_f64, int64_t win_s0,
                           int64_t win_s1, double scale, int64_t n0, int64_t n1, int64_t n2, int32_t axis,
                           void* stream) {
  if (!dst || !acc || !win || n0 <= 0 || n1 <= 0 || n2 <= 0 || axis < 0 || axis > 2) return AETHER_ERR_INVALID;
  View3 a{acc, acc_is_f64, acc_s0, acc_s1}, w{win, win_is_f64, win_s0, win_s1};
  blend_crossfade_kernel<<<stream_grid(n0 * n1 * n2), 256, 0, ST(stream)>>>(dst, dst_s0, dst_s1, a, w, scale, n0, n1,
                                                                            n2, axis);
  return cudaGetLastError() == cudaSuccess ? AETHER_OK : AETHER_ERR_CUDA;
}
int aether_scale_copy(double* dst, int64_t dst_s0, int64_t dst_s1, const void* src, int32_t src_is_f64, int64_t src_s0,
                      int64_t src_s1, double scale, int32_t apply_scale, int64_t n0, int64_t n1, int64_t n2,
                      void* stream) {
  if (!dst || !src || n0 <= 0 || n1 <= 0 || n2 <= 0) return AETHER_ERR_INVALID;
  View3 s{src, src_is_f64, src_s0, src_s1};
  scale_copy_kernel<<<stream_grid(n0 * n1 * n2), 256, 0, ST(stream)>>>(dst, dst_s0, dst_s1, s, scale, apply_scale, n0,
                                                                       n1, n2);
  return cudaGetLastError() == cudaSuccess ? AETHER_OK : AETHER_ERR_CUDA;
}
}
