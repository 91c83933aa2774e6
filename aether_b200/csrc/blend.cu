// K10: the arithmetic of the sliding-window blend chain on the device.
//   replaces  aether/utils/postprocess_utils.py:847-864 (compute_scale)  and
//             evaluation/video_depth/launch_aether.py:166-252 (spatial blend) / :259-285 (temporal blend).
//
// All kernels work on a 3-D region [n0, n1, n2] (frames x rows x cols, unit stride on the last axis) of
// disparity buffers that are fp32 (a raw pipeline window, pipeline :940) or fp64 (an already blended
// accumulation: the reference allocates `np.ones(result_shape)` = float64, launch_aether.py:211/:276).
//
// HBM-bound streaming kernels.  Layout of the work: a block walks whole ROWS (i0, i1) of the region in a
// grid-stride loop and its threads walk the contiguous columns, so the only integer division is one per row
// (round 1 did two int64 div/mod pairs per ELEMENT and reached 40 % of the copy bandwidth); every warp access is
// a run of consecutive 4- or 8-byte elements (rows of 853 columns are not 16-byte aligned, so wider vectors would
// need a peel loop for nothing: the run is already one 128/256-byte transaction per warp); each thread issues the
// loads of four column chunks (one 853-column row per pass) before it consumes any, which with eight resident blocks
// per SM keeps ~100 KB in flight per SM to cover the HBM latency.  Algorithmic bytes = region elements x (4|8 read per source + 8 written).
//
// The alignment scale never visits the host: aether_scale_reduce leaves {sum(p*t), sum(p*p)} in a device buffer
// and the cross-fade / scaled copy derive `scale` from it (fp32 quotient like the reference's torch.sum(...) /
// torch.sum(...).item()), so a whole 60-window chain is enqueued without a single synchronisation
// (round 1: one .tolist() per window).
#include "host_util.h"

namespace aether {

struct View3 {            // element strides of axis 0 and 1 (axis 2 is contiguous)
  const void* p;
  int is_f64;
  int64_t s0, s1;
};

constexpr int kBlendThreads = 256;
constexpr int kMaxReduceBlocks = 2048;
constexpr int kUnroll = 4;          // column chunks of a row whose loads are issued before any is consumed

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
// compute_scale's tail (postprocess_utils.py:853-864): both sums are fp32 tensors, the quotient is an fp32 division,
// `.item()` widens it to a python float; a zero denominator gives 0.
__device__ __forceinline__ double scale_from_sums(const double* sums) {
  const float num = static_cast<float>(sums[0]), den = static_cast<float>(sums[1]);
  return den != 0.f ? double(__fdiv_rn(num, den)) : 0.0;
}

// compute_scale with mask == 1 (np.ones_like, launch_aether.py:195/:272): both operands are cast to fp32
// (`torch.from_numpy(x).float()`, postprocess_utils.py:848-851), the products are fp32; the reference sums
// them in fp32 (torch.sum), we add the same fp32 products four at a time in fp32 and accumulate those in fp64, which is
// at least as accurate.
// Deterministic: per-block partials land in `work`, the last block to finish adds them in block order.
// Tolerance on the resulting scale vs the reference: rel 1e-6 (tests).
__global__ void __launch_bounds__(kBlendThreads)
scale_reduce_kernel(View3 pred, View3 target, int64_t n0, int64_t n1, int64_t n2, double* sums, double* partial,
                    unsigned int* counter) {
  double num = 0.0, den = 0.0;
  const int64_t rows = n0 * n1;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const int64_t i0 = r / n1, i1 = r - i0 * n1;
    const int64_t po = i0 * pred.s0 + i1 * pred.s1, to = i0 * target.s0 + i1 * target.s1;
    for (int64_t c0 = threadIdx.x; c0 < n2; c0 += kUnroll * kBlendThreads) {
      float p[kUnroll], t[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int64_t c = c0 + u * kBlendThreads;
        p[u] = c < n2 ? ld_f32(pred, po + c) : 0.f;
        t[u] = c < n2 ? ld_f32(target, to + c) : 0.f;
      }
      // the four fp32 products of a pass are summed pairwise in fp32 (like one level of torch.sum's own pairwise
      // tree), then widened: a quarter of the F2F.F64 conversions, which -- not the loads -- bounded this kernel
      // (ncu round 2: 54 % of the copy bandwidth with a widening per product).  Fixed order => deterministic.
      static_assert(kUnroll == 4, "pairwise tree below is written for four chunks");
      num += double(__fadd_rn(__fadd_rn(__fmul_rn(p[0], t[0]), __fmul_rn(p[1], t[1])),
                              __fadd_rn(__fmul_rn(p[2], t[2]), __fmul_rn(p[3], t[3]))));
      den += double(__fadd_rn(__fadd_rn(__fmul_rn(p[0], p[0]), __fmul_rn(p[1], p[1])),
                              __fadd_rn(__fmul_rn(p[2], p[2]), __fmul_rn(p[3], p[3]))));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    num += __shfl_xor_sync(0xffffffffu, num, o);
    den += __shfl_xor_sync(0xffffffffu, den, o);
  }
  __shared__ double snum[kBlendThreads / 32], sden[kBlendThreads / 32];
  __shared__ bool last;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { snum[w] = num; sden[w] = den; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < kBlendThreads / 32; ++i) { a += snum[i]; b += sden[i]; }
    partial[2 * blockIdx.x] = a;
    partial[2 * blockIdx.x + 1] = b;
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x < 32) {
    __threadfence();
    double a = 0, b = 0;                                  // lane-strided, then a fixed-order butterfly: deterministic
    for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) {
      a += __ldcg(&partial[2 * i]);
      b += __ldcg(&partial[2 * i + 1]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (threadIdx.x == 0) {
      sums[0] = a;
      sums[1] = b;
      *counter = 0;                                       // self-cleaning: the next launch on this stream starts at 0
    }
  }
}

// dst = acc * w + (scale*win) * (1 - w),  w = np.linspace(1, 0, n)[k], k = index along `axis`.
// numpy linspace: start + k*step with step = (stop-start)/(n-1), last sample forced to `stop`; n == 1 -> [1.0].
__device__ __forceinline__ double linspace_1_0(int64_t k, int64_t n) {
  if (n <= 1) return 1.0;
  if (k == n - 1) return 0.0;
  return 1.0 + double(k) * ((0.0 - 1.0) / double(n - 1));
}

__global__ void __launch_bounds__(kBlendThreads)
blend_crossfade_kernel(double* __restrict__ dst, int64_t d0, int64_t d1, View3 acc, View3 win, double scale_host,
                       const double* __restrict__ sums, int64_t n0, int64_t n1, int64_t n2, int axis) {
  const double scale = sums ? scale_from_sums(sums) : scale_host;
  const int64_t rows = n0 * n1;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const int64_t i0 = r / n1, i1 = r - i0 * n1;
    const int64_t ao = i0 * acc.s0 + i1 * acc.s1, wo = i0 * win.s0 + i1 * win.s1, dof = i0 * d0 + i1 * d1;
    const double w_row = axis == 0 ? linspace_1_0(i0, n0) : linspace_1_0(i1, n1);
    for (int64_t c0 = threadIdx.x; c0 < n2; c0 += kUnroll * kBlendThreads) {
      double a[kUnroll], b[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int64_t c = c0 + u * kBlendThreads;
        if (c < n2) {
          a[u] = ld_f64(acc, ao + c);
          b[u] = aligned(win, wo + c, scale);
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int64_t c = c0 + u * kBlendThreads;
        if (c < n2) {
          const double w = axis == 2 ? linspace_1_0(c, n2) : w_row;
          dst[dof + c] = a[u] * w + b[u] * (1.0 - w);
        }
      }
    }
  }
}

// dst (fp64) = apply_scale ? scale*src (numpy promotion as above) : src
__global__ void __launch_bounds__(kBlendThreads)
scale_copy_kernel(double* __restrict__ dst, int64_t d0, int64_t d1, View3 src, double scale_host,
                  const double* __restrict__ sums, int apply_scale, int64_t n0, int64_t n1, int64_t n2) {
  const double scale = sums ? scale_from_sums(sums) : scale_host;
  const int64_t rows = n0 * n1;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const int64_t i0 = r / n1, i1 = r - i0 * n1;
    const int64_t so = i0 * src.s0 + i1 * src.s1, dof = i0 * d0 + i1 * d1;
    for (int64_t c0 = threadIdx.x; c0 < n2; c0 += kUnroll * kBlendThreads) {
      double v[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int64_t c = c0 + u * kBlendThreads;
        if (c < n2) v[u] = apply_scale ? aligned(src, so + c, scale) : ld_f64(src, so + c);
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int64_t c = c0 + u * kBlendThreads;
        if (c < n2) dst[dof + c] = v[u];
      }
    }
  }
}

// depth = clip(1 / disparity, 0, 100) in float64 (launch_aether.py:347: np.clip(1.0 / disparity_video, 0, 1e2);
// 1/0 = inf -> 100, like numpy).  src fp32 or fp64 region, dst fp64.
__global__ void __launch_bounds__(kBlendThreads)
disparity_to_depth_kernel(double* __restrict__ dst, int64_t d0, int64_t d1, View3 src, int64_t n0, int64_t n1,
                          int64_t n2) {
  const int64_t rows = n0 * n1;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const int64_t i0 = r / n1, i1 = r - i0 * n1;
    const int64_t so = i0 * src.s0 + i1 * src.s1, dof = i0 * d0 + i1 * d1;
    for (int64_t c = threadIdx.x; c < n2; c += kBlendThreads) {
      const double d = __ddiv_rn(1.0, ld_f64(src, so + c));
      dst[dof + c] = fmin(fmax(d, 0.0), 100.0);           // NaN (0/0 never occurs: numerator is 1) propagates like np.clip
    }
  }
}

static unsigned row_grid(int64_t rows, int cap_per_sm) {
  const int64_t cap = int64_t(num_sms()) * cap_per_sm;
  return (unsigned)(rows < cap ? rows : cap);
}

}  // namespace aether

using namespace aether;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
extern "C" {
int64_t aether_scale_reduce_work_bytes(void) { return 16 + 8 + int64_t(kMaxReduceBlocks) * 16; }

int aether_scale_reduce(const void* pred, int32_t pred_is_f64, int64_t pred_s0, int64_t pred_s1, const void* target,
                        int32_t target_is_f64, int64_t target_s0, int64_t target_s1, int64_t n0, int64_t n1,
                        int64_t n2, void* work, void* stream) {
  if (!pred || !target || !work || n0 <= 0 || n1 <= 0 || n2 <= 0) return AETHER_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(work) & 7) != 0) return AETHER_ERR_INVALID;
  View3 p{pred, pred_is_f64, pred_s0, pred_s1}, t{target, target_is_f64, target_s0, target_s1};
  double* sums = reinterpret_cast<double*>(work);
  unsigned int* counter = reinterpret_cast<unsigned int*>(sums + 2);
  double* partial = sums + 3;
  unsigned grid = row_grid(n0 * n1, 8);
  if (grid > (unsigned)kMaxReduceBlocks) grid = kMaxReduceBlocks;
  scale_reduce_kernel<<<grid, kBlendThreads, 0, ST(stream)>>>(p, t, n0, n1, n2, sums, partial, counter);
  return cudaGetLastError() == cudaSuccess ? AETHER_OK : AETHER_ERR_CUDA;
}
int aether_blend_crossfade(double* dst, int64_t dst_s0, int64_t dst_s1, const void* acc, int32_t acc_is_f64,
                           int64_t acc_s0, int64_t acc_s1, const void* win, int32_t win_is_f64, int64_t win_s0,
                           int64_t win_s1, double scale, const double* scale_sums, int64_t n0, int64_t n1, int64_t n2,
                           int32_t axis, void* stream) {
  if (!dst || !acc || !win || n0 <= 0 || n1 <= 0 || n2 <= 0 || axis < 0 || axis > 2) return AETHER_ERR_INVALID;
  View3 a{acc, acc_is_f64, acc_s0, acc_s1}, w{win, win_is_f64, win_s0, win_s1};
  blend_crossfade_kernel<<<row_grid(n0 * n1, 8), kBlendThreads, 0, ST(stream)>>>(dst, dst_s0, dst_s1, a, w, scale,
                                                                                 scale_sums, n0, n1, n2, axis);
  return cudaGetLastError() == cudaSuccess ? AETHER_OK : AETHER_ERR_CUDA;
}
int aether_scale_copy(double* dst, int64_t dst_s0, int64_t dst_s1, const void* src, int32_t src_is_f64, int64_t src_s0,
                      int64_t src_s1, double scale, const double* scale_sums, int32_t apply_scale, int64_t n0,
                      int64_t n1, int64_t n2, void* stream) {
  if (!dst || !src || n0 <= 0 || n1 <= 0 || n2 <= 0) return AETHER_ERR_INVALID;
  View3 s{src, src_is_f64, src_s0, src_s1};
  scale_copy_kernel<<<row_grid(n0 * n1, 8), kBlendThreads, 0, ST(stream)>>>(dst, dst_s0, dst_s1, s, scale, scale_sums,
                                                                            apply_scale, n0, n1, n2);
  return cudaGetLastError() == cudaSuccess ? AETHER_OK : AETHER_ERR_CUDA;
}
// One link of the reference's blend chain as ONE call: window `win` ([n_win along `axis`] x the other two extents, fp32
// or fp64) is appended to the fp64 accumulation `buf` whose valid extent along `axis` ends at `prev_end`; the window
// starts at `start` (overlap = prev_end - start):   scale = compute_scale(win[:overlap], buf[start:prev_end]);
// buf[start:prev_end] = buf[start:prev_end] * w + scale * win[:overlap] * (1 - w);  buf[prev_end:start + n_win] = scale *
// win[overlap:]   (launch_aether.py:193-250 spatial, :268-284 temporal).  The SURVEY 8(b) `blend_tiles` entry point: three
// kernels on `stream`, the scale never leaves the device.  e0, e1, e2 = extents of `win`; strides in elements.
int aether_blend_link(double* buf, int64_t buf_s0, int64_t buf_s1, const void* win, int32_t win_is_f64, int64_t win_s0,
                      int64_t win_s1, int64_t e0, int64_t e1, int64_t e2, int32_t axis, int64_t start, int64_t prev_end,
                      void* work, void* stream) {
  if (!buf || !win || !work || axis < 0 || axis > 2 || e0 <= 0 || e1 <= 0 || e2 <= 0) return AETHER_ERR_INVALID;
  const int64_t ext[3] = {e0, e1, e2};
  const int64_t n_win = ext[axis], overlap = prev_end - start;
  if (overlap <= 0 || overlap > n_win || start < 0) return AETHER_ERR_INVALID;
  const int64_t bstr[3] = {buf_s0, buf_s1, 1}, wstr[3] = {win_s0, win_s1, 1};
  const int64_t wsz = win_is_f64 ? 8 : 4;
  double* acc = buf + start * bstr[axis];                                  // buf[start:prev_end] along `axis`
  int64_t n[3] = {e0, e1, e2};
  n[axis] = overlap;
  int rc = aether_scale_reduce(win, win_is_f64, win_s0, win_s1, acc, 1, buf_s0, buf_s1, n[0], n[1], n[2], work, stream);
  if (rc) return rc;
  const double* sums = reinterpret_cast<const double*>(work);
  rc = aether_blend_crossfade(acc, buf_s0, buf_s1, acc, 1, buf_s0, buf_s1, win, win_is_f64, win_s0, win_s1, 1.0, sums, n[0],
                              n[1], n[2], axis, stream);
  if (rc || overlap == n_win) return rc;
  n[axis] = n_win - overlap;
  const char* tail = reinterpret_cast<const char*>(win) + overlap * wstr[axis] * wsz;
  return aether_scale_copy(buf + prev_end * bstr[axis], buf_s0, buf_s1, tail, win_is_f64, win_s0, win_s1, 1.0, sums, 1, n[0],
                           n[1], n[2], stream);
}

int aether_disparity_to_depth(double* dst, int64_t dst_s0, int64_t dst_s1, const void* src, int32_t src_is_f64,
                              int64_t src_s0, int64_t src_s1, int64_t n0, int64_t n1, int64_t n2, void* stream) {
  if (!dst || !src || n0 <= 0 || n1 <= 0 || n2 <= 0) return AETHER_ERR_INVALID;
  View3 s{src, src_is_f64, src_s0, src_s1};
  disparity_to_depth_kernel<<<row_grid(n0 * n1, 8), kBlendThreads, 0, ST(stream)>>>(dst, dst_s0, dst_s1, s, n0, n1, n2);
  return cudaGetLastError() == cudaSuccess ? AETHER_OK : AETHER_ERR_CUDA;
}
}
