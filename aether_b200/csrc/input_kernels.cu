// Input side of the path on the device: what happens to the frames between the decoder and the
// VAE encoder in the reference --
//   evaluation/video_depth/launch_aether.py:388-403  prepare_input: aspect-preserving cv2.resize (INTER_LINEAR, uint8)
//                                                     to the 480 x 720 window, then / 255.0 (float64)
//   aether/pipelines/aetherv1_pipeline_cogvideox.py:451-512  centre crop, [F,H,W,3] -> [F,3,H,W], 2x - 1, cast to bf16
// as two streaming kernels over uint8 frames that never leave the GPU: 3 bytes per pixel cross PCIe once instead of a
// float64 clip (24 bytes per pixel) per tile.
//
// aether_resize_bilinear_u8 reproduces OpenCV's 8-bit INTER_LINEAR bit for bit (checked against cv2.resize on the
// reference's own fixture sizes): source coordinate fx = float((d + 0.5) * scale - 0.5), 11-bit fixed-point weights
// round-half-even(w * 2048), horizontal pass in int32 (borders: weight snapped to the edge pixel), vertical pass
// ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2 with the source ROW index clamped but the weights kept.
// aether_u8_frames_to_model_input: bf16( float( 2.0 * (double(k) / 255.0) - 1.0 ) ), which is also what the pipeline's
// own uint8 branch (:454, float32 division) yields for every k in 0..255 (tests/test_input_gpu.py enumerates them).
// HBM-bound: 3 B read + 3 B written per output pixel (resize), 3 B read + 6 B written (model input).
#include <cuda_bf16.h>

#include "host_util.h"

namespace aether {

struct Lin {
  int i0, i1;       // source indices
  int a0, a1;       // 11-bit fixed-point weights
};

__device__ __forceinline__ Lin lin_coeff(int d, double scale, int src, bool vertical) {
  float f = float((double(d) + 0.5) * scale - 0.5);
  int s = int(floorf(f));
  f -= float(s);
  Lin c;
  if (!vertical) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= src - 1) { f = 0.f; s = src - 1; }
    c.i0 = s;
    c.i1 = min(s + 1, src - 1);
  } else {
    c.i0 = min(max(s, 0), src - 1);
    c.i1 = min(max(s + 1, 0), src - 1);
  }
  c.a0 = __float2int_rn((1.0f - f) * 2048.0f);
  c.a1 = __float2int_rn(f * 2048.0f);
  return c;
}

__global__ void __launch_bounds__(256)
resize_bilinear_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int T, int h, int w, int H, int W,
                          double scale_y, double scale_x) {
  const int64_t total = int64_t(T) * H * W;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int x = int(i % W);
    int64_t r = i / W;
    const int y = int(r % H);
    const int t = int(r / H);
    const Lin cx = lin_coeff(x, scale_x, w, false), cy = lin_coeff(y, scale_y, h, true);
    const uint8_t* f = src + int64_t(t) * h * w * 3;
    const uint8_t* r0 = f + int64_t(cy.i0) * w * 3;
    const uint8_t* r1 = f + int64_t(cy.i1) * w * 3;
    uint8_t* o = dst + i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int h0 = int(r0[cx.i0 * 3 + c]) * cx.a0 + int(r0[cx.i1 * 3 + c]) * cx.a1;
      const int h1 = int(r1[cx.i0 * 3 + c]) * cx.a0 + int(r1[cx.i1 * 3 + c]) * cx.a1;
      int v = (((cy.a0 * (h0 >> 4)) >> 16) + ((cy.a1 * (h1 >> 4)) >> 16) + 2) >> 2;
      o[c] = uint8_t(min(max(v, 0), 255));
    }
  }
}

// dst[f, c, y, x] = bf16(2 * src[f, y0 + y, x0 + x, c] / 255 - 1); src frames have element strides (sT, sH) and 3
// interleaved channels.  One thread per output pixel (3 loads, 3 coalesced 2-byte stores in 3 planes).
__global__ void __launch_bounds__(256)
u8_frames_to_model_input_kernel(const uint8_t* __restrict__ src, int64_t sT, int64_t sH, __nv_bfloat16* __restrict__ dst,
                                int F, int H, int W) {
  const int64_t plane = int64_t(H) * W;
  const int64_t total = int64_t(F) * plane;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int x = int(i % W);
    int64_t r = i / W;
    const int y = int(r % H);
    const int f = int(r / H);
    const uint8_t* p = src + f * sT + y * sH + int64_t(x) * 3;
    __nv_bfloat16* o = dst + int64_t(f) * 3 * plane + int64_t(y) * W + x;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      o[c * plane] = __float2bfloat16_rn(float(2.0 * (double(p[c]) / 255.0) - 1.0));
  }
}

// world[t, y, x, :] = P_t[:, :3] . (depth * ((x + 0.5) * a + b, (y + 0.5) * c + d, 1)) + P_t[:, 3],  depth = 1 / clip(disp, 1e-8, 1e8)
// (scripts/demo.py:404-420 + aether/utils/postprocess_utils.py:381-403 `project`): cam[t] = {1/f, -cx/f, 1/f, -cy/f, P (3 x 4
// row-major)} as 16 doubles.  fp64 like the numpy path; 4|8 B read + 24 B written per pixel.
__global__ void __launch_bounds__(256)
project_points_kernel(const void* __restrict__ disp, int disp_f64, const double* __restrict__ cam, double* __restrict__ out,
                      int T, int H, int W) {
  const int64_t total = int64_t(T) * H * W;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int x = int(i % W);
    int64_t r = i / W;
    const int y = int(r % H);
    const int t = int(r / H);
    const double* c = cam + int64_t(t) * 16;
    double d = disp_f64 ? reinterpret_cast<const double*>(disp)[i] : double(reinterpret_cast<const float*>(disp)[i]);
    d = fmin(fmax(d, 1e-8), 1e8);
    const double depth = __ddiv_rn(1.0, d);
    const double px = double(float(x) + 0.5f), py = double(float(y) + 0.5f);       // float32 pixel centres like the ref
    const double cx = __dmul_rn(__dadd_rn(__dmul_rn(c[0], px), c[1]), depth);
    const double cy = __dmul_rn(__dadd_rn(__dmul_rn(c[2], py), c[3]), depth);
    const double cz = depth;
    double* o = out + i * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double* p = c + 4 + 4 * k;
      o[k] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(p[0], cx), __dmul_rn(p[1], cy)), __dmul_rn(p[2], cz)), p[3]);
    }
  }
}

static unsigned igrid(int64_t n) {
  int64_t g = ceil_div(n, 256);
  const int64_t cap = int64_t(num_sms()) * 16;
  return (unsigned)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace aether

using namespace aether;
extern "C" {
int aether_resize_bilinear_u8(const void* src, void* dst, int32_t T, int32_t h, int32_t w, int32_t H, int32_t W,
                              void* stream) {
  if (!src || !dst || T <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return AETHER_ERR_INVALID;
  // OpenCV: inv_scale = dsize / ssize (double); scale = 1 / inv_scale
  const double sy = 1.0 / (double(H) / double(h)), sx = 1.0 / (double(W) / double(w));
  resize_bilinear_u8_kernel<<<igrid(int64_t(T) * H * W), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint8_t*>(src), reinterpret_cast<uint8_t*>(dst), T, h, w, H, W, sy, sx);
  return cudaGetLastError() == cudaSuccess ? AETHER_OK : AETHER_ERR_CUDA;
}

int aether_project_points(const void* disparity, int32_t disparity_is_f64, const double* cam, double* points, int32_t T,
                          int32_t H, int32_t W, void* stream) {
  if (!disparity || !cam || !points || T <= 0 || H <= 0 || W <= 0) return AETHER_ERR_INVALID;
  project_points_kernel<<<igrid(int64_t(T) * H * W), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      disparity, disparity_is_f64, cam, points, T, H, W);
  return cudaGetLastError() == cudaSuccess ? AETHER_OK : AETHER_ERR_CUDA;
}

int aether_u8_frames_to_model_input(const void* src, int64_t stride_t, int64_t stride_h, void* dst, int32_t F, int32_t H,
                                    int32_t W, void* stream) {
  if (!src || !dst || F <= 0 || H <= 0 || W <= 0) return AETHER_ERR_INVALID;
  u8_frames_to_model_input_kernel<<<igrid(int64_t(F) * H * W), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint8_t*>(src), stride_t, stride_h, reinterpret_cast<__nv_bfloat16*>(dst), F, H, W);
  return cudaGetLastError() == cudaSuccess ? AETHER_OK : AETHER_ERR_CUDA;
}
}
