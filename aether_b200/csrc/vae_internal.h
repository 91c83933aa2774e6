// Internal C++ interface between vae_kernels.cu / conv_tcgen05.cu / gemm_tcgen05.cu and the handle-level VAE
// executor (vae_exec.cu).  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace aether {

// Small index map passed to kernels BY VALUE (frame -> source frame tables of the causal VAE: nearest temporal
// up-sampling, avg-pool pairs, SpatialNorm latent frame).  A frame batch never has more than 9 frames.
struct IMap {
  static constexpr int kMax = 32;
  int v[kMax];
  __host__ __device__ int operator[](int i) const { return v[i]; }
};

int conv3d_bf16(const void* x, int T_in, int H_in, int W_in, int Cin, const void* w_packed, const float* bias,
                const void* resid, void* y, int T_out, int H_out, int W_out, int Cout, int kt, int kh, int kw,
                int stride, int pad_h, int pad_w, cudaStream_t stream);
int gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N, int K,
              const float* bias, int epilogue, const float* gate_vid, const float* gate_txt, int64_t gate_bstride,
              int S, int St, int f16_from_col, cudaStream_t stream);
// `counter`: nullable zeroed device word (left zero); with it the statistics are ONE launch (see gn_partial_kernel)
int gn_stats(const void* x, int64_t N, int C, int G, float eps, float* workspace, float* mean_rstd, unsigned* counter,
             cudaStream_t stream);
int64_t gn_workspace_floats(int C);
int gn_apply_imap(const void* x, void* y, int64_t N, int C, int G, const float* mean_rstd, const float* gamma,
                  const float* beta, const void* zy, const void* zb, int zld, const IMap* tmap, int H, int W, int hz,
                  int wz, int silu, void* y2, int64_t y2_from, cudaStream_t stream);   // y2: rows >= y2_from also go there
int upsample_nearest_imap(const void* in, void* out, const IMap& tmap, int To, int Ho, int Wo, int Hi, int Wi, int sy,
                          int sx, int C, cudaStream_t stream);
int avgpool_time_imap(const void* in, void* out, const IMap& ia, const IMap& ib, int To, int64_t frame_elems,
                      cudaStream_t stream);
int crop_ncthw_to_thwc(const void* in, int64_t sC, int64_t sT, int64_t sH, void* out, int C, int Cp, int T, int H, int W,
                       cudaStream_t stream);
int copy_region_cl(const void* src, int Hs, int Ws, void* dst, int Hd, int Wd, int T, int h, int w, int C, int y0, int x0,
                   cudaStream_t stream);
int tile_blend(const void* a, void* b, int T, int Ha, int Wa, int Hb, int Wb, int C, int axis, int extent,
               cudaStream_t stream);
int thwc_to_ncthw(const void* in, void* out, int C, int Cp, int64_t thw, cudaStream_t stream);

}  // namespace aether
