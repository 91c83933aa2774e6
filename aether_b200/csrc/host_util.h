// Host-side helpers shared by the launchers: status codes, TMA tensor-map encoding through the driver
// entry point (resolved at run time so the library links only against the static CUDA runtime and
// loads on a machine without libcuda), and small integer utilities.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/aether_b200.h"

namespace aether {

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

#define AETHER_CUDA_OK(expr)                                                                        \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      fprintf(stderr, "[aether_b200] %s:%d CUDA error %s: %s\n", __FILE__, __LINE__,                \
              cudaGetErrorName(_e), cudaGetErrorString(_e));                                        \
      return AETHER_ERR_CUDA;                                                                       \
    }                                                                                               \
  } while (0)

#define AETHER_CHECK_ARG(cond)                                                                      \
  do {                                                                                              \
    if (!(cond)) {                                                                                  \
      fprintf(stderr, "[aether_b200] %s:%d invalid argument: %s\n", __FILE__, __LINE__, #cond);     \
      return AETHER_ERR_INVALID;                                                                    \
    }                                                                                               \
  } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// bf16 tensor map, up to 5 dims (dim 0 innermost, contiguous).  `strides_bytes[i]` is the byte stride of
// dim i+1.  The box is {box[0..rank)}; swizzle 128B requires box[0]*2 == 128 bytes.
inline int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) {
    fprintf(stderr, "[aether_b200] cuTensorMapEncodeTiled unavailable (no CUDA driver?)\n");
    return AETHER_ERR_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[aether_b200] cuTensorMapEncodeTiled failed: %d (rank %d dims %llu %llu box %u %u)\n", (int)r,
            rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
            rank > 1 ? box[1] : 0);
    return AETHER_ERR_CUDA;
  }
  return AETHER_OK;
}

// Row-major [rows, cols] bf16 matrix with leading dimension ld (elements); box = {box_cols, box_rows}.
inline int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                        uint32_t box_rows, uint32_t box_cols) {
  uint64_t dims[2] = {cols, rows};
  uint64_t str[1] = {ld * 2};
  uint32_t box[2] = {box_cols, box_rows};
  return make_tmap_bf16(out, base, 2, dims, str, box, true);
}

constexpr int kMaxDevices = 64;

inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}

// SM count of the CURRENT device (the one the caller's stream belongs to), cached per device ordinal.
inline int num_sms() {
  static std::atomic<int> cache[kMaxDevices];
  const int dev = current_device();
  std::atomic<int>* slot = (dev >= 0 && dev < kMaxDevices) ? &cache[dev] : nullptr;
  int n = slot ? slot->load(std::memory_order_relaxed) : 0;
  if (n) return n;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (slot) slot->store(n, std::memory_order_relaxed);
  return n;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE property of a kernel: remember, per launcher and per
// device ordinal, the size already granted.  Re-entrant (two threads may both set the same value; that is idempotent).
struct SmemGrant {
  std::atomic<int> bytes[kMaxDevices];
};

template <typename KernelT>
inline cudaError_t ensure_dynamic_smem(SmemGrant& g, KernelT kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return cudaSuccess;
  const int dev = current_device();
  std::atomic<int>* slot = (dev >= 0 && dev < kMaxDevices) ? &g.bytes[dev] : nullptr;
  if (slot && slot->load(std::memory_order_acquire) >= (int)bytes) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess && slot) slot->store((int)bytes, std::memory_order_release);
  return e;
}

}  // namespace aether
