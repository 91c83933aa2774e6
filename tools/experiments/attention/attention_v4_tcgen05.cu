// K1, variant 4 (mode 7 of aether_attention_bf16): ONE 128-row query tile per CTA, TWO CTAs resident per SM.
// Same per-tile algorithm as variant 3 (decoupled S / P TMEM buffers, S released as soon as it is in registers), but
// the two query tiles that share an SM are now independent CTAs with their own TMA / MMA issue threads: no ordering
// is imposed between them, so their TMEM-load and MUFU phases interleave freely.  Cost: K/V tiles are fetched from
// L2 once per CTA instead of once per tile pair (32 KB per ~2048 clk per CTA, ~4.7 KB/clk chip-wide, under the L2 cap).
//   CTA = 6 warps: warp 0 TMA (+ TMEM alloc), warp 1 MMA issuer, warps 2..5 softmax (lane quarter = warp % 4).
//   TMEM (256 columns per CTA, 512 per SM):  S [0,128)  P [128,192)  O [192,256)
//   smem per CTA: Q 16 KB + 3 K stages + 2 V stages = 96 KB  (two CTAs per SM).
#include "host_util.h"
#include "ptx.cuh"

namespace aether {
namespace attn4 {

constexpr int DH = 64, BQ = 128, BKV = 128, KSTAGES = 3, VSTAGES = 2;
constexpr int TILE_BYTES = BQ * DH * 2;
constexpr int SMEM_BYTES = 1024 + (1 + KSTAGES + VSTAGES) * TILE_BYTES + 256;
constexpr int THREADS = 192;
constexpr uint32_t COL_S = 0, COL_P = 128, COL_O = 192;

struct Params {
  int B, H, S;
  __nv_bfloat16* out;
  float scale_log2;
  int q_row0;      // first query row this launch covers (the tail launch of mode 5 starts at a multiple of 256)
};

__global__ void __launch_bounds__(THREADS, 2)
attention_v4_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + TILE_BYTES;
  uint8_t* smem_v = smem_k + KSTAGES * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + VSTAGES * TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = q_full + 1;
  uint64_t* k_empty = k_full + KSTAGES;
  uint64_t* v_full = k_empty + KSTAGES;
  uint64_t* v_empty = v_full + VSTAGES;
  uint64_t* s_full = v_empty + VSTAGES;
  uint64_t* s_free = s_full + 1;            // count 128
  uint64_t* p_full = s_free + 1;            // count 128
  uint64_t* p_free = p_full + 1;
  uint64_t* o_full = p_free + 1;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform (see gemm_tcgen05.cu)
  const int lane = threadIdx.x & 31;
  const int q_blk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = p.q_row0 + q_blk * BQ;
  const int n_kv = (p.S + BKV - 1) / BKV;
  const int H = p.H;

  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    mbar_init(q_full, 1);
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_full, 128);
    mbar_init(p_free, 1);
    mbar_init(o_full, 1);
    for (int i = 0; i < KSTAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < VSTAGES; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_base_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_ptr, 0);

  if (warp == 0) {
    {
      const bool lead = elect_one();
      if (lead) {
        mbar_arrive_expect_tx(q_full, TILE_BYTES);
        tma_load_4d(smem_q, &tmap_qkv, q_full, 0, h, q0, b);
      }
      __syncwarp();
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&k_empty[ks], kph ^ 1);
        if (lead) {
          mbar_arrive_expect_tx(&k_full[ks], TILE_BYTES);
          tma_load_4d(smem_k + ks * TILE_BYTES, &tmap_qkv, &k_full[ks], 0, H + h, j * BKV, b);
        }
        __syncwarp();
        if (++ks == KSTAGES) { ks = 0; kph ^= 1; }
        mbar_wait(&v_empty[vs], vph ^ 1);
        if (lead) {
          mbar_arrive_expect_tx(&v_full[vs], TILE_BYTES);
          tma_load_4d(smem_v + vs * TILE_BYTES, &tmap_qkv, &v_full[vs], 0, 2 * H + h, j * BKV, b);
        }
        __syncwarp();
        if (++vs == VSTAGES) { vs = 0; vph ^= 1; }
      }
    }
  } else if (warp == 1) {
    {
      const bool lead = elect_one();   // whole warp walks the schedule (uniform control flow), one lane issues
      constexpr uint32_t idesc_qk = make_idesc_f16kind(BQ, BKV, 1, 1, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16kind(BQ, DH, 1, 1, 0, 1);
      const uint32_t s_col = tmem_base + COL_S, p_col = tmem_base + COL_P, o_col = tmem_base + COL_O;
      const uint64_t q_desc = make_sw128_desc(smem_u32(smem_q));
      auto issue_qk = [&](int ks) {
        const uint64_t k_desc = make_sw128_desc(smem_u32(smem_k + ks * TILE_BYTES));
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) tc_mma_ss(s_col, q_desc + 2 * k, k_desc + 2 * k, idesc_qk, k > 0);
      };
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      mbar_wait(&k_full[0], 0);
      mbar_wait(q_full, 0);
      tc_fence_after();
      if (lead) {
        issue_qk(0);
        tc_commit(s_full);
        tc_commit(&k_empty[0]);
      }
      __syncwarp();
      ks = 1;
      if (ks == KSTAGES) { ks = 0; kph ^= 1; }
      for (int j = 0; j < n_kv; ++j) {
        const bool last = (j + 1 == n_kv);
        if (!last) {
          mbar_wait(&k_full[ks], kph);
          mbar_wait(s_free, j & 1);
          tc_fence_after();
          if (lead) {
            issue_qk(ks);
            tc_commit(s_full);
            tc_commit(&k_empty[ks]);
          }
          __syncwarp();
          if (++ks == KSTAGES) { ks = 0; kph ^= 1; }
        }
        mbar_wait(&v_full[vs], vph);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        if (lead) {
          const uint64_t v_desc = make_sw128_desc(smem_u32(smem_v + vs * TILE_BYTES));
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k)
            tc_mma_ts(o_col, p_col + 8 * k, v_desc + 128 * k, idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
          tc_commit(p_free);
          tc_commit(&v_empty[vs]);
          if (last) tc_commit(o_full);
        }
        __syncwarp();
        if (++vs == VSTAGES) { vs = 0; vph ^= 1; }
      }
    }
  } else {
    // (no setmaxnreg here: with 6 warps the second warpgroup is incomplete and the .sync.aligned rebalance hangs)
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_off + COL_S;
    const uint32_t p_addr = tmem_base + lane_off + COL_P;
    const uint32_t o_addr = tmem_base + lane_off + COL_O;
    const float sl2 = p.scale_log2;
    const float rescale_thresh = 8.0f / sl2;
    float m_used = -INFINITY, l = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t s[128];
      {
        uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
        uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
        uint32_t(&s2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[64]);
        uint32_t(&s3)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[96]);
        tmem_ld_32x32b_x32(s_addr, s0);
        tmem_ld_32x32b_x32(s_addr + 32, s1);
        tmem_ld_32x32b_x32(s_addr + 64, s2);
        tmem_ld_32x32b_x32(s_addr + 96, s3);
      }
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive(s_free);
      const int kv_valid = p.S - j * BKV;
      if (kv_valid < BKV) {
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if (c >= kv_valid) s[c] = 0xFF800000u;
      }
      float mx[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) mx[c] = __uint_as_float(s[c]);
#pragma unroll
      for (int c = 8; c < 128; ++c) mx[c & 7] = fmaxf(mx[c & 7], __uint_as_float(s[c]));
      const float m_tile =
          fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
      const float m_new = fmaxf(m_used, m_tile);
      const bool need = (m_new - m_used) > rescale_thresh;
      bool pv_prev_done = (j == 0);
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = fast_exp2((m_used - m_new) * sl2);
        l *= alpha;
        m_used = m_new;
        if (j > 0) {
          mbar_wait(p_free, (j - 1) & 1);
          pv_prev_done = true;
          tc_fence_after();
          uint32_t o0[32], o1[32];
          tmem_ld_32x32b_x32(o_addr, o0);
          tmem_ld_32x32b_x32(o_addr + 32, o1);
          tc_wait_ld();
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            o0[c] = __float_as_uint(__uint_as_float(o0[c]) * alpha);
            o1[c] = __float_as_uint(__uint_as_float(o1[c]) * alpha);
          }
          tmem_st_32x32b_x32(o_addr, o0);
          tmem_st_32x32b_x32(o_addr + 32, o1);
        }
      }
      const float neg_m = -m_used * sl2;
      float sum[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) sum[c] = 0.f;
      uint32_t pk[64];
#pragma unroll
      for (int c = 0; c < 128; c += 2) {
        const float p0 = fast_exp2(fmaf(__uint_as_float(s[c]), sl2, neg_m));
        const float p1 = fast_exp2(fmaf(__uint_as_float(s[c + 1]), sl2, neg_m));
        sum[c & 7] += p0;
        sum[(c + 1) & 7] += p1;
        pk[c >> 1] = pack_bf16x2(p0, p1);
      }
      l += ((sum[0] + sum[1]) + (sum[2] + sum[3])) + ((sum[4] + sum[5]) + (sum[6] + sum[7]));
      if (!pv_prev_done) {
        mbar_wait(p_free, (j - 1) & 1);
        tc_fence_after();
      }
      {
        const uint32_t(&p0)[32] = *reinterpret_cast<const uint32_t(*)[32]>(&pk[0]);
        const uint32_t(&p1)[32] = *reinterpret_cast<const uint32_t(*)[32]>(&pk[32]);
        tmem_st_32x32b_x32(p_addr, p0);
        tmem_st_32x32b_x32(p_addr + 32, p1);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(p_full);
    }

    mbar_wait(o_full, 0);
    tc_fence_after();
    uint32_t o0[32], o1[32];
    tmem_ld_32x32b_x32(o_addr, o0);
    tmem_ld_32x32b_x32(o_addr + 32, o1);
    tc_wait_ld();
    const int row = q0 + row_in_tile;
    if (row < p.S) {
      const float inv = 1.0f / l;
      __nv_bfloat16* dst = p.out + (int64_t(b) * p.S + row) * (int64_t(H) * DH) + h * DH;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(o0[v * 8 + 0]) * inv, __uint_as_float(o0[v * 8 + 1]) * inv);
        w.y = pack_bf16x2(__uint_as_float(o0[v * 8 + 2]) * inv, __uint_as_float(o0[v * 8 + 3]) * inv);
        w.z = pack_bf16x2(__uint_as_float(o0[v * 8 + 4]) * inv, __uint_as_float(o0[v * 8 + 5]) * inv);
        w.w = pack_bf16x2(__uint_as_float(o0[v * 8 + 6]) * inv, __uint_as_float(o0[v * 8 + 7]) * inv);
        reinterpret_cast<uint4*>(dst)[v] = w;
      }
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(o1[v * 8 + 0]) * inv, __uint_as_float(o1[v * 8 + 1]) * inv);
        w.y = pack_bf16x2(__uint_as_float(o1[v * 8 + 2]) * inv, __uint_as_float(o1[v * 8 + 3]) * inv);
        w.z = pack_bf16x2(__uint_as_float(o1[v * 8 + 4]) * inv, __uint_as_float(o1[v * 8 + 5]) * inv);
        w.w = pack_bf16x2(__uint_as_float(o1[v * 8 + 6]) * inv, __uint_as_float(o1[v * 8 + 7]) * inv);
        reinterpret_cast<uint4*>(dst)[4 + v] = w;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace attn4

int attention_v4_launch_rows(const CUtensorMap& tm, int B, int S, int H, void* out, float scale_log2, int q_row0,
                             cudaStream_t stream);
int attention_v4_launch(const CUtensorMap& tm, int B, int S, int H, void* out, float scale_log2, cudaStream_t stream) {
  return attention_v4_launch_rows(tm, B, S, H, out, scale_log2, 0, stream);
}

// Query rows [q_row0, S) only (all keys): the tail launch of the split mode-5 schedule.
int attention_v4_launch_rows(const CUtensorMap& tm, int B, int S, int H, void* out, float scale_log2, int q_row0,
                             cudaStream_t stream) {
  static SmemGrant grant;
  AETHER_CUDA_OK(ensure_dynamic_smem(grant, attn4::attention_v4_kernel, attn4::SMEM_BYTES));
  attn4::Params p;
  p.B = B; p.H = H; p.S = S;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.scale_log2 = scale_log2;
  p.q_row0 = q_row0;
  AETHER_CHECK_ARG(q_row0 >= 0 && q_row0 < S);
  dim3 grid((unsigned)ceil_div(S - q_row0, attn4::BQ), (unsigned)H, (unsigned)B);
  attn4::attention_v4_kernel<<<grid, attn4::THREADS, attn4::SMEM_BYTES, stream>>>(tm, p);
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}

}  // namespace aether
