// K1, variant 2 (mode 4 of aether_attention_bf16): same algorithm / TMEM layout / MMA issue order as
// attention_tcgen05.cu, but 16 softmax warps instead of 8: every query row is shared by TWO threads (columns 0..63 and
// 64..127 of the 128-key tile).  The profile of variant 1 (profiles/r1_attention_mode0_ncu_summary.txt) shows the
// softmax loop latency-bound at ~0.42 IPC with 2 softmax warps per scheduler and the MUFU at 62 %; four warps per
// scheduler give the issue slots something to overlap.  Costs: one 256-thread named barrier + a shared-memory
// exchange of the partial row maxima per tile.
//   warp 0: TMEM alloc + TMA producer   warp 1: MMA issuer   warps 2..9: query tile 0   warps 10..17: query tile 1
//   softmax warp w: TMEM lane quarter q = w % 4 (hardware rule), column half hh = ((w - 2) % 8) / 4.
#include "host_util.h"
#include "ptx.cuh"

namespace aether {
namespace attn2 {

constexpr int DH = 64, BQ = 128, BKV = 128, KSTAGES = 3, VSTAGES = 3;
constexpr int TILE_BYTES = BQ * DH * 2;
constexpr int XCHG_BYTES = 2 * 2 * 2 * 128 * 4;      // [tile][parity][half][row] fp32
constexpr int SMEM_BYTES = 1024 + (2 + KSTAGES + VSTAGES) * TILE_BYTES + XCHG_BYTES + 256;
constexpr int THREADS = 576;
constexpr uint32_t COL_S0 = 0, COL_S1 = 128, COL_O0 = 256, COL_O1 = 320;

struct Params {
  int B, H, S;
  __nv_bfloat16* out;
  float scale_log2;
};

__global__ void __launch_bounds__(THREADS, 1)
attention_v2_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + 2 * TILE_BYTES;
  uint8_t* smem_v = smem_k + KSTAGES * TILE_BYTES;
  float* xchg = reinterpret_cast<float*>(smem_v + VSTAGES * TILE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xchg) + XCHG_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = q_full + 2;
  uint64_t* k_empty = k_full + KSTAGES;
  uint64_t* v_full = k_empty + KSTAGES;
  uint64_t* v_empty = v_full + VSTAGES;
  uint64_t* s_full = v_empty + VSTAGES;
  uint64_t* p_full = s_full + 2;            // count 256
  uint64_t* o_full = p_full + 2;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_blk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = q_blk * 2 * BQ;
  const int n_kv = (p.S + BKV - 1) / BKV;
  const int H = p.H;

  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 256);
      mbar_init(&o_full[i], 1);
    }
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
  if (warp == 0) tmem_alloc<512>(tmem_base_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (int t = 0; t < 2; ++t) {
        mbar_arrive_expect_tx(&q_full[t], TILE_BYTES);
        tma_load_4d(smem_q + t * TILE_BYTES, &tmap_qkv, &q_full[t], 0, h, q0 + t * BQ, b);
      }
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&k_empty[ks], kph ^ 1);
        mbar_arrive_expect_tx(&k_full[ks], TILE_BYTES);
        tma_load_4d(smem_k + ks * TILE_BYTES, &tmap_qkv, &k_full[ks], 0, H + h, j * BKV, b);
        if (++ks == KSTAGES) { ks = 0; kph ^= 1; }
        mbar_wait(&v_empty[vs], vph ^ 1);
        mbar_arrive_expect_tx(&v_full[vs], TILE_BYTES);
        tma_load_4d(smem_v + vs * TILE_BYTES, &tmap_qkv, &v_full[vs], 0, 2 * H + h, j * BKV, b);
        if (++vs == VSTAGES) { vs = 0; vph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_f16kind(BQ, BKV, 1, 1, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16kind(BQ, DH, 1, 1, 0, 1);
      const uint32_t s_col[2] = {tmem_base + COL_S0, tmem_base + COL_S1};
      const uint32_t o_col[2] = {tmem_base + COL_O0, tmem_base + COL_O1};
      uint64_t q_desc[2];
      for (int t = 0; t < 2; ++t) q_desc[t] = make_sw128_desc(smem_u32(smem_q + t * TILE_BYTES));
      auto issue_qk = [&](int t, int ks) {
        const uint64_t k_desc = make_sw128_desc(smem_u32(smem_k + ks * TILE_BYTES));
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) tc_mma_ss(s_col[t], q_desc[t] + 2 * k, k_desc + 2 * k, idesc_qk, k > 0);
      };
      auto issue_pv = [&](int t, int vs, bool first) {
        const uint64_t v_desc = make_sw128_desc(smem_u32(smem_v + vs * TILE_BYTES));
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)
          tc_mma_ts(o_col[t], s_col[t] + 8 * k, v_desc + 128 * k, idesc_pv, (!first || k > 0) ? 1u : 0u);
      };
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0, pph = 0;
      mbar_wait(&k_full[0], 0);
      mbar_wait(&q_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      tc_commit(&s_full[0]);
      mbar_wait(&q_full[1], 0);
      tc_fence_after();
      issue_qk(1, 0);
      tc_commit(&s_full[1]);
      tc_commit(&k_empty[0]);
      ks = 1;
      if (ks == KSTAGES) { ks = 0; kph ^= 1; }
      for (int j = 0; j < n_kv; ++j) {
        const bool last = (j + 1 == n_kv);
        mbar_wait(&v_full[vs], vph);
        mbar_wait(&p_full[0], pph);
        tc_fence_after();
        issue_pv(0, vs, j == 0);
        if (!last) {
          mbar_wait(&k_full[ks], kph);
          tc_fence_after();
          issue_qk(0, ks);
          tc_commit(&s_full[0]);
        } else {
          tc_commit(&o_full[0]);
        }
        mbar_wait(&p_full[1], pph);
        tc_fence_after();
        issue_pv(1, vs, j == 0);
        tc_commit(&v_empty[vs]);
        if (!last) {
          issue_qk(1, ks);
          tc_commit(&s_full[1]);
          tc_commit(&k_empty[ks]);
          if (++ks == KSTAGES) { ks = 0; kph ^= 1; }
        } else {
          tc_commit(&o_full[1]);
        }
        if (++vs == VSTAGES) { vs = 0; vph ^= 1; }
        pph ^= 1;
      }
    }
  } else {
    const int idx = warp - 2;
    const int t = idx >> 3;
    const int hh = (idx & 7) >> 2;                  // column half handled by this warp
    const int q = warp & 3;                         // TMEM lane quarter (hardware: warp id % 4)
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_off + (t == 0 ? COL_S0 : COL_S1) + hh * 64;   // fp32 S columns of this half
    const uint32_t p_addr = tmem_base + lane_off + (t == 0 ? COL_S0 : COL_S1) + hh * 32;   // packed P columns of this half
    const uint32_t o_addr = tmem_base + lane_off + (t == 0 ? COL_O0 : COL_O1) + hh * 32;
    float* xt = xchg + t * (2 * 2 * 128);
    const float sl2 = p.scale_log2;
    const float rescale_thresh = 8.0f / sl2;
    float m_used = -INFINITY, l = 0.f;
    uint32_t sph = 0;

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[t], sph);
      sph ^= 1;
      tc_fence_after();
      uint32_t s[64];
      {
        uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
        uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
        tmem_ld_32x32b_x32(s_addr, s0);
        tmem_ld_32x32b_x32(s_addr + 32, s1);
      }
      tc_wait_ld();
      const int kv_valid = p.S - j * BKV - hh * 64;   // valid columns inside this half
      if (kv_valid < 64) {
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (c >= kv_valid) s[c] = 0xFF800000u;
      }
      float mx[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) mx[c] = __uint_as_float(s[c]);
#pragma unroll
      for (int c = 4; c < 64; ++c) mx[c & 3] = fmaxf(mx[c & 3], __uint_as_float(s[c]));
      const float pm = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      float* ex = xt + (j & 1) * 256;
      ex[hh * 128 + row_in_tile] = pm;
      named_bar_sync(1 + t, 256);          // both halves have read their S and published their partial max
      const float m_tile = fmaxf(pm, ex[(hh ^ 1) * 128 + row_in_tile]);
      const float m_new = fmaxf(m_used, m_tile);
      const bool need = (m_new - m_used) > rescale_thresh;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = fast_exp2((m_used - m_new) * sl2);
        l *= alpha;
        m_used = m_new;
        if (j > 0) {
          uint32_t o[32];
          tmem_ld_32x32b_x32(o_addr, o);
          tc_wait_ld();
#pragma unroll
          for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
          tmem_st_32x32b_x32(o_addr, o);
        }
      }
      const float neg_m = -m_used * sl2;
      float sum[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 64; c += 2) {
        const float p0 = fast_exp2(fmaf(__uint_as_float(s[c]), sl2, neg_m));
        const float p1 = fast_exp2(fmaf(__uint_as_float(s[c + 1]), sl2, neg_m));
        sum[(c >> 1) & 3] += p0 + p1;
        pk[c >> 1] = pack_bf16x2(p0, p1);
      }
      l += (sum[0] + sum[1]) + (sum[2] + sum[3]);
      tmem_st_32x32b_x32(p_addr, pk);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
    }

    mbar_wait(&o_full[t], 0);
    tc_fence_after();
    uint32_t o[32];
    tmem_ld_32x32b_x32(o_addr, o);
    tc_wait_ld();
    float* ex = xt + (n_kv & 1) * 256;      // a parity slot not used by the last iteration
    ex[hh * 128 + row_in_tile] = l;
    named_bar_sync(1 + t, 256);
    l += ex[(hh ^ 1) * 128 + row_in_tile];
    const int row = q0 + t * BQ + row_in_tile;
    if (row < p.S) {
      const float inv = 1.0f / l;
      __nv_bfloat16* dst = p.out + (int64_t(b) * p.S + row) * (int64_t(H) * DH) + h * DH + hh * 32;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(o[v * 8 + 0]) * inv, __uint_as_float(o[v * 8 + 1]) * inv);
        w.y = pack_bf16x2(__uint_as_float(o[v * 8 + 2]) * inv, __uint_as_float(o[v * 8 + 3]) * inv);
        w.z = pack_bf16x2(__uint_as_float(o[v * 8 + 4]) * inv, __uint_as_float(o[v * 8 + 5]) * inv);
        w.w = pack_bf16x2(__uint_as_float(o[v * 8 + 6]) * inv, __uint_as_float(o[v * 8 + 7]) * inv);
        reinterpret_cast<uint4*>(dst)[v] = w;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace attn2

int attention_v2_launch(const CUtensorMap& tm, int B, int S, int H, void* out, float scale_log2, cudaStream_t stream) {
  static SmemGrant grant;
  AETHER_CUDA_OK(ensure_dynamic_smem(grant, attn2::attention_v2_kernel, attn2::SMEM_BYTES));
  attn2::Params p;
  p.B = B; p.H = H; p.S = S;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.scale_log2 = scale_log2;
  dim3 grid((unsigned)ceil_div(S, 2 * attn2::BQ), (unsigned)H, (unsigned)B);
  attn2::attention_v2_kernel<<<grid, attn2::THREADS, attn2::SMEM_BYTES, stream>>>(tm, p);
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}

}  // namespace aether
