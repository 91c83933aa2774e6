// K1, variant 6 (mode 12 of aether_attention_bf16): THREE query tiles (384 rows) per CTA, 64-key tiles, twelve
// softmax warps -- three per scheduler instead of two.
//
// Why (ncu of mode 5 after the uniform-issue fix): the softmax warps no longer wait for the MMA warp, yet the MUFU --
// the binding pipe at head_dim 64 -- is only 77 % busy: with two warps per scheduler the XU idles whenever both are in
// their load / max / pack / store phases.  A third independent warp per scheduler fills those gaps, 64-column tiles
// keep the per-thread state at 64 scores so twelve softmax warps fit the register file (~90 of 128 registers), and
// 384-row CTAs turn the grid into 40 x 48 = 1920 CTAs = 12.97 waves of 148 SMs (mode 5: 19.1 -> 20 waves, 4 % tail).
//
// TMEM (480 of 512 columns), per query tile t at base t*160:   S_t +0 (64)   P_t +64 (32, bf16)   O_t +96 (64)
// smem: Q 3 x 16 KB, K ring 4 x 8 KB, V ring 4 x 8 KB (64-row boxes, 128B swizzle).
// warps: 0 = TMA producer, 1 = MMA issuer (+ TMEM alloc), 2..3 idle, 4..15 = softmax (tile (w-4)/4, lane quarter w%4).
// MMA issue order: QK_t(0) for t = 0..2, then per key tile j:  QK_0..2(j+1)  PV_0..2(j).
// barriers: s_full[t] MMA->softmax | s_free[t] softmax->MMA (S in registers) | p_full[t] softmax->MMA |
//           p_free[t] MMA->softmax (PV_t(j) retired: P_t and O_t may be touched) | o_full[t] (last PV retired).
#include "host_util.h"
#include "ptx.cuh"

namespace aether {
namespace attn6 {

constexpr int DH = 64, BQ = 128, NT = 3, BKV = 64, KSTAGES = 4, VSTAGES = 4;
constexpr int Q_BYTES = BQ * DH * 2, KV_BYTES = BKV * DH * 2;
constexpr int SMEM_BYTES = 1024 + NT * Q_BYTES + (KSTAGES + VSTAGES) * KV_BYTES + 512;
constexpr int THREADS = 512;
constexpr uint32_t COL_TILE = 160, COL_S = 0, COL_P = 64, COL_O = 96;

struct Params {
  int B, H, S;
  __nv_bfloat16* out;
  float scale_log2;
};

__global__ void __launch_bounds__(THREADS, 1)
attention_v6_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + NT * Q_BYTES;
  uint8_t* smem_v = smem_k + KSTAGES * KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + VSTAGES * KV_BYTES);
  uint64_t* q_full = bars;                  // [NT]
  uint64_t* k_full = q_full + NT;
  uint64_t* k_empty = k_full + KSTAGES;
  uint64_t* v_full = k_empty + KSTAGES;
  uint64_t* v_empty = v_full + VSTAGES;
  uint64_t* s_full = v_empty + VSTAGES;     // [NT]
  uint64_t* s_free = s_full + NT;           // count 128
  uint64_t* p_full = s_free + NT;           // count 128
  uint64_t* p_free = p_full + NT;
  uint64_t* o_full = p_free + NT;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(o_full + NT);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform (see gemm_tcgen05.cu)
  const int lane = threadIdx.x & 31;
  const int q_blk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = q_blk * NT * BQ;
  const int n_kv = (p.S + BKV - 1) / BKV;
  const int H = p.H;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    for (int i = 0; i < NT; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 128);
      mbar_init(&p_full[i], 128);
      mbar_init(&p_free[i], 1);
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
  if (warp == 1) tmem_alloc<512>(tmem_base_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_ptr, 0);

  if (warp < 4) {
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer (uniform control flow)
      const bool lead = elect_one();
      if (lead) {
        for (int t = 0; t < NT; ++t) {
          mbar_arrive_expect_tx(&q_full[t], Q_BYTES);
          tma_load_4d(smem_q + t * Q_BYTES, &tmap_qkv, &q_full[t], 0, h, q0 + t * BQ, b);
          tma_load_4d(smem_q + t * Q_BYTES + KV_BYTES, &tmap_qkv, &q_full[t], 0, h, q0 + t * BQ + BKV, b);
        }
      }
      __syncwarp();
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&k_empty[ks], kph ^ 1);
        if (lead) {
          mbar_arrive_expect_tx(&k_full[ks], KV_BYTES);
          tma_load_4d(smem_k + ks * KV_BYTES, &tmap_qkv, &k_full[ks], 0, H + h, j * BKV, b);
        }
        __syncwarp();
        if (++ks == KSTAGES) { ks = 0; kph ^= 1; }
        mbar_wait(&v_empty[vs], vph ^ 1);
        if (lead) {
          mbar_arrive_expect_tx(&v_full[vs], KV_BYTES);
          tma_load_4d(smem_v + vs * KV_BYTES, &tmap_qkv, &v_full[vs], 0, 2 * H + h, j * BKV, b);
        }
        __syncwarp();
        if (++vs == VSTAGES) { vs = 0; vph ^= 1; }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer: whole warp walks the schedule,
      // one elected lane issues (uniform control flow -> back-to-back UTCHMMA).
      const bool lead = elect_one();
      constexpr uint32_t idesc_qk = make_idesc_f16kind(BQ, BKV, 1, 1, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16kind(BQ, DH, 1, 1, 0, 1);
      auto issue_qk = [&](int t, int ks) {
        const uint64_t q_desc = make_sw128_desc(smem_u32(smem_q + t * Q_BYTES));
        const uint64_t k_desc = make_sw128_desc(smem_u32(smem_k + ks * KV_BYTES));
        const uint32_t d = tmem_base + t * COL_TILE + COL_S;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) tc_mma_ss(d, q_desc + 2 * k, k_desc + 2 * k, idesc_qk, k > 0);
      };
      auto issue_pv = [&](int t, int vs, bool first) {
        const uint64_t v_desc = make_sw128_desc(smem_u32(smem_v + vs * KV_BYTES));
        const uint32_t d = tmem_base + t * COL_TILE + COL_O;
        const uint32_t a = tmem_base + t * COL_TILE + COL_P;
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)
          tc_mma_ts(d, a + 8 * k, v_desc + 128 * k, idesc_pv, (!first || k > 0) ? 1u : 0u);
      };
      // In-order blocking schedule.  (An event-driven issuer -- per-tile program counters, mbarrier.test_wait polling,
      // staggered tile start -- was measured at 4.17 ms against 3.34 ms for this loop: the polling warp costs more
      // than the lock-step it removes, and the start stagger made no difference.)
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      mbar_wait(&k_full[0], 0);
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        mbar_wait(&q_full[t], 0);
        tc_fence_after();
        if (lead) {
          issue_qk(t, 0);
          tc_commit(&s_full[t]);
        }
        __syncwarp();
      }
      if (lead) tc_commit(&k_empty[0]);
      __syncwarp();
      ks = 1;
      for (int j = 0; j < n_kv; ++j) {
        const uint32_t ph = j & 1;
        const bool last = (j + 1 == n_kv);
        if (!last) {
          mbar_wait(&k_full[ks], kph);
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            mbar_wait(&s_free[t], ph);
            tc_fence_after();
            if (lead) {
              issue_qk(t, ks);
              tc_commit(&s_full[t]);
            }
            __syncwarp();
          }
          if (lead) tc_commit(&k_empty[ks]);
          __syncwarp();
          if (++ks == KSTAGES) { ks = 0; kph ^= 1; }
        }
        mbar_wait(&v_full[vs], vph);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          mbar_wait(&p_full[t], ph);
          tc_fence_after();
          if (lead) {
            issue_pv(t, vs, j == 0);
            tc_commit(&p_free[t]);
            if (last) tc_commit(&o_full[t]);
          }
          __syncwarp();
        }
        if (lead) tc_commit(&v_empty[vs]);
        __syncwarp();
        if (++vs == VSTAGES) { vs = 0; vph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps
    // no setmaxnreg: a softmax thread needs ~90 registers, below the 128 every thread of a 512-thread CTA gets
    const int t = (warp - 4) >> 2;
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t tile_addr = tmem_base + lane_off + t * COL_TILE;
    const uint32_t o_addr = tile_addr + COL_O;
    const float sl2 = p.scale_log2;
    const float rescale_thresh = 8.0f / sl2;
    float m_used = -INFINITY, l = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      uint32_t lo[32], hi[32];
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      tmem_ld_32x32b_x32(tile_addr + COL_S, lo);
      tmem_ld_32x32b_x32(tile_addr + COL_S + 32, hi);
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive(&s_free[t]);                      // S_t may be overwritten by QK_t(j+1) from now on
      const int kv_valid = p.S - j * BKV;
      if (kv_valid < BKV) {                         // last key tile only: TMA zero-filled keys must not contribute
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          if (c >= kv_valid) lo[c] = 0xFF800000u;
          if (c + 32 >= kv_valid) hi[c] = 0xFF800000u;
        }
      }
      float mx[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) mx[c] = fmaxf(__uint_as_float(lo[c]), __uint_as_float(hi[c]));
#pragma unroll
      for (int c = 4; c < 32; ++c) mx[c & 3] = fmaxf(mx[c & 3], fmaxf(__uint_as_float(lo[c]), __uint_as_float(hi[c])));
      const float m_new = fmaxf(m_used, fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])));
      const bool need = (m_new - m_used) > rescale_thresh;
      bool pv_prev_done = (j == 0);
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = fast_exp2((m_used - m_new) * sl2);
        l *= alpha;
        m_used = m_new;
        if (j > 0) {
          mbar_wait(&p_free[t], (j - 1) & 1);       // PV_t(j-1) retired: O_t is quiescent
          pv_prev_done = true;
          tc_fence_after();
#pragma unroll 1
          for (int ch = 0; ch < 4; ++ch) {          // rare path: 16 columns at a time keeps the register peak low
            uint32_t o[16];
            tmem_ld_32x32b_x16(o_addr + ch * 16, o);
            tc_wait_ld();
#pragma unroll
            for (int c = 0; c < 16; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
            tmem_st_32x32b_x16(o_addr + ch * 16, o);
          }
        }
      }
      const float neg_m = -m_used * sl2;
      float sum[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        const float p0 = fast_exp2(fmaf(__uint_as_float(lo[c]), sl2, neg_m));
        const float p1 = fast_exp2(fmaf(__uint_as_float(lo[c + 1]), sl2, neg_m));
        const float p2 = fast_exp2(fmaf(__uint_as_float(hi[c]), sl2, neg_m));
        const float p3 = fast_exp2(fmaf(__uint_as_float(hi[c + 1]), sl2, neg_m));
        sum[0] += p0;
        sum[1] += p1;
        sum[2] += p2;
        sum[3] += p3;
        pk[c >> 1] = pack_bf16x2(p0, p1);
        pk[16 + (c >> 1)] = pack_bf16x2(p2, p3);
      }
      l += (sum[0] + sum[1]) + (sum[2] + sum[3]);
      if (!pv_prev_done) {                          // P_t is still being read by PV_t(j-1) until p_free flips
        mbar_wait(&p_free[t], (j - 1) & 1);
        tc_fence_after();
      }
      tmem_st_32x32b_x32(tile_addr + COL_P, pk);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
    }

    mbar_wait(&o_full[t], 0);
    tc_fence_after();
    uint32_t o0[32], o1[32];
    tmem_ld_32x32b_x32(o_addr, o0);
    tmem_ld_32x32b_x32(o_addr + 32, o1);
    tc_wait_ld();
    const int row = q0 + t * BQ + row_in_tile;
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
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace attn6

int attention_v6_launch(const void* qkv, int B, int S, int H, void* out, float scale_log2, cudaStream_t stream) {
  static SmemGrant grant;
  AETHER_CUDA_OK(ensure_dynamic_smem(grant, attn6::attention_v6_kernel, attn6::SMEM_BYTES));
  CUtensorMap tm;
  const uint64_t dims[4] = {64, uint64_t(3 * H), uint64_t(S), uint64_t(B)};
  const uint64_t strides[3] = {128, uint64_t(3 * H) * 128, uint64_t(S) * uint64_t(3 * H) * 128};
  const uint32_t box[4] = {64, 1, uint32_t(attn6::BKV), 1};
  int rc = make_tmap_bf16(&tm, qkv, 4, dims, strides, box, true);
  if (rc) return rc;
  attn6::Params p;
  p.B = B; p.H = H; p.S = S;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.scale_log2 = scale_log2;
  dim3 grid((unsigned)ceil_div(S, attn6::NT * attn6::BQ), (unsigned)H, (unsigned)B);
  attn6::attention_v6_kernel<<<grid, attn6::THREADS, attn6::SMEM_BYTES, stream>>>(tm, p);
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}

}  // namespace aether
