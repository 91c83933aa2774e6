// K1, variant 3 (mode 5 of aether_attention_bf16): decoupled S / P buffers.
//
// Measurement-driven redesign of attention_tcgen05.cu (mode 0).  At head_dim 64 the MUFU (16 ex2 per clock per SM) needs
// 1024 clk per 128x128 score tile against 512 clk of tensor work, so the softmax warps must never wait for the MMA
// warp.  (This header first also blamed the TMEM -> register path; the ncu counter smsp__mem_tensor_reads_op_ldt later
// showed tcgen05.ld at 5 % of its bandwidth -- see DESIGN.md "Attention roofline".)  In mode 0 P aliases S, so
// S_t(j+1) can only be produced after PV_t(j): every softmax warpgroup idles for a full MMA round trip per key tile.
// Here P gets its own TMEM columns
//     S0 [0,128)  S1 [128,256)  P0 [256,320)  P1 [320,384)  O0 [384,448)  O1 [448,512)      (all 512 columns)
// and the softmax warps release S_t as soon as it sits in registers (s_free), so the MMA warp issues QK_t(j+1)
// DURING the exp phase of tile t.  Per warpgroup the steady state is  load S -> row max -> exp/pack/store
// (MUFU-bound) with no MMA wait; the issue order below starts the two warpgroups half a period apart.
//   MMA issue order per key tile j:  QK_0(j+1) | PV_0(j) | QK_1(j+1) | PV_1(j)
//   barriers:  s_full (MMA->softmax)  s_free (softmax->MMA, 128 arrivals)  p_full (softmax->MMA, 128)
//              p_free (MMA->softmax: PV_t(j) retired => P_t and O_t may be touched again)  o_full (last PV retired)
#include "host_util.h"
#include "ptx.cuh"

namespace aether {
namespace attn3 {

constexpr int DH = 64, BQ = 128, BKV = 128, KSTAGES = 3, VSTAGES = 3;
constexpr int TILE_BYTES = BQ * DH * 2;
constexpr int SMEM_BYTES = 1024 + (2 + KSTAGES + VSTAGES) * TILE_BYTES + 256;
constexpr int THREADS = 384;
constexpr uint32_t COL_S0 = 0, COL_S1 = 128, COL_P0 = 256, COL_P1 = 320, COL_O0 = 384, COL_O1 = 448;

struct Params {
  int B, H, S;
  __nv_bfloat16* out;
  float scale_log2;
  // Work decomposition (1-D grid).  CTA c works on item = item_begin + c / kv_splits, an item being one 256-row query
  // block of one head (q_blk fastest, then head, then batch); with kv_splits > 1 it only visits key tiles
  // [split * kv_chunk, (split + 1) * kv_chunk) and leaves an UNNORMALISED partial (O fp32, row max, row sum) in
  // part_o / part_ml for attention_combine_kernel -- the tail wave of the grid is cut along the keys so that it fills
  // every SM instead of 20 of 148 (see the launcher).
  int q_blocks, item_begin, kv_splits, kv_chunk;
  float* part_o;       // [ctas][256][64]
  float* part_ml;      // [ctas][256][2]
};

// Round-1 variants of this kernel (per-warp pipelined softmax, 12.5-37.5 % of the exponentials on the FMA pipe,
// interleaved consumers, a split schedule for the ragged rows) measured equal or slower and live on, unbuilt, under
// tools/experiments/attention/ together with the 16-softmax-warp, 1-tile-per-CTA, 64-key-tile and 3-query-tile kernels.
// Round 2: the scale (s * sl2 - m) and row-sum math run on fma.rn.f32x2 / add.rn.f32x2, two elements per issue slot
// (SASS per 128-key tile and row: 128 MUFU.EX2, 64 FFMA2, 64 FADD2, 64 FMNMX3, 64 F2FP instead of 128 FFMA + 128 FADD):
// 3.317 -> 3.159 ms isolated, 3.850 -> 3.752 ms inside the power-capped step.  Tried and dropped in the same experiment
// (profiles/r2_attn_experiments_lazy_packed.log): starting the exponentials of tile j with the stale offset of tile j-1
// and taking the tile maximum off the critical path ("lazy max", with a redo from registers for pathological jumps) --
// 3.93 ms alone, 3.70 ms combined with the packed math: S has to stay live in registers next to P; and, on top of the
// packed math, 12.5 % / 25 % of the exponentials as the FMA-pipe polynomial (profiles/r2_attn_experiments_packed_poly.log):
// 3.165 / 3.126 ms isolated against 3.160, no difference inside the power-capped step.
__global__ void __launch_bounds__(THREADS, 1)
attention_v3_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + 2 * TILE_BYTES;
  uint8_t* smem_v = smem_k + KSTAGES * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + VSTAGES * TILE_BYTES);
  uint64_t* q_full = bars;                  // [2]
  uint64_t* k_full = q_full + 2;
  uint64_t* k_empty = k_full + KSTAGES;
  uint64_t* v_full = k_empty + KSTAGES;
  uint64_t* v_empty = v_full + VSTAGES;
  uint64_t* s_full = v_empty + VSTAGES;     // [2]
  uint64_t* s_free = s_full + 2;            // [2] count 128
  uint64_t* p_full = s_free + 2;            // [2] count 128
  uint64_t* p_free = p_full + 2;            // [2]
  uint64_t* o_full = p_free + 2;            // [2]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform: uniform-datapath MMA issue
  const int lane = threadIdx.x & 31;
  const int H = p.H;
  const int item = p.item_begin + int(blockIdx.x) / p.kv_splits;
  const int split = int(blockIdx.x) % p.kv_splits;
  const int q_blk = item % p.q_blocks, h = (item / p.q_blocks) % H, b = item / (p.q_blocks * H);
  const int q0 = q_blk * 2 * BQ;
  const int n_kv_total = (p.S + BKV - 1) / BKV;
  const int kv_t0 = p.kv_splits > 1 ? split * p.kv_chunk : 0;                    // first key tile of this CTA
  const int n_kv = p.kv_splits > 1 ? min(p.kv_chunk, n_kv_total - kv_t0) : n_kv_total;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    for (int i = 0; i < 2; ++i) {
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
  if (warp == 2) tmem_alloc<512>(tmem_base_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_ptr, 0);

  if (warp < 4) {
    setmaxnreg_dec<48>();
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer (uniform control flow)
      const bool lead = elect_one();
      if (lead) {
        for (int t = 0; t < 2; ++t) {
          mbar_arrive_expect_tx(&q_full[t], TILE_BYTES);
          tma_load_4d(smem_q + t * TILE_BYTES, &tmap_qkv, &q_full[t], 0, h, q0 + t * BQ, b);
        }
      }
      __syncwarp();
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&k_empty[ks], kph ^ 1);
        if (lead) {
          mbar_arrive_expect_tx(&k_full[ks], TILE_BYTES);
          tma_load_4d(smem_k + ks * TILE_BYTES, &tmap_qkv, &k_full[ks], 0, H + h, (kv_t0 + j) * BKV, b);
        }
        __syncwarp();
        if (++ks == KSTAGES) { ks = 0; kph ^= 1; }
        mbar_wait(&v_empty[vs], vph ^ 1);
        if (lead) {
          mbar_arrive_expect_tx(&v_full[vs], TILE_BYTES);
          tma_load_4d(smem_v + vs * TILE_BYTES, &tmap_qkv, &v_full[vs], 0, 2 * H + h, (kv_t0 + j) * BKV, b);
        }
        __syncwarp();
        if (++vs == VSTAGES) { vs = 0; vph ^= 1; }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer: whole warp walks the schedule
      // (uniform control flow -> plain UTCHMMA, no per-instruction ELECT waterfall); one elected lane issues.
      const bool lead = elect_one();
      constexpr uint32_t idesc_qk = make_idesc_f16kind(BQ, BKV, 1, 1, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16kind(BQ, DH, 1, 1, 0, 1);
      const uint32_t s_col[2] = {tmem_base + COL_S0, tmem_base + COL_S1};
      const uint32_t p_col[2] = {tmem_base + COL_P0, tmem_base + COL_P1};
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
          tc_mma_ts(o_col[t], p_col[t] + 8 * k, v_desc + 128 * k, idesc_pv, (!first || k > 0) ? 1u : 0u);
      };
      // Skewed schedule: tile 1 runs HALF A PERIOD behind tile 0, so that one warpgroup is in its TMEM-load phase
      // while the other is in its MUFU phase (both pipes deliver 16 elements/clk; in phase they would queue on
      // the same pipe).  Per key tile j the MMA thread alternates
      //   A (tile 0 has S_0(j) in registers):  QK_0(j+1)            then  PV_1(j-1)   [j == 0: the late QK_1(0)]
      //   B (tile 1 has S_1(j) in registers):  QK_1(j+1)            then  PV_0(j)
      // K(j+1) is released after QK_1(j+1), V(j) after PV_1(j) (one iteration later).
      int ks = 0;                 // stage of K(j+1) inside the loop
      uint32_t kph = 0;
      int vs0 = 0, vs1 = 0;       // V stage of PV_0(j) / PV_1(j-1)
      uint32_t vph0 = 0;
      mbar_wait(&k_full[0], 0);
      mbar_wait(&q_full[0], 0);
      tc_fence_after();
      if (lead) {
        issue_qk(0, 0);
        tc_commit(&s_full[0]);
      }
      __syncwarp();
      ks = 1;
      if (ks == KSTAGES) { ks = 0; kph ^= 1; }
      for (int j = 0; j < n_kv; ++j) {
        const bool last = (j + 1 == n_kv);
        const uint32_t ph = j & 1;
        // ---- phase A
        if (!last) {
          mbar_wait(&k_full[ks], kph);
          mbar_wait(&s_free[0], ph);
          tc_fence_after();
          if (lead) {
            issue_qk(0, ks);
            tc_commit(&s_full[0]);
          }
          __syncwarp();
        }
        if (j == 0) {
          mbar_wait(&q_full[1], 0);
          tc_fence_after();
          if (lead) {
            issue_qk(1, 0);                    // K(0) is still resident: stage 0
            tc_commit(&s_full[1]);
            tc_commit(&k_empty[0]);
          }
          __syncwarp();
        } else {
          mbar_wait(&p_full[1], (j - 1) & 1);
          tc_fence_after();
          if (lead) {
            issue_pv(1, vs1, j == 1);
            tc_commit(&p_free[1]);
            tc_commit(&v_empty[vs1]);
          }
          __syncwarp();
          if (++vs1 == VSTAGES) vs1 = 0;
        }
        // ---- phase B
        if (!last) {
          mbar_wait(&s_free[1], ph);
          tc_fence_after();
          if (lead) {
            issue_qk(1, ks);
            tc_commit(&s_full[1]);
            tc_commit(&k_empty[ks]);
          }
          __syncwarp();
          if (++ks == KSTAGES) { ks = 0; kph ^= 1; }
        }
        mbar_wait(&v_full[vs0], vph0);
        mbar_wait(&p_full[0], ph);
        tc_fence_after();
        if (lead) {
          issue_pv(0, vs0, j == 0);
          tc_commit(&p_free[0]);
          if (last) tc_commit(&o_full[0]);
        }
        __syncwarp();
        if (++vs0 == VSTAGES) { vs0 = 0; vph0 ^= 1; }
      }
      // drain: PV_1(n-1)
      mbar_wait(&p_full[1], (n_kv - 1) & 1);
      tc_fence_after();
      if (lead) {
        issue_pv(1, vs1, n_kv == 1);
        tc_commit(&p_free[1]);
        tc_commit(&v_empty[vs1]);
        tc_commit(&o_full[1]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    setmaxnreg_inc<224>();
    const int t = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_off + (t == 0 ? COL_S0 : COL_S1);
    const uint32_t p_addr = tmem_base + lane_off + (t == 0 ? COL_P0 : COL_P1);
    const uint32_t o_addr = tmem_base + lane_off + (t == 0 ? COL_O0 : COL_O1);
    const float sl2 = p.scale_log2;
    const float rescale_thresh = 8.0f / sl2;
    float m_used = -INFINITY, l = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[t], j & 1);
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
      mbar_arrive(&s_free[t]);                     // S_t may be overwritten by QK_t(j+1) from now on
      const int kv_valid = p.S - (kv_t0 + j) * BKV;
      if (kv_valid < BKV) {
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if (c >= kv_valid) s[c] = 0xFF800000u;
      }
      // row maximum of the tile (3-input max tree), needed before the first exponential
      float mx[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) mx[c] = __uint_as_float(s[c]);
#pragma unroll
      for (int c = 8; c < 128; ++c) mx[c & 7] = fmaxf(mx[c & 7], __uint_as_float(s[c]));
      const float m_tile =
          fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
      bool pv_prev_done = (j == 0);
      auto scale_O = [&](float alpha) {            // O_t *= alpha (O_t must be quiescent: PV_t(j-1) retired)
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
      };
      auto wait_pv_prev = [&]() {
        if (!pv_prev_done) {
          mbar_wait(&p_free[t], (j - 1) & 1);      // PV_t(j-1) retired: O_t quiescent, P_t writable
          tc_fence_after();
          pv_prev_done = true;
        }
      };
      {
        const float m_new = fmaxf(m_used, m_tile);
        const bool need = (m_new - m_used) > rescale_thresh;
        if (__any_sync(0xffffffffu, need)) {       // lazy rescale: only when some row's maximum grew by more than 2^8
          const float alpha = fast_exp2((m_used - m_new) * sl2);
          l *= alpha;
          m_used = m_new;
          if (j > 0) {
            wait_pv_prev();
            scale_O(alpha);
          }
        }
      }
      const float neg_m = -m_used * sl2;
      uint32_t pk[64];
      // P = 2^(s * sl2 + neg_m) packed to bf16, row sum in fp32 -- packed f32x2 math (pairs of adjacent keys)
      const uint64_t sl2x2 = pack_f32x2(sl2, sl2), negx2 = pack_f32x2(neg_m, neg_m);
      uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};               // four (0.0f, 0.0f) accumulator pairs
#pragma unroll
      for (int c = 0; c < 128; c += 2) {
        float x0, x1;
        unpack_f32x2(ffma2(pack_f32x2(__uint_as_float(s[c]), __uint_as_float(s[c + 1])), sl2x2, negx2), x0, x1);
        const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
        acc[(c >> 1) & 3] = fadd2(acc[(c >> 1) & 3], pack_f32x2(p0, p1));
        pk[c >> 1] = pack_bf16x2(p0, p1);
      }
      float tile_sum;
      {
        float a0, a1;
        unpack_f32x2(fadd2(fadd2(acc[0], acc[1]), fadd2(acc[2], acc[3])), a0, a1);
        tile_sum = a0 + a1;
      }
      l += tile_sum;
      wait_pv_prev();                              // P_t is still being read by PV_t(j-1) until p_free flips
      {
        const uint32_t(&p0)[32] = *reinterpret_cast<const uint32_t(*)[32]>(&pk[0]);
        const uint32_t(&p1)[32] = *reinterpret_cast<const uint32_t(*)[32]>(&pk[32]);
        tmem_st_32x32b_x32(p_addr, p0);
        tmem_st_32x32b_x32(p_addr + 32, p1);
      }
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
    if (p.kv_splits > 1) {
      // partial result of this key range: O (relative to the offset m_used), m_used and l, combined later
      float4* po = reinterpret_cast<float4*>(p.part_o + (int64_t(blockIdx.x) * 2 * BQ + t * BQ + row_in_tile) * DH);
#pragma unroll
      for (int v = 0; v < 8; ++v)
        po[v] = make_float4(__uint_as_float(o0[4 * v]), __uint_as_float(o0[4 * v + 1]), __uint_as_float(o0[4 * v + 2]),
                            __uint_as_float(o0[4 * v + 3]));
#pragma unroll
      for (int v = 0; v < 8; ++v)
        po[8 + v] = make_float4(__uint_as_float(o1[4 * v]), __uint_as_float(o1[4 * v + 1]), __uint_as_float(o1[4 * v + 2]),
                                __uint_as_float(o1[4 * v + 3]));
      *reinterpret_cast<float2*>(p.part_ml + (int64_t(blockIdx.x) * 2 * BQ + t * BQ + row_in_tile) * 2) =
          make_float2(m_used, l);
    } else if (row < p.S) {
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
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}


// out[row, h*64 + c] = sum_i O_i[c] * w_i / sum_i l_i * w_i,  w_i = 2^((m_i - max_j m_j) * scale_log2): the flash-attention
// merge of the `splits` key-range partials of one item.  One warp per query row, two columns per lane.
__global__ void __launch_bounds__(256)
attention_combine_kernel(const Params p, int n_items) {
  const int lane = threadIdx.x & 31;
  const int wrow = (blockIdx.x * 256 + threadIdx.x) >> 5;           // global (item, row) index
  if (wrow >= n_items * 2 * BQ) return;
  const int it = wrow / (2 * BQ), r = wrow - it * (2 * BQ);
  const int item = p.item_begin + it;
  const int q_blk = item % p.q_blocks, h = (item / p.q_blocks) % p.H, b = item / (p.q_blocks * p.H);
  const int row = q_blk * 2 * BQ + r;
  if (row >= p.S) return;
  float m = -INFINITY;
  for (int i = 0; i < p.kv_splits; ++i)
    m = fmaxf(m, p.part_ml[((int64_t(it) * p.kv_splits + i) * 2 * BQ + r) * 2]);
  float l = 0.f, a0 = 0.f, a1 = 0.f;
  for (int i = 0; i < p.kv_splits; ++i) {
    const int64_t base = (int64_t(it) * p.kv_splits + i) * 2 * BQ + r;
    const float2 ml = *reinterpret_cast<const float2*>(p.part_ml + base * 2);
    const float w = fast_exp2((ml.x - m) * p.scale_log2);
    const float2 o = *reinterpret_cast<const float2*>(p.part_o + base * DH + 2 * lane);
    l = fmaf(ml.y, w, l);
    a0 = fmaf(o.x, w, a0);
    a1 = fmaf(o.y, w, a1);
  }
  const float inv = 1.0f / l;
  __nv_bfloat16* dst = p.out + (int64_t(b) * p.S + row) * (int64_t(p.H) * DH) + h * DH + 2 * lane;
  *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(a0 * inv, a1 * inv);
}

}  // namespace attn3

// Scratch for the split tail: (items % #SM) x splits CTAs x 256 rows x (64 + 2) floats; 0 when the grid divides evenly.
int64_t attention_v3_workspace_bytes(int B, int S, int H) {
  const int nsm = num_sms();
  if (nsm <= 0) return 0;
  const int64_t items = ceil_div(S, 2 * attn3::BQ) * int64_t(H) * B;
  const int rem = int(items % nsm);
  if (items <= nsm || rem == 0) return 0;
  const int splits = nsm / rem;
  if (splits < 2) return 0;
  return int64_t(rem) * splits * 2 * attn3::BQ * (attn3::DH + 2) * 4 + 256;
}

// One launch covers the items that fill whole waves (one CTA per SM, so a "wave" is #SM items); the remaining
// items % #SM would occupy that many SMs for a full item time while the others idle -- at S = 15076, H = 48 that is 20 of
// 148 SMs for the 20th of 19.14 waves (4.3 % of the kernel).  With scratch they run in a second launch cut along the keys
// into floor(#SM / rem) ranges each (7 x 17 key tiles) and a small merge kernel, so the tail costs ~1/7 of an item time.
int attention_v3_launch(const CUtensorMap& tm, int B, int S, int H, void* out, float scale_log2, void* workspace,
                        int64_t workspace_bytes, cudaStream_t stream) {
  static SmemGrant grant;
  AETHER_CUDA_OK(ensure_dynamic_smem(grant, attn3::attention_v3_kernel, attn3::SMEM_BYTES));
  const auto kern = attn3::attention_v3_kernel;
  attn3::Params p;
  p.B = B; p.H = H; p.S = S;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.scale_log2 = scale_log2;
  p.q_blocks = (int)ceil_div(S, 2 * attn3::BQ);
  p.item_begin = 0; p.kv_splits = 1; p.kv_chunk = 0; p.part_o = nullptr; p.part_ml = nullptr;
  const int64_t items = int64_t(p.q_blocks) * H * B;
  const int nsm = num_sms();
  const int n_kv = (int)ceil_div(S, attn3::BKV);
  const int rem = int(items % nsm);
  int splits = (items > nsm && rem != 0) ? nsm / rem : 1;
  if (splits > n_kv) splits = n_kv;
  const int64_t need = int64_t(rem) * splits * 2 * attn3::BQ * (attn3::DH + 2) * 4 + 256;
  if (splits < 2 || workspace == nullptr || workspace_bytes < need) {
    kern<<<(unsigned)items, attn3::THREADS, attn3::SMEM_BYTES, stream>>>(tm, p);
    AETHER_CUDA_OK(cudaGetLastError());
    return AETHER_OK;
  }
  const int64_t full = items - rem;
  kern<<<(unsigned)full, attn3::THREADS, attn3::SMEM_BYTES, stream>>>(tm, p);
  attn3::Params q = p;
  q.item_begin = (int)full;
  q.kv_splits = splits;
  q.kv_chunk = (int)ceil_div(n_kv, splits);
  if (int64_t(q.kv_chunk) * (splits - 1) >= n_kv) {          // the last range must not be empty
    q.kv_splits = splits = (int)ceil_div(n_kv, q.kv_chunk);
  }
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  q.part_o = reinterpret_cast<float*>(base);
  q.part_ml = q.part_o + int64_t(rem) * splits * 2 * attn3::BQ * attn3::DH;
  kern<<<(unsigned)(rem * splits), attn3::THREADS, attn3::SMEM_BYTES, stream>>>(tm, q);
  const int warps = rem * 2 * attn3::BQ;
  attn3::attention_combine_kernel<<<(unsigned)ceil_div(int64_t(warps) * 32, 256), 256, 0, stream>>>(q, rem);
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}

}  // namespace aether
