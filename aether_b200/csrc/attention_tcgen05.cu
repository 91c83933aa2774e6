// K1: joint text+video self-attention  O = softmax(Q K^T / sqrt(dh)) V,  dh = 64,
// non-causal, no mask, B x H heads x S tokens (S = 226 + 11*30*45 = 15076 at 41 frames 480x720).
//
// Replaces F.scaled_dot_product_attention inside diffusers' CogVideoXAttnProcessor2_0, which the
// reference reaches through aether/pipelines/aetherv1_pipeline_cogvideox.py:865-875 (42 layers x steps).
// Q and K arrive already QK-LayerNorm'ed and RoPE'd (qk_norm_rope kernel); V is read in place from the
// fused QKV GEMM output.  All three live in one buffer  qkv[B, S, 3, H, 64]  (16-bit elements) addressed
// by a single 4-D TMA tensor map {64, 3H, S, B}; O is written token-major  out[B, S, H*64]  (bf16) so it
// feeds to_out directly.
//
// CTA = 256 query rows (two 128-row tiles) of one (b, h); 12 warps:
//   warp 0      TMA producer: Q tiles once, then a 3-stage K ring and a 3-stage V ring (16 KB boxes)
//   warp 1      MMA issuer  : S_t = Q_t K_j^T   (tcgen05.mma SS, M=128 N=128 K=64 -> 128 fp32 TMEM columns)
//                             O_t += P_t V_j    (tcgen05.mma TS: A = P_t in TMEM (16-bit, aliases S_t),
//                                                B = V_j shared memory MN-major, M=128 N=64 K=128)
//               issue order PV_0(j) QK_0(j+1) PV_1(j) QK_1(j+1): the two query tiles ping-pong so the
//               tensor pipe works on one while the other is in softmax.
//   warps 4..7  softmax warpgroup for tile 0, warps 8..11 for tile 1: one thread per query row
//               (tcgen05.ld 32x32b => no shuffles), exp2 with the scale folded in, lazy O rescale
//               (only when the running max grew by > 2^8), P packed to 16 bit and tcgen05.st back.
//
// Two arithmetic modes (template F16PV):
//   false: P = bf16, V = bf16; exp2 = one fp32 MUFU op per element; row sums on the CUDA cores.
//   true : P = fp16, V = fp16 (the QKV GEMM epilogue emits the V third as fp16).  The kernel is bound by the
//          MUFU (16 ex2/clk/SM: 1024 clk per 128x128 tile against 512 clk of tensor work), so this mode
//          (a) computes two exponentials per MUFU op with ex2.approx.f16x2 -- whose packed output IS the P
//          operand, no separate convert -- after a packed fma.rn.f32x2 scale/subtract, and (b) moves the row
//          sums to the tensor core: L_t += P_t * Ones, a 16-column N-tile whose B operand is a constant
//          2 KB shared-memory tile (row 0 = 1.0), so l is the fp32 sum of exactly the P values used for O.
// TMEM columns: S0/P0 [0,128) S1/P1 [128,256) O0 [256,320) O1 [320,384) L0 [384,400) L1 [400,416) of 512.
// Roofline: tensor;  algorithmic flop / launch = 4 * B * H * S^2 * 64.
#include "host_util.h"
#include "ptx.cuh"

namespace aether {
namespace attn {

constexpr int DH = 64;
constexpr int BQ = 128;          // rows per query tile (2 tiles per CTA)
constexpr int BKV = 128;         // keys per tile
constexpr int KSTAGES = 3, VSTAGES = 3;
constexpr int TILE_BYTES = BQ * DH * 2;   // 16 KB (same for Q, K, V tiles)
constexpr int ONES_BYTES = 2048;          // 16 rows x 128 B, K-major SW128 (row 0 = fp16 ones)
constexpr int SMEM_BYTES = 1024 + (2 + KSTAGES + VSTAGES) * TILE_BYTES + ONES_BYTES + 256;
constexpr int THREADS = 384;

constexpr uint32_t COL_S0 = 0, COL_S1 = 128, COL_O0 = 256, COL_O1 = 320, COL_L0 = 384, COL_L1 = 400;

struct Params {
  int B, H, S;
  __nv_bfloat16* out;       // [B, S, H*64]
  float scale_log2;         // dh^-0.5 * log2(e)
};

template <int MODE>   // 0: bf16 P/V   (1: fp16 P/V, MUFU exp -- not built)   2: fp16 P/V, 40% of the exponentials on the FMA pipe
__global__ void __launch_bounds__(THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const Params p) {
  constexpr bool F16PV = (MODE == 1 || MODE == 2);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                               // 2 tiles
  uint8_t* smem_k = smem + 2 * TILE_BYTES;              // KSTAGES tiles
  uint8_t* smem_v = smem_k + KSTAGES * TILE_BYTES;      // VSTAGES tiles
  uint8_t* smem_ones = smem_v + VSTAGES * TILE_BYTES;   // 2 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_ones + ONES_BYTES);
  uint64_t* q_full = bars;                  // [2]
  uint64_t* k_full = q_full + 2;            // [KSTAGES]
  uint64_t* k_empty = k_full + KSTAGES;     // [KSTAGES]
  uint64_t* v_full = k_empty + KSTAGES;     // [VSTAGES]
  uint64_t* v_empty = v_full + VSTAGES;     // [VSTAGES]
  uint64_t* s_full = v_empty + VSTAGES;     // [2]  MMA -> softmax  (S_t ready; also implies PV_t(j-1) retired)
  uint64_t* p_full = s_full + 2;            // [2]  softmax -> MMA  (P_t written, O_t rescaled)   count 128
  uint64_t* o_full = p_full + 2;            // [2]  MMA -> softmax  (last PV_t retired)
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform (see gemm_tcgen05.cu)
  const int lane = threadIdx.x & 31;
  const int q_blk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = q_blk * 2 * BQ;
  const int n_kv = (p.S + BKV - 1) / BKV;
  const int H = p.H;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
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
  if (F16PV && warp == 3) {
    // constant B operand of the row-sum MMA: 16 (N) x 64 (K) fp16, K-major, row 0 = 1.0, rows 1..15 = 0.
    // The 128B swizzle only permutes 16-byte chunks inside a row, so a constant row needs no swizzling.
    uint32_t* o32 = reinterpret_cast<uint32_t*>(smem_ones);
    for (int i = lane; i < ONES_BYTES / 4; i += 32) o32[i] = (i < 32) ? 0x3C003C00u : 0u;
    fence_proxy_async_smem();
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
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer: the whole warp walks the schedule
      // (uniform control flow -> back-to-back UTCHMMA, see gemm_tcgen05.cu), one elected lane issues.
      const bool lead = elect_one();
      constexpr uint32_t PV_FMT = F16PV ? 0u : 1u;                                      // 0 = F16, 1 = BF16
      constexpr uint32_t idesc_qk = make_idesc_f16kind(BQ, BKV, 1, 1, 0, 0);            // Q, K bf16, both K-major
      constexpr uint32_t idesc_pv = make_idesc_f16kind(BQ, DH, PV_FMT, PV_FMT, 0, 1);   // P (TMEM), V MN-major
      constexpr uint32_t idesc_l = make_idesc_f16kind(BQ, 16, 0, 0, 0, 0);              // P (TMEM), Ones K-major
      const uint32_t s_col[2] = {tmem_base + COL_S0, tmem_base + COL_S1};
      const uint32_t o_col[2] = {tmem_base + COL_O0, tmem_base + COL_O1};
      const uint32_t l_col[2] = {tmem_base + COL_L0, tmem_base + COL_L1};
      const uint64_t ones_desc = make_sw128_desc(smem_u32(smem_ones));
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
        for (int k = 0; k < BKV / 16; ++k) {  // 16 keys per step: P advances 8 TMEM columns, V 16 rows = 2048 B
          const uint32_t acc = (!first || k > 0) ? 1u : 0u;
          tc_mma_ts(o_col[t], s_col[t] + 8 * k, v_desc + 128 * k, idesc_pv, acc);
          if (F16PV) tc_mma_ts(l_col[t], s_col[t] + 8 * k, ones_desc, idesc_l, acc);
        }
      };

      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0, pph = 0;
      // prologue: S_0(0), S_1(0)
      mbar_wait(&k_full[0], 0);
      mbar_wait(&q_full[0], 0);
      tc_fence_after();
      if (lead) {
        issue_qk(0, 0);
        tc_commit(&s_full[0]);
      }
      __syncwarp();
      mbar_wait(&q_full[1], 0);
      tc_fence_after();
      if (lead) {
        issue_qk(1, 0);
        tc_commit(&s_full[1]);
        tc_commit(&k_empty[0]);
      }
      __syncwarp();
      ks = 1;
      if (ks == KSTAGES) { ks = 0; kph ^= 1; }
      for (int j = 0; j < n_kv; ++j) {
        const bool last = (j + 1 == n_kv);
        mbar_wait(&v_full[vs], vph);
        // ---- tile 0
        mbar_wait(&p_full[0], pph);
        tc_fence_after();
        if (lead) issue_pv(0, vs, j == 0);
        __syncwarp();
        if (!last) {
          mbar_wait(&k_full[ks], kph);
          tc_fence_after();
          if (lead) {
            issue_qk(0, ks);
            tc_commit(&s_full[0]);
          }
        } else {
          if (lead) tc_commit(&o_full[0]);
        }
        __syncwarp();
        // ---- tile 1
        mbar_wait(&p_full[1], pph);
        tc_fence_after();
        if (lead) {
          issue_pv(1, vs, j == 0);
          tc_commit(&v_empty[vs]);
          if (!last) {
            issue_qk(1, ks);
            tc_commit(&s_full[1]);
            tc_commit(&k_empty[ks]);
          } else {
            tc_commit(&o_full[1]);
          }
        }
        __syncwarp();
        if (!last) {
          if (++ks == KSTAGES) { ks = 0; kph ^= 1; }
        }
        if (++vs == VSTAGES) { vs = 0; vph ^= 1; }
        pph ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    setmaxnreg_inc<224>();
    const int t = (warp - 4) >> 2;                 // query tile 0 / 1
    const int q = warp & 3;                        // TMEM lane quarter
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_off + (t == 0 ? COL_S0 : COL_S1);
    const uint32_t o_addr = tmem_base + lane_off + (t == 0 ? COL_O0 : COL_O1);
    const uint32_t l_addr = tmem_base + lane_off + (t == 0 ? COL_L0 : COL_L1);
    const float sl2 = p.scale_log2;
    const float rescale_thresh = 8.0f / sl2;       // in raw-score units
    float m_used = -INFINITY;                      // reference max currently baked into O and l
    float l = 0.f;                                 // row sum (F16PV: kept in TMEM by the Ones MMA instead)
    uint32_t sph = 0;

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[t], sph);
      sph ^= 1;
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
      const int kv_valid = p.S - j * BKV;          // >= 128 except on the ragged last tile
      if (kv_valid < BKV) {
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if (c >= kv_valid) s[c] = 0xFF800000u;   // -inf: TMA zero-filled keys must not contribute
      }
      // 8 independent max chains (only 2 softmax warps per scheduler: dependent-chain latency is exposed)
      float mx[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) mx[c] = __uint_as_float(s[c]);
#pragma unroll
      for (int c = 8; c < 128; ++c) mx[c & 7] = fmaxf(mx[c & 7], __uint_as_float(s[c]));
      const float m_tile =
          fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
      const float m_new = fmaxf(m_used, m_tile);
      const bool need = (m_new - m_used) > rescale_thresh;     // true on the first tile (m_used = -inf)
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = fast_exp2((m_used - m_new) * sl2);     // 0 on the first tile
        l *= alpha;
        m_used = m_new;
        if (j > 0) {
          uint32_t o0[32], o1[32];
          tmem_ld_32x32b_x32(o_addr, o0);
          tmem_ld_32x32b_x32(o_addr + 32, o1);
          uint32_t lv = 0;
          if (F16PV) tmem_ld_32x32b_x1(l_addr, lv);
          tc_wait_ld();
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            o0[c] = __float_as_uint(__uint_as_float(o0[c]) * alpha);
            o1[c] = __float_as_uint(__uint_as_float(o1[c]) * alpha);
          }
          tmem_st_32x32b_x32(o_addr, o0);
          tmem_st_32x32b_x32(o_addr + 32, o1);
          if (F16PV) tmem_st_32x32b_x1(l_addr, __float_as_uint(__uint_as_float(lv) * alpha));
        }
      }
      const float neg_m = -m_used * sl2;
      uint32_t pk[64];
      if (F16PV) {
        const uint64_t sl2x2 = pack_f32x2(sl2, sl2), negx2 = pack_f32x2(neg_m, neg_m);
#pragma unroll
        for (int c = 0; c < 128; c += 2) {
          const uint64_t x2 = ffma2(pack_f32x2(__uint_as_float(s[c]), __uint_as_float(s[c + 1])), sl2x2, negx2);
          // MODE 2: pairs 1 and 3 of every 5 (40 %) bypass the MUFU (FA4-style software exp2)
          const bool poly = (MODE == 2) && (((c >> 1) % 5 == 1) || ((c >> 1) % 5 == 3));
          pk[c >> 1] = poly ? exp2_poly_f16x2_from_f32x2(x2) : exp2_f16x2_from_f32x2(x2);
        }
      } else {
        float sum[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) sum[c] = 0.f;
#pragma unroll
        for (int c = 0; c < 128; c += 2) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(s[c]), sl2, neg_m));
          const float p1 = fast_exp2(fmaf(__uint_as_float(s[c + 1]), sl2, neg_m));
          sum[c & 7] += p0;
          sum[(c + 1) & 7] += p1;
          pk[c >> 1] = pack_bf16x2(p0, p1);
        }
        l += ((sum[0] + sum[1]) + (sum[2] + sum[3])) + ((sum[4] + sum[5]) + (sum[6] + sum[7]));
      }
      {
        const uint32_t(&p0)[32] = *reinterpret_cast<const uint32_t(*)[32]>(&pk[0]);
        const uint32_t(&p1)[32] = *reinterpret_cast<const uint32_t(*)[32]>(&pk[32]);
        tmem_st_32x32b_x32(s_addr, p0);
        tmem_st_32x32b_x32(s_addr + 32, p1);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
    }

    // ---- epilogue: O_t / l -> bf16 -> out[b, row, h*64 .. h*64+63]
    mbar_wait(&o_full[t], 0);
    tc_fence_after();
    uint32_t o0[32], o1[32];
    tmem_ld_32x32b_x32(o_addr, o0);
    tmem_ld_32x32b_x32(o_addr + 32, o1);
    if (F16PV) {
      uint32_t lv;
      tmem_ld_32x32b_x1(l_addr, lv);
      tc_wait_ld();
      l = __uint_as_float(lv);
    } else {
      tc_wait_ld();
    }
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
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int MODE>
static int launch(const CUtensorMap& tm, const Params& p, dim3 grid, cudaStream_t stream) {
  static SmemGrant grant;
  AETHER_CUDA_OK(ensure_dynamic_smem(grant, attention_kernel<MODE>, SMEM_BYTES));
  attention_kernel<MODE><<<grid, THREADS, SMEM_BYTES, stream>>>(tm, p);
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}

}  // namespace attn

int attention_v3_launch(const CUtensorMap& tm, int B, int S, int H, void* out, float scale_log2, void* workspace,
                        int64_t workspace_bytes, cudaStream_t stream);
int64_t attention_v3_workspace_bytes(int B, int S, int H);

// `v_fp16` is the variant id: 5 = decoupled S / P TMEM buffers (product default, attention_v3_tcgen05.cu), 0 = baseline
// (P aliases S), 2 = fp16 P/V (the V third of qkv then holds fp16 values, GEMM f16_from_col) with 40 % of the exponentials
// as an FMA-pipe polynomial.
int attention_bf16_ws(const void* qkv, void* out, int B, int S, int H, float softmax_scale, int v_fp16, void* workspace,
                      int64_t workspace_bytes, cudaStream_t stream) {
  AETHER_CHECK_ARG(B > 0 && S > 0 && H > 0);
  AETHER_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0);
  CUtensorMap tm;
  const uint64_t dims[4] = {64, uint64_t(3 * H), uint64_t(S), uint64_t(B)};
  const uint64_t strides[3] = {128, uint64_t(3 * H) * 128, uint64_t(S) * uint64_t(3 * H) * 128};
  const uint32_t box[4] = {64, 1, uint32_t(attn::BKV), 1};
  int rc = make_tmap_bf16(&tm, qkv, 4, dims, strides, box, true);   // 16-bit elements; the type only sizes the copy
  if (rc) return rc;
  attn::Params p;
  p.B = B; p.H = H; p.S = S;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  dim3 grid((unsigned)ceil_div(S, 2 * attn::BQ), (unsigned)H, (unsigned)B);
  switch (v_fp16) {
    case 0: return attn::launch<0>(tm, p, grid, stream);                             // baseline: P aliases S
    case 2: return attn::launch<2>(tm, p, grid, stream);                             // fp16 P/V + 40 % polynomial exp2
    case 5: return attention_v3_launch(tm, B, S, H, out, p.scale_log2, workspace, workspace_bytes, stream);   // default
    default: break;
  }
  AETHER_CHECK_ARG(!"unknown attention variant id (v_fp16 must be 0, 2 or 5)");
  return AETHER_ERR_INVALID;
}

int attention_bf16(const void* qkv, void* out, int B, int S, int H, float softmax_scale, int v_fp16,
                   cudaStream_t stream) {
  return attention_bf16_ws(qkv, out, B, S, H, softmax_scale, v_fp16, nullptr, 0, stream);
}

}  // namespace aether

extern "C" int64_t aether_attention_workspace_bytes(int32_t B, int32_t S, int32_t H, int32_t v_fp16) {
  return v_fp16 == 5 ? aether::attention_v3_workspace_bytes(B, S, H) : 0;
}
extern "C" int aether_attention_bf16_ws(const void* qkv, void* out, int32_t B, int32_t S, int32_t H, float softmax_scale,
                                        int32_t v_fp16, void* workspace, int64_t workspace_bytes, void* stream) {
  return aether::attention_bf16_ws(qkv, out, B, S, H, softmax_scale, v_fp16, workspace, workspace_bytes,
                                   reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int aether_attention_bf16(const void* qkv, void* out, int32_t B, int32_t S, int32_t H,
                                     float softmax_scale, int32_t v_fp16, void* stream) {
  return aether::attention_bf16(qkv, out, B, S, H, softmax_scale, v_fp16, reinterpret_cast<cudaStream_t>(stream));
}
