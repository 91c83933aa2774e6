// K1, variant 5 (mode 8 of aether_attention_bf16): 64-key tiles, S double-buffered per query tile, and every softmax
// warp keeps the tcgen05.ld of S(j+1) IN FLIGHT while it runs the exponentials of S(j).
//
// Why (hypothesis at the time): the MUFU (16 ex2/clk/SM) needs 1024 clk per 128x128 scores, the tensor pipe 512; mode
// 5 loads a whole 128-column tile, waits, then runs the MUFU phase and measured 1530 clk per tile (MUFU 67 % busy).
// Here the S load of the next 64 columns is asynchronous and retires during the exponentials, and S / P are
// double-buffered so the MMA warp can run two key tiles ahead.  Outcome: correct, 3.74 ms -- the real limiter of
// mode 5 was the MMA-issue waterfall, not the load; kept as a variant (mode 8).
//
// TMEM (512 columns), per query tile t in {0,1} at base t*256:
//     S_t[0] +0   S_t[1] +64   P_t[0] +128 (32 cols, bf16)   P_t[1] +160   O_t +192 (64 cols)
// smem: Q 2 x 16 KB, K ring 6 x 8 KB, V ring 6 x 8 KB (64-row boxes, 128B swizzle).
// warps: 0 = TMA producer, 1 = MMA issuer (+ TMEM alloc), 2..9 = two softmax warpgroups (thread = query row),
//        10..11 = idle register donors.
// MMA issue order: QK_0(0) QK_1(0) QK_0(1) QK_1(1), then per key tile j:  QK_0(j+2) QK_1(j+2)  PV_0(j) PV_1(j).
// barriers (b = j & 1, completion index j >> 1):
//     s_full[t][b]  MMA -> softmax        s_free[t][b]  softmax -> MMA (S(j) sits in registers)
//     p_full[t][b]  softmax -> MMA        p_free[t][b]  MMA -> softmax (PV_t(j) retired: P_t[b], O_t quiescent)
#include "host_util.h"
#include "ptx.cuh"

namespace aether {
namespace attn5 {

constexpr int DH = 64, BQ = 128, BKV = 64, KSTAGES = 6, VSTAGES = 6;
constexpr int Q_BYTES = BQ * DH * 2, KV_BYTES = BKV * DH * 2;
constexpr int SMEM_BYTES = 1024 + 2 * Q_BYTES + (KSTAGES + VSTAGES) * KV_BYTES + 512;
// 12 warps although 10 work: setmaxnreg can only redistribute the registers the CTA was launched with
// (threads x 168 as ptxas sizes this kernel), and 8 x 224 + 4 x 48 registers per lane need the pool of 384 threads.
constexpr int THREADS = 384;
constexpr uint32_t COL_TILE = 256, COL_S = 0, COL_P = 128, COL_O = 192;

struct Params {
  int B, H, S;
  __nv_bfloat16* out;
  float scale_log2;
};

__global__ void __launch_bounds__(THREADS, 1)
attention_v5_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + 2 * Q_BYTES;
  uint8_t* smem_v = smem_k + KSTAGES * KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + VSTAGES * KV_BYTES);
  uint64_t* q_full = bars;                  // [2]
  uint64_t* k_full = q_full + 2;
  uint64_t* k_empty = k_full + KSTAGES;
  uint64_t* v_full = k_empty + KSTAGES;
  uint64_t* v_empty = v_full + VSTAGES;
  uint64_t* s_full = v_empty + VSTAGES;     // [t*2 + b]
  uint64_t* s_free = s_full + 4;            // count 128
  uint64_t* p_full = s_free + 4;            // count 128
  uint64_t* p_free = p_full + 4;
  uint64_t* o_full = p_free + 4;            // [2]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(o_full + 2);

  // warp index and TMEM base through a lane-0 broadcast: the compiler then knows they are warp-uniform, keeps the
  // role branches and the MMA operands on the uniform datapath and emits plain UTCHMMA instead of a per-instruction
  // ELECT / BRA.U.ANY waterfall (measured: the waterfall made the single MMA thread the bottleneck of this kernel).
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int q_blk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = q_blk * 2 * BQ;
  const int n_kv = (p.S + BKV - 1) / BKV;
  const int H = p.H;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&o_full[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 128);
      mbar_init(&p_full[i], 128);
      mbar_init(&p_free[i], 1);
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

  if (warp < 2 || warp >= 10) {
    setmaxnreg_dec<48>();
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer (uniform control flow)
      const bool lead = elect_one();
      if (lead) {
        for (int t = 0; t < 2; ++t) {
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
      // ---------------------------------------------------------------- MMA issuer: the whole warp walks the
      // schedule (uniform control flow, every lane polls the barriers), one elected lane issues MMAs and commits.
      const bool lead = elect_one();
      constexpr uint32_t idesc_qk = make_idesc_f16kind(BQ, BKV, 1, 1, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16kind(BQ, DH, 1, 1, 0, 1);
      uint64_t q_desc[2];
      for (int t = 0; t < 2; ++t) q_desc[t] = make_sw128_desc(smem_u32(smem_q + t * Q_BYTES));
      auto issue_qk = [&](int t, int sb, int ks) {
        const uint64_t k_desc = make_sw128_desc(smem_u32(smem_k + ks * KV_BYTES));
        const uint32_t d = tmem_base + t * COL_TILE + COL_S + sb * 64;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) tc_mma_ss(d, q_desc[t] + 2 * k, k_desc + 2 * k, idesc_qk, k > 0);
      };
      auto issue_pv = [&](int t, int pb, int vs, bool first) {
        const uint64_t v_desc = make_sw128_desc(smem_u32(smem_v + vs * KV_BYTES));
        const uint32_t d = tmem_base + t * COL_TILE + COL_O;
        const uint32_t a = tmem_base + t * COL_TILE + COL_P + pb * 32;
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)
          tc_mma_ts(d, a + 8 * k, v_desc + 128 * k, idesc_pv, (!first || k > 0) ? 1u : 0u);
      };
      int ks = 0;                 // stage of the next K tile to consume
      uint32_t kph = 0;
      int vs = 0;
      uint32_t vph = 0;
      // prologue: S(0) and S(1) of both query tiles
      const int n_pro = n_kv < 2 ? n_kv : 2;
      for (int j = 0; j < n_pro; ++j) {
        mbar_wait(&k_full[ks], kph);
        if (j == 0) mbar_wait(&q_full[0], 0);
        tc_fence_after();
        if (lead) {
          issue_qk(0, j, ks);
          tc_commit(&s_full[0 * 2 + j]);
        }
        if (j == 0) {
          mbar_wait(&q_full[1], 0);
          tc_fence_after();
        }
        if (lead) {
          issue_qk(1, j, ks);
          tc_commit(&s_full[1 * 2 + j]);
          tc_commit(&k_empty[ks]);
        }
        __syncwarp();
        if (++ks == KSTAGES) { ks = 0; kph ^= 1; }
      }
      for (int j = 0; j < n_kv; ++j) {
        const int sb = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const bool more = (j + 2 < n_kv);
        const bool last = (j + 1 == n_kv);
        // QK_t(j+2) first: its only dependency (S_t(j) in registers, signalled at the END of softmax step j-1) is
        // met a whole step before p_full(j); issuing it behind PV_t(j) put the MMA round trip on the softmax
        // critical path (ncu: 41 % of the softmax samples on the s_full wait).
        if (more) {
          mbar_wait(&k_full[ks], kph);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            mbar_wait(&s_free[t * 2 + sb], ph);
            tc_fence_after();
            if (lead) {
              issue_qk(t, sb, ks);
              tc_commit(&s_full[t * 2 + sb]);
            }
          }
          if (lead) tc_commit(&k_empty[ks]);
          __syncwarp();
          if (++ks == KSTAGES) { ks = 0; kph ^= 1; }
        }
        mbar_wait(&v_full[vs], vph);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&p_full[t * 2 + sb], ph);
          tc_fence_after();
          if (lead) {
            issue_pv(t, sb, vs, j == 0);
            tc_commit(&p_free[t * 2 + sb]);
            if (last) tc_commit(&o_full[t]);
          }
        }
        if (lead) tc_commit(&v_empty[vs]);
        __syncwarp();
        if (++vs == VSTAGES) { vs = 0; vph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    setmaxnreg_inc<224>();
    const int t = (warp - 2) >> 2;
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t tile_addr = tmem_base + lane_off + t * COL_TILE;
    const uint32_t o_addr = tile_addr + COL_O;
    const float sl2 = p.scale_log2;
    const float rescale_thresh = 8.0f / sl2;
    float m_used = -INFINITY, l = 0.f;
    uint64_t* const sfull = s_full + t * 2;
    uint64_t* const sfree = s_free + t * 2;
    uint64_t* const pfull = p_full + t * 2;
    uint64_t* const pfree = p_free + t * 2;

    // One key tile: `lo`/`hi` hold S(j) (columns 0..31 / 32..63); the load of S(j+1) into `nlo`/`nhi` is issued
    // first and only awaited at the end.
    auto step = [&](int j, uint32_t (&lo)[32], uint32_t (&hi)[32], uint32_t (&nlo)[32], uint32_t (&nhi)[32]) {
      const int sb = j & 1;
      const bool have_next = (j + 1 < n_kv);
      if (have_next) {
        mbar_wait(&sfull[sb ^ 1], ((j + 1) >> 1) & 1);
        tc_fence_after();
        const uint32_t a = tile_addr + COL_S + (sb ^ 1) * 64;
        tmem_ld_32x32b_x32(a, nlo);
        tmem_ld_32x32b_x32(a + 32, nhi);
      }
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
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = fast_exp2((m_used - m_new) * sl2);
        l *= alpha;
        m_used = m_new;
        if (j > 0) {
          mbar_wait(&pfree[(j - 1) & 1], ((j - 1) >> 1) & 1);   // PV_t(j-1) retired: O_t is quiescent
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
      if (j >= 2) {                                 // P_t[sb] is read by PV_t(j-2) until p_free flips
        mbar_wait(&pfree[sb], ((j >> 1) - 1) & 1);
        tc_fence_after();
      }
      tmem_st_32x32b_x32(tile_addr + COL_P + sb * 32, pk);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&pfull[sb]);
      if (have_next) {
        tc_wait_ld();
        tie_regs(nlo);
        tie_regs(nhi);
        tc_fence_before();
        mbar_arrive(&sfree[sb ^ 1]);                // S_t[sb^1] may be overwritten by QK_t(j+3)
      }
    };

    uint32_t a_lo[32], a_hi[32], b_lo[32], b_hi[32];
    mbar_wait(&sfull[0], 0);
    tc_fence_after();
    tmem_ld_32x32b_x32(tile_addr + COL_S, a_lo);
    tmem_ld_32x32b_x32(tile_addr + COL_S + 32, a_hi);
    tc_wait_ld();
    tie_regs(a_lo);
    tie_regs(a_hi);
    tc_fence_before();
    mbar_arrive(&sfree[0]);
    for (int j = 0; j < n_kv; j += 2) {
      step(j, a_lo, a_hi, b_lo, b_hi);
      if (j + 1 < n_kv) step(j + 1, b_lo, b_hi, a_lo, a_hi);
    }

    mbar_wait(&o_full[t], 0);
    tc_fence_after();
    uint32_t(&o0)[32] = a_lo;
    uint32_t(&o1)[32] = a_hi;
    tmem_ld_32x32b_x32(o_addr, o0);
    tmem_ld_32x32b_x32(o_addr + 32, o1);
    tc_wait_ld();
    tie_regs(o0);
    tie_regs(o1);
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

}  // namespace attn5

int attention_v5_launch(const void* qkv, int B, int S, int H, void* out, float scale_log2, cudaStream_t stream) {
  static SmemGrant grant;
  AETHER_CUDA_OK(ensure_dynamic_smem(grant, attn5::attention_v5_kernel, attn5::SMEM_BYTES));
  CUtensorMap tm;
  const uint64_t dims[4] = {64, uint64_t(3 * H), uint64_t(S), uint64_t(B)};
  const uint64_t strides[3] = {128, uint64_t(3 * H) * 128, uint64_t(S) * uint64_t(3 * H) * 128};
  const uint32_t box[4] = {64, 1, uint32_t(attn5::BKV), 1};
  int rc = make_tmap_bf16(&tm, qkv, 4, dims, strides, box, true);
  if (rc) return rc;
  attn5::Params p;
  p.B = B; p.H = H; p.S = S;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.scale_log2 = scale_log2;
  dim3 grid((unsigned)ceil_div(S, 2 * attn5::BQ), (unsigned)H, (unsigned)B);
  attn5::attention_v5_kernel<<<grid, attn5::THREADS, attn5::SMEM_BYTES, stream>>>(tm, p);
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}

}  // namespace aether
