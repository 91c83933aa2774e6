// K3/K4: bf16 GEMM  C[M,N] = epi(A[M,K] * W[N,K]^T)  on 5th-gen tensor cores.
//
// Replaces every nn.Linear of CogVideoXBlock (to_q/k/v fused, to_out, ff.net.0.proj, ff.net.2) plus
// patch_embed.proj / text_proj / proj_out -- the third-party modules the reference calls at
// aether/pipelines/aetherv1_pipeline_cogvideox.py:865-875.
//
// Design (one CTA per SM, persistent over output tiles):
//   warp 0      : TMA producer   - cp.async.bulk.tensor 128B-swizzled {64 x 128} A and {64 x 256} W boxes
//                                  into a 4-stage shared-memory ring (48 KB / stage)
//   warp 1      : MMA issuer     - one elected thread issues tcgen05.mma.cta_group::1.kind::f16
//                                  (M=128, N=256, K=16) x4 per stage, accumulating fp32 in TMEM
//   warps 2..5  : epilogue       - tcgen05.ld the 128x256 fp32 tile (double-buffered: 2 x 256 TMEM columns,
//                                  so the epilogue of tile i overlaps the main loop of tile i+1), apply
//                                  bias / GELU-tanh / gate*x+residual, pack to bf16, stage a 128x64 chunk
//                                  in swizzled shared memory and TMA-store it.
// Roofline: tensor (2*M*N*K flop / launch); operands are re-read from L2, HBM traffic ~ (M*K + N*K + M*N)*2 B.
#include "host_util.h"
#include "ptx.cuh"

namespace aether {

namespace gemm {
constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int CCHUNK = 64;                       // epilogue column chunk (64 bf16 = one 128B swizzle row)
constexpr int C_BYTES = BM * CCHUNK * 2;
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 2 * C_BYTES + 256;
constexpr int THREADS = 192;
constexpr int GROUP_M = 8;
constexpr int DH_QK = 64;                        // head_dim of the fused QKV epilogue (= CCHUNK: one chunk per head)

struct Params {
  int M, N, K;
  int num_m, num_n;
  const float* bias;        // [N] or null
  const float* gate_vid;    // EPI 2: [B, N] (batch stride gate_bstride) for rows s >= St
  const float* gate_txt;    // EPI 2: same for rows s < St
  int64_t gate_bstride;
  int S, St;                // row -> (b = row / S, s = row % S)
  __nv_bfloat16* C;         // EPI 2 reads the residual from C (in place)
  int64_t ldc;
  int f16_from;             // output columns >= f16_from are written as fp16 instead of bf16 (V third of QKV)
  // EPI 3 (fused QKV projection): LayerNorm(64) of every q / k head + rotary embedding of the video tokens
  const float* qn_g; const float* qn_b; const float* kn_g; const float* kn_b;   // [64] each
  const float* rope_cos; const float* rope_sin;                                 // [S - St, 64] fp32 or null
  float qk_eps;
  int heads;                // N = 3 * heads * 64: columns [0, 64 heads) q, then k, then v
};

__device__ __forceinline__ void tile_coords(int t, const Params& p, int& m_blk, int& n_blk) {
  const int group_size = GROUP_M * p.num_n;
  const int g = t / group_size;
  const int first_m = g * GROUP_M;
  const int gm = min(GROUP_M, p.num_m - first_m);
  const int local = t - g * group_size;
  m_blk = first_m + local % gm;
  n_blk = local / gm;
}

__device__ __forceinline__ float gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float inner = k0 * (x + k1 * x * x * x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(inner));
  return 0.5f * x * (1.0f + t);
}

// Epilogue of one 128 x 256 accumulator tile, executed by the 128 epilogue threads of a CTA (thread = row):
// tcgen05.ld 64-column chunks, bias / GELU-tanh / gate*x + residual, pack, swizzled smem staging, TMA store.
// `taddr_tile` = TMEM address of the tile's column 0 in this warp's lane quarter.  Shared by the 1-CTA and 2-CTA kernels.
template <int EPI>
__device__ __forceinline__ void epilogue_tile(const Params& p, const CUtensorMap* tmap_c, uint32_t taddr_tile, int m_blk,
                                              int n_blk, int row_in_tile, uint8_t* smem_c, int& cbuf,
                                              bool store_leader) {
  const int row = m_blk * BM + row_in_tile;
  const bool row_ok = row < p.M;
  const float* gate = nullptr;
  if (EPI == 2) {
    const int b = row_ok ? row / p.S : 0;
    const int s = row_ok ? row - b * p.S : 0;
    gate = (s < p.St ? p.gate_txt : p.gate_vid) + b * p.gate_bstride;
  }
#pragma unroll 1
  for (int ch = 0; ch < BN / CCHUNK; ++ch) {
    const int n0 = n_blk * BN + ch * CCHUNK;
    if (n0 >= p.N) break;
    uint32_t v0[32], v1[32];
    const uint32_t taddr = taddr_tile + ch * CCHUNK;
    tmem_ld_32x32b_x32(taddr, v0);
    tmem_ld_32x32b_x32(taddr + 32, v1);
    // Gated-residual epilogue: the thread's 128 bytes of residual (its row, this chunk's 64 columns) are requested as
    // eight independent 16-byte loads BEFORE anything waits, together with the TMEM load.  They used to be issued one
    // per 8-column group between dependent math (ncu round 2, to_out shape: eight serial L2 latencies per chunk, tensor
    // pipe 68 % of the active cycles because the MMA warp waited for the epilogue to release its accumulator).
    uint4 resid[CCHUNK / 8];
    if (EPI == 2) {
#pragma unroll
      for (int g8 = 0; g8 < CCHUNK / 8; ++g8) {
        const int n = n0 + g8 * 8;
        resid[g8] = (row_ok && n < p.N) ? *reinterpret_cast<const uint4*>(p.C + int64_t(row) * p.ldc + n)
                                        : make_uint4(0u, 0u, 0u, 0u);
      }
    }
    tc_wait_ld();

    uint32_t packed[32];
    if (EPI == 3 && n0 < 2 * p.heads * DH_QK) {
      // One 64-column chunk = one q or k head of this thread's token: QK-LayerNorm and the rotary embedding happen on
      // the accumulator row in registers, with the roundings and the summation order of the stand-alone
      // qk_norm_rope kernel (dit_kernels.cu), which this path replaces inside the DiT (370 MB of HBM traffic and one
      // launch per layer less).
      const bool is_k = n0 >= p.heads * DH_QK;
      float f[64];
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        float x = __uint_as_float(c < 32 ? v0[c] : v1[c - 32]);
        if (p.bias != nullptr) x += __ldg(p.bias + n0 + c);
        f[c] = __bfloat162float(__float2bfloat16_rn(x));     // the projection output is a bf16 tensor upstream
      }
      float part[8];
#pragma unroll
      for (int g8 = 0; g8 < 8; ++g8) {
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += f[g8 * 8 + j];
        part[g8] = sum;
      }
      const float mean = (((part[0] + part[1]) + (part[2] + part[3])) + ((part[4] + part[5]) + (part[6] + part[7]))) *
                         (1.0f / 64);
#pragma unroll
      for (int g8 = 0; g8 < 8; ++g8) {
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = f[g8 * 8 + j] - mean;
          sq += d * d;
        }
        part[g8] = sq;
      }
      const float sq = ((part[0] + part[1]) + (part[2] + part[3])) + ((part[4] + part[5]) + (part[6] + part[7]));
      const float rstd = rsqrtf(sq * (1.0f / 64) + p.qk_eps);
      const float* g = is_k ? p.kn_g : p.qn_g;
      const float* bt = is_k ? p.kn_b : p.qn_b;
#pragma unroll
      for (int c = 0; c < 64; ++c) f[c] = (f[c] - mean) * rstd;
#pragma unroll
      for (int c = 0; c < 64; c += 4) {
        const float4 gm = __ldg(reinterpret_cast<const float4*>(g + c));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bt + c));
        f[c] = f[c] * gm.x + bb.x; f[c + 1] = f[c + 1] * gm.y + bb.y;
        f[c + 2] = f[c + 2] * gm.z + bb.z; f[c + 3] = f[c + 3] * gm.w + bb.w;
      }
      const int tok = row_ok ? row % p.S : 0;
      if (p.rope_cos != nullptr && tok >= p.St) {
        const float* cr = p.rope_cos + int64_t(tok - p.St) * 64;
        const float* sr = p.rope_sin + int64_t(tok - p.St) * 64;
#pragma unroll
        for (int c = 0; c < 64; c += 4) {
          const float4 cs = __ldg(reinterpret_cast<const float4*>(cr + c));
          const float4 sn = __ldg(reinterpret_cast<const float4*>(sr + c));
          // upstream rounds LayerNorm's output to bf16 before apply_rotary_emb upcasts it again
          const float a0 = __bfloat162float(__float2bfloat16_rn(f[c])), a1 = __bfloat162float(__float2bfloat16_rn(f[c + 1]));
          const float a2 = __bfloat162float(__float2bfloat16_rn(f[c + 2])), a3 = __bfloat162float(__float2bfloat16_rn(f[c + 3]));
          f[c] = a0 * cs.x - a1 * sn.x;
          f[c + 1] = a1 * cs.y + a0 * sn.y;
          f[c + 2] = a2 * cs.z - a3 * sn.z;
          f[c + 3] = a3 * cs.w + a2 * sn.w;
        }
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) packed[c] = pack_bf16x2(f[2 * c], f[2 * c + 1]);
    } else {
#pragma unroll
    for (int g8 = 0; g8 < 8; ++g8) {     // 8 groups of 8 columns (one 16-byte bf16 vector each)
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = g8 * 8 + j;
        x[j] = __uint_as_float(c < 32 ? v0[c] : v1[c - 32]);
      }
      const int n = n0 + g8 * 8;
      const bool col_ok = n < p.N;       // N % 8 == 0 is required
      if (p.bias != nullptr && col_ok) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
        x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
        x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
      }
      if (EPI == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = gelu_tanh(x[j]);
      }
      if (EPI == 2) {
        if (row_ok && col_ok) {
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(gate + n));
          const float4 g1 = __ldg(reinterpret_cast<const float4*>(gate + n + 4));
          const uint4 r = resid[g8];
          x[0] = bf16_lo(r.x) + g0.x * x[0]; x[1] = bf16_hi(r.x) + g0.y * x[1];
          x[2] = bf16_lo(r.y) + g0.z * x[2]; x[3] = bf16_hi(r.y) + g0.w * x[3];
          x[4] = bf16_lo(r.z) + g1.x * x[4]; x[5] = bf16_hi(r.z) + g1.y * x[5];
          x[6] = bf16_lo(r.w) + g1.z * x[6]; x[7] = bf16_hi(r.w) + g1.w * x[7];
        }
      }
      if (n >= p.f16_from) {
        // fp16 columns (V of the fp16-P/V attention variants): saturate instead of overflowing to inf -- bf16, which
        // the reference keeps V in, has fp32's range (NaN still propagates: both comparisons are false for it)
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = x[j] > 65504.f ? 65504.f : (x[j] < -65504.f ? -65504.f : x[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) packed[g8 * 4 + j] = pack_f16x2(x[2 * j], x[2 * j + 1]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) packed[g8 * 4 + j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
      }
    }
    }
    // staging buffer `cbuf` was last used two chunks ago; make sure that TMA store has read it
    if (store_leader) tma_store_wait_read<1>();
    named_bar_sync(1, 128);
    uint8_t* crow = smem_c + cbuf * C_BYTES + row_in_tile * 128;
#pragma unroll
    for (int c16 = 0; c16 < 8; ++c16) {
      const int phys = c16 ^ (row_in_tile & 7);
      *reinterpret_cast<uint4*>(crow + phys * 16) =
          make_uint4(packed[c16 * 4], packed[c16 * 4 + 1], packed[c16 * 4 + 2], packed[c16 * 4 + 3]);
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (store_leader) {
      tma_store_2d(tmap_c, smem_c + cbuf * C_BYTES, n0, m_blk * BM);
      tma_store_commit();
    }
    cbuf ^= 1;
  }
}

template <int EPI>
__global__ void __launch_bounds__(THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const __grid_constant__ CUtensorMap tmap_c, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_c = smem + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + 2 * C_BYTES);
  uint64_t* full = bars;                    // [STAGES]
  uint64_t* empty = bars + STAGES;          // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;     // [2]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  // lane-0 broadcasts make warp index and TMEM base provably warp-uniform: role branches and MMA operands stay on
  // the uniform datapath and ptxas emits back-to-back UTCHMMA instead of an ELECT / BRA.U.ANY waterfall per MMA.
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m * p.num_n;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_c);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_base_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_ptr, 0);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (uniform control flow, elected issue)
    {
      const bool lead = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(t, p, m_blk, n_blk);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if (lead) {
            mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
            tma_load_2d(sa, &tmap_a, &full[stage], kb * BK, m_blk * BM);
            tma_load_2d(sb, &tmap_b, &full[stage], kb * BK, n_blk * BN);
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: the whole warp walks the tile / k-block
    // schedule (uniform control flow, every lane polls the barriers); one elected lane issues MMAs and commits.
    {
      const bool lead = elect_one();
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t a_desc = make_sw128_desc(sa);
          const uint64_t b_desc = make_sw128_desc(sa + A_BYTES);
          if (lead) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // +32 bytes (encoded >>4 = 2) per K=16 step inside the 128B swizzle atom
              tc_mma_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            tc_commit(&empty[stage]);
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (lead) tc_commit(&tmem_full[acc]);
        __syncwarp();
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5, 128 threads)
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row_in_tile = q * 32 + lane;
    const bool store_leader = (threadIdx.x == 64);
    int acc = 0;
    uint32_t acc_phase = 0;
    int cbuf = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(t, p, m_blk, n_blk);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<EPI>(p, &tmap_c, tmem_base + (uint32_t(q * 32) << 16) + acc * BN, m_blk, n_blk, row_in_tile, smem_c,
                         cbuf, store_leader);
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (store_leader) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): a cluster of two CTAs on one TPC owns a 256 x 256 output tile.  CTA r loads
// its own 128 rows of A and HALF of the W tile (128 of the 256 output columns); the leader issues
// tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16) which feeds both tensor cores from both halves.  Per SM and
// k-block the shared-memory traffic drops from 48 KB written + 48 KB read to 32 + 32 KB -- with one CTA per SM that
// traffic (192 B/clk against the 128 B/clk/SM the tensor core itself needs) was the reason the 1-CTA kernel stalls
// at ~81 % tensor-pipe activity -- and the stage shrinks to 32 KB, so the ring holds 6 stages instead of 4.
//   full[s]       leader's barrier: 1 arrive.expect_tx (leader producer) + 64 KB of complete_tx from BOTH CTAs' TMA
//   empty[s]      per CTA, armed by the leader's multicast tcgen05.commit (the pair's MMAs have read both stages)
//   tmem_full[a]  per CTA, multicast commit;   tmem_empty[a]  leader's, 256 arrivals (both CTAs' epilogue threads)
constexpr int STAGES2 = 6;
constexpr int BH_BYTES = (BN / 2) * BK * 2;          // half of the W tile
constexpr int STAGE2_BYTES = A_BYTES + BH_BYTES;     // 32 KB
constexpr int SMEM2_BYTES = 1024 + STAGES2 * STAGE2_BYTES + 2 * C_BYTES + 256;

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_bh,
             const __grid_constant__ CUtensorMap tmap_c, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_c = smem + STAGES2 * STAGE2_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + 2 * C_BYTES);
  uint64_t* full = bars;                     // [STAGES2]  (used in the leader)
  uint64_t* empty = bars + STAGES2;          // [STAGES2]
  uint64_t* tmem_full = bars + 2 * STAGES2;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]        (used in the leader)
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();               // 0 = leader
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_tiles = p.num_m * p.num_n;               // num_m counts 256-row pair tiles here
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_bh);
    tma_prefetch_desc(&tmap_c);
    for (int i = 0; i < STAGES2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 256);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta<512>(tmem_base_ptr);
  tc_fence_before();
  __syncthreads();                                       // CTA-local ordering of the allocator's shared-memory write
  // (compute-sanitizer racecheck still flags the cta_group::2 allocator's write against the read below: the peer CTA's
  //  half of the collective alloc is ordered by the cluster barrier, which the tool does not model;
  //  profiles/r2_sanitizer_racecheck_gemm2.log.  memcheck is clean.)
  cluster_sync_all();                                    // barriers of BOTH CTAs initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_ptr, 0);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    const bool lead = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int t = pair; t < num_tiles; t += num_pairs) {
      int m2, n_blk;
      tile_coords(t, p, m2, n_blk);
      const int m_row = (2 * m2 + int(rank)) * BM;
      const int n_row = n_blk * BN + int(rank) * (BN / 2);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = smem + stage * STAGE2_BYTES;
        if (lead) {
          if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * STAGE2_BYTES);
          tma_load_2d_2cta(sa, &tmap_a, &full[stage], kb * BK, m_row);
          tma_load_2d_2cta(sa + A_BYTES, &tmap_bh, &full[stage], kb * BK, n_row);
        }
        __syncwarp();
        if (++stage == STAGES2) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0) {
      const bool lead = elect_one();
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE2_BYTES);
          const uint64_t a_desc = make_sw128_desc(sa);
          const uint64_t b_desc = make_sw128_desc(sa + A_BYTES);
          if (lead) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc_mma_ss_2cta(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            tc_commit_2cta(&empty[stage], 3);
          }
          __syncwarp();
          if (++stage == STAGES2) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (lead) tc_commit_2cta(&tmem_full[acc], 3);
        __syncwarp();
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5 of each CTA, own 128 rows)
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    const bool store_leader = (threadIdx.x == 64);
    int acc = 0;
    uint32_t acc_phase = 0;
    int cbuf = 0;
    for (int t = pair; t < num_tiles; t += num_pairs) {
      int m2, n_blk;
      tile_coords(t, p, m2, n_blk);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<EPI>(p, &tmap_c, tmem_base + (uint32_t(q * 32) << 16) + acc * BN, 2 * m2 + int(rank), n_blk,
                         row_in_tile, smem_c, cbuf, store_leader);
      tc_fence_before();
      mbar_arrive_cluster(&tmem_empty[acc], 0);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (store_leader) tma_store_wait<0>();
  }

  tc_fence_before();
  cluster_sync_all();                                    // the pair's MMAs and remote arrivals are done in both CTAs
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta<512>(tmem_base);
  }
}

template <int EPI>
int launch2(const CUtensorMap& ta, const CUtensorMap& tbh, const CUtensorMap& tc, const Params& p, cudaStream_t stream) {
  static SmemGrant grant;
  AETHER_CUDA_OK(ensure_dynamic_smem(grant, gemm2_kernel<EPI>, SMEM2_BYTES));
  const int tiles = p.num_m * p.num_n;
  const int pairs = num_sms() / 2;
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
  gemm2_kernel<EPI><<<grid, THREADS, SMEM2_BYTES, stream>>>(ta, tbh, tc, p);
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}

template <int EPI>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const Params& p, cudaStream_t stream) {
  static SmemGrant grant;
  AETHER_CUDA_OK(ensure_dynamic_smem(grant, gemm_kernel<EPI>, SMEM_BYTES));
  const int tiles = p.num_m * p.num_n;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_kernel<EPI><<<grid, THREADS, SMEM_BYTES, stream>>>(ta, tb, tc, p);
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}
}  // namespace gemm

// `qk` (epilogue 3 only) carries the QK-LayerNorm / rotary fields of Params.
static int gemm_impl(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N, int K,
                     const float* bias, int epilogue, const float* gate_vid, const float* gate_txt,
                     int64_t gate_bstride, int S, int St, int f16_from_col, const gemm::Params* qk,
                     cudaStream_t stream) {
  AETHER_CHECK_ARG(M > 0 && N > 0 && K > 0);
  AETHER_CHECK_ARG(f16_from_col < 0 || f16_from_col % 8 == 0);
  AETHER_CHECK_ARG(N % 8 == 0 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0);
  AETHER_CHECK_ARG(epilogue >= 0 && epilogue <= 3 && (epilogue == 3) == (qk != nullptr));
  AETHER_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(C) & 15) == 0);
  if (epilogue == 2) AETHER_CHECK_ARG(gate_vid != nullptr && gate_txt != nullptr && S > 0);
  CUtensorMap ta, tb, tc;
  int rc;
  if ((rc = make_tmap_2d(&ta, A, M, K, lda, gemm::BM, gemm::BK))) return rc;
  if ((rc = make_tmap_2d(&tb, W, N, K, ldw, gemm::BN, gemm::BK))) return rc;
  if ((rc = make_tmap_2d(&tc, C, M, N, ldc, gemm::BM, gemm::CCHUNK))) return rc;
  gemm::Params p;
  p.M = M; p.N = N; p.K = K;
  p.num_m = (int)ceil_div(M, gemm::BM);
  p.num_n = (int)ceil_div(N, gemm::BN);
  p.bias = bias;
  p.gate_vid = gate_vid; p.gate_txt = gate_txt; p.gate_bstride = gate_bstride;
  p.S = S > 0 ? S : M; p.St = St;
  p.C = reinterpret_cast<__nv_bfloat16*>(C);
  p.ldc = ldc;
  p.f16_from = f16_from_col < 0 ? 0x7fffffff : f16_from_col;
  p.qn_g = p.qn_b = p.kn_g = p.kn_b = p.rope_cos = p.rope_sin = nullptr;
  p.qk_eps = 0.f; p.heads = 0;
  if (qk != nullptr) {
    p.qn_g = qk->qn_g; p.qn_b = qk->qn_b; p.kn_g = qk->kn_g; p.kn_b = qk->kn_b;
    p.rope_cos = qk->rope_cos; p.rope_sin = qk->rope_sin; p.qk_eps = qk->qk_eps; p.heads = qk->heads;
  }
  // Large problems run on CTA pairs (cta_group::2).
  if (M >= 1024 && N >= gemm::BN && num_sms() >= 2) {
    CUtensorMap tbh;
    if ((rc = make_tmap_2d(&tbh, W, N, K, ldw, gemm::BN / 2, gemm::BK))) return rc;
    gemm::Params p2 = p;
    p2.num_m = (int)ceil_div(M, 2 * gemm::BM);
    switch (epilogue) {
      case 0: return gemm::launch2<0>(ta, tbh, tc, p2, stream);
      case 1: return gemm::launch2<1>(ta, tbh, tc, p2, stream);
      case 2: return gemm::launch2<2>(ta, tbh, tc, p2, stream);
      default: return gemm::launch2<3>(ta, tbh, tc, p2, stream);
    }
  }
  switch (epilogue) {
    case 0: return gemm::launch<0>(ta, tb, tc, p, stream);
    case 1: return gemm::launch<1>(ta, tb, tc, p, stream);
    case 2: return gemm::launch<2>(ta, tb, tc, p, stream);
    default: return gemm::launch<3>(ta, tb, tc, p, stream);
  }
}

int gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N, int K,
              const float* bias, int epilogue, const float* gate_vid, const float* gate_txt, int64_t gate_bstride,
              int S, int St, int f16_from_col, cudaStream_t stream) {
  AETHER_CHECK_ARG(epilogue >= 0 && epilogue <= 2);
  return gemm_impl(A, lda, W, ldw, C, ldc, M, N, K, bias, epilogue, gate_vid, gate_txt, gate_bstride, S, St,
                   f16_from_col, nullptr, stream);
}

// Fused QKV projection: qkv[rows, 3*H*64] = A . W^T + bias, then QK-LayerNorm(64) on every q / k head and the rotary
// embedding on the video tokens (s >= St), all inside the GEMM epilogue (bit-compatible with gemm_bf16 followed by
// qk_norm_rope).  Replaces to_q / to_k / to_v + norm_q / norm_k + apply_rotary_emb of CogVideoXAttnProcessor2_0.
int gemm_qkv_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* qkv, int rows, int K, const float* bias,
                  int S, int St, int H, const float* gq, const float* bq, const float* gk, const float* bk, float eps,
                  const float* cosb, const float* sinb, int f16_from_col, cudaStream_t stream) {
  AETHER_CHECK_ARG(H > 0 && S > 0 && St >= 0 && gq && bq && gk && bk);
  AETHER_CHECK_ARG((cosb == nullptr) == (sinb == nullptr));
  gemm::Params qk{};
  qk.qn_g = gq; qk.qn_b = bq; qk.kn_g = gk; qk.kn_b = bk;
  qk.rope_cos = cosb; qk.rope_sin = sinb; qk.qk_eps = eps; qk.heads = H;
  const int N = 3 * H * gemm::DH_QK;
  return gemm_impl(A, lda, W, ldw, qkv, N, rows, N, K, bias, 3, nullptr, nullptr, 0, S, St, f16_from_col, &qk, stream);
}

}  // namespace aether

extern "C" int aether_gemm_qkv_norm_rope_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* qkv,
                                              int32_t rows, int32_t K, const float* bias, int32_t S, int32_t St,
                                              int32_t H, const float* qn_g, const float* qn_b, const float* kn_g,
                                              const float* kn_b, float eps, const float* rope_cos,
                                              const float* rope_sin, int32_t f16_from_col, void* stream) {
  return aether::gemm_qkv_bf16(A, lda, W, ldw, qkv, rows, K, bias, S, St, H, qn_g, qn_b, kn_g, kn_b, eps, rope_cos,
                               rope_sin, f16_from_col, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int aether_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc,
                                int32_t M, int32_t N, int32_t K, const float* bias, int32_t epilogue,
                                const float* gate_vid, const float* gate_txt, int64_t gate_bstride, int32_t S,
                                int32_t St, int32_t f16_from_col, void* stream) {
  return aether::gemm_bf16(A, lda, W, ldw, C, ldc, M, N, K, bias, epilogue, gate_vid, gate_txt, gate_bstride, S, St,
                           f16_from_col, reinterpret_cast<cudaStream_t>(stream));
}
