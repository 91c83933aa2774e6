// K9: the convolutions of AutoencoderKLCogVideoX as an implicit GEMM on tcgen05.
//
// Replaces nn.Conv3d inside CogVideoXCausalConv3d (3x3x3, temporal padding = first frame / conv_cache, spatial
// zero padding), the nn.Conv2d of CogVideoXUpsample3D (3x3, pad 1) / CogVideoXDownsample3D (3x3, stride 2, pad
// (0,1,0,1)) and the 1x1x1 shortcuts -- third-party diffusers modules the reference reaches through
// vae.encode / vae.decode (aether/pipelines/aetherv1_pipeline_cogvideox.py:557-620, :931, :936).
//
// Layout: activations are channels-last  x[T, H, W, C]  (bf16); the caller hands in the TIME-PADDED input
// (T_in = T_out + kt - 1 leading frames: replicated first frame or the previous frame batch's cache), so the
// kernel itself is a plain "valid in time / padded in space" convolution.
// GEMM view:  M = T*H*W output positions (tiles of 8 rows x 16 cols of one frame = 128),  N = Cout,
//             K = taps * Cin_pad (tap-major, channels padded to a multiple of 64 with zero weights).
// For every (tap, 64-channel chunk) the A tile is ONE 4-D TMA box {64 ch, 16 x, 8 y, 1 t} of the input taken at
// the tap's offset; out-of-range rows/cols/channels are zero-filled by the TMA unit, which is exactly the
// convolution's zero padding (and the asymmetric (0,1) padding of the stride-2 down-sampler, fetched with
// element strides {1,2,2,1}).  B tiles come from the packed weight matrix [Cout, K]; accumulation is fp32 in
// TMEM (double-buffered); the epilogue adds bias (+ residual), packs bf16 and TMA-stores a {64 ch,16,8,1} box.
// Same warp roles / pipeline as gemm_tcgen05.cu.   Roofline: tensor, 2*M*Cout*taps*Cin flop per launch.
#include "host_util.h"
#include "ptx.cuh"

namespace aether {
namespace conv {

constexpr int BH = 8, BW = 16, BM = BH * BW, BK = 64;
constexpr int A_BYTES = BM * BK * 2;
constexpr int CCHUNK = 64;
constexpr int C_BYTES = BM * CCHUNK * 2;
constexpr int THREADS = 192;

template <int BN>
struct Cfg {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : 6;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 2 * C_BYTES + 256;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
};

struct Params {
  int T, H, W;              // output frames / rows / cols
  int Cout, cin_chunks;     // cin_chunks = Cin_pad / 64
  int kt, kh, kw, stride, pad_h, pad_w;
  int tiles_t, tiles_x, tiles_y, num_n;   // tile walk: frames x 8-row bands x 16-column bands x channel blocks
  const float* bias;
  const __nv_bfloat16* resid;   // optional [T, H, W, Cout]
  // CTA-pair kernel only: the pair owns two output tiles that are neighbours along `pair_axis` (0 = t, 1 = y, 2 = x);
  // tiles_t / tiles_y / tiles_x then count PAIR tiles along that axis
  int pair_axis;
};

__device__ __forceinline__ void tile_coords(int tile, const Params& p, int& t, int& ty, int& tx, int& nb) {
  nb = tile % p.num_n;
  int r = tile / p.num_n;
  tx = r % p.tiles_x;
  r /= p.tiles_x;
  ty = r % p.tiles_y;
  t = r / p.tiles_y;
}

// Epilogue of one 128-position x BN accumulator tile (thread = output position): tcgen05.ld 64-channel chunks, bias
// (+ residual), pack to bf16, swizzled shared-memory staging, TMA store of a {64 ch, 16, 8, 1} box (clipped at the
// tensor's edges by the TMA unit).  Shared by the one-CTA and the CTA-pair kernel.
template <int BN>
__device__ __forceinline__ void epilogue_tile(const Params& p, const CUtensorMap* tmap_y, uint32_t taddr_tile, int t, int ty,
                                              int tx, int nb, int row_in_tile, uint8_t* smem_c, int& cbuf,
                                              bool store_leader) {
  constexpr int NCH = (BN + CCHUNK - 1) / CCHUNK;      // 64-channel chunks per tile (BN = 32 -> one 32-wide chunk)
  constexpr int CW = BN < CCHUNK ? BN : CCHUNK;        // chunk width
  const int y = ty * BH + row_in_tile / BW;
  const int x = tx * BW + row_in_tile % BW;
  const bool pos_ok = (t < p.T) && (y < p.H) && (x < p.W);
  const int64_t pos = (int64_t(t) * p.H + y) * p.W + x;
#pragma unroll 1
  for (int ch = 0; ch < NCH; ++ch) {
    const int n0 = nb * BN + ch * CW;
    if (n0 >= p.Cout) break;
    uint32_t v0[32], v1[32];
    const uint32_t taddr = taddr_tile + ch * CW;
    tmem_ld_32x32b_x32(taddr, v0);
    if (CW > 32) tmem_ld_32x32b_x32(taddr + 32, v1);
    // residual (resnet skip connection): all 16-byte loads of this thread's row are issued before the TMEM wait
    // instead of one per 8-channel group between dependent math (same finding as the GEMM's gated-residual epilogue)
    uint4 rres[CW / 8];
    if (p.resid != nullptr) {
#pragma unroll
      for (int g8 = 0; g8 < CW / 8; ++g8) {
        const int n = n0 + g8 * 8;
        rres[g8] = (pos_ok && n < p.Cout) ? *reinterpret_cast<const uint4*>(p.resid + pos * p.Cout + n)
                                          : make_uint4(0u, 0u, 0u, 0u);
      }
    }
    tc_wait_ld();
    uint32_t packed[32];
#pragma unroll
    for (int g8 = 0; g8 < CW / 8; ++g8) {
      float xv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = g8 * 8 + j;
        xv[j] = __uint_as_float(c < 32 ? v0[c] : v1[c - 32]);
      }
      const int n = n0 + g8 * 8;
      const bool col_ok = n < p.Cout;      // Cout % 8 == 0 (the caller pads tiny Cout)
      if (p.bias != nullptr && col_ok) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
        xv[0] += b0.x; xv[1] += b0.y; xv[2] += b0.z; xv[3] += b0.w;
        xv[4] += b1.x; xv[5] += b1.y; xv[6] += b1.z; xv[7] += b1.w;
      }
      if (p.resid != nullptr && pos_ok && col_ok) {
        const uint4 r = rres[g8];
        xv[0] += bf16_lo(r.x); xv[1] += bf16_hi(r.x); xv[2] += bf16_lo(r.y); xv[3] += bf16_hi(r.y);
        xv[4] += bf16_lo(r.z); xv[5] += bf16_hi(r.z); xv[6] += bf16_lo(r.w); xv[7] += bf16_hi(r.w);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) packed[g8 * 4 + j] = pack_bf16x2(xv[2 * j], xv[2 * j + 1]);
    }
    if (store_leader) tma_store_wait_read<1>();
    named_bar_sync(1, 128);
    // staging tile [128 positions][CW channels]: 128-byte rows with the 128B swizzle when CW == 64,
    // dense 64-byte rows (no swizzle) when CW == 32
    if (CW == 64) {
      uint8_t* crow = smem_c + cbuf * C_BYTES + row_in_tile * 128;
#pragma unroll
      for (int c16 = 0; c16 < 8; ++c16) {
        const int phys = c16 ^ (row_in_tile & 7);
        *reinterpret_cast<uint4*>(crow + phys * 16) =
            make_uint4(packed[c16 * 4], packed[c16 * 4 + 1], packed[c16 * 4 + 2], packed[c16 * 4 + 3]);
      }
    } else {
      uint8_t* crow = smem_c + cbuf * C_BYTES + row_in_tile * (CW * 2);
#pragma unroll
      for (int c16 = 0; c16 < CW / 8; ++c16)
        *reinterpret_cast<uint4*>(crow + c16 * 16) =
            make_uint4(packed[c16 * 4], packed[c16 * 4 + 1], packed[c16 * 4 + 2], packed[c16 * 4 + 3]);
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (store_leader) {
      tma_store_4d(tmap_y, smem_c + cbuf * C_BYTES, n0, tx * BW, ty * BH, t);
      tma_store_commit();
    }
    cbuf ^= 1;
  }
}

template <int BN>
__global__ void __launch_bounds__(THREADS, 1)
conv_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
            const __grid_constant__ CUtensorMap tmap_y, const Params p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_c = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + 2 * C_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::STAGES;
  uint64_t* tmem_full = bars + 2 * C::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform (see gemm_tcgen05.cu)
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_t * p.tiles_y * p.tiles_x * p.num_n;
  const int taps = p.kt * p.kh * p.kw;
  const int num_kb = taps * p.cin_chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_y);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_base_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_ptr, 0);

  if (warp == 0) {
    // TMA producer: uniform control flow for the whole warp, one elected lane issues the copies
    {
      const bool lead = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int t, ty, tx, nb;
        tile_coords(tile, p, t, ty, tx, nb);
        const int x0 = tx * BW * p.stride - p.pad_w;
        const int y0 = ty * BH * p.stride - p.pad_h;
        int kb = 0;
        for (int dt = 0; dt < p.kt; ++dt)
          for (int dy = 0; dy < p.kh; ++dy)
            for (int dx = 0; dx < p.kw; ++dx)
              for (int cc = 0; cc < p.cin_chunks; ++cc, ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* sa = smem + stage * C::STAGE_BYTES;
                if (lead) {
                  mbar_arrive_expect_tx(&full[stage], C::STAGE_BYTES);
                  tma_load_4d(sa, &tmap_x, &full[stage], cc * BK, x0 + dx, y0 + dy, t + dt);
                  tma_load_2d(sa + A_BYTES, &tmap_w, &full[stage], kb * BK, nb * BN);
                }
                __syncwarp();
                if (++stage == C::STAGES) {
                  stage = 0;
                  phase ^= 1;
                }
              }
      }
    }
  } else if (warp == 1) {
    // whole warp walks the schedule (uniform control flow -> back-to-back UTCHMMA); one elected lane issues
    {
      const bool lead = elect_one();
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t a_desc = make_sw128_desc(sa);
          const uint64_t b_desc = make_sw128_desc(sa + A_BYTES);
          if (lead) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc_mma_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            tc_commit(&empty[stage]);
          }
          __syncwarp();
          if (++stage == C::STAGES) {
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
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    const bool store_leader = (threadIdx.x == 64);
    int acc = 0;
    uint32_t acc_phase = 0;
    int cbuf = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int t, ty, tx, nb;
      tile_coords(tile, p, t, ty, tx, nb);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<BN>(p, &tmap_y, tmem_base + (uint32_t(q * 32) << 16) + acc * BN, t, ty, tx, nb, row_in_tile, smem_c,
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
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

template <int BN>
int launch(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& ty, const Params& p, cudaStream_t stream) {
  static SmemGrant grant;
  AETHER_CUDA_OK(ensure_dynamic_smem(grant, conv_kernel<BN>, Cfg<BN>::SMEM_BYTES));
  const int64_t tiles = int64_t(p.tiles_t) * p.tiles_y * p.tiles_x * p.num_n;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  conv_kernel<BN><<<grid, THREADS, Cfg<BN>::SMEM_BYTES, stream>>>(tx, tw, ty, p);
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}


// ------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2, same scheme as gemm2_kernel): a cluster of two CTAs owns two output tiles that
// are neighbours along t, y or x (host picks the axis that wastes the fewest out-of-range positions) and the same BN
// output channels.  CTA r fetches the activation boxes of ITS tile and HALF of the weight tile; the leader issues
// M = 256 MMAs that read both halves.  Round-2 ncu of the one-CTA kernel at the decoder's 128 -> 128 convs: 72 % tensor
// activity with 128 B/clk/SM of L2 -> shared-memory operand traffic (16 KB of activations + 16 KB of weights per 256
// tensor clocks); the pair needs 96 B/clk (24 KB), the 256-channel convs 64 instead of 96, and the smaller stages
// make the ring 8 / 6 deep.
template <int BN>
struct Cfg2 {
  static constexpr int BH_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + BH_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 6 : 8;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 2 * C_BYTES + 256;
  static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
conv2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_wh,
             const __grid_constant__ CUtensorMap tmap_y, const Params p) {
  using C = Cfg2<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_c = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + 2 * C_BYTES);
  uint64_t* full = bars;                       // [STAGES]  (used in the leader)
  uint64_t* empty = bars + C::STAGES;          // [STAGES]
  uint64_t* tmem_full = bars + 2 * C::STAGES;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]        (used in the leader)
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();               // 0 = leader
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_tiles = p.tiles_t * p.tiles_y * p.tiles_x * p.num_n;      // pair tiles
  const int taps = p.kt * p.kh * p.kw;
  const int num_kb = taps * p.cin_chunks;
  // this CTA's tile of pair tile (t, ty, tx)
  const int rt = p.pair_axis == 0 ? int(rank) : 0, ry = p.pair_axis == 1 ? int(rank) : 0, rx = p.pair_axis == 2 ? int(rank) : 0;
  const int mt = p.pair_axis == 0 ? 2 : 1, my = p.pair_axis == 1 ? 2 : 1, mx = p.pair_axis == 2 ? 2 : 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_wh);
    tma_prefetch_desc(&tmap_y);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 256);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta<C::TMEM_COLS>(tmem_base_ptr);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                    // barriers of BOTH CTAs initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_ptr, 0);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    const bool lead = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      int t, ty, tx, nb;
      tile_coords(tile, p, t, ty, tx, nb);
      t = t * mt + rt; ty = ty * my + ry; tx = tx * mx + rx;
      const int x0 = tx * BW * p.stride - p.pad_w;
      const int y0 = ty * BH * p.stride - p.pad_h;
      const int n_row = nb * BN + int(rank) * (BN / 2);
      int kb = 0;
      for (int dt = 0; dt < p.kt; ++dt)
        for (int dy = 0; dy < p.kh; ++dy)
          for (int dx = 0; dx < p.kw; ++dx)
            for (int cc = 0; cc < p.cin_chunks; ++cc, ++kb) {
              mbar_wait(&empty[stage], phase ^ 1);
              uint8_t* sa = smem + stage * C::STAGE_BYTES;
              if (lead) {
                if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * C::STAGE_BYTES);
                // a tile beyond the last frame / row / column (odd counts) reads zero-filled boxes and stores nothing
                tma_load_4d_2cta(sa, &tmap_x, &full[stage], cc * BK, x0 + dx, y0 + dy, t + dt);
                tma_load_2d_2cta(sa + A_BYTES, &tmap_wh, &full[stage], kb * BK, n_row);
              }
              __syncwarp();
              if (++stage == C::STAGES) {
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
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t a_desc = make_sw128_desc(sa);
          const uint64_t b_desc = make_sw128_desc(sa + A_BYTES);
          if (lead) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc_mma_ss_2cta(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            tc_commit_2cta(&empty[stage], 3);
          }
          __syncwarp();
          if (++stage == C::STAGES) {
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
    // ------------------------------------------------------------ epilogue (warps 2..5 of each CTA, own tile)
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    const bool store_leader = (threadIdx.x == 64);
    int acc = 0;
    uint32_t acc_phase = 0;
    int cbuf = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      int t, ty, tx, nb;
      tile_coords(tile, p, t, ty, tx, nb);
      t = t * mt + rt; ty = ty * my + ry; tx = tx * mx + rx;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<BN>(p, &tmap_y, tmem_base + (uint32_t(q * 32) << 16) + acc * BN, t, ty, tx, nb, row_in_tile, smem_c,
                        cbuf, store_leader);
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
    tmem_dealloc_2cta<C::TMEM_COLS>(tmem_base);
  }
}

template <int BN>
int launch2(const CUtensorMap& tx, const CUtensorMap& twh, const CUtensorMap& ty, const Params& p, cudaStream_t stream) {
  static SmemGrant grant;
  AETHER_CUDA_OK(ensure_dynamic_smem(grant, conv2_kernel<BN>, Cfg2<BN>::SMEM_BYTES));
  const int64_t tiles = int64_t(p.tiles_t) * p.tiles_y * p.tiles_x * p.num_n;
  const int pairs = num_sms() / 2;
  const int grid = 2 * (int)(tiles < pairs ? tiles : pairs);
  conv2_kernel<BN><<<grid, THREADS, Cfg2<BN>::SMEM_BYTES, stream>>>(tx, twh, ty, p);
  AETHER_CUDA_OK(cudaGetLastError());
  return AETHER_OK;
}

}  // namespace conv

int conv3d_bf16_impl(const void* x, int T_in, int H_in, int W_in, int Cin, const void* w_packed, const float* bias,
                const void* resid, void* y, int T_out, int H_out, int W_out, int Cout, int kt, int kh, int kw,
                int stride, int pad_h, int pad_w, bool allow_pairs, cudaStream_t stream) {
  AETHER_CHECK_ARG(x && w_packed && y);
  AETHER_CHECK_ARG(Cin % 8 == 0 && Cout % 8 == 0 && (stride == 1 || stride == 2));
  AETHER_CHECK_ARG(T_in == T_out + kt - 1 && kt >= 1 && kh >= 1 && kw >= 1);
  const int cin_pad = (int)ceil_div(Cin, 64) * 64;
  const int taps = kt * kh * kw;
  CUtensorMap tx, tw, ty;
  int rc;
  {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return AETHER_ERR_CUDA;
    cuuint64_t gdim[4] = {cuuint64_t(Cin), cuuint64_t(W_in), cuuint64_t(H_in), cuuint64_t(T_in)};
    cuuint64_t gstr[3] = {cuuint64_t(Cin) * 2, cuuint64_t(W_in) * Cin * 2, cuuint64_t(H_in) * W_in * Cin * 2};
    cuuint32_t box[4] = {64, cuuint32_t(conv::BW * stride), cuuint32_t(conv::BH * stride), 1};
    cuuint32_t estr[4] = {1, cuuint32_t(stride), cuuint32_t(stride), 1};
    CUresult r = enc(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      fprintf(stderr, "[aether_b200] conv input tensor map failed: %d (Cin %d W %d H %d T %d stride %d)\n", (int)r, Cin,
              W_in, H_in, T_in, stride);
      return AETHER_ERR_CUDA;
    }
  }
  // N tile: 256 when Cout allows it (A tiles re-read half as often, 96 B/clk of shared-memory operand traffic
  // instead of 128 B/clk), 128 otherwise, 32 for the tiny conv_out / moments heads.
  const int bn = Cout <= 32 ? 32 : (Cout % 256 == 0 ? 256 : 128);
  if ((rc = make_tmap_2d(&tw, w_packed, Cout, uint64_t(taps) * cin_pad, uint64_t(taps) * cin_pad, bn, conv::BK)))
    return rc;
  {
    const uint32_t cw = bn < 64 ? bn : 64;
    const uint64_t dims[4] = {uint64_t(Cout), uint64_t(W_out), uint64_t(H_out), uint64_t(T_out)};
    const uint64_t str[3] = {uint64_t(Cout) * 2, uint64_t(W_out) * Cout * 2, uint64_t(H_out) * W_out * Cout * 2};
    const uint32_t box[4] = {cw, uint32_t(conv::BW), uint32_t(conv::BH), 1};
    if ((rc = make_tmap_bf16(&ty, y, 4, dims, str, box, cw == 64))) return rc;
  }
  conv::Params p;
  p.T = T_out; p.H = H_out; p.W = W_out; p.Cout = Cout; p.cin_chunks = cin_pad / 64;
  p.kt = kt; p.kh = kh; p.kw = kw; p.stride = stride; p.pad_h = pad_h; p.pad_w = pad_w;
  p.tiles_t = T_out;
  p.tiles_x = (int)ceil_div(W_out, conv::BW);
  p.tiles_y = (int)ceil_div(H_out, conv::BH);
  p.num_n = (int)ceil_div(Cout, bn);
  p.bias = bias;
  p.resid = reinterpret_cast<const __nv_bfloat16*>(resid);
  p.pair_axis = -1;
  if (allow_pairs && bn >= 128 && num_sms() >= 2) {
    // CTA pairs: two neighbouring tiles along the axis whose (rounded-up-to-even) tile count wastes the least
    const int n_ax[3] = {p.tiles_t, p.tiles_y, p.tiles_x};
    int best = -1;
    double best_waste = 1e9;
    for (int a = 2; a >= 0; --a) {                       // ties: x, then y, then t (neighbouring boxes share L2 lines)
      const double waste = double((n_ax[a] + 1) / 2 * 2) / n_ax[a];
      if (waste < best_waste - 1e-9) { best_waste = waste; best = a; }
    }
    CUtensorMap twh;
    if ((rc = make_tmap_2d(&twh, w_packed, Cout, uint64_t(taps) * cin_pad, uint64_t(taps) * cin_pad, bn / 2, conv::BK)))
      return rc;
    conv::Params p2 = p;
    p2.pair_axis = best;
    if (best == 0) p2.tiles_t = (p.tiles_t + 1) / 2;
    else if (best == 1) p2.tiles_y = (p.tiles_y + 1) / 2;
    else p2.tiles_x = (p.tiles_x + 1) / 2;
    return bn == 256 ? conv::launch2<256>(tx, twh, ty, p2, stream) : conv::launch2<128>(tx, twh, ty, p2, stream);
  }
  if (bn == 32) return conv::launch<32>(tx, tw, ty, p, stream);
  if (bn == 256) return conv::launch<256>(tx, tw, ty, p, stream);
  return conv::launch<128>(tx, tw, ty, p, stream);
}

int conv3d_bf16(const void* x, int T_in, int H_in, int W_in, int Cin, const void* w_packed, const float* bias,
                const void* resid, void* y, int T_out, int H_out, int W_out, int Cout, int kt, int kh, int kw,
                int stride, int pad_h, int pad_w, cudaStream_t stream) {
  return conv3d_bf16_impl(x, T_in, H_in, W_in, Cin, w_packed, bias, resid, y, T_out, H_out, W_out, Cout, kt, kh, kw, stride,
                          pad_h, pad_w, true, stream);
}

}  // namespace aether

// Same operation on the one-CTA kernel only (the CTA-pair kernel's reference in the parity tests and the A/B timing).
extern "C" int aether_conv3d_bf16_1cta(const void* x, int32_t T_in, int32_t H_in, int32_t W_in, int32_t Cin,
                                       const void* w_packed, const float* bias, const void* resid, void* y,
                                       int32_t T_out, int32_t H_out, int32_t W_out, int32_t Cout, int32_t kt, int32_t kh,
                                       int32_t kw, int32_t stride, int32_t pad_h, int32_t pad_w, void* stream) {
  return aether::conv3d_bf16_impl(x, T_in, H_in, W_in, Cin, w_packed, bias, resid, y, T_out, H_out, W_out, Cout, kt, kh,
                                  kw, stride, pad_h, pad_w, false, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int aether_conv3d_bf16(const void* x, int32_t T_in, int32_t H_in, int32_t W_in, int32_t Cin,
                                  const void* w_packed, const float* bias, const void* resid, void* y, int32_t T_out,
                                  int32_t H_out, int32_t W_out, int32_t Cout, int32_t kt, int32_t kh, int32_t kw,
                                  int32_t stride, int32_t pad_h, int32_t pad_w, void* stream) {
  return aether::conv3d_bf16(x, T_in, H_in, W_in, Cin, w_packed, bias, resid, y, T_out, H_out, W_out, Cout, kt, kh, kw,
                             stride, pad_h, pad_w, reinterpret_cast<cudaStream_t>(stream));
}
