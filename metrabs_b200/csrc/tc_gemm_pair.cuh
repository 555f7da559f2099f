// tcgen05 CTA-PAIR GEMM (cta_group::2) for the flat 1x1 convolutions:  D[M, N] = A[M, K] * W[N, K]^T  (+ bias, activation,
// residual, bf16 NHWC out) with a 256-row tile per pair of CTAs on the two SMs of a TPC.
//
// STATUS: written at the end of round 1 WITHOUT access to a GPU - it compiles for sm_100a and follows the 2-SM recipes of
// the CUTLASS sm100 headers (cute/arch/copy_sm100_tma.hpp SM100_TMA_2SM_LOAD_2D, cute/arch/mma_sm100_umma.hpp
// SM100_MMA_F16BF16_2x1SM_SS, cute/arch/tmem_allocator_sm100.hpp Allocator2Sm, cutlass/arch/barrier.h
// umma_arrive_multicast_2x1SM), but it has never run.  It is OFF unless MTB_TC_PAIR=1; validate it with
//   MTB_TC_PAIR=1 python -m pytest tests/test_gpu_tc.py -k tc_ops      (tensor-core ops vs the CUDA-core kernels)
// before trusting a number from it.
//
// Why (DESIGN.md, round-1 measurements): the projection GEMMs stream A from HBM through a 3-4 stage ring of 44 KB stages
// (16 KB of A + the whole 28 KB B tile) and are bound by bytes in flight (~2100 cycles per stage round trip against ~460
// cycles of MMA work); the large 3x3 convs are bound by L2->SM operand traffic because every CTA re-reads the whole B
// tile per k-block.  In a CTA pair each CTA stages only HALF of the B tile (the MMA reads the other half from the peer's
// shared memory), so a stage is 16 + 14 KB (6 in flight in the same budget) and the B traffic per SM halves.
//
// Roles (320 threads per CTA):  warps 0-3 epilogue (TMEM lanes 32w..32w+31 = tile rows of THIS CTA, direct 64-byte row
// stores) | warp 4 TMA producer (own A half + own B half; the transaction bytes of both CTAs land on the LEADER's full
// barrier) | warp 5 TMEM allocation (both CTAs, cta_group::2) and, in the leader CTA only, the single-thread MMA issuer |
// warps 6-9 (SCALE variant, MTB_TC_PAIR_SCALE=1) squeeze-excitation scalers: each CTA multiplies ITS A tile in shared
// memory by s[crop(row)][k] once it has landed (A then completes on a CTA-local barrier) and arrives on the leader's
// `scaled` barrier; with 6 stages in flight the extra hand-off should no longer cost what the in-place se_scale_kernel
// pass costs (2.4 ms per step) - in the single-CTA kernel it did.
#pragma once
#include "tc_gemm.cuh"

namespace mtb {

constexpr int TP_THREADS = 320;
constexpr int TP_BM = 128;                  // rows per CTA (256 per pair)
constexpr int TP_BK = 64;
constexpr int TP_A_BYTES = TP_BM * TP_BK * 2;   // 16 KB
constexpr int TP_MAX_STAGES = 12;
constexpr int TP_RING_BYTES = 200 * 1024;
constexpr int TP_BAR_OFF = TP_RING_BYTES;
constexpr int TP_SMEM_BYTES = TP_BAR_OFF + 512 + 1024 /*align slack*/;
constexpr uint32_t TP_PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address: the even (leader) CTA

struct TcPairParams {
  const void* res;
  const float* bias;
  __nv_bfloat16* out;
  int M, Cout, Cin;
  int bn;         // N-tile stride (multiple of 32, <= 256); each CTA stages bn / 2 weight rows per k-block
  int n_tiles, m_pairs, kchunks;
  int nstages, stage_stride;
  const float* a_scale;  // SCALE variant: squeeze-excitation scale [B][Cin] fp32
  int a_scale_P;         // pixels per crop (row / P = crop index)
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-SM TMA load: data into THIS CTA's shared memory, transaction bytes onto the mbarrier at the same offset in the LEADER CTA
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar & TP_PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
// D[tmem, both CTAs] (+)= A[smem of each CTA: its 128 rows] * B[smem: N/2 rows from each CTA], issued by ONE thread of the leader
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all prior tcgen05.mma of this thread have completed) on the mbarrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
// arrive on the mbarrier at this offset in the LEADER CTA (rank 0 of the pair), from either CTA
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(bar));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// kind::f16 instruction descriptor for a pair: D fp32, A/B bf16, both K-major, M = 256, N = n
__host__ __device__ inline uint32_t umma_idesc_bf16_m256(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <int ACT, int RES, bool SCALE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TP_THREADS, 1)
tc_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcPairParams p) {
  extern __shared__ uint8_t tp_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)tp_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + TP_BAR_OFF);
  uint64_t* full = bars;                               // [12] used in the leader: bytes of BOTH CTAs' loads
  uint64_t* empty = bars + TP_MAX_STAGES;              // [12] per CTA: the pair's MMAs have read this slot
  uint64_t* tmem_full = bars + 2 * TP_MAX_STAGES;      // [2]  per CTA: accumulator complete
  uint64_t* tmem_empty = tmem_full + 2;                // [2]  used in the leader: all 8 epilogue warps of the pair drained it
  uint64_t* a_full = tmem_empty + 2;                   // [12] SCALE: per CTA, its own A tile has landed
  uint64_t* scaled = a_full + TP_MAX_STAGES;           // [12] SCALE: used in the leader, all 8 scaler warps of the pair are done
  uint32_t* tmem_slot = (uint32_t*)(scaled + TP_MAX_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < TP_MAX_STAGES; ++i) {
      mbar_init(&full[i], 1);    // the leader producer's arrive.expect_tx
      mbar_init(&empty[i], 1);   // tcgen05.commit (multicast to both CTAs)
      mbar_init(&a_full[i], 1);  // SCALE: this CTA's producer
      mbar_init(&scaled[i], 8);  // SCALE: 4 scaler warps x 2 CTAs
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);  // 4 epilogue warps x 2 CTAs (arrivals land in the leader)
    }
    fence_barrier_init();
  }
  if (warp == 5) {  // same logical warp in both CTAs (Allocator2Sm contract)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised and the TMEM allocation visible before any cross-CTA traffic
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  const int npairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const int pair_tiles = p.m_pairs * p.n_tiles;
  const int bnh = p.bn >> 1;                               // weight rows staged per CTA and k-block
  const uint32_t bh_bytes = (uint32_t)bnh * TP_BK * 2;
  // bytes landing on the leader's full barrier per stage: both CTAs' loads (SCALE: only the weight halves; A completes locally)
  const uint32_t stage_tx = SCALE ? 2u * bh_bytes : 2u * ((uint32_t)TP_A_BYTES + bh_bytes);
  const uint32_t a_full0 = smem_u32(a_full), scaled0 = smem_u32(scaled);
  const int nstages = p.nstages;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
  const uint32_t tmem_full0 = smem_u32(tmem_full), tmem_empty0 = smem_u32(tmem_empty);

  if (warp == 4) {
    // ===== TMA producer (both CTAs): own 128 rows of A, own half of the weight tile =====
    uint32_t stage = 0, phase = 0;
    for (int u = pair; u < pair_tiles; u += npairs) {
      const int m_pair = u / p.n_tiles, n_blk = u - m_pair * p.n_tiles;
      const int row0 = (m_pair * 2 + (int)rank) * TP_BM;
      const int n0 = n_blk * p.bn;
      const int n_mma = (min(p.bn, p.Cout - n0) + 15) & ~15;       // MMA N of this tile; each CTA supplies n_mma / 2 rows
      const int brow = n0 + (int)rank * (n_mma >> 1);
#pragma unroll 1
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait_a(empty0 + stage * 8, phase ^ 1);
        if (elect_one()) {
          const uint32_t sa = smem_base + stage * p.stage_stride;
          if (leader) mbar_expect_tx_a(full0 + stage * 8, stage_tx);
          if constexpr (SCALE) {
            mbar_expect_tx_a(a_full0 + stage * 8, (uint32_t)TP_A_BYTES);
            tma_load_2d_a(sa, &tmA, a_full0 + stage * 8, kc * TP_BK, row0);
          } else {
            tma_load_2d_2sm(sa, &tmA, full0 + stage * 8, kc * TP_BK, row0);
          }
          tma_load_2d_2sm(sa + TP_A_BYTES, &tmB, full0 + stage * 8, kc * TP_BK, brow);
        }
        __syncwarp();
        if (++stage == (uint32_t)nstages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5 && leader) {
    // ===== MMA issuer (leader CTA only): M = 256 across the pair =====
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    constexpr uint32_t hi_sw = (uint32_t)((8 * TP_BK * 2) >> 4) | (1u << 14) | (2u << 29);
    const uint32_t base16 = smem_base >> 4, stride16 = (uint32_t)p.stage_stride >> 4, b_off16 = TP_A_BYTES >> 4;
    for (int u = pair; u < pair_tiles; u += npairs) {
      const int n_blk = u % p.n_tiles;
      const int n_mma = (min(p.bn, p.Cout - n_blk * p.bn) + 15) & ~15;
      const uint32_t idesc = umma_idesc_bf16_m256(n_mma);
      mbar_wait_a(tmem_empty0 + acc * 8, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * TC_MAX_BN;
#pragma unroll 1
      for (int kb = 0; kb < p.kchunks; ++kb) {
        mbar_wait_a(full0 + stage * 8, phase);
        if constexpr (SCALE) mbar_wait_a(scaled0 + stage * 8, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a16 = base16 + stage * stride16;
#pragma unroll
          for (int k = 0; k < TP_BK / 16; ++k)
            umma_bf16_2sm(d_tmem, make_desc(a16 + 2 * k, hi_sw), make_desc(a16 + b_off16 + 2 * k, hi_sw), idesc, (uint32_t)(kb | k));
          umma_commit_2sm(empty0 + stage * 8);                            // both CTAs may refill this slot
          if (kb == p.kchunks - 1) umma_commit_2sm(tmem_full0 + acc * 8);  // both CTAs' epilogues may read the accumulator
        }
        __syncwarp();
        if (++stage == (uint32_t)nstages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp < 4) {
    // ===== epilogue (both CTAs): this CTA's 128 tile rows; lane = one row, 32-column chunks =====
    uint32_t acc = 0, acc_phase = 0;
    const __nv_bfloat16* __restrict__ res = (const __nv_bfloat16*)p.res;
    for (int u = pair; u < pair_tiles; u += npairs) {
      const int m_pair = u / p.n_tiles, n_blk = u - m_pair * p.n_tiles;
      const int m = (m_pair * 2 + (int)rank) * TP_BM + warp * 32 + lane;
      const bool valid = m < p.M;
      const int n0 = n_blk * p.bn;
      const int n_valid = min(p.bn, p.Cout - n0);
      const size_t off = (size_t)m * p.Cout + n0;
      mbar_wait_a(tmem_full0 + acc * 8, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * TC_MAX_BN;
      for (int c0 = 0; c0 < n_valid; c0 += 32) {
        uint4 rv[4];
        if constexpr (RES != 0) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            rv[g] = make_uint4(0u, 0u, 0u, 0u);
            if (valid && c0 + g * 8 < n_valid) rv[g] = *reinterpret_cast<const uint4*>(res + off + c0 + g * 8);
          }
        }
        uint32_t v[32];
        tmem_ld16_issue(taddr + c0, v);
        if (c0 + 16 < n_valid) tmem_ld16_issue(taddr + c0 + 16, v + 16);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (c0 + g * 8 >= n_valid) continue;
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + g * 8));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + g * 8 + 4));
          float o[8];
          o[0] = __uint_as_float(v[g * 8 + 0]) + b0.x; o[1] = __uint_as_float(v[g * 8 + 1]) + b0.y;
          o[2] = __uint_as_float(v[g * 8 + 2]) + b0.z; o[3] = __uint_as_float(v[g * 8 + 3]) + b0.w;
          o[4] = __uint_as_float(v[g * 8 + 4]) + b1.x; o[5] = __uint_as_float(v[g * 8 + 5]) + b1.y;
          o[6] = __uint_as_float(v[g * 8 + 6]) + b1.z; o[7] = __uint_as_float(v[g * 8 + 7]) + b1.w;
          if constexpr (RES != 0) {
            const unsigned wd[4] = {rv[g].x, rv[g].y, rv[g].z, rv[g].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float r0 = __uint_as_float(wd[i] << 16), r1 = __uint_as_float(wd[i] & 0xffff0000u);
              o[2 * i] = RES == 2 ? tc_act<ACT>(o[2 * i] + r0) : tc_act<ACT>(o[2 * i]) + r0;
              o[2 * i + 1] = RES == 2 ? tc_act<ACT>(o[2 * i + 1] + r1) : tc_act<ACT>(o[2 * i + 1]) + r1;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = tc_act<ACT>(o[i]);
          }
          uint4 ov;
          __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
          for (int i = 0; i < 4; ++i) o2[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
          if (valid) *reinterpret_cast<uint4*>(p.out + off + c0 + g * 8) = ov;
        }
      }
      // this warp is done with the accumulator: tell the leader's MMA issuer (8 arrivals release it)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tmem_empty0 + acc * 8);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  if constexpr (SCALE) {
    if (warp >= 6) {
      // ===== squeeze-excitation scalers (both CTAs): same mapping as tc_conv_kernel's - warp w owns rows [32w, 32w+32), a
      // quarter-warp covers one 128-byte row (bank-conflict free); fp32 product rounded once to bf16 =====
      const int sw = warp - 6;
      const int sub = lane >> 3, pos = lane & 7;
      uint32_t stage = 0, phase = 0;
      for (int u = pair; u < pair_tiles; u += npairs) {
        const int m_pair = u / p.n_tiles;
        const int m0 = (m_pair * 2 + (int)rank) * TP_BM + sw * 32 + sub;
        int soff[8];
        uint32_t rok = 0;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int m = m0 + it * 4;
          const bool ok = m < p.M;
          rok |= ok ? (1u << it) : 0u;
          soff[it] = (ok ? m / p.a_scale_P : 0) * p.Cin;
        }
#pragma unroll 1
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait_a(a_full0 + stage * 8, phase);
          uint8_t* sa = smem + stage * p.stage_stride + (sw * 32 + sub) * 128 + pos * 16;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int k = kc * 64 + ((pos ^ (((it & 1) << 2) | sub)) << 3);  // 128B swizzle: logical chunk at position pos
            if (!((rok >> it) & 1u) || k >= p.Cin) continue;
            uint4* ptr = reinterpret_cast<uint4*>(sa + it * 512);
            uint4 v = *ptr;
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.a_scale + soff[it] + k));
            const float4 s1 = __ldg(reinterpret_cast<const float4*>(p.a_scale + soff[it] + k + 4));
            const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
            unsigned wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float lo = __uint_as_float(wd[i] << 16) * sc[2 * i];
              const float hi = __uint_as_float(wd[i] & 0xffff0000u) * sc[2 * i + 1];
              __nv_bfloat162 pk = __floats2bfloat162_rn(lo, hi);
              wd[i] = *reinterpret_cast<unsigned*>(&pk);
            }
            *ptr = make_uint4(wd[0], wd[1], wd[2], wd[3]);
          }
          fence_proxy_async();  // generic-proxy writes -> visible to the tensor cores (async proxy) of the pair
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(scaled0 + stage * 8);
          if (++stage == (uint32_t)nstages) { stage = 0; phase ^= 1; }
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();  // neither CTA may exit (or free TMEM) while its peer can still address its shared memory / barriers
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

struct TcPairMaps {
  CUtensorMap a, b;
  const void* in = nullptr;
  int B = -1, bn = 0;
};

inline bool tc_pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MTB_TC_PAIR");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}
// flat 1x1 stride-1 GEMMs with 128-byte A rows and no fused SE scale
inline bool tc_pair_eligible(const ConvParams& p) {
  return tc_pair_enabled() && p.R == 1 && p.S == 1 && p.stride == 1 && p.Cin > 32 && p.Cin % 8 == 0 && p.Cout % 8 == 0 &&
         (long)p.B * p.Hout * p.Wout >= 256;
}

template <int ACT, int RES, bool SCALE>
inline const char* tc_pair_launch_k(int grid, const CUtensorMap& a, const CUtensorMap& b, const TcPairParams& q, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(tc_gemm_pair_kernel<ACT, RES, SCALE>, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM_BYTES) !=
        cudaSuccess)
      return "cannot raise dynamic shared memory for tc_gemm_pair_kernel";
    attr_set = true;
  }
  launch_k(tc_gemm_pair_kernel<ACT, RES, SCALE>, dim3(grid), dim3(TP_THREADS), TP_SMEM_BYTES, st, a, b, q);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
template <int ACT, int RES>
inline const char* tc_pair_launch_t(int grid, const CUtensorMap& a, const CUtensorMap& b, const TcPairParams& q, cudaStream_t st) {
  if constexpr (ACT == ACT_NONE) {  // the scaled GEMMs are the MBConv projections: no activation
    if (q.a_scale) return tc_pair_launch_k<ACT, RES, true>(grid, a, b, q, st);
  }
  if (q.a_scale) return "fused SE scale in the pair GEMM is only built for activation-free projections";
  return tc_pair_launch_k<ACT, RES, false>(grid, a, b, q, st);
}
template <int ACT>
inline const char* tc_pair_launch_res(int res_mode, int grid, const CUtensorMap& a, const CUtensorMap& b, const TcPairParams& q,
                                      cudaStream_t st) {
  switch (res_mode) {
    case 0: return tc_pair_launch_t<ACT, 0>(grid, a, b, q, st);
    case 1: return tc_pair_launch_t<ACT, 1>(grid, a, b, q, st);
    default: return tc_pair_launch_t<ACT, 2>(grid, a, b, q, st);
  }
}

// w: the op's tensor-core weights ([Cout][Cin] bf16 K-major + fp32 bias); maps cached per (input pointer, batch)
inline bool tc_pair_scale_enabled() {  // MTB_TC_PAIR_SCALE=1 (with MTB_TC_PAIR=1): SE scale fused into the pair GEMM
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MTB_TC_PAIR_SCALE");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}
// the projection behind a squeeze-excitation runs as a pair GEMM that applies the scale itself (no se_scale_kernel pass)
inline bool tc_pair_fuses_scale(const ConvParams& p, bool has_scale) {
  return has_scale && tc_pair_scale_enabled() && tc_pair_eligible(p) && p.act == ACT_NONE;
}

inline const char* tc_pair_launch(const TcWeights& w, TcPairMaps& maps, const ConvParams& p, bool res_first, bool fuse_scale,
                                  cudaStream_t st) {
  TcPairParams q;
  q.a_scale = fuse_scale ? p.a_scale : nullptr;
  q.a_scale_P = p.Hin * p.Win;
  q.res = p.res; q.bias = w.d_bias; q.out = (__nv_bfloat16*)p.out;
  q.M = p.B * p.Hout * p.Wout; q.Cout = p.Cout; q.Cin = p.Cin;
  const int nt = (p.Cout + 255) / 256;
  q.bn = (((p.Cout + nt - 1) / nt) + 31) / 32 * 32;   // even split, multiple of 32 (each CTA stages bn / 2 rows: multiple of 16)
  if (q.bn > 256) q.bn = 256;
  q.n_tiles = (p.Cout + q.bn - 1) / q.bn;
  q.m_pairs = (q.M + 2 * TP_BM - 1) / (2 * TP_BM);
  q.kchunks = (p.Cin + TP_BK - 1) / TP_BK;
  q.stage_stride = (TP_A_BYTES + (q.bn / 2) * TP_BK * 2 + 1023) / 1024 * 1024;
  q.nstages = TP_RING_BYTES / q.stage_stride;
  if (q.nstages > TP_MAX_STAGES) q.nstages = TP_MAX_STAGES;
  if (q.nstages < 2) return "operand ring too small for the pair tile";
  if (maps.in != p.in || maps.B != p.B || maps.bn != q.bn) {
    const char* e = make_tmap_2d(&maps.a, p.in, (uint64_t)q.M, (uint64_t)p.Cin, TP_BM, TP_BK);
    if (e) return e;
    e = make_tmap_2d(&maps.b, w.d_w, (uint64_t)p.Cout, (uint64_t)p.Cin, (uint32_t)(q.bn / 2), TP_BK);
    if (e) return e;
    maps.in = p.in; maps.B = p.B; maps.bn = q.bn;
  }
  const int pair_tiles = q.m_pairs * q.n_tiles;
  const int npairs = pair_tiles < 74 ? pair_tiles : 74;
  const int grid = 2 * npairs;
  const int res_mode = p.res ? (res_first ? 2 : 1) : 0;
  switch (p.act) {
    case ACT_NONE: return tc_pair_launch_res<ACT_NONE>(res_mode, grid, maps.a, maps.b, q, st);
    case ACT_SILU: return tc_pair_launch_res<ACT_SILU>(res_mode, grid, maps.a, maps.b, q, st);
    case ACT_RELU: return tc_pair_launch_res<ACT_RELU>(res_mode, grid, maps.a, maps.b, q, st);
    case ACT_HSWISH: return tc_pair_launch_res<ACT_HSWISH>(res_mode, grid, maps.a, maps.b, q, st);
    default: return "unsupported activation in tc_gemm_pair_kernel";
  }
}

}  // namespace mtb
