// tcgen05 GEMM with a RESIDENT weight panel for the short-K flat 1x1 convolutions (the MBConv expand GEMMs):
//   D[M, N] = A[M, K] * W[N, K]^T (+ bias, activation, residual, bf16 NHWC out),   K <= ~640, N large.
//
// STATUS: written at the end of round 1 WITHOUT access to a GPU - it compiles for sm_100a, reuses the descriptor / barrier
// helpers the measured tc_conv_kernel runs on, but has never run.  OFF unless MTB_TC_BRES=1; validate with
//   MTB_TC_BRES=1 python -m pytest tests/test_gpu_tc.py -k tc_ops
//
// Why: counted with the weights once per launch (not once per crop, as the round-1 per-op table did), the expand GEMMs move
// 2.2-2.8 TB/s and reach 0.4-0.7 PFLOP/s - neither roofline; they are bound by L2->SM operand traffic: tc_conv_kernel
// streams [A 16 KB | B 28 KB] per k-block, i.e. 176 KB per 128 x 224 output tile of the 224 -> 1344 expand (4100 cycles at
// the 42.6 B/clk/SM L2 cap against 1800 cycles of MMA).  With K this short the whole weight panel of one N tile fits in
// shared memory (4 k-blocks x 28 KB), so here a CTA keeps ITS N tile's panel resident (grid = a multiple of the number of
// N tiles, so a CTA's N tile never changes) and streams only the A tiles: 64 KB per output tile.
//
// Roles (320 threads): warps 0-7 epilogue (TMEM lane quarter w & 3, alternating 32-column chunks by w >> 2; direct 64-byte
// row stores) | warp 8 TMA producer (the weight panel once, then the A ring) | warp 9 TMEM allocator + MMA issuer.
#pragma once
#include "tc_gemm.cuh"

namespace mtb {

constexpr int TR_THREADS = 320;
constexpr int TR_BM = 128, TR_BK = 64;
constexpr int TR_A_BYTES = TR_BM * TR_BK * 2;  // 16 KB
constexpr int TR_MAX_KB = 12;                   // k-blocks of the resident panel
constexpr int TR_MAX_ASTAGES = 8;
constexpr int TR_DATA_BYTES = 200 * 1024;       // weight panel + A ring
constexpr int TR_BAR_OFF = TR_DATA_BYTES;
constexpr int TR_SMEM_BYTES = TR_BAR_OFF + 512 + 1024 /*align slack*/;

struct TcBresParams {
  const void* res;
  const float* bias;
  __nv_bfloat16* out;
  int M, Cout, Cin;
  int bn;                 // N tile (multiple of 16, <= 256)
  int n_tiles, m_tiles, kchunks;
  int panel_stride;       // bytes per k-block of the panel (bn * 128, 1024-aligned)
  int a_off, a_stages;    // A ring behind the panel
};

template <int ACT, int RES>
__global__ void __launch_bounds__(TR_THREADS, 1)
tc_gemm_bres_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcBresParams p) {
  extern __shared__ uint8_t tr_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)tr_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + TR_BAR_OFF);
  uint64_t* b_full = bars;                          // [12] k-block kb of the panel has landed (used once)
  uint64_t* a_full = bars + TR_MAX_KB;              // [8]
  uint64_t* a_empty = a_full + TR_MAX_ASTAGES;      // [8]
  uint64_t* tmem_full = a_empty + TR_MAX_ASTAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;             // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < TR_MAX_KB; ++i) mbar_init(&b_full[i], 1);
    for (int i = 0; i < TR_MAX_ASTAGES; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  // this CTA's N tile is fixed (gridDim.x is a multiple of n_tiles); it walks the M tiles m_first, m_first + m_step, ...
  const int n_blk = (int)(blockIdx.x % (unsigned)p.n_tiles);
  const int m_first = (int)(blockIdx.x / (unsigned)p.n_tiles), m_step = (int)(gridDim.x / (unsigned)p.n_tiles);
  const int n0 = n_blk * p.bn;
  const int n_valid = min(p.bn, p.Cout - n0);
  const int n_mma = (n_valid + 15) & ~15;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t b_full0 = smem_u32(b_full), a_full0 = smem_u32(a_full), a_empty0 = smem_u32(a_empty);
  const uint32_t tmem_full0 = smem_u32(tmem_full), tmem_empty0 = smem_u32(tmem_empty);
  const uint32_t b_bytes = (uint32_t)p.bn * TR_BK * 2;

  if (warp == 8) {
    // ===== TMA producer: the weight panel of this N tile once, then the A tiles =====
    if (elect_one()) {
      for (int kb = 0; kb < p.kchunks; ++kb) {
        mbar_expect_tx_a(b_full0 + kb * 8, b_bytes);
        tma_load_2d_a(smem_base + kb * p.panel_stride, &tmB, b_full0 + kb * 8, kb * TR_BK, n0);
      }
    }
    __syncwarp();
    uint32_t stage = 0, phase = 0;
    for (int m_blk = m_first; m_blk < p.m_tiles; m_blk += m_step) {
#pragma unroll 1
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait_a(a_empty0 + stage * 8, phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx_a(a_full0 + stage * 8, (uint32_t)TR_A_BYTES);
          tma_load_2d_a(smem_base + p.a_off + stage * TR_A_BYTES, &tmA, a_full0 + stage * 8, kc * TR_BK, m_blk * TR_BM);
        }
        __syncwarp();
        if (++stage == (uint32_t)p.a_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer =====
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    constexpr uint32_t hi_sw = (uint32_t)((8 * TR_BK * 2) >> 4) | (1u << 14) | (2u << 29);
    const uint32_t base16 = smem_base >> 4, a_off16 = (uint32_t)p.a_off >> 4, panel16 = (uint32_t)p.panel_stride >> 4;
    const uint32_t idesc = umma_idesc_bf16(n_mma);
    bool first = true;
    for (int m_blk = m_first; m_blk < p.m_tiles; m_blk += m_step) {
      mbar_wait_a(tmem_empty0 + acc * 8, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * TC_MAX_BN;
#pragma unroll 1
      for (int kb = 0; kb < p.kchunks; ++kb) {
        if (first) mbar_wait_a(b_full0 + kb * 8, 0);
        mbar_wait_a(a_full0 + stage * 8, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a16 = base16 + a_off16 + stage * (TR_A_BYTES >> 4);
          const uint32_t b16 = base16 + (uint32_t)kb * panel16;
#pragma unroll
          for (int k = 0; k < TR_BK / 16; ++k)
            umma_bf16(d_tmem, make_desc(a16 + 2 * k, hi_sw), make_desc(b16 + 2 * k, hi_sw), idesc, (uint32_t)(kb | k));
          umma_commit_a(a_empty0 + stage * 8);
          if (kb == p.kchunks - 1) umma_commit_a(tmem_full0 + acc * 8);
        }
        __syncwarp();
        if (++stage == (uint32_t)p.a_stages) { stage = 0; phase ^= 1; }
      }
      first = false;
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp < 8) {
    // ===== epilogue: lane = one tile row; warps w and w + 4 share a TMEM lane quarter and alternate 32-column chunks =====
    const int q = warp & 3, par = warp >> 2;
    uint32_t acc = 0, acc_phase = 0;
    const __nv_bfloat16* __restrict__ res = (const __nv_bfloat16*)p.res;
    for (int m_blk = m_first; m_blk < p.m_tiles; m_blk += m_step) {
      const int m = m_blk * TR_BM + q * 32 + lane;
      const bool valid = m < p.M;
      const size_t off = (size_t)m * p.Cout + n0;
      mbar_wait_a(tmem_full0 + acc * 8, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * TC_MAX_BN;
      for (int c0 = par * 32; c0 < n_valid; c0 += 64) {
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

struct TcBresMaps {
  CUtensorMap a, b;
  const void* in = nullptr;
  int B = -1, bn = 0;
};

inline bool tc_bres_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MTB_TC_BRES");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// N tile for which the whole [kchunks][bn][64] weight panel plus >= 3 A stages fit the 200 KB budget; 0 = not eligible
inline int tc_bres_pick_bn(int cin, int cout) {
  const int kchunks = (cin + TR_BK - 1) / TR_BK;
  if (kchunks > TR_MAX_KB) return 0;
  const int budget = TR_DATA_BYTES - 3 * TR_A_BYTES;
  int bn_max = budget / (kchunks * TR_BK * 2);
  bn_max = bn_max / 16 * 16;
  if (bn_max > 256) bn_max = 256;
  if (bn_max < 128 && bn_max < (cout + 15) / 16 * 16) return 0;   // narrow MMAs would give the gain back
  const int nt = (cout + bn_max - 1) / bn_max;
  const int bn = (((cout + nt - 1) / nt) + 15) / 16 * 16;        // even split, multiple of 16
  return bn;
}
inline bool tc_bres_eligible(const ConvParams& p) {
  if (!tc_bres_enabled() || p.R != 1 || p.S != 1 || p.stride != 1 || p.Cin <= 32 || p.Cin % 8 != 0 || p.Cout % 8 != 0) return false;
  const int bn = tc_bres_pick_bn(p.Cin, p.Cout);
  if (bn == 0) return false;
  const int n_tiles = (p.Cout + bn - 1) / bn;
  const long m_tiles = ((long)p.B * p.Hout * p.Wout + TR_BM - 1) / TR_BM;
  return n_tiles <= 148 && m_tiles * n_tiles >= 148;  // every CTA gets several M tiles for its resident panel
}

template <int ACT, int RES>
inline const char* tc_bres_launch_t(int grid, const CUtensorMap& a, const CUtensorMap& b, const TcBresParams& q, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(tc_gemm_bres_kernel<ACT, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, TR_SMEM_BYTES) != cudaSuccess)
      return "cannot raise dynamic shared memory for tc_gemm_bres_kernel";
    attr_set = true;
  }
  launch_k(tc_gemm_bres_kernel<ACT, RES>, dim3(grid), dim3(TR_THREADS), TR_SMEM_BYTES, st, a, b, q);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
template <int ACT>
inline const char* tc_bres_launch_res(int res_mode, int grid, const CUtensorMap& a, const CUtensorMap& b, const TcBresParams& q,
                                      cudaStream_t st) {
  switch (res_mode) {
    case 0: return tc_bres_launch_t<ACT, 0>(grid, a, b, q, st);
    case 1: return tc_bres_launch_t<ACT, 1>(grid, a, b, q, st);
    default: return tc_bres_launch_t<ACT, 2>(grid, a, b, q, st);
  }
}

inline const char* tc_bres_launch(const TcWeights& w, TcBresMaps& maps, const ConvParams& p, bool res_first, cudaStream_t st) {
  TcBresParams q;
  q.res = p.res; q.bias = w.d_bias; q.out = (__nv_bfloat16*)p.out;
  q.M = p.B * p.Hout * p.Wout; q.Cout = p.Cout; q.Cin = p.Cin;
  q.bn = tc_bres_pick_bn(p.Cin, p.Cout);
  if (q.bn == 0) return "weight panel does not fit shared memory";
  q.n_tiles = (p.Cout + q.bn - 1) / q.bn;
  q.m_tiles = (q.M + TR_BM - 1) / TR_BM;
  q.kchunks = (p.Cin + TR_BK - 1) / TR_BK;
  q.panel_stride = (q.bn * TR_BK * 2 + 1023) / 1024 * 1024;
  q.a_off = q.kchunks * q.panel_stride;
  q.a_stages = (TR_DATA_BYTES - q.a_off) / TR_A_BYTES;
  if (q.a_stages > TR_MAX_ASTAGES) q.a_stages = TR_MAX_ASTAGES;
  if (q.a_stages < 2) return "no room for the A ring behind the weight panel";
  if (maps.in != p.in || maps.B != p.B || maps.bn != q.bn) {
    const char* e = make_tmap_2d(&maps.a, p.in, (uint64_t)q.M, (uint64_t)p.Cin, TR_BM, TR_BK);
    if (e) return e;
    e = make_tmap_2d(&maps.b, w.d_w, (uint64_t)p.Cout, (uint64_t)p.Cin, (uint32_t)q.bn, TR_BK);
    if (e) return e;
    maps.in = p.in; maps.B = p.B; maps.bn = q.bn;
  }
  // grid: the largest multiple of n_tiles that fits the SMs, so that (blockIdx.x mod n_tiles) is a CTA's only N tile
  int per_tile = 148 / q.n_tiles;
  if (per_tile > q.m_tiles) per_tile = q.m_tiles;
  if (per_tile < 1) return "more N tiles than SMs";
  const int grid = per_tile * q.n_tiles;
  const int res_mode = p.res ? (res_first ? 2 : 1) : 0;
  switch (p.act) {
    case ACT_NONE: return tc_bres_launch_res<ACT_NONE>(res_mode, grid, maps.a, maps.b, q, st);
    case ACT_SILU: return tc_bres_launch_res<ACT_SILU>(res_mode, grid, maps.a, maps.b, q, st);
    case ACT_RELU: return tc_bres_launch_res<ACT_RELU>(res_mode, grid, maps.a, maps.b, q, st);
    case ACT_HSWISH: return tc_bres_launch_res<ACT_HSWISH>(res_mode, grid, maps.a, maps.b, q, st);
    default: return "unsupported activation in tc_gemm_bres_kernel";
  }
}

}  // namespace mtb
