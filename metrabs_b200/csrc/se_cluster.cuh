// Squeeze-excitation FCs in ONE launch on a thread-block cluster (bf16 throughput mode):
//   scale[b, c] = act2(W2^T act1(W1^T mean[b] + b1) + b2)     (SqueezeExcitation.forward via backbones/efficientnet.py:110-173)
// The three-launch path (split-K fc1, reduce, fc2) cost ~25 us per MBConv block whatever the batch (launch / drain
// latency of tiny kernels, 2 ms of a 23 ms step); the one-CTA-per-crop kernel (se_fused_kernel) made every CTA stream
// both weight matrices through one SM.  Here a cluster of 8 CTAs owns 8 crops: CTA r takes the r-th K slice of fc1 and
// the r-th N slice of fc2, so each cluster reads the weights once (28 MB of L2 traffic per block at 256 crops), and the
// fc1 partial sums are exchanged through distributed shared memory in a fixed order (deterministic).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace mtb {

constexpr int SEC_CL = 8;        // CTAs per cluster
constexpr int SEC_CB_MAX = 16;   // crops per cluster: template argument CB = 8 or 16
constexpr int SEC_THREADS = 256;
constexpr int SEC_MAX_JPL = 5;   // hidden units per lane: csq <= 160
constexpr int SEC_MAX_C = 4096;

inline int sec_slice(int C) { return ((C + 4 * SEC_CL - 1) / (4 * SEC_CL)) * 4; }
inline size_t sec_base_floats(int C, int csq, int cb) {
  return (size_t)cb * sec_slice(C) + (size_t)(SEC_THREADS / 32) * cb * csq + 2 * (size_t)cb * csq;
}
// the CTA's fc2 weight slice [csq][KS] is prefetched into shared memory by bulk copies when it fits (<= 128 KB)
inline bool sec_stage_w2(int C, int csq) { return (size_t)csq * sec_slice(C) * sizeof(float) <= 128 * 1024; }
inline size_t sec_smem_bytes(int C, int csq, int cb, bool stage_w2) {
  return (sec_base_floats(C, csq, cb) + (stage_w2 ? (size_t)csq * sec_slice(C) : 0)) * sizeof(float) + 16;
}
inline bool sec_eligible(int C, int csq) { return C % 4 == 0 && csq % 4 == 0 && csq <= 32 * SEC_MAX_JPL && C <= SEC_MAX_C; }

template <int SEC_CB>
__global__ void __cluster_dims__(SEC_CL, 1, 1) __launch_bounds__(SEC_THREADS)
se_cluster_kernel(const float* __restrict__ pooled, int slices, size_t slice_stride, const float* __restrict__ w1,
                  const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                  float* __restrict__ scale, int B, int C, int csq, int act1, int act2, int stage_w2) {
  pdl_trigger();
  pdl_wait();
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b0 = (int)(blockIdx.x / SEC_CL) * SEC_CB;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NW = SEC_THREADS / 32;
  extern __shared__ float sec_smem[];
  const int KS = ((C + 4 * SEC_CL - 1) / (4 * SEC_CL)) * 4;  // K slice of fc1 == N slice of fc2 (multiple of 4)
  float* xs = sec_smem;                        // [CB][KS]   squeezed input, this CTA's K slice
  float* hp = xs + SEC_CB * KS;                // [NW][CB][csq] per-warp fc1 partials
  float* hcta = hp + NW * SEC_CB * csq;        // [CB][csq]  this CTA's fc1 partial (read by the whole cluster)
  float* hid = hcta + SEC_CB * csq;            // [CB][csq]  hidden activations (complete)
  const int k0 = rank * KS, k1 = min(k0 + KS, C);
  // fc2 weight slice w2[:, k0:k1] -> shared memory, one bulk copy per hidden unit, in flight during the squeeze and fc1
  // (the fc2 loop was a chain of dependent L2 round trips: 96-160 rows x ~800 cycles with four loads in flight)
  float* w2s = hid + SEC_CB * csq;             // [csq][KS]
  __shared__ __align__(8) unsigned long long w2_bar;
  const int nw2 = k1 > k0 ? k1 - k0 : 0;       // floats per row of the slice (multiple of 4)
  if (stage_w2 && nw2 > 0) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&w2_bar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&w2_bar)),
                   "r"((unsigned)(csq * nw2 * 4))
                   : "memory");
    }
    __syncthreads();
    for (int j = tid; j < csq; j += SEC_THREADS)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       (unsigned)__cvta_generic_to_shared(w2s + (size_t)j * KS)),
                   "l"(w2 + (size_t)j * C + k0), "r"((unsigned)(nw2 * 4)), "r"((unsigned)__cvta_generic_to_shared(&w2_bar))
                   : "memory");
  }

  // squeezed input: sum of the depthwise kernel's partial-mean slices
  for (int i = tid; i < SEC_CB * KS; i += SEC_THREADS) {
    const int cb = i / KS, k = k0 + (i - cb * KS);
    float v = 0.f;
    if (k < k1 && b0 + cb < B) {
      const float* src = pooled + (size_t)(b0 + cb) * C + k;
      for (int s = 0; s < slices; ++s) v += __ldg(src + (size_t)s * slice_stride);
    }
    xs[i] = v;
  }
  __syncthreads();
  // fc1 over this CTA's K slice: warp w takes rows k0 + w, k0 + w + NW, ... (csq contiguous floats: coalesced)
  {
    float acc[SEC_CB][SEC_MAX_JPL];
#pragma unroll
    for (int cb = 0; cb < SEC_CB; ++cb)
#pragma unroll
      for (int i = 0; i < SEC_MAX_JPL; ++i) acc[cb][i] = 0.f;
    // Every cluster streams the SAME weight rows; started together they all hit the same L2 lines at the same time and the
    // slices holding them serialise the requests (measured in round 1: 37 us per launch, no faster than three launches).
    // Each cluster therefore starts its row loop at its own rotation (fixed per cluster: still deterministic).
    const int nrows = k1 > k0 + warp ? (k1 - k0 - warp + NW - 1) / NW : 0;
    int ri = nrows > 0 ? (int)((blockIdx.x / SEC_CL) * 5u % (unsigned)nrows) : 0;
#pragma unroll 8
    for (int n = 0; n < nrows; ++n) {
      const int k = k0 + warp + ri * NW;
      if (++ri == nrows) ri = 0;
      const float* wr = w1 + (size_t)k * csq;
      float wv[SEC_MAX_JPL];
#pragma unroll
      for (int i = 0; i < SEC_MAX_JPL; ++i) wv[i] = (lane + 32 * i < csq) ? __ldg(wr + lane + 32 * i) : 0.f;
#pragma unroll
      for (int cb = 0; cb < SEC_CB; ++cb) {
        const float xv = xs[cb * KS + (k - k0)];
#pragma unroll
        for (int i = 0; i < SEC_MAX_JPL; ++i) acc[cb][i] = fmaf(xv, wv[i], acc[cb][i]);
      }
    }
#pragma unroll
    for (int cb = 0; cb < SEC_CB; ++cb)
#pragma unroll
      for (int i = 0; i < SEC_MAX_JPL; ++i)
        if (lane + 32 * i < csq) hp[(warp * SEC_CB + cb) * csq + lane + 32 * i] = acc[cb][i];
  }
  __syncthreads();
  for (int i = tid; i < SEC_CB * csq; i += SEC_THREADS) {
    const int cb = i / csq, j = i - cb * csq;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) v += hp[(w * SEC_CB + cb) * csq + j];
    hcta[i] = v;
  }
  cluster.sync();
  // hidden = act1(b1 + sum over the cluster's K slices), ranks in a fixed order
  for (int i = tid; i < SEC_CB * csq; i += SEC_THREADS) {
    const int j = i % csq;
    float v = b1[j];
#pragma unroll
    for (int r = 0; r < SEC_CL; ++r) v += cluster.map_shared_rank(hcta, r)[i];
    hid[i] = apply_act(v, act1);
  }
  cluster.sync();  // every remote read of hcta is done (a CTA may exit) and hid is visible to this CTA's threads
  if (stage_w2 && nw2 > 0) {  // the weight slice has landed
    unsigned ok = 0;
    while (!ok)
      asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0;\nselp.u32 %0, 1, 0, P1;\n}\n"
                   : "=r"(ok)
                   : "r"((unsigned)__cvta_generic_to_shared(&w2_bar))
                   : "memory");
  }
  // fc2 over this CTA's N slice: thread = (channel quad, half of the crops); rows of w2 are C contiguous floats
  const int n1 = min(k0 + KS, C);
  const int nq = n1 > k0 ? (n1 - k0) >> 2 : 0;
  const int cbase = (tid >> 7) * (SEC_CB / 2);
  for (int q = tid & 127; q < nq; q += 128) {
    const int n = k0 + q * 4;
    const float4 bv = *reinterpret_cast<const float4*>(b2 + n);
    float4 o[SEC_CB / 2];
#pragma unroll
    for (int c = 0; c < SEC_CB / 2; ++c) o[c] = bv;
    int j = (int)((blockIdx.x / SEC_CL) * 7u % (unsigned)csq);  // per-cluster rotation of the row order (see fc1)
#pragma unroll 8
    for (int jj = 0; jj < csq; ++jj) {
      const float4 wv = stage_w2 ? *reinterpret_cast<const float4*>(w2s + (size_t)j * KS + q * 4)
                                 : __ldg(reinterpret_cast<const float4*>(w2 + (size_t)j * C + n));
#pragma unroll
      for (int c = 0; c < SEC_CB / 2; ++c) {
        const float hv = hid[(cbase + c) * csq + j];
        o[c].x = fmaf(hv, wv.x, o[c].x); o[c].y = fmaf(hv, wv.y, o[c].y);
        o[c].z = fmaf(hv, wv.z, o[c].z); o[c].w = fmaf(hv, wv.w, o[c].w);
      }
      if (++j == csq) j = 0;
    }
#pragma unroll
    for (int c = 0; c < SEC_CB / 2; ++c) {
      const int b = b0 + cbase + c;
      if (b >= B) continue;
      *reinterpret_cast<float4*>(scale + (size_t)b * C + n) =
          make_float4(apply_act(o[c].x, act2), apply_act(o[c].y, act2), apply_act(o[c].z, act2), apply_act(o[c].w, act2));
    }
  }
}

}  // namespace mtb
