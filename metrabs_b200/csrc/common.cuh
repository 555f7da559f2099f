// Shared helpers for the metrabs_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <utility>

namespace mtb {

enum Act : int { ACT_NONE = 0, ACT_SILU = 1, ACT_RELU = 2, ACT_HSWISH = 3, ACT_SIGMOID = 4, ACT_HSIGMOID = 5 };

// Programmatic dependent launch (PDL): every kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization.
// pdl_trigger() lets the NEXT kernel in the stream start launching once all CTAs of this grid have started (its prologue
// then overlaps this grid's tail); pdl_wait() blocks until the PREVIOUS grid has completed and its writes are visible - it
// must precede the first global-memory access of a kernel.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Measured on B200 (EfficientNetV2-L, 256 crops/step): 8.41 k crops/s without the attribute, 8.01 k with it (dependents
// that launch early hold SM slots while spinning in griddepcontrol.wait), so PDL is OFF unless MTB_ENABLE_PDL=1; without
// the launch attribute griddepcontrol.* are no-ops.
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MTB_ENABLE_PDL");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// Scoped PDL for a chain of tiny dependent launches (the squeeze-excitation FCs): their launch latency, not their work, is
// what the step pays for, and a small early-launched grid does not hold the SM slots a 210 KB tensor-core CTA would.
inline int& pdl_force_depth() {
  static thread_local int d = 0;
  return d;
}
struct PdlScope {
  bool on;
  explicit PdlScope(bool enable) : on(enable) { if (on) ++pdl_force_depth(); }
  ~PdlScope() { if (on) --pdl_force_depth(); }
};

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_enabled() || pdl_force_depth() > 0) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case ACT_SILU: return x * sigmoidf_(x);
    case ACT_RELU: return fmaxf(x, 0.0f);
    case ACT_HSWISH: return x * fminf(fmaxf(x + 3.0f, 0.0f), 6.0f) * (1.0f / 6.0f);
    case ACT_SIGMOID: return sigmoidf_(x);
    case ACT_HSIGMOID: return fminf(fmaxf(x + 3.0f, 0.0f), 6.0f) * (1.0f / 6.0f);
    default: return x;
  }
}

// compile-time activation (exact expf forms: used by the fp32 parity kernels).  A runtime switch inside per-element code
// gets if-converted into every branch, so kernels dispatch ONCE per thread with act_dispatch and run a templated body.
template <int ACT>
__device__ __forceinline__ float act_t(float x) {
  if constexpr (ACT == ACT_SILU) return x * sigmoidf_(x);
  else if constexpr (ACT == ACT_RELU) return fmaxf(x, 0.0f);
  else if constexpr (ACT == ACT_HSWISH) return x * fminf(fmaxf(x + 3.0f, 0.0f), 6.0f) * (1.0f / 6.0f);
  else if constexpr (ACT == ACT_SIGMOID) return sigmoidf_(x);
  else if constexpr (ACT == ACT_HSIGMOID) return fminf(fmaxf(x + 3.0f, 0.0f), 6.0f) * (1.0f / 6.0f);
  else return x;
}
template <int V>
struct IntTag {
  static constexpr int value = V;
};
template <typename F>
__device__ __forceinline__ void act_dispatch(int act, F&& f) {
  switch (act) {
    case ACT_SILU: f(IntTag<ACT_SILU>{}); break;
    case ACT_RELU: f(IntTag<ACT_RELU>{}); break;
    case ACT_HSWISH: f(IntTag<ACT_HSWISH>{}); break;
    case ACT_SIGMOID: f(IntTag<ACT_SIGMOID>{}); break;
    case ACT_HSIGMOID: f(IntTag<ACT_HSIGMOID>{}); break;
    default: f(IntTag<ACT_NONE>{}); break;
  }
}

// ---- 4-wide vector load/store for fp32 and bf16 activation storage -------------------------------------
template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T>
__device__ __forceinline__ void store4(T* p, float4 v);
template <>
__device__ __forceinline__ void store4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
template <typename T>
__device__ __forceinline__ float load1(const T* p);
template <>
__device__ __forceinline__ float load1<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float load1<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <>
__device__ __forceinline__ float load1<__half>(const __half* p) { return __half2float(*p); }
template <typename T>
__device__ __forceinline__ void store1(T* p, float v);
template <>
__device__ __forceinline__ void store1<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store1<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// Packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2): two IEEE fp32 operations per issued instruction, bit-identical to the
// scalar fmaf / fmul / fadd.  The kernel is issue-bound (ncu: 56 % issue-active with 2 warps per scheduler), so halving
// the FMA instruction count is worth more than any memory-side change.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace mtb
