// Microbenchmark: tcgen05.mma (kind::f16, M=128, K=16) issue-to-completion rate of ONE SM as a function of the shared-memory
// operand layout (128B / 64B swizzle, no-swizzle core matrices with aligned and unaligned strides) and of N.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/bin/mma_rate scripts/mma_rate.cu
// Values in shared memory are zeros: only the timing matters.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <string>
#include <cuda_runtime.h>

struct Cfg {
  uint32_t a_lo, a_hi, a_step, a_wrap;  // descriptor low word (start address>>4 | LBO<<16), high word, per-MMA step (16 B units), wrap
  uint32_t b_lo, b_hi, b_step, b_wrap;
  uint32_t idesc, count;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(416, 1) mma_rate_kernel(const Cfg* cfgs, int ncfg, long long* out, int nspin, int spin_mode) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(smem + 192 * 1024);
  uint32_t* slot = (uint32_t*)(bar + 2);
  for (int i = threadIdx.x; i < 192 * 1024 / 16; i += blockDim.x) ((uint4*)smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 1)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t base = smem_u32(smem) >> 4;
    uint32_t parity = 0;
    for (int c = 0; c < ncfg; ++c) {
      const Cfg g = cfgs[c];
      for (int rep = 0; rep < 3; ++rep) {
        const long long t0 = clock64();
        uint32_t ai = 0, bi = 0;
        for (uint32_t i = 0; i < g.count; ++i) {
          const uint32_t alo = ((g.a_lo & 0x3FFF) + base + ai * g.a_step) & 0x3FFF;
          const uint32_t blo = ((g.b_lo & 0x3FFF) + base + bi * g.b_step) & 0x3FFF;
          uint64_t ad, bd;
          asm("mov.b64 %0, {%1, %2};" : "=l"(ad) : "r"(alo | (g.a_lo & 0xFFFF0000u)), "r"(g.a_hi));
          asm("mov.b64 %0, {%1, %2};" : "=l"(bd) : "r"(blo | (g.b_lo & 0xFFFF0000u)), "r"(g.b_hi));
          asm volatile(
              "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
              "l"(ad), "l"(bd), "r"(g.idesc), "r"(i)
              : "memory");
          if (++ai == g.a_wrap) ai = 0;
          if (++bi == g.b_wrap) bi = 0;
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) {
          asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
                       : "=r"(ok)
                       : "r"(smem_u32(bar)), "r"(parity)
                       : "memory");
        }
        parity ^= 1;
        const long long t2 = clock64();
        out[(c * 3 + rep) * 2 + 0] = t1 - t0;
        out[(c * 3 + rep) * 2 + 1] = t2 - t0;
      }
    }
  }
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar + 1)) : "memory");
  const int warp = threadIdx.x >> 5;
  if (warp >= 1 && warp <= nspin) {
    // spinning waiters, as the epilogue / producer warps of the conv kernel do while the MMA thread works
    uint32_t ok = 0;
    while (!ok) {
      if (spin_mode == 2) {
        asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\nselp.u32 %0, 1, 0, P1;\n}\n"
                     : "=r"(ok)
                     : "r"(smem_u32(bar + 1)), "r"(0u), "r"(1000000u)
                     : "memory");
      } else {
        asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
                     : "=r"(ok)
                     : "r"(smem_u32(bar + 1)), "r"(0u)
                     : "memory");
        if (!ok && spin_mode == 1) __nanosleep(100);
        if (!ok && spin_mode == 3) { if (clock64() == 0) printf("x"); }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

static uint32_t idesc(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
static uint32_t hi_sw(int row_bytes) {  // SBO = 8 rows, version 1, layout by row width
  uint32_t layout = row_bytes == 128 ? 2u : row_bytes == 64 ? 4u : 6u;
  return (uint32_t)((8 * row_bytes) >> 4) | (1u << 14) | (layout << 29);
}
static uint32_t hi_nosw(int sbo_bytes) { return (uint32_t)(sbo_bytes >> 4) | (1u << 14); }

int main() {
  std::vector<Cfg> cfgs;
  std::vector<std::string> names;
  const uint32_t B_OFF = (128 * 1024) >> 4;  // B operands live in [128K, 192K)
  auto add = [&](const std::string& name, uint32_t a_lo, uint32_t a_hi, uint32_t a_step, uint32_t a_wrap, uint32_t b_lo, uint32_t b_hi,
                 uint32_t b_step, uint32_t b_wrap, int n) {
    for (uint32_t count : {64u, 256u}) {
      cfgs.push_back({a_lo, a_hi, a_step, a_wrap, b_lo + B_OFF, b_hi, b_step, b_wrap, idesc(n), count});
      names.push_back(name + " N=" + std::to_string(n) + " count=" + std::to_string(count));
    }
  };
  for (int n : {32, 256}) {
    // both operands 128B-swizzled, BK = 64: 4 MMAs per stage, +32 B each
    add("A sw128 / B sw128", 0, hi_sw(128), 2, 4, 0, hi_sw(128), 2, 4, n);
    // both 64B-swizzled, BK = 32
    add("A sw64  / B sw64 ", 0, hi_sw(64), 2, 2, 0, hi_sw(64), 2, 2, n);
    // both 32B-swizzled, BK = 16
    add("A sw32  / B sw32 ", 0, hi_sw(32), 2, 1, 0, hi_sw(32), 2, 1, n);
    // A no-swizzle, canonical dense: core matrices 128 B, SBO 128, LBO 2048 (16 row groups), K chunk pairs 4096 B apart
    add("A nosw sbo128 lbo2048 / B sw128", (2048u >> 4) << 16, hi_nosw(128), 4096 >> 4, 4, 0, hi_sw(128), 2, 4, n);
    // A no-swizzle, 8x16 patch rows of 10 pixels: SBO 160, LBO 2880 (the mode-2 layout), start at pixel 0
    add("A nosw sbo160 lbo2880 / B sw128", (2880u >> 4) << 16, hi_nosw(160), (2 * 2880) >> 4, 4, 0, hi_sw(128), 2, 4, n);
    // same, start shifted by one pixel (16 B): core matrices straddle 128 B lines
    add("A nosw sbo160 lbo2880 +16B / B sw128", ((2880u >> 4) << 16) | 1u, hi_nosw(160), (2 * 2880) >> 4, 4, 0, hi_sw(128), 2, 4, n);
    // no-swizzle with a 256 B row-group stride (aligned) and 16 B shifted
    add("A nosw sbo256 lbo8192 / B sw128", (8192u >> 4) << 16, hi_nosw(256), 16384 >> 4, 2, 0, hi_sw(128), 2, 4, n);
    add("A nosw sbo256 lbo8192 +16B / B sw128", ((8192u >> 4) << 16) | 1u, hi_nosw(256), 16384 >> 4, 2, 0, hi_sw(128), 2, 4, n);
    // swapped roles: A = weights (sw128), B = pixels from the no-swizzle patch
    add("A sw128 / B nosw sbo160 lbo2880 +16B", 0, hi_sw(128), 2, 4, ((2880u >> 4) << 16) | 1u, hi_nosw(160), (2 * 2880) >> 4, 4, n);
    add("A sw128 / B nosw sbo128 lbo4096", 0, hi_sw(128), 2, 4, (4096u >> 4) << 16, hi_nosw(128), 8192 >> 4, 2, n);
  }
  Cfg* d_cfg;
  long long* d_out;
  cudaMalloc(&d_cfg, cfgs.size() * sizeof(Cfg));
  cudaMalloc(&d_out, cfgs.size() * 6 * sizeof(long long));
  cudaMemcpy(d_cfg, cfgs.data(), cfgs.size() * sizeof(Cfg), cudaMemcpyHostToDevice);
  const int smem = 192 * 1024 + 1024 + 256;
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int spin_cfg[][2] = {{0, 0}, {4, 0}, {8, 0}, {12, 0}, {12, 3}, {12, 1}, {12, 2}};
  for (auto& sc : spin_cfg) {
  printf("---- %d spinning warps, spin mode %d (0 tight try_wait, 3 try_wait + clock64 check, 1 try_wait + nanosleep(100), 2 try_wait with suspend hint)\n", sc[0], sc[1]);
  mma_rate_kernel<<<1, 416, smem>>>(d_cfg, (int)cfgs.size(), d_out, sc[0], sc[1]);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("error: %s\n", cudaGetErrorString(e));
    return 1;
  }
  std::vector<long long> out(cfgs.size() * 6);
  cudaMemcpy(out.data(), d_out, out.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  printf("%-60s %12s %12s\n", "config (best of 3)", "issue cyc/MMA", "total cyc/MMA");
  for (size_t c = 0; c < cfgs.size(); c += 2) {
    long long best64 = 1LL << 60, best256 = 1LL << 60, iss256 = 1LL << 60;
    for (int r = 0; r < 3; ++r) {
      if (out[(c * 3 + r) * 2 + 1] < best64) best64 = out[(c * 3 + r) * 2 + 1];
      if (out[((c + 1) * 3 + r) * 2 + 1] < best256) best256 = out[((c + 1) * 3 + r) * 2 + 1];
      if (out[((c + 1) * 3 + r) * 2 + 0] < iss256) iss256 = out[((c + 1) * 3 + r) * 2 + 0];
    }
    std::string nm = names[c].substr(0, names[c].find(" count"));
    printf("%-60s issue %7.1f  total@256 %7.1f  slope(64->256) %7.1f cyc/MMA\n", nm.c_str(), iss256 / 256.0, best256 / 256.0,
           (best256 - best64) / 192.0);
  }
  }
  return 0;
}
