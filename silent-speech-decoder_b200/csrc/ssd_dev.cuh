// Device-side primitives used by the CTC kernels (sm_100a).
//
// Everything Blackwell/Hopper specific that has no CUDA C++ spelling is wrapped here as a
// small inline-PTX function: mbarrier (init / arrive / expect_tx / parity wait), 1-D TMA
// bulk copies global->shared (cp.async.bulk, SASS UBLKCP), named barriers, cache-hinted
// global accesses.  When compiled with -DSSD_EMU (tests only, see tests/emu/ssd_emu.h)
// the same names resolve to a CPU emulation so kernel logic can be exercised without a GPU.
#pragma once

#include <cstdint>
#include <cstddef>

#ifdef SSD_EMU
#include "ssd_emu.h"
#define SSD_DYN_SMEM(name) unsigned char* name = emu::dyn_smem()
#define SSD_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu::launch(emu::dim3e(grid), emu::dim3e(block), (size_t)(smem), [=]() { kernel(__VA_ARGS__); })
#define SSD_ASSERT(c)                                                                  \
  do {                                                                                 \
    if (!(c)) {                                                                        \
      std::fprintf(stderr, "SSD_ASSERT failed: %s  (%s:%d)\n", #c, __FILE__, __LINE__); \
      std::abort();                                                                    \
    }                                                                                  \
  } while (0)
#else
#include <cuda_runtime.h>
#define SSD_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#define SSD_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define SSD_ASSERT(c) ((void)0)
#endif

namespace ssd {

constexpr unsigned FULL = 0xffffffffu;

// ------------------------------------------------------------------ named barriers
// bar.sync id, n : only the n threads that name barrier `id` take part (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
#ifdef SSD_EMU
  emu::bar_sync(id, nthreads);
#else
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
#endif
}

// bar.arrive id, n : count towards barrier `id` without waiting (the waiters use bar_sync with the same n)
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
#ifdef SSD_EMU
  emu::bar_arrive(id, nthreads);
#else
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
#endif
}

// ------------------------------------------------------------------ mbarrier
#ifndef SSD_EMU
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#endif

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
#ifdef SSD_EMU
  emu::mbar_init(bar, count);
#else
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
#endif
}
// make mbarrier inits visible to the async (TMA) proxy
__device__ __forceinline__ void mbar_fence_init() {
#ifndef SSD_EMU
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
#ifdef SSD_EMU
  emu::mbar_arrive(bar);
#else
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
#endif
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
#ifdef SSD_EMU
  emu::mbar_arrive_expect_tx(bar, bytes);
#else
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
#endif
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
#ifdef SSD_EMU
  return emu::mbar_test(bar, parity);
#else
  uint32_t ok;
  asm volatile(
      // suspend-time hint (ns): the thread sleeps in hardware until the phase completes or the
      // hint expires, instead of coming back to the issue slot every few hundred cycles
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
#endif
}
// Blocks until the phase with the given parity has completed (try_wait suspends in hardware).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#ifdef SSD_EMU
  emu::mbar_wait(bar, parity);
#else
  while (!mbar_test(bar, parity)) {
  }
#endif
}

// For waits that are expected to be long (a producer far ahead of its consumer, a worker waiting for
// its first block): back off between polls so that the issue slots go to the warps with work.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, unsigned ns = 500u) {
#ifdef SSD_EMU
  (void)ns;
  emu::mbar_wait(bar, parity);
#else
  while (!mbar_test(bar, parity)) __nanosleep(ns);
#endif
}

// ------------------------------------------------------------------ TMA 1-D bulk copy
// global -> shared::cta, completion counted in bytes on `bar`.  dst, src and bytes must be
// multiples of 16.  One thread issues; the copy engine moves the data.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
#ifdef SSD_EMU
  emu::bulk_g2s(dst_smem, src_gmem, bytes, bar);
#else
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
#endif
}
// The same copy with an L2 eviction priority for the lines it touches (cp.async.bulk ... .L2::cache_hint):
// keep = 1: evict_last, the data will be read again by this kernel (the loss reads every log-prob three times: both
// sweeps and the block workers); keep = 0: evict_first, this was the last reader.
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, int keep) {
#ifdef SSD_EMU
  (void)keep;
  emu::bulk_g2s(dst_smem, src_gmem, bytes, bar);
#else
  uint64_t pol;
  if (keep) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
#endif
}
// Hint: pull `bytes` (multiple of 16) at `src_gmem` into L2 ahead of the bulk copy that will fetch them
// (cp.async.bulk.prefetch.L2, SASS UBLKPF).  No shared memory, no completion to wait for.
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
#ifdef SSD_EMU
  (void)src_gmem;
  (void)bytes;
#else
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
#endif
}
// generic-proxy writes to smem that the async proxy will overwrite later must be ordered first
__device__ __forceinline__ void fence_proxy_async() {
#ifndef SSD_EMU
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}

// ------------------------------------------------------------------ cache-hinted global access
// streaming store: written once, not re-read by this kernel
__device__ __forceinline__ void st_stream(float* p, float v) {
#ifdef SSD_EMU
  *p = v;
#else
#if defined(SSD_ST_ASM) && SSD_ST_ASM
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
#else
  __stcs(p, v);  // st.global.cs without a compiler barrier: the surrounding shared-memory loads may move across it
#endif
#endif
}
__device__ __forceinline__ void st_stream4(float4* p, float4 v) {
#ifdef SSD_EMU
  *p = v;
#else
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
#endif
}
__device__ __forceinline__ float ld_stream(const float* p) {
#ifdef SSD_EMU
  return *p;
#else
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
#endif
}

__device__ __forceinline__ int lane_id() {
#ifdef SSD_EMU
  return emu::lane_id();
#else
  return (int)(threadIdx.x & 31u);
#endif
}

}  // namespace ssd

