// SIMT emulator -- TEST INFRASTRUCTURE ONLY (never built into or loaded by the product).
//
// Lets the *same* kernel sources under silent-speech-decoder_b200/csrc/ be compiled by g++
// (-DSSD_EMU) and executed on the CPU, one CTA at a time, each CUDA thread as a ucontext
// fiber.  It exists because the build container has no GPU: kernel logic (barrier
// placement, warp collectives, mbarrier pipelines, index arithmetic) is debugged here under
// the parity tests before a GPU minute is spent.  The product library libssd_ctc.so is built
// by nvcc only; nothing in the package imports the emulated build.
//
// Fidelity: threads of a CTA run cooperatively and switch only at synchronising calls
// (__syncthreads, named barriers, warp collectives, mbarrier waits).  The visiting order is
// configurable (SSD_EMU_ORDER=fwd|rev|rnd:<seed>) so that a missing barrier shows up as
// order-dependent output.  Bulk async copies are deferred until the scheduler's next round
// so that a missing mbarrier wait reads stale data.
#pragma once
#ifndef SSD_EMU
#error "ssd_emu.h is only for -DSSD_EMU builds"
#endif

#include <ucontext.h>
#include <sys/mman.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <unordered_map>
#include <vector>

namespace emu {

struct dim3e {
  unsigned x = 1, y = 1, z = 1;
  dim3e() {}
  dim3e(unsigned x_, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  dim3e tidx;
  int linear = 0;
  bool done = false;
  std::function<bool()> ready;  // empty => runnable
  const char* wait_what = "";
  const void* wait_obj = nullptr;
};

struct NamedBar {
  int arrived = 0;
  uint64_t gen = 0;
};

struct WarpSlot {
  uint64_t vals[2][32];
  unsigned arrived = 0;
  uint64_t gen = 0;
};

struct MBar {
  int count = 0;        // arrivals per phase
  int pending = 0;      // arrivals still missing in the current phase
  long long tx = 0;     // bytes still expected in the current phase
  unsigned phase = 0;   // parity of the phase in progress
};

struct PendingCopy {
  void* dst;
  const void* src;
  size_t bytes;
  void* bar;
};

struct State {
  dim3e grid, block, bidx;
  std::vector<Fiber> fibers;
  Fiber* cur = nullptr;
  ucontext_t sched;
  std::vector<unsigned char> dyn_smem;
  NamedBar bars[16];
  std::vector<WarpSlot> warps;
  std::unordered_map<void*, MBar> mbars;
  std::vector<PendingCopy> copies;
  std::function<void()> body;
  uint64_t rng = 0x9E3779B97F4A7C15ull;
  int order_mode = 0;  // 0 fwd, 1 rev, 2 rnd
  size_t stack_bytes = 256 * 1024;
  std::vector<char*> stack_pool;
};

inline State& S() {
  static State s;
  return s;
}

inline void die(const char* msg) {
  std::fprintf(stderr, "[ssd_emu] fatal: %s\n", msg);
  std::abort();
}

inline void fiber_entry() {
  State& s = S();
  s.body();
  s.cur->done = true;
  swapcontext(&s.cur->ctx, &s.sched);
}

inline void yield_until(std::function<bool()> cond, const char* what = "", const void* obj = nullptr) {
  State& s = S();
  if (cond()) return;
  s.cur->ready = std::move(cond);
  s.cur->wait_what = what;
  s.cur->wait_obj = obj;
  swapcontext(&s.cur->ctx, &s.sched);
}

inline unsigned char* dyn_smem();
inline void flush_copies() {
  State& s = S();
  if (s.copies.empty()) return;
  std::vector<PendingCopy> todo;
  todo.swap(s.copies);
  for (auto& c : todo) {
    std::memcpy(c.dst, c.src, c.bytes);
    auto it = s.mbars.find(c.bar);
    if (it == s.mbars.end()) die("bulk copy completes on an uninitialised mbarrier");
    MBar& m = it->second;
    m.tx -= (long long)c.bytes;
    if (m.pending == 0 && m.tx == 0) {
      m.phase ^= 1u;
      m.pending = m.count;
    }
  }
}

inline void run_block() {
  State& s = S();
  const int n = (int)(s.block.x * s.block.y * s.block.z);
  s.fibers.assign(n, Fiber());
  while ((int)s.stack_pool.size() < n) {
    void* p = mmap(nullptr, s.stack_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) die("mmap stack");
    s.stack_pool.push_back((char*)p);
  }
  for (auto& b : s.bars) b = NamedBar();
  s.warps.assign((n + 31) / 32, WarpSlot());
  s.mbars.clear();
  s.copies.clear();
  for (int i = 0; i < n; ++i) {
    Fiber& f = s.fibers[i];
    f.linear = i;
    f.tidx.x = i % s.block.x;
    f.tidx.y = (i / s.block.x) % s.block.y;
    f.tidx.z = i / (s.block.x * s.block.y);
    f.stack = s.stack_pool[i];
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = s.stack_bytes;
    f.ctx.uc_link = &s.sched;
    makecontext(&f.ctx, (void (*)())fiber_entry, 0);
  }
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  int remaining = n;
  while (remaining > 0) {
    if (s.order_mode == 1) {
      for (int i = 0; i < n; ++i) order[i] = n - 1 - i;
    } else if (s.order_mode == 2) {
      for (int i = n - 1; i > 0; --i) {
        s.rng ^= s.rng << 13; s.rng ^= s.rng >> 7; s.rng ^= s.rng << 17;
        std::swap(order[i], order[(int)(s.rng % (uint64_t)(i + 1))]);
      }
    }
    bool progress = false;
    for (int k = 0; k < n; ++k) {
      Fiber& f = s.fibers[order[k]];
      if (f.done) continue;
      if (f.ready && !f.ready()) continue;
      f.ready = nullptr;
      s.cur = &f;
      swapcontext(&s.sched, &f.ctx);
      s.cur = nullptr;
      progress = true;
      if (f.done) --remaining;
    }
    if (!s.copies.empty()) {
      flush_copies();
      progress = true;
    }
    if (!progress) {
      std::fprintf(stderr, "[ssd_emu] deadlock in block (%u,%u,%u): %d threads blocked\n", s.bidx.x, s.bidx.y, s.bidx.z, remaining);
      for (int w0 = 0; w0 < n; w0 += 32) {
        std::fprintf(stderr, "  warp %d:", w0 / 32);
        for (int i = w0; i < w0 + 32 && i < n; ++i) {
          if (s.fibers[i].done) { std::fprintf(stderr, " done"); continue; }
          long off = (long)((const unsigned char*)s.fibers[i].wait_obj - dyn_smem());
          if (i == w0 || s.fibers[i].wait_obj != s.fibers[i - 1].wait_obj || s.fibers[i].wait_what != s.fibers[i - 1].wait_what)
            std::fprintf(stderr, " [lane %d..] %s @%ld", i - w0, s.fibers[i].wait_what, off);
        }
        std::fprintf(stderr, "\n");
      }
      std::abort();
    }
  }
}

inline void launch(dim3e grid, dim3e block, size_t smem, std::function<void()> body) {
  State& s = S();
  const char* om = std::getenv("SSD_EMU_ORDER");
  s.order_mode = 0;
  if (om) {
    if (!std::strncmp(om, "rev", 3)) s.order_mode = 1;
    else if (!std::strncmp(om, "rnd", 3)) {
      s.order_mode = 2;
      if (om[3] == ':') s.rng = 0x9E3779B97F4A7C15ull ^ (uint64_t)std::strtoull(om + 4, nullptr, 10);
    }
  }
  s.grid = grid;
  s.block = block;
  s.body = std::move(body);
  s.dyn_smem.assign(smem + 1024, 0xCD);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        s.bidx = dim3e(bx, by, bz);
        std::fill(s.dyn_smem.begin(), s.dyn_smem.end(), (unsigned char)0xCD);
        run_block();
      }
}

inline size_t dyn_smem_size() { return S().dyn_smem.size() - 1024; }
inline unsigned char* dyn_smem() {
  State& s = S();
  uintptr_t p = (uintptr_t)s.dyn_smem.data();
  p = (p + 1023) & ~(uintptr_t)1023;
  return (unsigned char*)p;
}

// ------------------------------------------------------------------ barriers
inline void bar_sync(int id, int count) {
  State& s = S();
  NamedBar& b = s.bars[id];
  if (++b.arrived == count) {
    b.arrived = 0;
    ++b.gen;
    return;
  }
  const uint64_t g = b.gen;
  NamedBar* bp = &b;
  yield_until([bp, g] { return bp->gen != g; }, "bar.sync", bp);
}

// bar.arrive: counts like bar.sync but does not wait
inline void bar_arrive(int id, int count) {
  NamedBar& b = S().bars[id];
  if (++b.arrived == count) {
    b.arrived = 0;
    ++b.gen;
  }
}

inline int lane_id() { return S().cur->linear & 31; }
inline int warp_id() { return S().cur->linear >> 5; }

inline unsigned full_mask_of_warp() {
  State& s = S();
  const int n = (int)(s.block.x * s.block.y * s.block.z);
  const int w = warp_id();
  const int cnt = std::min(32, n - w * 32);
  return cnt == 32 ? 0xffffffffu : ((1u << cnt) - 1u);
}

// All lanes named in mask exchange one 64-bit value; returns pointer to the 32-slot array
// of this collective (valid until the *next but one* collective of the warp).
inline const uint64_t* warp_exchange(unsigned mask, uint64_t mine) {
  State& s = S();
  WarpSlot& w = s.warps[warp_id()];
  const int lane = lane_id();
  mask &= full_mask_of_warp();
  if (!(mask >> lane & 1u)) die("warp collective called by a lane outside its mask");
  const uint64_t g = w.gen;
  uint64_t* buf = w.vals[g & 1];
  buf[lane] = mine;
  w.arrived |= 1u << lane;
  if ((w.arrived & mask) == mask) {
    w.arrived &= ~mask;
    ++w.gen;
  } else {
    WarpSlot* wp = &w;
    yield_until([wp, g] { return wp->gen != g; }, "warp collective", wp);
  }
  return buf;
}

// ------------------------------------------------------------------ mbarrier / bulk copy
inline MBar& mbar_get(void* p) {
  auto it = S().mbars.find(p);
  if (it == S().mbars.end()) die("mbarrier used before init");
  return it->second;
}
inline void mbar_init(void* p, int count) {
  MBar m;
  m.count = count;
  m.pending = count;
  S().mbars[p] = m;
}
inline void mbar_settle(MBar& m) {
  if (m.pending == 0 && m.tx == 0) {
    m.phase ^= 1u;
    m.pending = m.count;
  }
}
inline void mbar_arrive(void* p) {
  MBar& m = mbar_get(p);
  if (m.pending <= 0) die("mbarrier over-arrival");
  --m.pending;
  mbar_settle(m);
}
inline void mbar_arrive_expect_tx(void* p, unsigned bytes) {
  MBar& m = mbar_get(p);
  if (m.pending <= 0) die("mbarrier over-arrival");
  m.tx += bytes;
  --m.pending;
  mbar_settle(m);
}
inline void mbar_expect_tx(void* p, unsigned bytes) { mbar_get(p).tx += bytes; }
inline bool mbar_test(void* p, unsigned parity) { return mbar_get(p).phase != (parity & 1u); }
inline void mbar_wait(void* p, unsigned parity) {
  MBar* m = &mbar_get(p);
  parity &= 1u;
  yield_until([m, parity] { return m->phase != parity; }, parity ? "mbarrier parity 1" : "mbarrier parity 0", p);
}
inline void bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
  if (((uintptr_t)dst & 15) || ((uintptr_t)src & 15) || (bytes & 15)) die("bulk copy needs 16-byte aligned src/dst/size");
  S().copies.push_back(PendingCopy{dst, src, bytes, bar});
}

}  // namespace emu

// ====================================================================== CUDA surface
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static
#define __restrict__ __restrict
#define __align__(n) __attribute__((aligned(n)))

using dim3 = emu::dim3e;
using std::max;
using std::min;
struct float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct uint2 { unsigned x, y; };
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
struct float2 { float x, y; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }
#define threadIdx (emu::S().cur->tidx)
#define blockIdx (emu::S().bidx)
#define blockDim (emu::S().block)
#define gridDim (emu::S().grid)
#define warpSize 32

typedef void* cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaPeekAtLastError() { return 0; }

inline void __syncthreads() {
  emu::State& s = emu::S();
  emu::bar_sync(0, (int)(s.block.x * s.block.y * s.block.z));
}
inline void __syncwarp(unsigned mask = 0xffffffffu) { (void)emu::warp_exchange(mask, 0); }
inline void __threadfence() {}
inline void __threadfence_block() {}

template <typename T>
inline T emu_from_bits(uint64_t b) {
  T v;
  std::memcpy(&v, &b, sizeof(T));
  return v;
}
template <typename T>
inline uint64_t emu_to_bits(T v) {
  static_assert(sizeof(T) <= 8, "shuffle payload too wide");
  uint64_t b = 0;
  std::memcpy(&b, &v, sizeof(T));
  return b;
}
template <typename T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  const int lane = emu::lane_id();
  const uint64_t* buf = emu::warp_exchange(mask, emu_to_bits(v));
  const int base = lane & ~(width - 1);
  const int from = base + (src & (width - 1));
  return emu_from_bits<T>(buf[from]);
}
template <typename T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const int lane = emu::lane_id();
  const uint64_t* buf = emu::warp_exchange(mask, emu_to_bits(v));
  const int base = lane & ~(width - 1);
  const int from = lane - (int)delta;
  return from < base ? v : emu_from_bits<T>(buf[from]);
}
template <typename T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const int lane = emu::lane_id();
  const uint64_t* buf = emu::warp_exchange(mask, emu_to_bits(v));
  const int base = lane & ~(width - 1);
  const int from = lane + (int)delta;
  return from >= base + width ? v : emu_from_bits<T>(buf[from]);
}
template <typename T>
inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
  const int lane = emu::lane_id();
  const uint64_t* buf = emu::warp_exchange(mask, emu_to_bits(v));
  const int from = lane ^ lanemask;
  (void)width;
  return emu_from_bits<T>(buf[from]);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
  const uint64_t* buf = emu::warp_exchange(mask, pred ? 1u : 0u);
  unsigned m = mask & emu::full_mask_of_warp(), r = 0;
  for (int i = 0; i < 32; ++i)
    if ((m >> i & 1u) && buf[i]) r |= 1u << i;
  return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == (mask & emu::full_mask_of_warp()); }
template <typename T, typename F>
inline T emu_reduce(unsigned mask, T v, F f) {
  const uint64_t* buf = emu::warp_exchange(mask, emu_to_bits(v));
  unsigned m = mask & emu::full_mask_of_warp();
  bool first = true;
  T acc = v;
  for (int i = 0; i < 32; ++i)
    if (m >> i & 1u) {
      T x = emu_from_bits<T>(buf[i]);
      acc = first ? x : f(acc, x);
      first = false;
    }
  return acc;
}
inline int __reduce_max_sync(unsigned mask, int v) { return emu_reduce(mask, v, [](int a, int b) { return a > b ? a : b; }); }
inline int __reduce_min_sync(unsigned mask, int v) { return emu_reduce(mask, v, [](int a, int b) { return a < b ? a : b; }); }
inline int __reduce_add_sync(unsigned mask, int v) { return emu_reduce(mask, v, [](int a, int b) { return a + b; }); }
inline unsigned __reduce_max_sync(unsigned mask, unsigned v) { return emu_reduce(mask, v, [](unsigned a, unsigned b) { return a > b ? a : b; }); }
inline unsigned __reduce_min_sync(unsigned mask, unsigned v) { return emu_reduce(mask, v, [](unsigned a, unsigned b) { return a < b ? a : b; }); }
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) { return emu_reduce(mask, v, [](unsigned a, unsigned b) { return a + b; }); }
inline unsigned __reduce_or_sync(unsigned mask, unsigned v) { return emu_reduce(mask, v, [](unsigned a, unsigned b) { return a | b; }); }

inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __double2hiint(double d) { return (int)(emu_to_bits(d) >> 32); }
inline int __double2loint(double d) { return (int)(emu_to_bits(d) & 0xffffffffu); }
inline double __hiloint2double(int hi, int lo) { return emu_from_bits<double>(((uint64_t)(unsigned)hi << 32) | (unsigned)lo); }
inline long long __double_as_longlong(double d) { return (long long)emu_to_bits(d); }
inline double __longlong_as_double(long long v) { return emu_from_bits<double>((uint64_t)v); }
inline int __float_as_int(float f) { return (int)emu_to_bits(f); }
inline float __int_as_float(int v) { return emu_from_bits<float>((uint64_t)(unsigned)v); }
inline unsigned __float_as_uint(float f) { return (unsigned)emu_to_bits(f); }
inline float __uint_as_float(unsigned v) { return emu_from_bits<float>((uint64_t)v); }
template <typename T>
inline T __ldg(const T* p) { return *p; }
inline int __float2int_rn(float f) { return (int)std::lrintf(f); }
inline float __expf(float x) { return std::exp(x); }
inline float __logf(float x) { return std::log(x); }

template <typename T>
inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T>
inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T>
inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T>
inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <typename T>
inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
template <typename T>
inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
