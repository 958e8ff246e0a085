// Minimal CPU emulation of the CUDA execution model for kernel logic checks without a GPU:
// the threads of a block are user-level fibers (a 10-instruction x86-64 stack switch) run round-robin by one OS thread,
// switching at __syncthreads / __syncwarp / warp collectives; blocks run one after another.
// Test infrastructure only (tests/emu/): it checks indices, barrier uniformity (a divergent
// barrier deadlocks and is reported) and data flow of a kernel, not memory-model subtleties.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(x) __attribute__((aligned(x)))
#define __shared__ static
#define __constant__ static const
#define __noinline__

struct emu_dim3 {
  unsigned x = 1, y = 1, z = 1;
};
struct uint2 {
  uint32_t x, y;
};
struct uint4 {
  uint32_t x, y, z, w;
};
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
struct alignas(16) float4 {
  float x, y, z, w;
};
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
template <typename T>
static inline T min(T a, T b) {
  return b < a ? b : a;
}
template <typename T>
static inline T max(T a, T b) {
  return a < b ? b : a;
}
struct alignas(16) double2 {
  double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

static emu_dim3 threadIdx, blockIdx, blockDim, gridDim;   // threadIdx: the running fiber's

struct EmuBarrier {
  unsigned expected = 0, arrived = 0;
  uint64_t gen = 0;
};
struct EmuWarp {
  EmuBarrier bar;
  uint32_t val[2][32];
  unsigned op = 0;   // collective operations completed (parity selects the slot array)
};
#if !defined(__x86_64__)
#error "tests/emu/cuda_emu.h: the fiber switch is written for x86-64"
#endif
// Saves the callee-saved registers on the current stack, stores its pointer, continues on the other stack.
extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch,.-emu_switch
)");

struct EmuFiber {
  void *sp = nullptr;
  std::vector<unsigned char> stack;
  bool done = false;
  EmuBarrier *wait = nullptr;
  uint64_t wait_gen = 0;
};
static void *emu_sched_sp = nullptr;
static std::vector<EmuFiber> emu_fibers;
static std::vector<EmuWarp> emu_warps;
static EmuBarrier emu_block_bar;
static unsigned emu_cur = 0;
static const std::function<void()> *emu_body = nullptr;
alignas(128) static unsigned char emu_dyn_smem[232448];

static inline void emu_wait(EmuBarrier &b) {
  EmuFiber &f = emu_fibers[emu_cur];
  const uint64_t g = b.gen;
  if (++b.arrived == b.expected) {
    b.arrived = 0;
    ++b.gen;
    return;   // the last arrival runs on
  }
  f.wait = &b;
  f.wait_gen = g;
  emu_switch(&f.sp, emu_sched_sp);
}
// A polling loop gives the other fibers a turn (the fiber stays runnable)
static uint64_t emu_yields = 0;
static inline void emu_yield() {
  EmuFiber &f = emu_fibers[emu_cur];
  f.wait = nullptr;
  ++emu_yields;
  emu_switch(&f.sp, emu_sched_sp);
}
static inline EmuWarp &emu_warp() { return emu_warps[threadIdx.x >> 5]; }
static inline void __syncthreads() { emu_wait(emu_block_bar); }
static inline void __syncwarp() { emu_wait(emu_warp().bar); }
// bar.sync id, count: a named barrier over `count` threads (a kernel's inline PTX is replaced by this call)
static EmuBarrier emu_named[16];
static inline void emu_named_barrier(int id, int count) {
  EmuBarrier &b = emu_named[id & 15];
  if (b.arrived == 0) b.expected = (unsigned)count;
  emu_wait(b);
}
// Block barrier that also ORs a predicate over the block.  The accumulator of a barrier generation is reset by
// the generation's first arrival; it was last used two generations earlier, and every thread read that result
// before it could arrive at the generation in between.
static uint32_t emu_or_acc[2];
static inline int __syncthreads_or(int pred) {
  const unsigned slot = (unsigned)(emu_block_bar.gen & 1u);
  if (emu_block_bar.arrived == 0) emu_or_acc[slot] = 0;
  if (pred) emu_or_acc[slot] = 1;
  emu_wait(emu_block_bar);
  return (int)emu_or_acc[slot];
}
// A collective: every lane deposits its value in the slot array of the operation's parity and
// waits for the warp; the array is reused two operations later, after every lane has read it.
static inline const uint32_t *emu_exchange(uint32_t v) {
  EmuWarp &w = emu_warp();
  const unsigned lane = threadIdx.x & 31;
  // the operation index a lane is at = completed ops of the warp as seen before it arrives
  const unsigned slot = (unsigned)(w.bar.gen & 1u);
  w.val[slot][lane] = v;
  emu_wait(w.bar);
  return w.val[slot];
}
static inline uint32_t __ballot_sync(uint32_t, bool pred) {
  const uint32_t *v = emu_exchange(pred ? 1u : 0u);
  uint32_t m = 0;
  const unsigned nl = std::min(32u, blockDim.x - (threadIdx.x & ~31u));
  for (unsigned i = 0; i < nl; ++i) m |= v[i] << i;
  return m;
}
static inline int __any_sync(uint32_t, int pred) {
  const uint32_t *v = emu_exchange(pred ? 1u : 0u);
  const unsigned nl = std::min(32u, blockDim.x - (threadIdx.x & ~31u));
  int r = 0;
  for (unsigned i = 0; i < nl; ++i) r |= (int)v[i];
  return r;
}
static inline uint32_t __shfl_sync(uint32_t, uint32_t x, int src) { return emu_exchange(x)[src & 31]; }
static inline uint32_t __shfl_up_sync(uint32_t, uint32_t x, int delta) {
  const int lane = threadIdx.x & 31;
  const uint32_t *v = emu_exchange(x);
  return lane >= delta ? v[lane - delta] : x;
}
static inline uint32_t __shfl_xor_sync(uint32_t, uint32_t x, int m) { return emu_exchange(x)[(threadIdx.x & 31) ^ m]; }
static inline unsigned long long __shfl_xor_sync(uint32_t, unsigned long long x, int m) {
  const uint32_t lo = emu_exchange((uint32_t)x)[(threadIdx.x & 31) ^ m];
  const uint32_t hi = emu_exchange((uint32_t)(x >> 32))[(threadIdx.x & 31) ^ m];
  return ((unsigned long long)hi << 32) | lo;
}
static inline unsigned int atomicOr(unsigned int *p, unsigned int v) {
  const unsigned int o = *p;
  *p = o | v;
  return o;
}
static inline int atomicAdd(int *p, int v) {
  const int o = *p;
  *p = o + v;
  return o;
}
static inline unsigned long long atomicOr(unsigned long long *p, unsigned long long v) {
  const unsigned long long o = *p;
  *p = o | v;
  return o;
}
static inline unsigned long long atomicAnd(unsigned long long *p, unsigned long long v) {
  const unsigned long long o = *p;
  *p = o & v;
  return o;
}
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline float __uint_as_float(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline uint32_t __float_as_uint(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
static inline float __int_as_float(int i) { return __uint_as_float((uint32_t)i); }
// IEEE double operations, round to nearest (compile the harness with -ffp-contract=off: no fused multiply-add)
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline int __shfl_sync(uint32_t m, int x, int src) { return (int)__shfl_sync(m, (uint32_t)x, src); }
static inline unsigned long long __shfl_sync(uint32_t m, unsigned long long x, int src) {
  const uint32_t lo = __shfl_sync(m, (uint32_t)x, src), hi = __shfl_sync(m, (uint32_t)(x >> 32), src);
  return ((unsigned long long)hi << 32) | lo;
}
static inline uint32_t __shfl_down_sync(uint32_t, uint32_t x, int delta) {
  const int lane = threadIdx.x & 31;
  const uint32_t *v = emu_exchange(x);
  return lane + delta < 32 ? v[lane + delta] : x;
}
static inline double __shfl_down_sync(uint32_t m, double x, int delta) {
  uint64_t u;
  memcpy(&u, &x, 8);
  const uint32_t lo = __shfl_down_sync(m, (uint32_t)u, delta), hi = __shfl_down_sync(m, (uint32_t)(u >> 32), delta);
  u = ((uint64_t)hi << 32) | lo;
  double r;
  memcpy(&r, &u, 8);
  return r;
}
static inline unsigned long __shfl_sync(uint32_t m, unsigned long x, int src) {
  return (unsigned long)__shfl_sync(m, (unsigned long long)x, src);
}
static inline unsigned long long __shfl_up_sync(uint32_t m, unsigned long long x, int delta) {
  const uint32_t lo = __shfl_up_sync(m, (uint32_t)x, delta), hi = __shfl_up_sync(m, (uint32_t)(x >> 32), delta);
  return ((unsigned long long)hi << 32) | lo;
}
static inline int __all_sync(uint32_t, int pred) {
  const uint32_t *v = emu_exchange(pred ? 1u : 0u);
  const unsigned nl = std::min(32u, blockDim.x - (threadIdx.x & ~31u));
  int r = 1;
  for (unsigned i = 0; i < nl; ++i) r &= (int)v[i];
  return r;
}
template <typename T>
static inline T __ldcs(const T *p) {
  return *p;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift) {
  return (uint32_t)((((uint64_t)hi << 32) | lo) >> (shift & 31u));
}
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
  const unsigned long long o = *p;
  *p = o + v;
  return o;
}
template <typename T>
static inline T __ldg(const T *p) {
  return *p;
}
static inline uint32_t atomicCAS(uint32_t *p, uint32_t expect, uint32_t v) {
  const uint32_t o = *p;
  if (o == expect) *p = v;
  return o;
}
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) {
  const unsigned long long o = *p;
  if (v > o) *p = v;
  return o;
}
static inline uint32_t atomicMax(uint32_t *p, uint32_t v) {
  const uint32_t o = *p;
  if (v > o) *p = v;
  return o;
}
static inline uint32_t atomicAdd(uint32_t *p, uint32_t v) {
  const uint32_t o = *p;
  *p = o + v;
  return o;
}

static void emu_trampoline() {
  (*emu_body)();
  emu_fibers[emu_cur].done = true;
  emu_switch(&emu_fibers[emu_cur].sp, emu_sched_sp);
  abort();   // a finished fiber is never resumed
}

// Runs `body` as a grid of `grid` blocks of `block` threads, block after block.  Returns false
// on a deadlock (some threads wait at a barrier the others never reach).
static inline bool emu_launch(unsigned grid, unsigned block, const std::function<void()> &body) {
  blockDim.x = block;
  gridDim.x = grid;
  emu_body = &body;
  const unsigned nwarps = (block + 31) / 32;
  for (unsigned b = 0; b < grid; ++b) {
    blockIdx.x = b;
    emu_block_bar = EmuBarrier();
    emu_block_bar.expected = block;
    for (EmuBarrier &nb : emu_named) nb = EmuBarrier();
    emu_warps.assign(nwarps, EmuWarp());
    for (unsigned w = 0; w < nwarps; ++w) emu_warps[w].bar.expected = std::min(32u, block - 32 * w);
    emu_fibers.clear();
    emu_fibers.resize(block);
    for (unsigned t = 0; t < block; ++t) {
      EmuFiber &f = emu_fibers[t];
      f.stack.resize(96 * 1024);
      // initial frame: six zeroed callee-saved registers, then the entry point as return address
      uintptr_t top = (reinterpret_cast<uintptr_t>(f.stack.data()) + f.stack.size()) & ~(uintptr_t)15;
      uint64_t *frame = reinterpret_cast<uint64_t *>(top - 64);
      memset(frame, 0, 64);
      frame[6] = reinterpret_cast<uint64_t>(&emu_trampoline);
      f.sp = frame;
    }
    unsigned live = block;
    while (live) {
      bool progressed = false;
      for (unsigned t = 0; t < block; ++t) {
        EmuFiber &f = emu_fibers[t];
        if (f.done) continue;
        if (f.wait && f.wait->gen == f.wait_gen) continue;   // still blocked
        f.wait = nullptr;
        emu_cur = t;
        threadIdx.x = t;
        emu_switch(&emu_sched_sp, f.sp);
        progressed = true;
        if (f.done) --live;
      }
      if (!progressed) {
        fprintf(stderr, "emu: deadlock in block %u (%u threads blocked at divergent barriers)\n", b, live);
        return false;
      }
    }
  }
  return true;
}
