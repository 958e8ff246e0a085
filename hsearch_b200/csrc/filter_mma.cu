// Pipelined tcgen05 candidate filter for the Euclidean metric (K3 / K4 bulk path).
//
// The filter bound of a (query q, member m) pair is the squared distance
//     d2(q, m) = |x_m|^2 + |q|^2 - 2 <x_m, q>,       x_m, q in R^(8*len)
// (the quantity PairwiseDistance_square accumulates, motif_both_points.cpp:176-183).
// Over one bucket <x_m, q> is a dense [members x 8*len] x [8*len x queries]
// contraction, so it runs on the 5th-generation tensor cores:
//
//   A  [128 members][kp]  FP16 embedding rows, built in shared memory from the
//                         1-byte residue codes of the bucket-ordered store (one
//                         16-byte table row per residue), plus two columns of 1.0,
//   B  [<= qmax queries][kp]  FP16 query coordinates plus two columns holding
//                         c_q = -|q|^2/2 + margin (split hi/lo), resident per item,
//   D  [128][<= 256]      FP32 accumulators in TMEM, double buffered:
//                         D = <x~_m, q~> + c_q,
//
// and a pair survives iff  D >= rowthr_m = ((1 - beta)|x_m|^2 - thr) / 2, i.e. iff
// its FP16/FP32 distance estimate is within the threshold plus a rigorous error
// margin (derivation in DESIGN.md, "filter margins").  The filter never decides
// a hit: survivors go to the exact FP64 stage (verify.cu), so hits and distances
// stay bit-exact.
//
// One persistent CTA per SM, warp specialised:
//   warps 0-7   epilogue: tcgen05.ld the accumulators, 3-input max per 8 columns,
//               warp-uniform branch into the rare survivor scan, per-warp staging
//               of survivors in shared memory, one global atomic per flush;
//   warps 8-11  producers: build A tiles (3-4 stages), load B when the item changes;
//   warp 12     one thread issues tcgen05.mma (M=128, N<=256, K=16) and commits to
//               the mbarriers that free A stages / publish accumulators.
// Shared-memory operand layout (both operands K-major, no swizzle): 8x8 FP16
// core matrices of 128 contiguous bytes; core matrices of one 8-column group are
// stacked over the row groups (SBO = 128 B), column groups follow each other at
// LBO = 2048 B (A, 128 rows) or qmax*16 B (B).
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "internal.cuh"
#include "verify.cuh"

namespace hs {

constexpr int kMmaEpiWarps = 16;     // 4 per TMEM lane quadrant
constexpr int kMmaEpiPerStage = kMmaEpiWarps;  // all epilogue warps drain every accumulator stage
constexpr int kMmaProdWarps = 4;
// The producer warps can work in kMmaProdGroups independent groups, group i building the tiles
// i, i + groups, ... with 128 / groups threads (several rows per thread), so that several tiles
// are in flight on the producer side.  Measured slower (39.2 against 36.2 ms at bench C2) and not
// validated for more than one group: one group is the supported configuration.
#ifndef HS_MMA_PROD_GROUPS
#define HS_MMA_PROD_GROUPS 1
#endif
constexpr int kMmaProdGroups = HS_MMA_PROD_GROUPS;
constexpr int kMmaProdGroupThreads = kMmaProdWarps * 32 / kMmaProdGroups;
constexpr int kMmaThreads = (kMmaEpiWarps + kMmaProdWarps + 2) * 32;  // 704
constexpr int kMmaProdThread0 = kMmaEpiWarps * 32;                    // 512
constexpr int kMmaIssueWarp = kMmaEpiWarps + kMmaProdWarps;           // 20
constexpr int kMmaLoadWarp = kMmaIssueWarp + 1;                       // 21: unit scheduler + code loader
constexpr int kMmaUnitRing = 4;      // units published ahead by the scheduler
#ifndef HS_MMA_UNIT_BATCH
#define HS_MMA_UNIT_BATCH 4
#endif
constexpr int kMmaUnitBatch = HS_MMA_UNIT_BATCH;  // consecutive units taken per atomic (keeps B resident)
constexpr int kMmaMaxCodeRing = 8;   // tiles of residue codes in flight (bulk async copies)
constexpr int kMmaM = 128;           // members per tile (UMMA M)
#ifndef HS_MMA_N
#define HS_MMA_N 256
#endif
constexpr int kMmaN = HS_MMA_N;      // accumulator columns per stage (<= 256, the UMMA N limit)
constexpr int kMmaAccStages = 512 / kMmaN;  // the stages fill the 512 TMEM columns
constexpr uint32_t kMmaTmemCols = 512;
constexpr int kMmaMaxStages = 4;     // A stages
#ifdef HS_MMA_PACK
constexpr int kMmaRowRing = 16;      // >= A stages + tiles in the accumulator stages (kMmaAccStages * 4) + 1
#else
constexpr int kMmaRowRing = 8;       // >= A stages + accumulator stages
#endif
#ifdef HS_MMA_PACK
constexpr int kMmaRing = 6;          // (the larger row ring takes the shared memory of two slots)
constexpr int kMmaRingFlush = 4;
#else
constexpr int kMmaRing = 8;          // survivors staged per epilogue LANE between flushes (lane-private slots, no atomics)
constexpr int kMmaRingFlush = 6;     // a lane holding this many makes its warp flush at the next group boundary
#endif
constexpr int kMmaAGroupBytes = 2048;

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
#ifdef HS_MBAR_HINT
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(0x989680u)  // suspend-time hint
      : "memory");
#else
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
#endif
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#ifdef HS_MBAR_SPIN
  while (!mbar_test_wait(bar, parity)) {
  }
#else
  while (!mbar_try_wait(bar, parity)) {
  }
#endif
}
// single-thread roles (MMA issuer, loader): poll without suspending
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
#if defined(HS_MBAR_SPIN1) || defined(HS_MBAR_SPIN)
  while (!mbar_test_wait(bar, parity)) {
  }
#else
  while (!mbar_try_wait(bar, parity)) {
  }
#endif
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// one arrival per warp: every lane's prior work is ordered before it by the warp barrier
__device__ __forceinline__ void mbar_arrive_warp(uint32_t bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bulk async copy global -> shared (TMA engine, no tensor map); 16-byte aligned, size multiple of 16
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
#ifdef HS_DIAG_NOPROXYFENCE
#define HS_PROXY_FENCE()
#else
#define HS_PROXY_FENCE() asm volatile("fence.proxy.async.shared::cta;" ::: "memory")
#endif
#ifdef HS_DIAG_NOFENCE
__device__ __forceinline__ void tc_before() {}
__device__ __forceinline__ void tc_after() {}
#else
__device__ __forceinline__ void tc_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
#endif
__device__ __forceinline__ void fence_async_shared() { HS_PROXY_FENCE(); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (sm_100: version 1).
__device__ __forceinline__ uint64_t mma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= 1ull << 46;
  return d;
}
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
#ifdef HS_DIAG_ARRIVE
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
  return;
#endif
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// Append the survivors the lanes of a warp hold in their private slots to the global list: a warp
// prefix sum of the per-lane counts, one global atomic, every lane writes its own run.  Returns 0
// (the lane's new count).
__device__ __noinline__ uint32_t mma_flush(const uint2 (*ring)[32], uint32_t lc, uint32_t table, Survivor *surv,
                                           unsigned long long cap, unsigned long long *count, int lane) {
  uint32_t inc = lc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
  if (total == 0u) return 0u;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(count, (unsigned long long)total);
  base = __shfl_sync(0xffffffffu, base, 0) + (inc - lc);
  for (uint32_t j = 0; j < lc; ++j) {
    const uint2 e = ring[j][lane];
    if (base + j < cap) {
      Survivor sv;
      sv.query = e.x;  // index into the query list; the exact stage resolves it
      sv.table = table;
      sv.pos = e.y;    // fragment id
      sv.pad = 3;
      surv[base + j] = sv;
    }
  }
  __syncwarp();
  return 0u;
}

__device__ __noinline__ void mma_emit_global(Survivor *surv, unsigned long long cap, unsigned long long *count, uint32_t qidx,
                                             uint32_t table, uint32_t rid) {
  const unsigned long long gi = atomicAdd(count, 1ull);
  if (gi < cap) {
    Survivor sv;
    sv.query = qidx;
    sv.table = table;
    sv.pos = rid;
    sv.pad = 3;  // query = index into the query list, pos = fragment id
    surv[gi] = sv;
  }
}

// Copy a warp's queued events (n * 5 sixteen-byte words) to the global event list: one atomic, coalesced.
__device__ __noinline__ void mma_flush_events(const uint4 *q, uint32_t n, uint4 *events, unsigned long long cap,
                                              unsigned long long *count, int lane) {
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(count, (unsigned long long)n);
  base = __shfl_sync(0xffffffffu, base, 0);
  const uint32_t nw = base + n <= cap ? n * 5u : (base < cap ? (uint32_t)(cap - base) * 5u : 0u);
  for (uint32_t i = lane; i < nw; i += 32) events[base * 5u + i] = q[i];
  __syncwarp();
}

#ifdef HS_MMA_PROF
#define PROF_DECL(n) unsigned long long n = 0
#define PROF_T0() const long long _t0 = clock64()
#define PROF_ADD(n) n += (unsigned long long)(clock64() - _t0)
#define PROF_OUT(i, n) atomicAdd(a.prof + (i), n)
#else
#define PROF_DECL(n)
#define PROF_T0()
#define PROF_ADD(n)
#define PROF_OUT(i, n)
#endif

// Event staging (HS_MMA_EVENTS).  Finding WHICH of a chunk's 16 columns pass is ~75 instructions
// executed by the one or two lanes of a warp whose row reached its threshold -- 3-6 % SIMD
// efficiency, 40 % of the kernel's instructions, 9 ms of a 35 ms launch (profiles/r02_filter_
// experiments.md).  Instead the lane stores the 16 raw accumulators with the row's threshold and
// ids (80 bytes, a dozen instructions) into its warp's queue; the queues are flushed to global
// memory and resolve_events_kernel applies the very same test `value >= threshold`, one event
// per lane at full SIMD width, and writes the survivors.
struct MmaEvent {
  uint32_t v[16];        // the accumulators of columns qidx0 .. qidx0 + 15 (FP32 bit patterns)
  uint32_t qidx0;        // index of column 0 in the query list
  uint32_t rid;          // fragment id of the row
  uint32_t rt;           // the row threshold (FP32 bit pattern)
  uint32_t ncol_table;   // valid columns (1..16) | table << 8
};
static_assert(sizeof(MmaEvent) == 80, "five 16-byte words");
constexpr int kMmaEvCap = 25;      // events queued per epilogue warp (2000 bytes)
constexpr int kMmaEvFlushAt = 10;  // queued events that make the warp flush at the next group boundary

struct MmaItem {
  uint32_t table;
  uint32_t q_begin, q_end;  // range of qlist
  uint32_t pad;
};
struct MmaUnit {
  uint32_t item;
  uint32_t m_begin, m_end;  // member positions in the table's bucket-ordered store
  uint32_t pad;
};

struct MmaArgs {
  const MmaItem *items;
  const MmaUnit *units;
  uint32_t nunits;
  uint32_t *unit_counter;     // zeroed before the launch; CTAs take batches of units from it
  const uint32_t *qlist;
  const __half *qb16;         // [Q][kp] query rows (coordinates, c_q hi/lo, zero padding)
  uint32_t tq_base;           // qb16 row of query id x is x - tq_base
  const uint8_t *const *stores;
  const uint32_t *const *sorted_ids;  // per table: ids in bucket order (nullptr entry: identity)
  uint64_t npad;
  int len, kp, nstages, qmax, cring;
  float thr, beta;
  unsigned long long *prof;   // HS_MMA_PROF builds only: per-role cycle counters [16]
  uint32_t debug;             // HS_MMA_PROF builds only (HS_MMA_DEBUG): 1 = epilogue loads but does not scan
  const uint4 *tab16;         // [20] FP16 embedding rows (8 halves each)
  const float *nx32;          // [20] squared row norms, rounded down
  Survivor *surv;
  unsigned long long surv_cap;
  unsigned long long *surv_count;
  MmaEvent *events;           // HS_MMA_EVENTS: rows that reached their threshold, with their 16 accumulators
  unsigned long long ev_cap;
  unsigned long long *ev_count;
};

struct MmaShared {
  uint64_t a_full[kMmaMaxStages], a_empty[kMmaMaxStages];
  uint64_t t_full[kMmaAccStages], t_empty[kMmaAccStages];
  uint64_t b_full;
  uint64_t u_full[kMmaUnitRing], u_empty[kMmaUnitRing];
  uint64_t c_full[kMmaMaxCodeRing], c_empty[kMmaMaxCodeRing];
  uint32_t uring[kMmaUnitRing];
  uint32_t tmem_base;
  uint32_t pad;
  // every embedding row stored 8 times, copy j in bank group j (16 bytes = 4 banks): lane i of a
  // quarter warp reads copy i & 7, so the 8 lanes of one LDS.128 phase never share a bank
  uint4 tab16[HS_AA][8];
  float nx32[HS_AA];
  float rowthr[kMmaRowRing][kMmaM];
  uint32_t rowid[kMmaRowRing][kMmaM];   // fragment id of every row (the exact stage then skips the id gather)
  // survivors staged per lane: (index into the query list, fragment id); slot-major so that a
  // warp-wide access touches consecutive words
#if HS_MMA_EVENTS
  uint4 evq[kMmaEpiWarps][kMmaEvCap * 5];
  uint32_t evcount[kMmaEpiWarps];
#else
  uint2 ring[kMmaEpiWarps][kMmaRing][32];
#endif
};

// Tile packing (-DHS_MMA_PACK; off: measured 36.1 ms against 35.0 at bench C2, profiles/r02_filter_
// experiments.md).  An item whose queries fit one group (<= kMmaN) narrower than half a stage shares
// each accumulator stage between 2 or 4 consecutive tiles of the unit: tile j of the batch gets
// the column block [j * kMmaN / P, ...) (same B operand, its own A stage), so the hand-off between
// the MMA issuer and the epilogue is paid once per P tiles.  Issuer and epilogue derive P from the
// item the same way.
__device__ __forceinline__ uint32_t mma_pack(uint32_t nq) {
#ifdef HS_MMA_PACK
  return nq <= (uint32_t)kMmaN / 4u ? 4u : nq <= (uint32_t)kMmaN / 2u ? 2u : 1u;
#else
  return 1u;
#endif
}

// Every role walks the same sequence of units, published by the scheduler lane
// through a small ring: returns the next unit index, or >= nunits at the end.
__device__ __forceinline__ uint32_t mma_next_unit(MmaShared &sh, uint32_t k, int lane, bool whole_warp) {
  const uint32_t slot = k % kMmaUnitRing;
  mbar_wait(smem_addr(&sh.u_full[slot]), (k / kMmaUnitRing) & 1u);
  const uint32_t u = sh.uring[slot];
  if (whole_warp) mbar_arrive_warp(smem_addr(&sh.u_empty[slot]), lane);
  else mbar_arrive(smem_addr(&sh.u_empty[slot]));
  return u;
}

// Survivors carry the index of their query in the query list (pad = 1); the exact stage
// resolves it and, for all-pairs runs, keeps only pairs with query id < member position.
// LENB: compile-time bound of the fragment length (unroll bound of the A-tile producers).
// (22 warps are allocated like 24: 80 registers per thread is the limit, 88 does not launch)
template <int LENB>
__global__ void __launch_bounds__(kMmaThreads, 1)
filter_mma_kernel(MmaArgs a) {
  extern __shared__ __align__(1024) unsigned char mma_smem[];
  __shared__ __align__(16) MmaShared sh;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int len = a.len, kp = a.kp, S = a.nstages, qmax = a.qmax, D = a.cring;
  const int ngrp = kp >> 3;                          // 8-column groups
  const uint32_t a_stage_bytes = (uint32_t)ngrp * kMmaAGroupBytes;
  const uint32_t b_lbo = (uint32_t)qmax * 16u;
  unsigned char *sB = mma_smem;
  unsigned char *sA = sB + (size_t)ngrp * b_lbo;
  unsigned char *sC = sA + (size_t)S * a_stage_bytes;  // code ring: [D][len][128] bytes
  const uint32_t nunits = a.nunits;

  if (tid == 0) {
    for (int s = 0; s < kMmaMaxStages; ++s) {
      mbar_init(smem_addr(&sh.a_full[s]), kMmaProdWarps / kMmaProdGroups);
      mbar_init(smem_addr(&sh.a_empty[s]), 1);
    }
    for (int s = 0; s < kMmaAccStages; ++s) {
      mbar_init(smem_addr(&sh.t_full[s]), 1);
      mbar_init(smem_addr(&sh.t_empty[s]), kMmaEpiPerStage);
    }
    mbar_init(smem_addr(&sh.b_full), kMmaProdWarps);
    for (int s = 0; s < kMmaUnitRing; ++s) {
      mbar_init(smem_addr(&sh.u_full[s]), 1);
      mbar_init(smem_addr(&sh.u_empty[s]), kMmaEpiWarps + kMmaProdWarps + 1);
    }
    for (int s = 0; s < kMmaMaxCodeRing; ++s) {
      mbar_init(smem_addr(&sh.c_full[s]), 1);
      mbar_init(smem_addr(&sh.c_empty[s]), kMmaProdWarps / kMmaProdGroups);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < HS_AA * 8) sh.tab16[tid >> 3][tid & 7] = a.tab16[tid >> 3];
  if (tid < HS_AA) sh.nx32[tid] = a.nx32[tid];
#if HS_MMA_EVENTS
  if (tid < kMmaEpiWarps) sh.evcount[tid] = 0u;
#endif
  if (warp == kMmaIssueWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&sh.tmem_base)),
                 "r"(kMmaTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // A stages start as zeros; the constant columns (1.0, 1.0 at 8*len, 8*len+1) are written once
  {
    uint4 *p = reinterpret_cast<uint4 *>(sA);
    const int n = (int)((size_t)S * a_stage_bytes / 16);
    for (int i = tid; i < n; i += kMmaThreads) p[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  for (int i = tid; i < S * kMmaM; i += kMmaThreads) {
    const int s = i / kMmaM, r = i - s * kMmaM;
    *reinterpret_cast<uint32_t *>(sA + (size_t)s * a_stage_bytes + (size_t)len * kMmaAGroupBytes + (r >> 3) * 128 +
                                  (r & 7) * 16) = 0x3C003C00u;  // (1.0h, 1.0h)
  }
  fence_async_shared();
  tc_before();
  __syncthreads();
  tc_after();
  const uint32_t tmem_base = sh.tmem_base;

  // Tiles of a unit start at its first member rounded down to 16 so that every 128-byte
  // row segment of the position-major store is 16-byte aligned for the bulk copies; rows
  // outside [m_begin, m_end) are masked by an infinite row threshold.
  if (warp == kMmaLoadWarp) {
    // ============================ scheduler + code loader ==============================
    if (lane == 0) {
      uint32_t k = 0, ct = 0, batch_next = 0, batch_end = 0;
      for (;; ++k) {
        const uint32_t slot = k % kMmaUnitRing;
        mbar_wait(smem_addr(&sh.u_empty[slot]), ((k / kMmaUnitRing) & 1u) ^ 1u);
        if (batch_next == batch_end) {
          batch_next = atomicAdd(a.unit_counter, (uint32_t)kMmaUnitBatch);
          batch_end = batch_next + kMmaUnitBatch;
        }
        const uint32_t u = batch_next < nunits ? batch_next : 0xffffffffu;
        ++batch_next;
        sh.uring[slot] = u;
        mbar_arrive(smem_addr(&sh.u_full[slot]));
        if (u >= nunits) break;
        const MmaUnit un = a.units[u];
        const MmaItem it = a.items[un.item];
        const uint8_t *store = a.stores[it.table];
        const uint32_t base = un.m_begin & ~15u;
        const uint32_t ntiles = (un.m_end - base + kMmaM - 1) / kMmaM;
#if defined(HS_DIAG_NOLOAD)
        if (ntiles) continue;
#endif
        for (uint32_t t = 0; t < ntiles; ++t, ++ct) {
          const uint32_t d = ct % D;
          mbar_wait(smem_addr(&sh.c_empty[d]), ((ct / D) & 1u) ^ 1u);
          const uint32_t bar = smem_addr(&sh.c_full[d]);
          mbar_arrive_expect_tx(bar, (uint32_t)len * 128u);
          const uint32_t dst = smem_addr(sC) + d * (uint32_t)len * 128u;
          const uint8_t *src = store + base + t * kMmaM;
          for (int p = 0; p < len; ++p) bulk_g2s(dst + (uint32_t)p * 128u, src + (uint64_t)p * a.npad, 128u, bar);
        }
      }
    }
    __syncwarp();
  } else if (warp >= kMmaEpiWarps && warp < kMmaIssueWarp) {
    // ============================ producers ============================================
    const int r = tid - kMmaProdThread0;  // 0 .. 127 (B rows are loaded by all producer threads together)
    constexpr int NR = kMmaProdGroups;     // member rows of a tile per thread
    const int grp = r / kMmaProdGroupThreads, lr = r % kMmaProdGroupThreads;
    uint32_t row_off[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int ri = lr + i * kMmaProdGroupThreads;
      row_off[i] = (uint32_t)(ri >> 3) * 128u + (uint32_t)(ri & 7) * 16u;
    }
    uint32_t pt = 0;                       // tiles of all units so far (this group builds pt % groups == grp)
    uint32_t prev_item = 0xffffffffu;
    PROF_DECL(p_wait_c);
    PROF_DECL(p_wait_a);
    PROF_DECL(p_bload);
    PROF_DECL(p_build);
    PROF_DECL(p_fence);
#ifdef HS_MMA_PROF
    const long long p_start = clock64();
#endif
    for (uint32_t k = 0;; ++k) {
      const uint32_t u = mma_next_unit(sh, k, lane, true);
      if (u >= nunits) break;
      const MmaUnit un = a.units[u];
      const MmaItem it = a.items[un.item];
      if (un.item != prev_item) {
        PROF_T0();
        // every MMA that reads the old B has completed once the latest A stage was released
        if (pt > 0) mbar_wait(smem_addr(&sh.a_empty[(pt - 1) % S]), ((pt - 1) / S) & 1u);
        const uint32_t nq = it.q_end - it.q_begin;
        const uint32_t nqp = (nq + 15u) & ~15u;
        for (uint32_t q = (uint32_t)r; q < nqp; q += kMmaProdWarps * 32) {
          unsigned char *drow = sB + (q >> 3) * 128 + (q & 7) * 16;
          if (q < nq) {
            const uint4 *src = reinterpret_cast<const uint4 *>(a.qb16 + (size_t)(a.qlist[it.q_begin + q] - a.tq_base) * kp);
            for (int g0 = 0; g0 < ngrp; g0 += 6) {  // six 16-byte loads in flight per step
              uint4 v[6];
#pragma unroll
              for (int j = 0; j < 6; ++j)
                if (g0 + j < ngrp) v[j] = __ldg(src + g0 + j);
#pragma unroll
              for (int j = 0; j < 6; ++j)
                if (g0 + j < ngrp) *reinterpret_cast<uint4 *>(drow + (size_t)(g0 + j) * b_lbo) = v[j];
            }
          } else {
            for (int g = 0; g < ngrp; ++g)  // padding column: c = -57344 never passes
              *reinterpret_cast<uint4 *>(drow + (size_t)g * b_lbo) = make_uint4(g == len ? 0x0000FB00u : 0u, 0u, 0u, 0u);
          }
        }
        fence_async_shared();
        mbar_arrive_warp(smem_addr(&sh.b_full), lane);
        prev_item = un.item;
        PROF_ADD(p_bload);
      }
      const uint32_t base = un.m_begin & ~15u;
      const uint32_t ntiles = (un.m_end - base + kMmaM - 1) / kMmaM;
      const uint32_t *ids = a.sorted_ids[it.table];
      for (uint32_t t = 0; t < ntiles; ++t, ++pt) {
        if ((int)(pt % (uint32_t)kMmaProdGroups) != grp) continue;
        const uint32_t s = pt % S, d = pt % D;
        uint32_t myid[NR];
        bool valid[NR];
        // the rows' fragment ids (coalesced; issued before the waits below, used after the A build)
#pragma unroll
        for (int i = 0; i < NR; ++i) {
          const uint32_t mypos = base + t * kMmaM + (uint32_t)(lr + i * kMmaProdGroupThreads);
          valid[i] = mypos >= un.m_begin && mypos < un.m_end;
          myid[i] = (valid[i] && ids) ? __ldg(ids + mypos) : mypos;
        }
        // the rows' residue codes from the ring the loader fills
        {
          PROF_T0();
          mbar_wait(smem_addr(&sh.c_full[d]), (pt / D) & 1u);
          PROF_ADD(p_wait_c);
        }
        uint8_t code[NR][LENB];
#pragma unroll
        for (int i = 0; i < NR; ++i) {
          const unsigned char *crow = sC + (size_t)d * len * 128 + lr + i * kMmaProdGroupThreads;
#pragma unroll
          for (int p = 0; p < LENB; ++p)
            if (p < len) code[i][p] = crow[p * 128];
        }
        mbar_arrive_warp(smem_addr(&sh.c_empty[d]), lane);
        {
          PROF_T0();
          mbar_wait(smem_addr(&sh.a_empty[s]), ((pt / S) & 1u) ^ 1u);
          PROF_ADD(p_wait_a);
        }
#ifdef HS_MMA_PROF
        const long long _pb0 = clock64();
#endif
        unsigned char *dst = sA + (size_t)s * a_stage_bytes;
        float nx[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) nx[i] = 0.f;
        // batches of 5 positions: their table rows are loaded before the first is stored, so that the
        // shared-memory loads overlap (one register quad reused for every position serialises
        // LDS -> STS -> LDS ...: measured 1.7 k cycles per tile)
#pragma unroll
        for (int i = 0; i < NR; ++i)
#pragma unroll
          for (int p0 = 0; p0 < LENB; p0 += 5) {
            uint4 rowv[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) {
              const int p = p0 + j;
              if (p < LENB && p < len) {
#if defined(HS_DIAG_NOPROD)
                continue;
#endif
#ifdef HS_MMA_PROF
                if (a.debug & 4u) continue;  // experiment: no A build
#endif
                const int c = min((int)code[i][p] / kCodeScale, HS_AA - 1);  // (rows outside the unit hold foreign bytes)
                rowv[j] = sh.tab16[c][lane & 7];
                nx[i] += sh.nx32[c];
              }
            }
#pragma unroll
            for (int j = 0; j < 5; ++j) {
              const int p = p0 + j;
              if (p < LENB && p < len) {
#if defined(HS_DIAG_NOPROD)
                continue;
#endif
#ifdef HS_MMA_PROF
                if (a.debug & 4u) continue;
#endif
                *reinterpret_cast<uint4 *>(dst + row_off[i] + (size_t)p * kMmaAGroupBytes) = rowv[j];
              }
            }
          }
#pragma unroll
        for (int i = 0; i < NR; ++i) {
          // rowthr = ((1 - beta) nx - thr) / 2, rounded down; +inf rows never pass
          // (nx is a sum of <= 32 FP32 terms: 4e-6 covers its rounding)
          float rt = 0.5f * (nx[i] * (1.0f - 4e-6f) * (1.0f - a.beta) - a.thr);
          rt -= (nx[i] + a.thr) * 2.4e-7f + 1e-6f;
          const int ri = lr + i * kMmaProdGroupThreads;
          sh.rowthr[pt % kMmaRowRing][ri] = valid[i] ? rt : __int_as_float(0x7f800000);
          sh.rowid[pt % kMmaRowRing][ri] = myid[i];
        }
#ifdef HS_MMA_PROF
        const long long _pb1 = clock64();
        p_build += (unsigned long long)(_pb1 - _pb0);
#endif
        fence_async_shared();
        mbar_arrive_warp(smem_addr(&sh.a_full[s]), lane);
#ifdef HS_MMA_PROF
        p_fence += (unsigned long long)(clock64() - _pb1);
#endif
      }
    }
#ifdef HS_MMA_PROF
    if (tid == kMmaProdThread0) {
      PROF_OUT(0, (unsigned long long)(clock64() - p_start));
      PROF_OUT(1, p_wait_c);
      PROF_OUT(2, p_wait_a);
      PROF_OUT(3, p_bload);
      PROF_OUT(16, p_build);
      PROF_OUT(17, p_fence);
    }
#endif
  } else if (warp == kMmaIssueWarp) {
    // ============================ MMA issuer ===========================================
    if (lane == 0) {
      uint32_t mt = 0, mg = 0, nb = 0;
      uint32_t prev_item = 0xffffffffu;
      PROF_DECL(m_wait_a);
      PROF_DECL(m_wait_t);
      PROF_DECL(m_wait_b);
      PROF_DECL(m_groups);
      const uint32_t sA_u32 = smem_addr(sA), sB_u32 = smem_addr(sB);
      const int ksteps = kp >> 4;
      PROF_DECL(m_unit);
      PROF_DECL(m_issue);
      PROF_DECL(m_commit);
#ifdef HS_MMA_PROF
      const long long m_start = clock64();
#endif
      for (uint32_t k = 0;; ++k) {
#ifdef HS_MMA_PROF
        const long long _mu0 = clock64();
#endif
        const uint32_t u = mma_next_unit(sh, k, lane, false);
        if (u >= nunits) break;
        const MmaUnit un = a.units[u];
        const MmaItem it = a.items[un.item];
#ifdef HS_MMA_PROF
        m_unit += (unsigned long long)(clock64() - _mu0) + (it.table == 0xffffffffu ? 1 : 0);
#endif
        if (un.item != prev_item) {
          {
            PROF_T0();
            mbar_wait_spin(smem_addr(&sh.b_full), nb & 1u);
            PROF_ADD(m_wait_b);
          }
          ++nb;
          prev_item = un.item;
        }
        const uint32_t nq = it.q_end - it.q_begin;
        const uint32_t ngroups = (nq + kMmaN - 1) / kMmaN;
        const uint32_t base = un.m_begin & ~15u;
        const uint32_t ntiles = (un.m_end - base + kMmaM - 1) / kMmaM;
        const uint32_t P = mma_pack(nq);            // tiles per accumulator stage (1 unless the group is narrow)
        const uint32_t blockw = (uint32_t)kMmaN / P;   // columns of a tile's block
        for (uint32_t t = 0; t < ntiles;) {
          const uint32_t nbatch = min(P, ntiles - t);
          for (uint32_t g = 0; g < ngroups; ++g, ++mg) {   // (P > 1 only with ngroups == 1)
            const uint32_t as = mg % kMmaAccStages;
            {
              PROF_T0();
              mbar_wait_spin(smem_addr(&sh.t_empty[as]), ((mg / kMmaAccStages) & 1u) ^ 1u);
              PROF_ADD(m_wait_t);
            }
#ifdef HS_MMA_PROF
            ++m_groups;
#endif
            tc_after();
            const uint32_t ng = min((uint32_t)kMmaN, nq - g * kMmaN);
            const uint32_t ngp = (ng + 15u) & ~15u;
            const uint32_t idesc = (1u << 4) | ((ngp >> 3) << 17) | ((uint32_t)(kMmaM >> 4) << 24);  // F16xF16->F32
            const uint32_t b0 = sB_u32 + (g * kMmaN >> 3) * 128u;
            for (uint32_t j = 0; j < nbatch; ++j) {
              const uint32_t s = (mt + j) % S;
              if (g == 0) {
                PROF_T0();
                mbar_wait_spin(smem_addr(&sh.a_full[s]), ((mt + j) / S) & 1u);
                PROF_ADD(m_wait_a);
                tc_after();
              }
              const uint32_t d_tmem = tmem_base + as * kMmaN + j * blockw;
              const uint32_t a0 = sA_u32 + s * a_stage_bytes;
#ifdef HS_MMA_PROF
              const long long _mi0 = clock64();
#endif
              for (int kk = 0; kk < ksteps; ++kk) {
                const uint64_t ad = mma_desc(a0 + (uint32_t)kk * 2u * kMmaAGroupBytes, kMmaAGroupBytes, 128u);
                const uint64_t bd = mma_desc(b0 + (uint32_t)kk * 2u * b_lbo, b_lbo, 128u);
#ifdef HS_MMA_PROF
                if ((a.debug & 2u) && kk > 0) continue;  // experiment: one MMA per group instead of kp/16
#endif
#ifdef HS_DIAG_NMMA
                if (kk >= HS_DIAG_NMMA) continue;        // experiment: fewer MMAs per group
#endif
                mma_f16_ss(d_tmem, ad, bd, idesc, kk > 0 ? 1u : 0u);
              }
#ifdef HS_MMA_PROF
              m_issue += (unsigned long long)(clock64() - _mi0);
#endif
              if (g + 1 == ngroups) mma_commit(smem_addr(&sh.a_empty[s]));
            }
            {
              PROF_T0();
              mma_commit(smem_addr(&sh.t_full[as]));
              PROF_ADD(m_commit);
            }
          }
          t += nbatch;
          mt += nbatch;
        }
      }
#ifdef HS_MMA_PROF
      PROF_OUT(4, m_wait_a);
      PROF_OUT(5, m_wait_t);
      PROF_OUT(6, m_wait_b);
      PROF_OUT(7, m_groups);
      PROF_OUT(10, m_unit);
      PROF_OUT(12, m_issue);
      PROF_OUT(13, m_commit);
      PROF_OUT(15, (unsigned long long)(clock64() - m_start));
#endif
    }
    __syncwarp();
  } else {
    // ============================ epilogue =============================================
    // Lane = member row; four warps per TMEM lane quadrant share the columns of a stage in
    // 32-column chunks.  Per 16 accumulator columns: their maximum (tree of 3-input max)
    // against the row threshold and one warp vote.  The rare rows that reach it hand their
    // 16 accumulators to the warp through a slot in shared memory; 16 lanes test one column
    // each (two rows per pass) and passing pairs are staged per warp, one global atomic per
    // flush.
    const int quad = warp & 3, sub = warp >> 2;
    const int row = quad * 32 + lane;
#if HS_MMA_EVENTS
    uint4 *evq = sh.evq[warp];
    uint32_t *evcount = &sh.evcount[warp];
#else
    uint2 (*ring)[32] = sh.ring[warp];
    uint32_t lc = 0;  // survivors this lane holds in its slots
#endif
    uint32_t et = 0, eg = 0;
    PROF_DECL(e_wait_t);
    PROF_DECL(e_rare);
    PROF_DECL(e_unit);
#ifdef HS_MMA_PROF
    const long long e_start = clock64();
#endif
    for (uint32_t k = 0;; ++k) {
#ifdef HS_MMA_PROF
      const long long _u0 = clock64();
#endif
      const uint32_t u = mma_next_unit(sh, k, lane, true);
      if (u >= nunits) break;
      const MmaUnit un = a.units[u];
      const MmaItem it = a.items[un.item];
#ifdef HS_MMA_PROF
      e_unit += (unsigned long long)(clock64() - _u0);
#endif
      const uint32_t nq = it.q_end - it.q_begin;
      const uint32_t ngroups = (nq + kMmaN - 1) / kMmaN;
      const uint32_t base = un.m_begin & ~15u;
      const uint32_t ntiles = (un.m_end - base + kMmaM - 1) / kMmaM;
      const uint32_t P = mma_pack(nq);            // tiles per accumulator stage (see mma_pack)
      const uint32_t blockw = (uint32_t)kMmaN / P;   // columns of a tile's block
      for (uint32_t t = 0; t < ntiles;) {
        const uint32_t nbatch = min(P, ntiles - t);
        for (uint32_t g = 0; g < ngroups; ++g, ++eg) {
          const uint32_t as = eg % kMmaAccStages;
          {
            PROF_T0();
            mbar_wait(smem_addr(&sh.t_full[as]), (eg / kMmaAccStages) & 1u);
            PROF_ADD(e_wait_t);
          }
          {
            PROF_T0();
            tc_after();
            PROF_ADD(e_unit);
          }
          float rt = 0.f;
          uint32_t rid = 0u;           // threshold and fragment id of this lane's row in tile rj of the batch
          uint32_t rj = 0xffffffffu;
          const uint32_t ng = min((uint32_t)kMmaN, nq - g * kMmaN);
          const uint32_t ngp = (ng + 15u) & ~15u;  // columns the MMA wrote per tile
          const uint32_t nchunks = ngp >> 4;  // 16-column chunks per tile; this warp takes c = sub, sub + 4, ...
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * kMmaN;
          const uint32_t qbase = it.q_begin + g * kMmaN;  // index into the query list of column 0
          // The 16 accumulators in vv = columns col0 .. col0+15 of tile j's block: their maximum (tree
          // of 3-input max) against the row threshold.  A row that reaches it (about 2 % of the rows
          // of a chunk) builds a branch-free column mask and puts its passing columns into its own
          // slots in shared memory -- no atomic, no other lane involved; the warp flushes the slots
          // at a group boundary once some lane runs full.
          auto scan16 = [&](const uint32_t (&vv)[16], uint32_t j, uint32_t col0) {
            if (j != rj) {
              rt = sh.rowthr[(et + j) % kMmaRowRing][row];
              rid = sh.rowid[(et + j) % kMmaRowRing][row];
              rj = j;
            }
            float vf[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) vf[i] = __uint_as_float(vv[i]);
            const float t0 = fmax3(vf[0], vf[1], vf[2]), t1 = fmax3(vf[3], vf[4], vf[5]), t2 = fmax3(vf[6], vf[7], vf[8]);
            const float t3 = fmax3(vf[9], vf[10], vf[11]), t4 = fmax3(vf[12], vf[13], vf[14]);
            const float m = fmaxf(fmax3(t0, t1, t2), fmax3(t3, t4, vf[15]));
            if (m >= rt) {
#ifdef HS_MMA_PROF
              const long long _r0 = clock64();
#endif
#if HS_MMA_EVENTS
              // the 16 accumulators, untested, into the warp's event queue (resolve_events_kernel finds the columns)
              const uint32_t slot = atomicAdd(evcount, 1u);
              const uint4 meta = make_uint4(qbase + col0, rid, __float_as_uint(rt), min(ng - col0, 16u) | (it.table << 8));
              uint4 *dst;
              if (slot < (uint32_t)kMmaEvCap) {
                dst = evq + slot * 5u;
              } else {  // queue full (dense hit regions): straight to the global list
                const unsigned long long gi = atomicAdd(a.ev_count, 1ull);
                dst = gi < a.ev_cap ? reinterpret_cast<uint4 *>(a.events + gi) : nullptr;
              }
              if (dst) {
                dst[0] = make_uint4(vv[0], vv[1], vv[2], vv[3]);
                dst[1] = make_uint4(vv[4], vv[5], vv[6], vv[7]);
                dst[2] = make_uint4(vv[8], vv[9], vv[10], vv[11]);
                dst[3] = make_uint4(vv[12], vv[13], vv[14], vv[15]);
                dst[4] = meta;
              }
#else
              uint32_t pm = 0u;
#pragma unroll
              for (int c = 0; c < 16; ++c) pm |= (vf[c] >= rt ? 1u : 0u) << c;
              const uint32_t ncol = ng - col0;  // valid columns of this chunk (>= 1)
              if (ncol < 16u) pm &= (1u << ncol) - 1u;
#ifndef HS_MMA_NO_RARE
              while (pm) {
                const uint32_t c = (uint32_t)__ffs((int)pm) - 1u;
                pm &= pm - 1u;
                if (lc < (uint32_t)kMmaRing) {
                  ring[lc][lane] = make_uint2(qbase + col0 + c, rid);
                  ++lc;
                } else {  // slots full (dense hit regions): straight to the global list
                  mma_emit_global(a.surv, a.surv_cap, a.surv_count, qbase + col0 + c, it.table, rid);
                }
              }
#else
              if (pm == 0xdeadbeefu) lc = pm;   // (timing experiment: survivors are dropped)
#endif
#endif
#ifdef HS_MMA_PROF
              e_rare += (unsigned long long)(clock64() - _r0);
#endif
            }
          };
          // The warp's chunks, tile block by tile block: the load of the next chunk is in flight
          // while the current one is scanned (two register buffers); the stage is released as soon
          // as the warp's last load has landed.
          auto release = [&]() {
            tc_before();
            mbar_arrive_warp(smem_addr(&sh.t_empty[as]), lane);
          };
          auto advance = [&](uint32_t &j, uint32_t &c) {
            c += 4u;
            if (c >= nchunks) {
              c = (uint32_t)sub;
              ++j;
            }
          };
#ifdef HS_MMA_PROF
          const bool do_scan = !(a.debug & 1u);
#else
          const bool do_scan = true;
#endif
#if defined(HS_DIAG_NOEPI)
          release();   // (timing experiment: the accumulators are never read)
#else
          uint32_t va[16], vb[16];
          uint32_t j = 0, c = (uint32_t)sub;
          if (c < nchunks) {
            tmem_ld16_issue(taddr + j * blockw + c * 16u, va);
            for (;;) {
              tmem_ld_wait();
              uint32_t j1 = j, c1 = c;
              advance(j1, c1);
              if (j1 < nbatch) tmem_ld16_issue(taddr + j1 * blockw + c1 * 16u, vb);
              else release();
              if (do_scan) scan16(va, j, c * 16u);
              if (j1 >= nbatch) break;
              tmem_ld_wait();
              uint32_t j2 = j1, c2 = c1;
              advance(j2, c2);
              if (j2 < nbatch) tmem_ld16_issue(taddr + j2 * blockw + c2 * 16u, va);
              else release();
              if (do_scan) scan16(vb, j1, c1 * 16u);
              if (j2 >= nbatch) break;
              j = j2;
              c = c2;
            }
          } else {
            release();  // no chunk for this warp in this group: still release the stage
          }
#endif
#if HS_MMA_EVENTS
          __syncwarp();
          const uint32_t wc = *(volatile uint32_t *)evcount;  // warp-uniform
          if (wc >= (uint32_t)kMmaEvFlushAt) {
            mma_flush_events(evq, min(wc, (uint32_t)kMmaEvCap), reinterpret_cast<uint4 *>(a.events), a.ev_cap, a.ev_count, lane);
            if (lane == 0) *(volatile uint32_t *)evcount = 0u;
            __syncwarp();
          }
#else
          if (__any_sync(0xffffffffu, lc >= (uint32_t)kMmaRingFlush))
            lc = mma_flush(ring, lc, it.table, a.surv, a.surv_cap, a.surv_count, lane);
#endif
        }
        t += nbatch;
        et += nbatch;
      }
#if !HS_MMA_EVENTS
      // the slots do not carry the table: emptied before the next unit (which may belong to another)
      lc = mma_flush(ring, lc, it.table, a.surv, a.surv_cap, a.surv_count, lane);
#endif
    }
#if HS_MMA_EVENTS
    __syncwarp();
    {
      const uint32_t wc = min(*(volatile uint32_t *)evcount, (uint32_t)kMmaEvCap);
      if (wc) mma_flush_events(evq, wc, reinterpret_cast<uint4 *>(a.events), a.ev_cap, a.ev_count, lane);
    }
#endif
#ifdef HS_MMA_PROF
    if (tid == 0) {
      PROF_OUT(8, (unsigned long long)(clock64() - e_start));
      PROF_OUT(9, e_wait_t);
      PROF_OUT(11, e_rare);
      PROF_OUT(14, e_unit);
    }
#endif
  }

  tc_before();
  __syncthreads();
  if (warp == kMmaIssueWarp) {
    tc_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kMmaTmemCols) : "memory");
  }
}

// One event per lane: the column test the filter's epilogue skipped (value >= row threshold on
// the same FP32 values, so the survivors are exactly those of the in-kernel test), survivors
// appended with one atomic per warp.
__global__ void __launch_bounds__(256)
resolve_events_kernel(const MmaEvent *__restrict__ events, const unsigned long long *__restrict__ ev_count,
                      unsigned long long ev_cap, Survivor *__restrict__ surv, unsigned long long surv_cap,
                      unsigned long long *__restrict__ surv_count) {
  const unsigned long long n = min(*ev_count, ev_cap);
  const int lane = threadIdx.x & 31;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i0 = (unsigned long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += stride) {
    const unsigned long long i = i0 + lane;
    uint32_t pm = 0u, qidx0 = 0u, rid = 0u, table = 0u;
    if (i < n) {
      const uint4 *e = reinterpret_cast<const uint4 *>(events + i);
      const uint4 w0 = __ldg(e), w1 = __ldg(e + 1), w2 = __ldg(e + 2), w3 = __ldg(e + 3), meta = __ldg(e + 4);
      const float rt = __uint_as_float(meta.z);
      const uint32_t v[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
      for (int c = 0; c < 16; ++c) pm |= (__uint_as_float(v[c]) >= rt ? 1u : 0u) << c;
      const uint32_t ncol = meta.w & 0xffu;
      if (ncol < 16u) pm &= (1u << ncol) - 1u;
      qidx0 = meta.x;
      rid = meta.y;
      table = meta.w >> 8;
    }
    const uint32_t cnt = (uint32_t)__popc(pm);
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += x;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    if (total == 0u) continue;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(surv_count, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 0) + (inc - cnt);
    while (pm) {
      const uint32_t c = (uint32_t)__ffs((int)pm) - 1u;
      pm &= pm - 1u;
      if (base < surv_cap) {
        Survivor sv;
        sv.query = qidx0 + c;  // index into the query list; the exact stage resolves it
        sv.table = table;
        sv.pos = rid;          // fragment id
        sv.pad = 3;
        surv[base] = sv;
      }
      ++base;
    }
  }
}

// ---- query rows ---------------------------------------------------------------------------
// qb16[q][k]: k < 8*len: fp16(q_k); k = 8*len, 8*len+1: c_q = -|q|^2/2 + margin as hi + lo;
// the rest 0.  A query the FP16 path cannot represent gets c_q = +60000: every member of
// its buckets survives the filter and the exact stage decides.
// Returns false when the FP16 path cannot represent the query (the caller then zeroes the
// coordinates: D = c_q = +60000 passes every row).
__device__ __forceinline__ bool write_cq(__half *row, int dim, int kp, double nq, double qmaxabs, double beta) {
  // margin: beta/2 * |q|^2 (products + accumulation), the hi/lo split and the accumulation of c itself
  double c = -0.5 * nq + 0.5 * beta * nq + (double)kp * 4.8e-7 * (0.5 * nq) + 2e-6 * (0.5 * nq) + 1e-3;
  const bool ok = (qmaxabs <= 3.0e4) && (nq <= 1.0e5) && (c == c);
  if (!ok) c = 60000.0;
  const __half hi = __double2half(c);
  const double rest = c - (double)__half2float(hi);
  row[dim] = hi;
  row[dim + 1] = __double2half(rest);  // its rounding (<= 2^-22 |c|) is inside the 2e-6 term
  for (int k = dim + 2; k < kp; ++k) row[k] = __float2half(0.f);
  return ok;
}

__global__ void build_qb_points_kernel(const double *__restrict__ q64, uint32_t Q, int dim, int kp, double beta,
                                       __half *__restrict__ qb16) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const double *src = q64 + (size_t)q * dim;
  __half *row = qb16 + (size_t)q * kp;
  double nq = 0.0, mx = 0.0;
  for (int k = 0; k < dim; ++k) {
    const double v = src[k];
    nq += v * v;
    mx = fmax(mx, fabs(v));
  }
  const bool ok = write_cq(row, dim, kp, nq, mx, beta);
  for (int k = 0; k < dim; ++k) row[k] = __double2half(ok ? src[k] : 0.0);
}

// queries = DB fragments q0 .. q0+nq-1 (all pairs)
__global__ void build_qb_codes_kernel(const uint8_t *__restrict__ codes, uint64_t q0, uint32_t nq, int len, int kp,
                                      const double *__restrict__ table64, double beta, __half *__restrict__ qb16) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const uint8_t *c = codes + (q0 + q) * (uint64_t)len;
  __half *row = qb16 + (size_t)q * kp;
  double n2 = 0.0, mx = 0.0;
  for (int p = 0; p < len; ++p)
    for (int j = 0; j < HS_CDIM; ++j) {
      const double v = table64[(int)c[p] * HS_CDIM + j];
      n2 += v * v;
      mx = fmax(mx, fabs(v));
      row[p * HS_CDIM + j] = __double2half(v);
    }
  write_cq(row, len * HS_CDIM, kp, n2, mx, beta);
}

// queries = the members at positions pos0 .. pos0+nq-1 of a bucket-ordered, position-major code
// store (cluster self-join of a large bucket)
__global__ void build_qb_store_kernel(const uint8_t *__restrict__ store, uint64_t npad, uint32_t pos0, uint32_t nq,
                                      int len, int kp, const double *__restrict__ table64, double beta,
                                      __half *__restrict__ qb16) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  __half *row = qb16 + (size_t)q * kp;
  double n2 = 0.0, mx = 0.0;
  for (int p = 0; p < len; ++p) {
    const int c = min((int)store[(uint64_t)p * npad + pos0 + q] / kCodeScale, HS_AA - 1);
    for (int j = 0; j < HS_CDIM; ++j) {
      const double v = table64[c * HS_CDIM + j];
      n2 += v * v;
      mx = fmax(mx, fabs(v));
      row[p * HS_CDIM + j] = __double2half(v);
    }
  }
  write_cq(row, len * HS_CDIM, kp, n2, mx, beta);
}

// ---- host side ----------------------------------------------------------------------------
static inline double mma_beta(int kp) {
  // |x~q~ - xq| <= (2^-10 + 2^-22)|x||q| per product (both operands rounded to FP16), FP32
  // accumulation of kp terms <= kp * 2^-21 * sum|x~q~|; sum|x||q| <= (|x|^2 + |q|^2)/2
  return (ldexp(1.0, -10) + ldexp(1.0, -22)) * 1.001 + (double)kp * ldexp(1.0, -21) * 1.01;
}

int mma_geometry(const hs_ctx *ctx, MmaGeometry *g) {
  const int len = (int)ctx->prm.len;
  g->kp = (8 * len + 2 + 15) & ~15;
  const size_t a_stage = (size_t)(g->kp >> 3) * kMmaAGroupBytes;
  const size_t budget = 168 * 1024;  // operands; the code ring (<= 16 KB) and ~38 KB static come on top
  int S = 3;
  if ((size_t)S * a_stage > budget / 2) S = 2;
  if ((size_t)S * a_stage > budget - 32 * 1024) return HS_ERR_UNSUPPORTED;   // (len 30: two 64 KB stages, 80 queries)
  size_t rows = (budget - (size_t)S * a_stage) / ((size_t)(g->kp >> 3) * 16);
  int qmax = (int)std::min<size_t>(512, rows) & ~15;
  if (qmax < 16) return HS_ERR_UNSUPPORTED;
  g->nstages = S;
  g->qmax = qmax;
  g->cring = len <= 16 ? kMmaMaxCodeRing : 4;
  g->smem = (size_t)(g->kp >> 3) * qmax * 16 + (size_t)S * a_stage + (size_t)g->cring * len * 128;
  g->beta = mma_beta(g->kp);
  return HS_OK;
}

// The tensor path needs every table entry representable in FP16 without overflow.
bool mma_filter_usable(const hs_ctx *ctx) {
  if (!ctx->have_ftable) return false;   // (integer metric: no contracting embedding could be built)
  if (ctx->prm.metric == HS_METRIC_BLOSUM_INT && ctx->no_mma_int) return false;
  if (ctx->prm.flags & HS_FLAG_SCALAR_FILTER) return false;
  if (ctx->no_mma_filter) return false;
  MmaGeometry g;
  if (mma_geometry(ctx, &g) != HS_OK) return false;
  for (int i = 0; i < HS_AA * HS_CDIM; ++i)
    if (!(fabs(ctx->ftable64[i]) <= 1.0e3)) return false;
  return true;
}

// Upload the FP16 embedding rows and the (rounded-down) squared row norms.
int mma_upload_tables(hs_ctx *ctx) {
  __half h[HS_AA * HS_CDIM];
  float nx[HS_AA];
  for (int c = 0; c < HS_AA; ++c) {
    double s = 0.0;
    for (int j = 0; j < HS_CDIM; ++j) {
      const double v = ctx->have_ftable ? ctx->ftable64[c * HS_CDIM + j] : 0.0;
      h[c * HS_CDIM + j] = __double2half(v);
      s += v * v;
    }
    float f = (float)s;
    if ((double)f > s) f = nextafterf(f, -INFINITY);
    nx[c] = f;
  }
  HS_TRY(ctx->d_tab16.reserve(sizeof h + sizeof nx));
  HS_CUDA(cudaMemcpyAsync(ctx->d_tab16.p, h, sizeof h, cudaMemcpyHostToDevice, ctx->stream));
  HS_CUDA(cudaMemcpyAsync((char *)ctx->d_tab16.p + sizeof h, nx, sizeof nx, cudaMemcpyHostToDevice, ctx->stream));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  return HS_OK;
}

int launch_build_qb_points(hs_ctx *ctx, const double *d_q64, uint32_t Q, void *d_qb16) {
  if (Q == 0) return HS_OK;
  MmaGeometry g;
  HS_TRY(mma_geometry(ctx, &g));
  build_qb_points_kernel<<<(Q + 127) / 128, 128, 0, ctx->stream>>>(d_q64, Q, (int)ctx->dim, g.kp, g.beta,
                                                                   reinterpret_cast<__half *>(d_qb16));
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches++;
  return HS_OK;
}

int launch_build_qb_codes(hs_ctx *ctx, uint64_t q0, uint32_t nq, void *d_qb16) {
  if (nq == 0) return HS_OK;
  MmaGeometry g;
  HS_TRY(mma_geometry(ctx, &g));
  build_qb_codes_kernel<<<(nq + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_codes.as<uint8_t>(), q0, nq,
                                                                   (int)ctx->prm.len, g.kp, ctx->d_ftable64.as<double>(),
                                                                   g.beta, reinterpret_cast<__half *>(d_qb16));
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches++;
  return HS_OK;
}

// queries = nq residue strings at d_qcodes (integer metric: the embedded rows come from the filter table)
int launch_build_qb_qcodes(hs_ctx *ctx, const uint8_t *d_qcodes, uint32_t nq, void *d_qb16) {
  if (nq == 0) return HS_OK;
  MmaGeometry g;
  HS_TRY(mma_geometry(ctx, &g));
  build_qb_codes_kernel<<<(nq + 127) / 128, 128, 0, ctx->stream>>>(d_qcodes, 0, nq, (int)ctx->prm.len, g.kp,
                                                                   ctx->d_ftable64.as<double>(), g.beta,
                                                                   reinterpret_cast<__half *>(d_qb16));
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches++;
  return HS_OK;
}

int launch_build_qb_store(hs_ctx *ctx, uint32_t table, uint32_t pos0, uint32_t nq, void *d_qb16) {
  if (nq == 0) return HS_OK;
  MmaGeometry g;
  HS_TRY(mma_geometry(ctx, &g));
  build_qb_store_kernel<<<(nq + 127) / 128, 128, 0, ctx->stream>>>(
      ctx->tables[table].codes_sorted.as<uint8_t>(), ctx->npad, pos0, nq, (int)ctx->prm.len, g.kp,
      ctx->d_ftable64.as<double>(), g.beta, reinterpret_cast<__half *>(d_qb16));
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches++;
  return HS_OK;
}

// items / units / qlist are device arrays; CTAs take units dynamically from *d_unit_counter.
int launch_filter_mma(hs_ctx *ctx, const FilterArgs &fa, const void *d_items, const void *d_units, uint32_t nunits,
                      uint32_t *d_unit_counter, uint32_t grid, const void *d_qb16, int mode) {
  if (grid == 0 || nunits == 0) return HS_OK;
  MmaGeometry g;
  HS_TRY(mma_geometry(ctx, &g));
  MmaArgs a;
  memset(&a, 0, sizeof a);
  a.items = reinterpret_cast<const MmaItem *>(d_items);
  a.units = reinterpret_cast<const MmaUnit *>(d_units);
  a.nunits = nunits;
  a.unit_counter = d_unit_counter;
  HS_CUDA(cudaMemsetAsync(d_unit_counter, 0, sizeof(uint32_t), ctx->stream));
  a.qlist = fa.qlist;
  a.qb16 = reinterpret_cast<const __half *>(d_qb16);
  a.tq_base = fa.tq_base;
  a.stores = fa.stores;
  a.sorted_ids = dev_sorted_ids(ctx);
  a.npad = fa.npad;
  a.len = fa.len;
  a.kp = g.kp;
  a.nstages = g.nstages;
  a.qmax = g.qmax;
  a.cring = g.cring;
  // a reference hit has d2 <= R^2 (1 + 1e-12) in exact arithmetic; with the integer metric a hit has
  // integer distance <= (int)R, hence embedded squared distance <= (int)R (contracting embedding)
  const double rr = ctx->prm.metric == HS_METRIC_BLOSUM_INT ? (double)(int)ctx->prm.R : ctx->prm.R * ctx->prm.R;
  const double r2 = rr * (1.0 + 1e-12) + 1e-30;
  float thr = (float)r2;
  if ((double)thr < r2) thr = nextafterf(thr, INFINITY);
  a.thr = thr;
  a.beta = (float)(g.beta * 1.0001);
#ifdef HS_MMA_PROF
  HS_TRY(ctx->d_misc.reserve(sizeof(unsigned long long) * 24));
  HS_CUDA(cudaMemsetAsync(ctx->d_misc.p, 0, sizeof(unsigned long long) * 24, ctx->stream));
  a.prof = ctx->d_misc.as<unsigned long long>();
  if (const char *e = getenv("HS_MMA_DEBUG")) a.debug = (uint32_t)atoi(e);
#endif
  a.tab16 = ctx->d_tab16.as<uint4>();
  a.nx32 = reinterpret_cast<const float *>((const char *)ctx->d_tab16.p + sizeof(__half) * HS_AA * HS_CDIM);
  a.surv = fa.surv;
  a.surv_cap = fa.surv_cap;
  a.surv_count = fa.surv_count;
#if HS_MMA_EVENTS
  // event list: as many entries as the survivor list (an event holds at least one survivor, apart from
  // rows whose only passing columns are padding); run_filter grows both together on overflow
  HS_TRY(ctx->d_events.reserve(sizeof(MmaEvent) * (size_t)fa.surv_cap));
  a.events = ctx->d_events.as<MmaEvent>();
  a.ev_cap = fa.surv_cap;
  a.ev_count = ctx->d_counters.as<unsigned long long>() + 15;
  HS_CUDA(cudaMemsetAsync(a.ev_count, 0, sizeof(unsigned long long), ctx->stream));
#endif
  (void)mode;
#define HS_MMA(LB)                                                                                            \
  do {                                                                                                        \
    HS_CUDA(cudaFuncSetAttribute(filter_mma_kernel<LB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem)); \
    filter_mma_kernel<LB><<<grid, kMmaThreads, g.smem, ctx->stream>>>(a);                                     \
  } while (0)
  if (a.len <= 8) HS_MMA(8);
  else if (a.len <= 10) HS_MMA(10);
  else if (a.len <= 12) HS_MMA(12);
  else if (a.len <= 16) HS_MMA(16);
  else if (a.len <= 20) HS_MMA(20);
  else if (a.len <= 25) HS_MMA(25);
  else HS_MMA(32);
#undef HS_MMA
#if HS_MMA_EVENTS
  resolve_events_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(a.events, a.ev_count, a.ev_cap, a.surv, a.surv_cap, a.surv_count);
  ctx->stats.kernel_launches++;
#endif
#ifdef HS_MMA_PROF
  {
    unsigned long long h[24];
    cudaMemcpyAsync(h, a.prof, sizeof h, cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    const double g = (double)grid;
    fprintf(stderr,
            "[mma prof] grid %u units %u | producer: total %.0f wait_codes %.0f wait_a_empty %.0f bload %.0f | "
            "issuer: wait_a_full %.0f wait_t_empty %.0f wait_b %.0f groups %.0f | epilogue(warp0): total %.0f "
            "wait_t_full %.0f issuer_unit %.0f rare %.0f issuer_mma %.0f issuer_commit %.0f unit_fetch %.0f issuer_total %.0f | producer build %.0f fence+arrive %.0f (cycles per CTA)\n",
            grid, nunits, h[0] / g, h[1] / g, h[2] / g, h[3] / g, h[4] / g, h[5] / g, h[6] / g, h[7] / g, h[8] / g,
            h[9] / g, h[10] / g, h[11] / g, h[12] / g, h[13] / g, h[14] / g, h[15] / g, h[16] / g, h[17] / g);
  }
#endif
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches++;
  return HS_OK;
}

}  // namespace hs
