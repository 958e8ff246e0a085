// Tensor-core candidate filter (tcgen05 / TMEM), the bulk path of K3/K4.
//
// The filter bound of a (query q, member m) pair is
//     s(q, m) = sum_pos Tq[q][pos][code_m[pos]]
// (verify.cu).  Over one bucket this is a dense contraction: with the member
// written as a one-hot vector x_m in {0,1}^(20*len) (x_m[20*pos + code] = 1) and
// the query's table flattened to t_q in R^(20*len),  s(q, m) = <x_m, t_q>.  ncu
// shows the scalar filter bound by shared-memory lookups (LSU pipe 76 %, DRAM
// 0.6 %; profiles/r01a_ncu_full.json), i.e. compute-bound on the wrong pipe, so
// buckets probed by many queries go through the 5th-generation tensor cores:
//
//   A  [128 members][Kp]  one-hot FP16, built in shared memory from the 1-byte
//                         residue codes of the bucket-ordered store,
//   B  [N queries][Kp]    FP16 tables rounded DOWN (so the bound never grows),
//   D  [128][N]           FP32 accumulator in TMEM (tcgen05.mma kind::f16),
//
// then every accumulator is compared with the threshold straight out of TMEM
// (tcgen05.ld) and the rare survivors are appended to the list the exact FP64
// stage consumes.  Products 1.0 * t are exact; the only error is the FP32
// accumulation, covered by the 2^-16 relative margin in filter_tc_threshold().
// The filter still never decides a hit.
//
// Shared-memory operand layout (both operands K-major, no swizzle): 8x8 FP16
// core matrices of 128 contiguous bytes (row r at byte 16*r); core matrices of
// one 8-column group are stacked over the row groups (stride 128 B = SBO), the
// column groups follow each other at 128*16 = 2048 B (= LBO).  One tcgen05.mma
// consumes 16 columns = two column groups.
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "internal.cuh"
#include "verify.cuh"

namespace hs {

constexpr int kTcThreads = 256;
constexpr int kTcM = 128;                 // members per tile (UMMA M)
constexpr int kTcColGroupBytes = 2048;    // 16 row groups * 128 B
constexpr uint32_t kTcTmemCols = 128;     // FP32 accumulator columns (N <= 128)
constexpr int kTcStage = 192;             // survivors staged per tile before the global flush

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (sm_100: version 1).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= 1ull << 46;
  return d;
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
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
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- FP16 query tables -------------------------------------------------------------
// tq16[q][k] = round_down_fp16(tq[q][k]) for k < 20*len, 0 for the padding columns.
__global__ void tq_to_half_kernel(const float *__restrict__ tq, uint64_t nq, int klen, int kp,
                                  __half *__restrict__ tq16) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq * (uint64_t)kp) return;
  const uint64_t q = i / kp;
  const int k = (int)(i - q * kp);
  tq16[i] = k < klen ? __float2half_rd(tq[q * klen + k]) : __float2half(0.f);
}

struct TcArgs {
  const WorkItem *items;  // block_begin counts blocks of this kernel
  uint32_t nitems;
  const uint32_t *qlist;
  const __half *tq16;     // [Q][kp]
  uint32_t tq_base;
  const uint8_t *const *stores;
  uint64_t npad;
  int len, kp, kc;        // kc: A columns resident at once (multiple of 16)
  uint32_t tiles_per_block;
  uint32_t lbo, sbo;
  float thr;
  Survivor *surv;
  unsigned long long surv_cap;
  unsigned long long *surv_count;
};

template <int MODE>
__global__ void __launch_bounds__(kTcThreads)
filter_tc_kernel(TcArgs a) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  __shared__ WorkItem s_item;
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) Survivor s_stage[kTcStage];
  __shared__ uint32_t s_nstage, s_nflush;
  __shared__ unsigned long long s_flush_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    uint32_t lo = 0, hi = a.nitems;  // last item with block_begin <= blockIdx.x
    while (hi - lo > 1) {
      const uint32_t m = (lo + hi) >> 1;
      if (a.items[m].block_begin <= blockIdx.x) lo = m; else hi = m;
    }
    s_item = a.items[lo];
    s_nstage = 0;
    mbar_init(smem_u32(&s_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                 "r"(kTcTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  const WorkItem it = s_item;
  const int len = a.len, kp = a.kp, kc = a.kc;
  const uint32_t nq = it.q_end - it.q_begin;   // <= 128
  const uint32_t npadq = (nq + 15u) & ~15u;    // UMMA N
  const int ngroups = kp >> 3;                 // 8-column groups of B

  unsigned char *sB = tc_smem;                                  // ngroups * 2048 B
  unsigned char *sA = tc_smem + (size_t)ngroups * kTcColGroupBytes;  // (kc/8) * 2048 B

  // ---- B: the item's query tables, q fastest so that the smem stores are contiguous
  {
    const int total = ngroups * 128;
    for (int i = tid; i < total; i += kTcThreads) {
      const int g = i >> 7, q = i & 127;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if ((uint32_t)q < nq) {
        const uint32_t row = a.qlist[it.q_begin + q] - a.tq_base;
        v = *reinterpret_cast<const uint4 *>(a.tq16 + (size_t)row * kp + (g << 3));
      }
      *reinterpret_cast<uint4 *>(sB + (size_t)g * kTcColGroupBytes + (q >> 3) * 128 + (q & 7) * 16) = v;
    }
  }

  const uint32_t idesc = (1u << 4) | ((npadq >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);  // F16 x F16 -> F32, K-major
  const uint32_t bar = smem_u32(&s_bar);
  const uint32_t sA_u32 = smem_u32(sA), sB_u32 = smem_u32(sB);
  const uint8_t *store = a.stores[it.table];
  const uint32_t chunk = blockIdx.x - it.block_begin;
  const uint32_t first = it.m_begin + chunk * a.tiles_per_block * kTcM;
  uint32_t parity = 0;

  // thread (r, h): member row r = tid & 127, column-group parity h = tid >> 7
  const int r = tid & 127, h = tid >> 7;
  const uint32_t row_off = (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
  uint8_t code[HS_MAX_LEN];
  auto load_codes = [&](uint32_t pos0) {
    const uint32_t mypos = pos0 + (uint32_t)r;
    const bool ok = pos0 < it.m_end && mypos < it.m_end;
#pragma unroll
    for (int p = 0; p < HS_MAX_LEN; ++p)
      if (p < len) code[p] = ok ? __ldg(store + (uint64_t)p * a.npad + mypos) : (uint8_t)0xff;
  };
  load_codes(first);

  for (uint32_t t = 0; t < a.tiles_per_block; ++t) {
    const uint32_t pos0 = first + t * kTcM;
    if (pos0 >= it.m_end) break;
    const bool mvalid = pos0 + (uint32_t)r < it.m_end;

    for (int k0 = 0; k0 < kp; k0 += kc) {
      const int kcur = min(kc, kp - k0);  // multiple of 16
      const int ng = kcur >> 3;
      // zero my 16-byte pieces (row r, column groups of parity h), then set the ones
      for (int g = h; g < ng; g += 2)
        *reinterpret_cast<uint4 *>(sA + (size_t)g * kTcColGroupBytes + row_off) = make_uint4(0u, 0u, 0u, 0u);
      if (mvalid) {
#pragma unroll
        for (int p = 0; p < HS_MAX_LEN; ++p) {
          if (p < len) {
            const int kk = p * HS_AA + (int)(code[p] / kCodeScale) - k0;
            if (kk >= 0 && kk < kcur && ((kk >> 3) & 1) == h)
              *reinterpret_cast<unsigned short *>(sA + (size_t)(kk >> 3) * kTcColGroupBytes + row_off + (kk & 7) * 2) =
                  (unsigned short)0x3C00;  // 1.0
          }
        }
      }
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      // the whole of warp 0 takes the issuer role so that no lane of the issuing
      // warp sits in the mbarrier wait while lane 0 is still issuing
      if (warp == 0) {
        if (lane == 0) {
          tc_fence_after();
          const int nsteps = kcur >> 4;
          for (int s = 0; s < nsteps; ++s) {
            const uint64_t ad = umma_desc(sA_u32 + (uint32_t)s * 2u * kTcColGroupBytes, a.lbo, a.sbo);
            const uint64_t bd = umma_desc(sB_u32 + (uint32_t)((k0 >> 4) + s) * 2u * kTcColGroupBytes, a.lbo, a.sbo);
            umma_f16(tmem_base, ad, bd, idesc, (k0 > 0 || s > 0) ? 1u : 0u);
          }
          umma_commit(bar);
        }
        __syncwarp();
      }
      if (k0 + kc >= kp) load_codes(pos0 + kTcM);  // next tile's codes travel while the MMAs run
      mbar_wait(bar, parity);
      parity ^= 1u;
    }
    tc_fence_after();

    // ---- epilogue: my TMEM lane is member row 32*(warp&3)+lane; column groups of 16.
    // Survivors are staged in shared memory and flushed with one global atomic per tile.
    {
      const int erow = (warp & 3) * 32 + lane;
      const uint32_t epos = pos0 + (uint32_t)erow;
      const bool evalid = epos < it.m_end;
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
      const int ncg = (int)(npadq >> 4);
      for (int cg = warp >> 2; cg < ncg; cg += 2) {
        float v[16];
        tmem_ld16(taddr + (uint32_t)cg * 16u, v);
        float m = v[0];
#pragma unroll
        for (int j = 1; j < 16; ++j) m = fminf(m, v[j]);
        if (evalid && m <= a.thr) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t qi = (uint32_t)cg * 16u + (uint32_t)j;
            if (qi < nq && v[j] <= a.thr) {
              const uint32_t qid = a.qlist[it.q_begin + qi];
              if (MODE == kModeSearch || qid < epos) {
                Survivor sv;
                sv.query = qid;
                sv.table = it.table;
                sv.pos = epos;
                sv.pad = 0;
                const uint32_t slot = atomicAdd(&s_nstage, 1u);
                if (slot < (uint32_t)kTcStage) {
                  s_stage[slot] = sv;
                } else {  // staging full: straight to the global list
                  const unsigned long long idx = atomicAdd(a.surv_count, 1ull);
                  if (idx < a.surv_cap) a.surv[idx] = sv;
                }
              }
            }
          }
        }
      }
    }
    tc_fence_before();  // the next tile's first MMA overwrites D: order the loads before it
    __syncthreads();
    if (tid == 0) {
      const uint32_t n = min(s_nstage, (uint32_t)kTcStage);
      s_nflush = n;
      s_nstage = 0;
      s_flush_base = n ? atomicAdd(a.surv_count, (unsigned long long)n) : 0ull;
    }
    __syncthreads();
    {
      const uint32_t n = s_nflush;
      const unsigned long long base = s_flush_base;
      for (uint32_t i = tid; i < n; i += kTcThreads)
        if (base + i < a.surv_cap) a.surv[base + i] = s_stage[i];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTcTmemCols) : "memory");
  }
}

// ---- host side ------------------------------------------------------------------------
float filter_tc_threshold(const hs_ctx *ctx) {
  const double t = (double)filter_threshold(ctx) * (1.0 + ldexp(1.0, -16));
  float f = (float)t;
  if ((double)f < t) f = nextafterf(f, INFINITY);
  return f;
}

// The largest number of A columns (multiple of 16) resident beside the full B.
static int tc_geometry(int len, int *kp_out, int *kc_out, size_t *smem_out) {
  const int kp = (len * HS_AA + 15) & ~15;
  const size_t bbytes = (size_t)(kp >> 3) * kTcColGroupBytes;
  const size_t two_per_sm = 108 * 1024, one_per_sm = 216 * 1024;
  int kc = kp;
  if (bbytes + (size_t)(kp >> 3) * kTcColGroupBytes > two_per_sm) {
    if (bbytes + 2 * kTcColGroupBytes > one_per_sm) return HS_ERR_UNSUPPORTED;
    kc = (int)(((one_per_sm - bbytes) / kTcColGroupBytes) * 8) & ~15;
    if (kc > kp) kc = kp;
  }
  *kp_out = kp;
  *kc_out = kc;
  *smem_out = bbytes + (size_t)(kc >> 3) * kTcColGroupBytes;
  return HS_OK;
}

int tc_min_queries() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("HS_TC_MIN_QUERIES");
    v = e ? atoi(e) : 8;
    if (v < 1) v = 1;
  }
  return v;
}

int launch_tq_to_half(hs_ctx *ctx, const float *d_tq, uint64_t nq, void *d_tq16_v) {
  __half *d_tq16 = reinterpret_cast<__half *>(d_tq16_v);
  int kp, kc;
  size_t smem;
  HS_TRY(tc_geometry((int)ctx->prm.len, &kp, &kc, &smem));
  const uint64_t n = nq * (uint64_t)kp;
  if (n == 0) return HS_OK;
  tq_to_half_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_tq, nq, (int)ctx->prm.len * HS_AA, kp,
                                                                        d_tq16);
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches++;
  return HS_OK;
}

uint32_t tc_padded_k(uint32_t len) { return (len * HS_AA + 15u) & ~15u; }

// items[].block_begin must count blocks of `tiles_per_block` 128-member tiles.
int launch_filter_tc(hs_ctx *ctx, const FilterArgs &fa, const void *d_tq16_v, uint32_t tiles_per_block,
                     uint32_t nblocks, int mode) {
  const __half *d_tq16 = reinterpret_cast<const __half *>(d_tq16_v);
  if (nblocks == 0) return HS_OK;
  TcArgs a;
  memset(&a, 0, sizeof a);
  size_t smem;
  HS_TRY(tc_geometry(fa.len, &a.kp, &a.kc, &smem));
  a.items = fa.items;
  a.nitems = fa.nitems;
  a.qlist = fa.qlist;
  a.tq16 = d_tq16;
  a.tq_base = fa.tq_base;
  a.stores = fa.stores;
  a.npad = fa.npad;
  a.len = fa.len;
  a.tiles_per_block = tiles_per_block;
  a.lbo = kTcColGroupBytes;
  a.sbo = 128;
  if (getenv("HS_TC_SWAP_LBO_SBO")) {  // bring-up switch
    a.lbo = 128;
    a.sbo = kTcColGroupBytes;
  }
  a.thr = filter_tc_threshold(ctx);
  a.surv = fa.surv;
  a.surv_cap = fa.surv_cap;
  a.surv_count = fa.surv_count;
  if (mode == kModeSearch || mode == kModeBrute) {
    auto kern = filter_tc_kernel<kModeSearch>;
    HS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<nblocks, kTcThreads, smem, ctx->stream>>>(a);
  } else {
    auto kern = filter_tc_kernel<kModeAllPairs>;
    HS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<nblocks, kTcThreads, smem, ctx->stream>>>(a);
  }
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches++;
  return HS_OK;
}

}  // namespace hs
