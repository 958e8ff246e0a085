// K3 (probe + candidate verification), K4 (brute force) and the pair stage of
// K5 (in-bucket near pairs) share one pair engine:
//
//   filter kernel  -- for every (query, member) pair of a work item, a cheap
//                     bound on the distance: `len` shared-memory lookups of a
//                     per-query residue table Tq[pos][code] and FP32 adds
//                     (exact integer arithmetic for the BLOSUM metric).  Pairs
//                     that cannot be within R are dropped here.
//   exact kernel   -- the survivors are re-evaluated exactly as the reference
//                     does (PairwiseDistance_square, motif_both_points.cpp:
//                     176-183: sequential FP64 subtract, multiply, add; or
//                     DistanceScore, evaluate_correlation.cpp:34-41), the
//                     reference's threshold predicate is applied, the
//                     first-table-wins rule of label[] (:232-238) is enforced
//                     by comparing the pair's keys in earlier tables, and hits
//                     (or union-find edges) are emitted.
//
// The filter never decides a hit; it only rejects pairs whose distance is
// provably above the threshold (margin derivation in DESIGN.md), so hit sets
// and distances are bit-exact.
#include <math.h>

#include <algorithm>

#include "verify.cuh"

namespace hs {

// ---- probe: query key -> bucket range (HashTable::find, :228-231) -------------
template <int KW>
__global__ void probe_kernel(const uint64_t *__restrict__ qkeys /* [Q][KW] of this table */,
                             const uint8_t *__restrict__ qvalid, uint32_t Q,
                             const uint64_t *__restrict__ ukeys /* [KW][nb] */, uint64_t nb,
                             const uint32_t *__restrict__ bstart, uint2 *__restrict__ qrange,
                             uint32_t *__restrict__ qrank) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  uint2 r = make_uint2(0u, 0u);
  uint32_t slot = 0xffffffffu;
  if (qvalid[q] && nb > 0) {
    uint64_t k[KW];
#pragma unroll
    for (int w = 0; w < KW; ++w) k[w] = qkeys[(size_t)q * KW + w];
    uint64_t lo = 0, hi = nb;  // lower bound
    while (lo < hi) {
      const uint64_t m = (lo + hi) >> 1;
      bool less = false, decided = false;
#pragma unroll
      for (int w = KW - 1; w >= 0; --w) {
        const uint64_t u = ukeys[(uint64_t)w * nb + m];
        if (!decided && u != k[w]) {
          less = u < k[w];
          decided = true;
        }
      }
      if (less) lo = m + 1; else hi = m;
    }
    if (lo < nb) {
      bool eq = true;
#pragma unroll
      for (int w = 0; w < KW; ++w) eq = eq && (ukeys[(uint64_t)w * nb + lo] == k[w]);
      if (eq) {
        r = make_uint2(bstart[lo], bstart[lo + 1]);
        slot = (uint32_t)lo;
      }
    }
  }
  qrange[q] = r;
  qrank[q] = slot;
}

// Hashed-key path (radix_sort.cu): binary search over the ascending key hashes, then the full key
// of the slot -- the key of its first member -- decides (a query whose key merely shares a
// bucket's hash finds nothing).
template <int KW>
__global__ void probe_hashed_kernel(const uint64_t *__restrict__ qkeys, const uint8_t *__restrict__ qvalid, uint32_t Q,
                                    const uint64_t *__restrict__ uhash /* [nb] */,
                                    const uint64_t *__restrict__ dbkeys /* [KW][N] by fragment id */, uint64_t N,
                                    const uint32_t *__restrict__ ids, uint64_t nb,
                                    const uint32_t *__restrict__ bstart, uint2 *__restrict__ qrange,
                                    uint32_t *__restrict__ qrank) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  uint2 r = make_uint2(0u, 0u);
  uint32_t slot = 0xffffffffu;
  if (qvalid[q] && nb > 0) {
    uint64_t k[KW];
#pragma unroll
    for (int w = 0; w < KW; ++w) k[w] = qkeys[(size_t)q * KW + w];
    const uint64_t h = key_hash<KW>(k);
    uint64_t lo = 0, hi = nb;
    while (lo < hi) {
      const uint64_t m = (lo + hi) >> 1;
      if (uhash[m] < h) lo = m + 1; else hi = m;
    }
    if (lo < nb && uhash[lo] == h) {
      const uint32_t b0 = bstart[lo];
      const uint32_t head = ids[b0];
      bool eq = true;
#pragma unroll
      for (int w = 0; w < KW; ++w) eq = eq && (dbkeys[(uint64_t)w * N + head] == k[w]);
      if (eq) {
        r = make_uint2(b0, bstart[lo + 1]);
        slot = (uint32_t)lo;
      }
    }
  }
  qrange[q] = r;
  qrank[q] = slot;
}

int launch_probe(hs_ctx *ctx, uint32_t table, const uint64_t *d_qkeys, const uint8_t *d_qvalid, uint32_t Q,
                 uint2 *d_qrange, uint32_t *d_qrank) {
  if (Q == 0) return HS_OK;
  const TableIndex &T = ctx->tables[table];
  const unsigned grid = (Q + 127) / 128;
  const uint64_t *qk = d_qkeys + (size_t)table * Q * ctx->key_words;
  const uint8_t *qv = d_qvalid + (size_t)table * Q;
  uint2 *qr = d_qrange + (size_t)table * Q;
  uint32_t *qs = d_qrank + (size_t)table * Q;
  if (T.hashed_keys) {
#define HS_PROBE_H(KWV)                                                                                    \
  probe_hashed_kernel<KWV><<<grid, 128, 0, ctx->stream>>>(qk, qv, Q, T.ukeys.as<uint64_t>(),                \
                                                          ctx->d_keys[table].as<uint64_t>(), ctx->N,         \
                                                          T.sorted_ids.as<uint32_t>(), T.nslots,            \
                                                          T.bstart.as<uint32_t>(), qr, qs)
    switch (ctx->key_words) {
      case 2: HS_PROBE_H(2); break;
      case 3: HS_PROBE_H(3); break;
      default: HS_PROBE_H(4); break;
    }
#undef HS_PROBE_H
    HS_CUDA(cudaGetLastError());
    ctx->stats.kernel_launches++;
    return HS_OK;
  }
#define HS_PROBE(KWV)                                                                                      \
  probe_kernel<KWV><<<grid, 128, 0, ctx->stream>>>(qk, qv, Q, T.ukeys.as<uint64_t>(), T.nslots,            \
                                                   T.bstart.as<uint32_t>(), qr, qs)
  switch (ctx->key_words) {
    case 1: HS_PROBE(1); break;
    case 2: HS_PROBE(2); break;
    case 3: HS_PROBE(3); break;
    default: HS_PROBE(4); break;
  }
#undef HS_PROBE
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches++;
  return HS_OK;
}

// ---- per-query filter tables ----------------------------------------------------
// Tq[q][pos][c] = fl32( sum_j (table[c][j] - q[8*pos+j])^2 ): the squared distance
// contribution of residue c at position pos.
__global__ void build_tq_points_kernel(const double *__restrict__ q64, uint32_t Q, int len,
                                       const double *__restrict__ table64, float *__restrict__ tq) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t n = (uint64_t)Q * len * HS_AA;
  if (i >= n) return;
  const int c = (int)(i % HS_AA);
  const uint64_t qp = i / HS_AA;  // q*len + pos
  const double *qv = q64 + qp * HS_CDIM;
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < HS_CDIM; ++j) {
    const double r = table64[c * HS_CDIM + j] - qv[j];
    s += r * r;
  }
  tq[i] = (float)s;
}
// Integer metric: Tq[q][pos][c] = D[c][qcode[pos]] (exact in FP32).
__global__ void build_tq_int_kernel(const uint8_t *__restrict__ qcodes, uint32_t Q, int len,
                                    const int32_t *__restrict__ metric, float *__restrict__ tq) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t n = (uint64_t)Q * len * HS_AA;
  if (i >= n) return;
  const int c = (int)(i % HS_AA);
  const uint64_t qp = i / HS_AA;
  tq[i] = (float)metric[c * HS_AA + qcodes[qp]];
}

// A dense query whose every 8-vector is bit-identical to a row of the embedding table is
// an embedded residue string (what protein2datapoints writes): recover its codes so that the
// exact stage reads 1 byte per position instead of 8 doubles.  qrow[q] = 1 when all positions
// matched.  The arithmetic downstream is unchanged (same values, same operation order).
__global__ void detect_query_codes_kernel(const double *__restrict__ q64, uint32_t Q, int len,
                                          const double *__restrict__ table64, uint8_t *__restrict__ qcodes,
                                          uint8_t *__restrict__ qrow) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const unsigned long long *tab = reinterpret_cast<const unsigned long long *>(table64);
  const unsigned long long *src = reinterpret_cast<const unsigned long long *>(q64) + (size_t)q * len * HS_CDIM;
  bool all = true;
  for (int p = 0; p < len; ++p) {
    int found = -1;
    for (int c = 0; c < HS_AA && found < 0; ++c) {
      bool same = true;
#pragma unroll
      for (int j = 0; j < HS_CDIM; ++j) same = same && (src[p * HS_CDIM + j] == tab[c * HS_CDIM + j]);
      if (same) found = c;
    }
    qcodes[(size_t)q * len + p] = found < 0 ? (uint8_t)0 : (uint8_t)found;
    all = all && found >= 0;
  }
  qrow[q] = all ? 1 : 0;
}

int launch_detect_query_codes(hs_ctx *ctx, const double *d_q64, uint32_t Q, uint8_t *d_qcodes, uint8_t *d_qrow) {
  if (Q == 0) return HS_OK;
  detect_query_codes_kernel<<<(Q + 127) / 128, 128, 0, ctx->stream>>>(d_q64, Q, (int)ctx->prm.len,
                                                                      ctx->d_table64.as<double>(), d_qcodes, d_qrow);
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches++;
  return HS_OK;
}

int launch_build_tq_points(hs_ctx *ctx, const double *d_q64, uint32_t Q, float *d_tq) {
  const uint64_t n = (uint64_t)Q * ctx->prm.len * HS_AA;
  if (n == 0) return HS_OK;
  build_tq_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
      d_q64, Q, (int)ctx->prm.len, ctx->d_table64.as<double>(), d_tq);
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches++;
  return HS_OK;
}
int launch_build_tq_int(hs_ctx *ctx, const uint8_t *d_qcodes, uint32_t Q, float *d_tq) {
  const uint64_t n = (uint64_t)Q * ctx->prm.len * HS_AA;
  if (n == 0) return HS_OK;
  build_tq_int_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
      d_qcodes, Q, (int)ctx->prm.len, ctx->d_metric.as<int32_t>(), d_tq);
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches++;
  return HS_OK;
}

// ---- filter -----------------------------------------------------------------------
template <int MODE, int LENB>
__global__ void __launch_bounds__(kFilterThreads)
filter_kernel(FilterArgs a) {
  extern __shared__ __align__(16) float s_tq[];  // [kFilterQChunk][len*20]
  __shared__ WorkItem s_item;
  const int tid = threadIdx.x;
  if (tid == 0) {
    uint32_t lo = 0, hi = a.nitems;  // last item with block_begin <= blockIdx.x
    while (hi - lo > 1) {
      const uint32_t m = (lo + hi) >> 1;
      if (a.items[m].block_begin <= blockIdx.x) lo = m; else hi = m;
    }
    s_item = a.items[lo];
  }
  __syncthreads();
  const WorkItem it = s_item;
  const int len = a.len;
  const int rowlen = len * HS_AA;
  const uint32_t tile = blockIdx.x - it.block_begin;
  const uint32_t pos0 = (it.m_begin & ~3u) + tile * kFilterTile + (uint32_t)tid * kFilterMembers;
  const uint8_t *store = a.stores[it.table];

  uint32_t cw[LENB];
#pragma unroll
  for (int p = 0; p < LENB; ++p)
    cw[p] = (p < len && pos0 < it.m_end) ? *reinterpret_cast<const uint32_t *>(store + (uint64_t)p * a.npad + pos0) : 0u;
  bool mvalid[kFilterMembers];
#pragma unroll
  for (int m = 0; m < kFilterMembers; ++m) mvalid[m] = (pos0 + m >= it.m_begin) && (pos0 + m < it.m_end);

  // self join: only queries positioned before the last member of this tile matter
  const uint32_t q_end = (MODE == kModeSelfJoin)
                             ? min(it.q_end, (it.m_begin & ~3u) + (tile + 1) * kFilterTile)
                             : it.q_end;
  for (uint32_t qc = it.q_begin; qc < q_end; qc += kFilterQChunk) {
    const int nq = (int)min((uint32_t)kFilterQChunk, q_end - qc);
    __syncthreads();
    for (int i = tid; i < nq * rowlen; i += kFilterThreads) {
      const int qq = i / rowlen, r = i - qq * rowlen;
      if (MODE == kModeSelfJoin) {
        const int p = r / HS_AA, c = r - p * HS_AA;
        const int cq = store[(uint64_t)p * a.npad + (qc + qq)] / kCodeScale;
        s_tq[i] = a.dsq32[cq * HS_AA + c];
      } else {
        s_tq[i] = a.tq[(uint64_t)(a.qlist[qc + qq] - a.tq_base) * rowlen + r];
      }
    }
    __syncthreads();
    for (int qq = 0; qq < nq; ++qq) {
      const char *row = reinterpret_cast<const char *>(s_tq + qq * rowlen);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int p = 0; p < LENB; ++p) {
        if (p < len) {
          const uint32_t w = cw[p];
          const char *rp = row + p * (HS_AA * 4);
          s0 += *reinterpret_cast<const float *>(rp + (w & 0xffu));
          s1 += *reinterpret_cast<const float *>(rp + ((w >> 8) & 0xffu));
          s2 += *reinterpret_cast<const float *>(rp + ((w >> 16) & 0xffu));
          s3 += *reinterpret_cast<const float *>(rp + (w >> 24));
        }
      }
      const float sv[kFilterMembers] = {s0, s1, s2, s3};
      const uint32_t qid = (MODE == kModeSelfJoin) ? (qc + qq) : a.qlist[qc + qq];
#pragma unroll
      for (int m = 0; m < kFilterMembers; ++m) {
        bool ok = mvalid[m] && sv[m] <= a.thr;
        if (MODE != kModeSearch) ok = ok && (qid < pos0 + m);  // each unordered pair once
        if (ok) {
          const unsigned long long idx = atomicAdd(a.surv_count, 1ull);
          if (idx < a.surv_cap) {
            Survivor s;
            s.query = qid;
            s.table = it.table;
            s.pos = pos0 + m;
            s.pad = 0;
            a.surv[idx] = s;
          }
        }
      }
    }
  }
}

template <int MODE>
static int launch_filter_mode(hs_ctx *ctx, const FilterArgs &a, uint32_t nblocks) {
  const size_t smem = (size_t)kFilterQChunk * a.len * HS_AA * sizeof(float);
#define HS_FILT(LB)                                                                     \
  do {                                                                                  \
    auto kern = filter_kernel<MODE, LB>;                                                \
    if (smem > 48 * 1024)                                                               \
      HS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<nblocks, kFilterThreads, smem, ctx->stream>>>(a);                            \
  } while (0)
  if (a.len <= 8) HS_FILT(8);
  else if (a.len <= 10) HS_FILT(10);
  else if (a.len <= 12) HS_FILT(12);
  else if (a.len <= 16) HS_FILT(16);
  else if (a.len <= 20) HS_FILT(20);
  else if (a.len <= 25) HS_FILT(25);
  else HS_FILT(32);
#undef HS_FILT
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches++;
  return HS_OK;
}

int launch_filter(hs_ctx *ctx, const FilterArgs &args, uint32_t nblocks, int mode) {
  if (nblocks == 0) return HS_OK;
  switch (mode) {
    case kModeSearch:
    case kModeBrute: return launch_filter_mode<kModeSearch>(ctx, args, nblocks);
    case kModeAllPairs: return launch_filter_mode<kModeAllPairs>(ctx, args, nblocks);
    default: return launch_filter_mode<kModeSelfJoin>(ctx, args, nblocks);
  }
}

// ---- exact stage --------------------------------------------------------------------
// `len` bytes at an arbitrarily aligned address as little-endian words: aligned 32-bit
// loads realigned by funnel shifts (reads at most 7 bytes past the end; buffers carry slack).
template <int NWORDS>
__device__ __forceinline__ void load_bytes(const uint8_t *p, int len, uint32_t (&w)[NWORDS]) {
  const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
  const uint32_t *base = reinterpret_cast<const uint32_t *>(addr & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(addr & 3u) * 8u;
  const int nw = (len + 3) >> 2;
  uint32_t prev = __ldg(base);
#pragma unroll
  for (int i = 0; i < NWORDS; ++i) {
    w[i] = 0u;
    if (i < nw) {
      const uint32_t next = __ldg(base + i + 1);
      w[i] = __funnelshift_r(prev, next, sh);
      prev = next;
    }
  }
}
// fragment record (16-byte aligned): NV 16-byte loads
template <int NV>
__device__ __forceinline__ void load_record(const uint8_t *p, uint32_t (&w)[4 * NV]) {
  const uint4 *src = reinterpret_cast<const uint4 *>(p);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const uint4 r = __ldg(src + v);
    w[4 * v + 0] = r.x;
    w[4 * v + 1] = r.y;
    w[4 * v + 2] = r.z;
    w[4 * v + 3] = r.w;
  }
}

constexpr int kExactThreads = 256;
constexpr int kSqStride = 10;  // REP == 1: doubles per residue-pair row (8 + 2 padding: spreads rows over the banks)
// REP == 8: the residue-pair table is stored eight times, copy c in the 16-byte bank group c, and lane
// l reads copy l & 7: the eight lanes of one LDS.128 phase never share a bank whatever rows they
// want (ncu, round 1: 59 % of the kernel's shared-memory wavefronts were bank conflicts and the
// l1tex data pipe was at 82 % of its peak).  200 KB of shared memory, one 1024-thread block per SM.
constexpr int kExactThreadsRep = 1024;
constexpr size_t kSqBytes1 = sizeof(double) * HS_AA * HS_AA * kSqStride;
constexpr size_t kSqBytes8 = sizeof(double2) * HS_AA * HS_AA * (HS_CDIM / 2) * 8;

// One thread per survivor, grid-stride.  NV: 16-byte words of residue codes per fragment
// (len <= 16 * NV).  The member's codes (and, on the rank path, its bucket ranks) come from
// its fragment record: one 32-byte sector per survivor.  Distances between residue strings
// read the squared coordinate differences from a shared residue-pair table, in the
// reference's summation order; hits are appended with one atomic per warp.
template <int NV, int REP>
__global__ void __launch_bounds__(REP == 8 ? kExactThreadsRep : kExactThreads) exact_kernel(ExactArgs a) {
  extern __shared__ __align__(16) unsigned char exact_smem[];
  double2 *s_sq2 = reinterpret_cast<double2 *>(exact_smem);  // (table[x][j] - table[q][j])^2
  __shared__ __align__(16) double s_table[HS_AA * HS_CDIM];
  __shared__ int s_metric[HS_AA * HS_AA];
  for (int i = threadIdx.x; i < HS_AA * HS_CDIM; i += blockDim.x) s_table[i] = a.table64[i];
  for (int i = threadIdx.x; i < HS_AA * HS_AA; i += blockDim.x) s_metric[i] = a.metric_tab[i];
  if (a.metric != HS_METRIC_BLOSUM_INT) {
    for (int i = threadIdx.x; i < HS_AA * HS_AA * (HS_CDIM / 2); i += blockDim.x) {
      const int j2 = i % (HS_CDIM / 2), pair = i / (HS_CDIM / 2);
      const int qc = pair / HS_AA, xc = pair - qc * HS_AA;
      // r = a - b; r * r (motif_both_points.cpp:180-181); (-r) * (-r) is the same double
      const double r0 = __dsub_rn(a.table64[xc * HS_CDIM + 2 * j2], a.table64[qc * HS_CDIM + 2 * j2]);
      const double r1 = __dsub_rn(a.table64[xc * HS_CDIM + 2 * j2 + 1], a.table64[qc * HS_CDIM + 2 * j2 + 1]);
      const double2 v = make_double2(__dmul_rn(r0, r0), __dmul_rn(r1, r1));
      if (REP == 8) {
#pragma unroll
        for (int c = 0; c < 8; ++c) s_sq2[(pair * (HS_CDIM / 2) + j2) * 8 + c] = v;
      } else {
        s_sq2[pair * (kSqStride / 2) + j2] = v;
      }
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int len = a.len;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  // all lanes of a warp run the same number of iterations (warp-level hit aggregation)
  for (unsigned long long i0 = (unsigned long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < a.nsurv;
       i0 += stride) {
    const unsigned long long i = i0 + lane;
    bool hit = false;
    double d2 = 0.0;
    uint64_t id = 0, qid = 0;
    Survivor s;
    s.table = 0;
    // the record of the survivor this thread takes next: pulled into L2 now, so that the gather
    // of the next iteration does not wait for DRAM (the exact stage stalls mostly on that load)
    if (i + stride < a.nsurv) {
      const Survivor nx = a.surv[i + stride];
      if (nx.pad & 2u)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.rec + (uint64_t)nx.pos * a.rec_stride));
    }
    if (i < a.nsurv) {
      s = a.surv[i];
      const uint32_t *ids = a.sorted_ids[s.table];
      // pad bit 1: the tensor filter already resolved the member's fragment id
      id = ((s.pad & 2u) || !ids) ? (uint64_t)s.pos : (uint64_t)__ldg(ids + s.pos);
      qid = (s.pad & 1u) ? (uint64_t)a.qlist_mma[s.query] : (uint64_t)s.query;
      bool live = !(a.mode == kModeAllPairs && !(qid < id));  // each unordered pair once
      const uint8_t *recp = a.rec + id * a.rec_stride;
      uint32_t mw[4 * NV], qw[4 * NV];
      if (live) load_record<NV>(recp, mw);
      bool have_qc = false;
      if (live) {
        if (a.mode == kModeSelfJoin) {
          // s.query: position of the query member (scalar filter) or, from the tensor filter,
          // the index of that position in its query list
          const uint32_t qpos = (s.pad & 1u) ? a.qlist_mma[s.query] : s.query;
          qid = ids ? (uint64_t)__ldg(ids + qpos) : (uint64_t)qpos;
          live = (s.pad & 1u) ? (qid < id) : true;  // tensor filter: keep each unordered pair once
          if (live) load_record<NV>(a.rec + qid * a.rec_stride, qw);
          have_qc = true;
        } else if (a.mode == kModeAllPairs && a.q64 == nullptr && a.qcodes == nullptr) {
          load_record<NV>(a.rec + qid * a.rec_stride, qw);  // all pairs of the DB: the query is a DB fragment
          have_qc = true;
        } else if (a.qcodes) {
          load_bytes<4 * NV>(a.qcodes + qid * len, len, qw);
          have_qc = true;
        }
      }
      // dense query that is an embedded residue string: its rows come from the shared table
      const bool dense = a.q64 && a.mode != kModeSelfJoin && !(a.qrow && a.qrow[qid] && have_qc);
      if (live) {
        if (a.metric == HS_METRIC_BLOSUM_INT) {
          int d = 0;
#pragma unroll
          for (int p = 0; p < 16 * NV; ++p)
            if (p < len) {
              const int xc = (mw[p >> 2] >> (8 * (p & 3))) & 0xff, qc = (qw[p >> 2] >> (8 * (p & 3))) & 0xff;
              d += s_metric[qc * HS_AA + xc];
            }
          d2 = (double)d;
          hit = d <= (int)a.R;
        } else {
          // PairwiseDistance_square (motif_both_points.cpp:176-183): r = a - b; dis += r * r,
          // strictly sequential, separate multiply and add
          double dis = 0.0;
          if (dense) {
            const double2 *qp = reinterpret_cast<const double2 *>(a.q64 + qid * a.dim);
#pragma unroll
            for (int p = 0; p < 16 * NV; ++p) {
              if (p >= len) break;
              const int xc = (mw[p >> 2] >> (8 * (p & 3))) & 0xff;
              const double2 *row = reinterpret_cast<const double2 *>(s_table + xc * HS_CDIM);
#pragma unroll
              for (int j = 0; j < HS_CDIM / 2; ++j) {
                const double2 t = row[j];
                const double2 q = __ldg(qp + p * (HS_CDIM / 2) + j);
                const double r0 = __dsub_rn(t.x, q.x);
                dis = __dadd_rn(dis, __dmul_rn(r0, r0));
                const double r1 = __dsub_rn(t.y, q.y);
                dis = __dadd_rn(dis, __dmul_rn(r1, r1));
              }
            }
          } else {
#pragma unroll
            for (int p = 0; p < 16 * NV; ++p)
              if (p < len) {
                const int xc = (mw[p >> 2] >> (8 * (p & 3))) & 0xff, qc = (qw[p >> 2] >> (8 * (p & 3))) & 0xff;
                const double2 *row = REP == 8 ? s_sq2 + (qc * HS_AA + xc) * (HS_CDIM / 2) * 8 + (lane & 7)
                                              : s_sq2 + (qc * HS_AA + xc) * (kSqStride / 2);
#pragma unroll
                for (int j = 0; j < HS_CDIM / 2; ++j) {
                  const double2 t = row[REP == 8 ? j * 8 : j];
                  dis = __dadd_rn(dis, t.x);
                  dis = __dadd_rn(dis, t.y);
                }
              }
          }
          d2 = dis;
          if (a.predicate == HS_PRED_D2_LE_R2) hit = dis <= __dmul_rn(a.R, a.R);
          else hit = !(sqrt(dis) > a.R);
        }
      }
      if (hit && a.mode == kModeSelfJoin) {
        uf_union(a.parent, (uint32_t)qid, (uint32_t)id);
        atomicAdd(a.edge_count, 1ull);
        hit = false;
      }
      if (hit && a.mode == kModeSearch) {
        // label[] (motif_both_points.cpp:232-238): the pair was already handled if an
        // earlier table put this fragment in the query's bucket.
        if (a.qrank) {
          const uint16_t *rk = reinterpret_cast<const uint16_t *>(recp + a.rec_rank_off);
          for (uint32_t l = 0; l < s.table; ++l) {
            const uint32_t qr = __ldg(a.qrank + (size_t)l * a.Q + qid);
            if (qr != 0xffffffffu && qr == (uint32_t)__ldg(rk + l)) hit = false;
          }
        } else {
          for (uint32_t l = 0; l < s.table && hit; ++l) {
            if (!a.qvalid[(size_t)l * a.Q + qid]) continue;
            bool same = true;
            for (int w = 0; w < a.key_words; ++w)
              same = same && (a.keys[l][(uint64_t)w * a.N + id] == a.qkeys[((size_t)l * a.Q + qid) * a.key_words + w]);
            if (same) hit = false;
          }
        }
      }
    }
    const uint32_t hm = __ballot_sync(0xffffffffu, hit);
    if (hm) {
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(a.hit_count, (unsigned long long)__popc(hm));
      base = __shfl_sync(0xffffffffu, base, 0);
      const unsigned long long idx = base + __popc(hm & ((1u << lane) - 1u));
      if (hit && idx < a.hit_cap) {
        hs_hit h;
        h.query = (uint32_t)qid;
        h.table_first = a.mode == kModeSearch ? s.table : 0u;
        h.db_id = a.id_base + id;
        h.dist2 = d2;
        a.hits[idx] = h;
      }
    }
  }
}

int launch_exact(hs_ctx *ctx, const ExactArgs &args) {
  if (args.nsurv == 0) return HS_OK;
  // grid-stride kernel: exactly one wave of resident blocks (a partial second wave would start
  // when the first ends and leave most SMs idle for its duration)
  static int per_sm[2][2] = {{0, 0}, {0, 0}};
  const int v = args.len <= 16 ? 0 : 1;
  // the replicated table pays once the survivors outnumber its set-up (200 KB per block)
  const int rep = (ctx->exact_rep && args.metric != HS_METRIC_BLOSUM_INT && args.nsurv >= (1ull << 20)) ? 1 : 0;
  const int threads = rep ? kExactThreadsRep : kExactThreads;
  const size_t smem = rep ? kSqBytes8 : kSqBytes1;
  const unsigned long long want = (args.nsurv + threads - 1) / threads;
  if (!per_sm[v][rep]) {
    HS_CUDA(cudaFuncSetAttribute(exact_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSqBytes8));
    HS_CUDA(cudaFuncSetAttribute(exact_kernel<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSqBytes8));
    HS_CUDA(cudaFuncSetAttribute(exact_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSqBytes1));
    HS_CUDA(cudaFuncSetAttribute(exact_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSqBytes1));
    if (v == 0 && rep == 0) HS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[v][rep], exact_kernel<1, 1>, threads, smem));
    if (v == 1 && rep == 0) HS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[v][rep], exact_kernel<2, 1>, threads, smem));
    if (v == 0 && rep == 1) HS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[v][rep], exact_kernel<1, 8>, threads, smem));
    if (v == 1 && rep == 1) HS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[v][rep], exact_kernel<2, 8>, threads, smem));
    if (per_sm[v][rep] < 1) per_sm[v][rep] = 1;
  }
  const unsigned grid = (unsigned)std::min<unsigned long long>(want, (unsigned long long)ctx->num_sms * per_sm[v][rep]);
  if (v == 0 && rep == 0) exact_kernel<1, 1><<<grid, threads, smem, ctx->stream>>>(args);
  else if (v == 1 && rep == 0) exact_kernel<2, 1><<<grid, threads, smem, ctx->stream>>>(args);
  else if (v == 0) exact_kernel<1, 8><<<grid, threads, smem, ctx->stream>>>(args);
  else exact_kernel<2, 8><<<grid, threads, smem, ctx->stream>>>(args);
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches++;
  return HS_OK;
}

}  // namespace hs
