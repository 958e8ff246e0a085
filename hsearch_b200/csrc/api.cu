// C ABI of libhsearch_b200.so: orchestration of the kernels (include/hsearch_b200.h).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <thread>
#include <vector>

#include "common.cuh"
#include "hash.cuh"
#include "host_tables.h"
#include "sort.cuh"
#include "verify.cuh"
#include "internal.cuh"
#include "hitsort.cuh"

namespace hs {
int setup_projection(hs_ctx *ctx, const double *a, const double *b);
int cluster_impl(hs_ctx *ctx, uint32_t *label_out);
int greedy_cluster_impl(hs_ctx *ctx, uint32_t *center_out, uint32_t *round_out, uint8_t *state_out);
int union_find_impl(hs_ctx *ctx, uint32_t n, const uint32_t *eu, const uint32_t *ev, uint64_t ne, uint32_t *label_out);
int extract_windows_impl(hs_ctx *ctx, const uint8_t *residues, const uint32_t *start_index, uint32_t nprot,
                         uint32_t stride, uint64_t id_base, uint32_t *pos_out, uint64_t pos_cap, uint64_t *nfrag);
int comm_gather_start(hs_ctx *ctx, const hs_hit *d_hits, uint64_t n, const uint64_t *d_keys, int tshift, uint32_t Q,
                      bool local_overflow);
int comm_gather_flush(hs_ctx *ctx);
int comm_broadcast(hs_ctx *ctx, void *d_buf, size_t bytes);
void comm_destroy(hs_ctx *ctx);

// Control traffic -- counters, bucket ranges, work lists, a few KB to a few MB per call -- does
// not use cudaMemcpy: a DMA copy queues on the copy engine of its direction behind whatever bulk
// transfer is in flight there (this context's hit list, or another context's database load on the
// same GPU), which stalls the pipeline for the length of that transfer (measured: 19 ms per
// step).  Instead a kernel moves the bytes between device memory and a mapped pinned staging
// area over PCIe, on the ctx stream.
__global__ void ctl_copy_kernel(void *__restrict__ dst, const void *__restrict__ src, size_t nbytes) {
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  const uintptr_t al = reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | nbytes;
  if ((al & 15) == 0) {
    for (size_t i = i0; i < nbytes / 16; i += stride) reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(src)[i];
  } else if ((al & 3) == 0) {
    for (size_t i = i0; i < nbytes / 4; i += stride) reinterpret_cast<uint32_t *>(dst)[i] = reinterpret_cast<const uint32_t *>(src)[i];
  } else {
    for (size_t i = i0; i < nbytes; i += stride) reinterpret_cast<uint8_t *>(dst)[i] = reinterpret_cast<const uint8_t *>(src)[i];
  }
  __threadfence_system();
}
static int ctl_launch(hs_ctx *ctx, void *dst, const void *src, size_t bytes) {
  const unsigned grid = (unsigned)std::min<size_t>((bytes / 16 + 255) / 256 + 1, 64);
  ctl_copy_kernel<<<grid, 256, 0, ctx->stream>>>(dst, src, bytes);
  HS_CUDA(cudaGetLastError());
  return HS_OK;
}
constexpr size_t kCtlDownMax = 8u << 20, kCtlUpArena = 16u << 20;

// Device -> host, synchronising.  (Also the point where the upload arena is known to be drained.)
int read_back(hs_ctx *ctx, const void *d_src, void *h_dst, size_t bytes) {
  if (bytes == 0) return HS_OK;
  if (bytes > kCtlDownMax) {
    HS_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    HS_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->up_used = 0;
    return HS_OK;
  }
  const size_t need = (bytes + 15) & ~(size_t)15;
  if (ctx->h_pinned_cap < need) {
    HS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    ctx->h_pinned = nullptr;
    ctx->h_pinned_cap = 0;
    const size_t want = std::max<size_t>(need + need / 4, 4096);
    HS_CUDA(cudaHostAlloc(&ctx->h_pinned, want, cudaHostAllocMapped));
    ctx->h_pinned_cap = want;
    HS_CUDA(cudaHostGetDevicePointer(&ctx->d_pinned, ctx->h_pinned, 0));
  }
  HS_TRY(ctl_launch(ctx, ctx->d_pinned, d_src, bytes));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->up_used = 0;
  memcpy(h_dst, ctx->h_pinned, bytes);
  return HS_OK;
}

// Host -> device, asynchronous: the bytes are copied into the pinned arena now (the caller's
// buffer may be a temporary) and moved by a kernel in stream order.
int upload(hs_ctx *ctx, void *d_dst, const void *h_src, size_t bytes) {
  if (bytes == 0) return HS_OK;
  if (bytes > kCtlUpArena / 2) {
    HS_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    HS_CUDA(cudaStreamSynchronize(ctx->stream));  // h_src may be a temporary
    ctx->up_used = 0;
    return HS_OK;
  }
  if (!ctx->h_up) {
    HS_CUDA(cudaHostAlloc(&ctx->h_up, kCtlUpArena, cudaHostAllocMapped));
    HS_CUDA(cudaHostGetDevicePointer(&ctx->d_up, ctx->h_up, 0));
    ctx->up_used = 0;
  }
  const size_t need = (bytes + 15) & ~(size_t)15;
  if (ctx->up_used + need > kCtlUpArena) {  // kernels may still read the arena: wait for them
    HS_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->up_used = 0;
  }
  memcpy((char *)ctx->h_up + ctx->up_used, h_src, bytes);
  HS_TRY(ctl_launch(ctx, d_dst, (const char *)ctx->d_up + ctx->up_used, bytes));
  ctx->up_used += need;
  return HS_OK;
}
// the ctx stream was synchronised by the caller: the arena is free again
void upload_drained(hs_ctx *ctx) { ctx->up_used = 0; }

// Host -> device of a caller's buffer that may be large (queries): read by a kernel straight from
// the caller's memory when that is pinned, else an ordinary copy.
static int upload_user(hs_ctx *ctx, void *d_dst, const void *h_src, size_t bytes) {
  if (bytes == 0) return HS_OK;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, h_src) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer &&
      bytes <= (64u << 20))
    return ctl_launch(ctx, d_dst, at.devicePointer, bytes);
  cudaGetLastError();
  HS_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return HS_OK;
}

static int upload_tables(hs_ctx *ctx) {
  HS_TRY(ctx->d_table64.reserve(sizeof(double) * HS_AA * HS_CDIM));
  HS_CUDA(cudaMemcpyAsync(ctx->d_table64.p, ctx->table64, sizeof(double) * HS_AA * HS_CDIM,
                          cudaMemcpyHostToDevice, ctx->stream));
  float dsq[HS_AA * HS_AA];
  for (int c = 0; c < HS_AA; ++c)
    for (int d = 0; d < HS_AA; ++d) {
      double s = 0.0;
      for (int j = 0; j < HS_CDIM; ++j) {
        const double r = ctx->table64[c * HS_CDIM + j] - ctx->table64[d * HS_CDIM + j];
        s += r * r;
      }
      dsq[c * HS_AA + d] = (float)s;
    }
  HS_TRY(ctx->d_dsq32.reserve(sizeof dsq));
  HS_CUDA(cudaMemcpyAsync(ctx->d_dsq32.p, dsq, sizeof dsq, cudaMemcpyHostToDevice, ctx->stream));
  int32_t metric[HS_AA * HS_AA];
  blosum_metric(metric);
  HS_TRY(ctx->d_metric.reserve(sizeof metric));
  HS_CUDA(cudaMemcpyAsync(ctx->d_metric.p, metric, sizeof metric, cudaMemcpyHostToDevice, ctx->stream));
  float metric32[HS_AA * HS_AA];
  for (int i = 0; i < HS_AA * HS_AA; ++i) metric32[i] = (float)metric[i];
  HS_TRY(ctx->d_metric32.reserve(sizeof metric32));
  HS_CUDA(cudaMemcpyAsync(ctx->d_metric32.p, metric32, sizeof metric32, cudaMemcpyHostToDevice, ctx->stream));
  ctx->have_ftable = true;
  if (ctx->prm.metric == HS_METRIC_BLOSUM_INT) ctx->have_ftable = blosum_filter_embedding(ctx->ftable64);
  else memcpy(ctx->ftable64, ctx->table64, sizeof ctx->ftable64);
  HS_TRY(ctx->d_ftable64.reserve(sizeof ctx->ftable64));
  if (ctx->have_ftable)
    HS_CUDA(cudaMemcpyAsync(ctx->d_ftable64.p, ctx->ftable64, sizeof ctx->ftable64, cudaMemcpyHostToDevice, ctx->stream));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  return mma_upload_tables(ctx);
}

void stats_begin(hs_ctx *ctx) {
  hs_stats &s = ctx->stats;
  const uint64_t nf = ctx->N;
  const uint32_t kw = ctx->key_words;
  memset(&s, 0, sizeof s);
  s.n_fragments = nf;
  s.key_words = kw;
  s.rank_path = ctx->rank_mode ? 1u : 0u;
}

float ev_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) {
    cudaGetLastError();
    return 0.f;
  }
  return ms;
}

// Filter threshold: pass iff filter distance <= thr.  Euclid: the FP32 table sum
// s^ satisfies s^ <= s_true * (1 + (len+1) * 2^-24), and a reference hit has
// s_true <= R^2 * (1 + 1e-13); thr = R^2 * (1 + (len+2) * 2^-23) rounded up
// doubles that margin.  Integer metric: the filter sum is exact.
float filter_threshold(const hs_ctx *ctx) {
  if (ctx->prm.metric == HS_METRIC_BLOSUM_INT) return (float)(int)ctx->prm.R;
  const double r2 = ctx->prm.R * ctx->prm.R;
  const double t = r2 * (1.0 + (double)(ctx->prm.len + 2) * ldexp(1.0, -23)) + 1e-30;
  float f = (float)t;
  if ((double)f < t) f = nextafterf(f, INFINITY);
  return f;
}

// Device tables of per-table pointers; entry L is the identity-order store.
static int upload_table_pointers(hs_ctx *ctx) {
  const uint32_t L = ctx->prm.L;
  std::vector<const void *> h(3 * (HS_MAX_L + 1), nullptr);
  for (uint32_t l = 0; l < L; ++l) {
    h[l] = ctx->tables[l].codes_sorted.p;
    h[(HS_MAX_L + 1) + l] = ctx->tables[l].sorted_ids.p;
    h[2 * (HS_MAX_L + 1) + l] = ctx->rank_mode ? nullptr : ctx->d_keys[l].p;
  }
  h[L] = ctx->d_codes_pm.p;
  h[(HS_MAX_L + 1) + L] = nullptr;
  HS_TRY(ctx->d_tabptrs.reserve(h.size() * sizeof(void *)));
  return upload(ctx, ctx->d_tabptrs.p, h.data(), h.size() * sizeof(void *));
}
const uint8_t *const *dev_stores(hs_ctx *ctx) { return ctx->d_tabptrs.as<const uint8_t *>(); }
const uint32_t *const *dev_sorted_ids(hs_ctx *ctx) {
  return reinterpret_cast<const uint32_t *const *>(ctx->d_tabptrs.as<const void *>() + (HS_MAX_L + 1));
}
const uint64_t *const *dev_keys(hs_ctx *ctx) {
  return reinterpret_cast<const uint64_t *const *>(ctx->d_tabptrs.as<const void *>() + 2 * (HS_MAX_L + 1));
}

int ensure_identity_store(hs_ctx *ctx) {
  if (ctx->have_codes_pm) return HS_OK;
  HS_TRY(build_code_store(ctx, nullptr, ctx->d_codes_pm));
  ctx->have_codes_pm = true;
  return upload_table_pointers(ctx);
}

// Bucket-ordered code stores of all tables (what the filters stream).  hs_build_index leaves them
// out for an index of tiny buckets (large K, small W: millions of buckets of one or two members)
// -- there a search examines a few thousand candidates and the stores, one random record gather per
// fragment and table, were most of the build -- and whoever needs them builds them on demand.
int ensure_code_stores(hs_ctx *ctx) {
  if (ctx->stores_built || ctx->N == 0) return HS_OK;
  bool blocked = false;
  HS_TRY(build_code_stores_blocked(ctx, &blocked));
  if (!blocked)
    for (uint32_t l = 0; l < ctx->prm.L; ++l) HS_TRY(build_table_store(ctx, l));
  ctx->stores_built = true;
  return upload_table_pointers(ctx);
}

// Filter bypass: every member of every probed bucket becomes a survivor (the exact stage decides).
// One warp per (table, query) range; off[] = exclusive prefix sums of the range sizes.
__global__ void expand_candidates_kernel(const uint2 *__restrict__ qrange, const uint32_t *__restrict__ off, uint32_t Q,
                                         uint32_t nranges, Survivor *__restrict__ surv) {
  const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= nranges) return;
  const uint2 rg = qrange[r];
  const uint32_t table = r / Q, q = r - table * Q;
  Survivor *dst = surv + off[r];
  for (uint32_t i = rg.x + lane; i < rg.y; i += 32) {
    Survivor sv;
    sv.query = q;
    sv.table = table;
    sv.pos = i;
    sv.pad = 0;
    dst[i - rg.x] = sv;
  }
}
__global__ void range_sizes_kernel(const uint2 *__restrict__ qrange, uint32_t n, uint32_t *__restrict__ sz) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sz[i] = qrange[i].y - qrange[i].x;
}
constexpr uint64_t kBypassMaxCandidates = 1ull << 22;

// Run the filter (scalar leg and, when given, the tensor-core leg; both append to
// the same survivor list), growing the survivor buffer and retrying on overflow.
int run_filter(hs_ctx *ctx, FilterArgs &fa, uint32_t nblocks, int mode, uint64_t *nsurv_out, FilterArgs *fa_tc,
               uint32_t nblocks_tc, const MmaLaunch *ml) {
  unsigned long long *cnt = ctx->d_counters.as<unsigned long long>() + 8;
  for (int attempt = 0; attempt < 3; ++attempt) {
    if (ctx->d_surv.cap < sizeof(Survivor) * (1u << 20)) HS_TRY(ctx->d_surv.reserve(sizeof(Survivor) * (1u << 22)));
    fa.surv = ctx->d_surv.as<Survivor>();
    fa.surv_cap = ctx->d_surv.cap / sizeof(Survivor);
    fa.surv_count = cnt;
    HS_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), ctx->stream));
    HS_CUDA(cudaMemsetAsync(cnt + 7, 0, sizeof(unsigned long long), ctx->stream));
    // the previous search's hit merge (multi-GPU) starts its bulk transfer now, beside this filter (comm.cu)
    if (ctx->nranks > 1) HS_TRY(comm_gather_flush(ctx));
    HS_TRY(launch_filter(ctx, fa, nblocks, mode));
    if (fa_tc && nblocks_tc) {
      fa_tc->surv = fa.surv;
      fa_tc->surv_cap = fa.surv_cap;
      fa_tc->surv_count = cnt;
      HS_CUDA(cudaEventRecord(ctx->ev[10], ctx->stream));
      HS_TRY(launch_filter_tc(ctx, *fa_tc, ctx->d_tq16.p, kTcTilesPerBlock, nblocks_tc, mode));
      HS_CUDA(cudaEventRecord(ctx->ev[11], ctx->stream));
    }
    if (ml && ml->grid) {
      FilterArgs fm = fa;
      fm.qlist = ml->qlist;
      HS_CUDA(cudaEventRecord(ctx->ev[13], ctx->stream));
      HS_TRY(launch_filter_mma(ctx, fm, ctx->d_mma_items.p, ctx->d_mma_units.p, ml->nunits,
                               reinterpret_cast<uint32_t *>(ctx->d_counters.as<unsigned long long>() + 13), ml->grid,
                               ctx->d_qb16.p, mode));
      HS_CUDA(cudaEventRecord(ctx->ev[14], ctx->stream));
    }
    // slot 8: survivors; slot 15: threshold events staged by the pipelined tensor filter (the event
    // list has the survivor list's capacity: either overflowing means a bigger buffer and a rerun)
    unsigned long long h8[8] = {0};
    HS_TRY(read_back(ctx, cnt, h8, sizeof h8));
    const unsigned long long n = h8[0], n_ev = (ml && ml->grid && HS_MMA_EVENTS) ? h8[7] : 0ull;
    if (fa_tc && nblocks_tc) ctx->stats.ms_filter_tc += ev_ms(ctx->ev[10], ctx->ev[11]);
    if (ml && ml->grid) ctx->stats.ms_filter_tc += ev_ms(ctx->ev[13], ctx->ev[14]);
    if (n <= fa.surv_cap && n_ev <= fa.surv_cap) {
      *nsurv_out = n;
      return HS_OK;
    }
    const unsigned long long need = std::max(n, n_ev);
    HS_TRY(ctx->d_surv.reserve(sizeof(Survivor) * (size_t)(need + need / 8 + 1024)));
  }
  set_error("filter: survivor buffer kept overflowing");
  return HS_ERR_NOMEM;
}

// ---- filter work list ------------------------------------------------------------------
// One group = the members [mb, me) of one bucket (or of the whole store) against a
// list of queries.  Groups probed by enough queries go to the tensor-core filter
// in items of <= kTcQueriesPerItem queries; the rest stay on the scalar filter.
struct FilterPlan {
  std::vector<WorkItem> items, items_tc;
  std::vector<uint32_t> qlist, qlist_tc;
  uint32_t nblocks = 0, nblocks_tc = 0;
  uint64_t ncand = 0, ncand_tc = 0;
  // pipelined tensor filter (Euclidean metric): items of <= qmax queries, units of member chunks
  bool mma = false;
  uint32_t mma_qmax = 0;
  std::vector<MmaItemHost> mma_items;
  std::vector<MmaUnitHost> mma_units;
  std::vector<uint32_t> qlist_mma;
};

static void plan_init(const hs_ctx *ctx, FilterPlan &P) {
  P.mma = mma_filter_usable(ctx);
  if (P.mma) {
    MmaGeometry g;
    mma_geometry(ctx, &g);
    P.mma_qmax = (uint32_t)g.qmax;
  }
}

static int plan_add(const hs_ctx *ctx, FilterPlan &P, uint32_t table, uint32_t mb, uint32_t me, const uint32_t *q,
                    size_t nq, bool allpairs_block) {
  if (me <= mb || nq == 0) return HS_OK;
  const bool tc = !(ctx->prm.flags & HS_FLAG_SCALAR_FILTER) && nq >= (size_t)tc_min_queries() &&
                  (me - mb) >= kTcMinMembers;
  if (tc && P.mma) {
    const size_t nchunks = (nq + P.mma_qmax - 1) / P.mma_qmax;
    size_t per = (nq + nchunks - 1) / nchunks;
    per = (per + 15) & ~(size_t)15;  // whole 16-column MMA chunks
    for (size_t c = 0; c < nq; c += per) {
      const size_t ce = std::min(nq, c + per);
      // all pairs i<j: members at or below the chunk's first query never pair with it
      const uint32_t m0 = allpairs_block ? std::max(mb, q[c] + 1) : mb;
      if (me <= m0) continue;
      MmaItemHost it;
      it.table = table;
      it.q_begin = (uint32_t)P.qlist_mma.size();
      P.qlist_mma.insert(P.qlist_mma.end(), q + c, q + ce);
      it.q_end = (uint32_t)P.qlist_mma.size();
      it.pad = 0;
      const uint32_t item = (uint32_t)P.mma_items.size();
      P.mma_items.push_back(it);
      // unit = chunk of the bucket's members handed to one CTA at a time: larger units amortise
      // the per-unit hand-off on large DBs (measured at 100 M fragments: 35.5 ms with 32 tiles,
      // 34.9 ms with 128), smaller ones keep all SMs busy on small DBs
      const uint32_t unit_tiles = ctx->N >= (1ull << 25) ? 4 * kMmaUnitTiles : ctx->N >= (1ull << 23) ? 2 * kMmaUnitTiles
                                                                                                       : kMmaUnitTiles;
      const uint32_t step = unit_tiles * 128u;
      for (uint64_t m = m0; m < me; m += step) {
        MmaUnitHost un;
        un.item = item;
        un.m_begin = (uint32_t)m;
        un.m_end = (uint32_t)std::min<uint64_t>(me, m + step);
        un.pad = 0;
        P.mma_units.push_back(un);
      }
      const uint64_t pairs = (uint64_t)(me - m0) * (ce - c);
      P.ncand += pairs;
      P.ncand_tc += pairs;
    }
    return HS_OK;
  }
  if (tc) {
    const size_t nchunks = (nq + kTcQueriesPerItem - 1) / kTcQueriesPerItem;
    const size_t per = (nq + nchunks - 1) / nchunks;
    for (size_t c = 0; c < nq; c += per) {
      const size_t ce = std::min(nq, c + per);
      WorkItem it;
      it.table = table;
      // all pairs i<j: members at or below the chunk's first query never pair with it
      it.m_begin = allpairs_block ? std::max(mb, q[c] + 1) : mb;
      it.m_end = me;
      if (it.m_end <= it.m_begin) continue;
      it.q_begin = (uint32_t)P.qlist_tc.size();
      P.qlist_tc.insert(P.qlist_tc.end(), q + c, q + ce);
      it.q_end = (uint32_t)P.qlist_tc.size();
      it.block_begin = P.nblocks_tc;
      const uint64_t per_block = (uint64_t)kTcTilesPerBlock * kTcTileMembers;
      const uint64_t nb = ((uint64_t)(it.m_end - it.m_begin) + per_block - 1) / per_block;
      if ((uint64_t)P.nblocks_tc + nb > 0x7fffffffull) {
        set_error("filter work list exceeds 2^31 blocks");
        return HS_ERR_UNSUPPORTED;
      }
      P.nblocks_tc += (uint32_t)nb;
      P.items_tc.push_back(it);
      const uint64_t pairs = (uint64_t)(it.m_end - it.m_begin) * (ce - c);
      P.ncand += pairs;
      P.ncand_tc += pairs;
    }
    return HS_OK;
  }
  for (size_t c = 0; c < nq; c += kQueriesPerItem) {
    const size_t ce = std::min(nq, c + (size_t)kQueriesPerItem);
    WorkItem it;
    it.table = table;
    it.m_begin = allpairs_block ? std::max(mb, q[c] + 1) : mb;
    it.m_end = me;
    if (it.m_end <= it.m_begin) continue;
    it.q_begin = (uint32_t)P.qlist.size();
    P.qlist.insert(P.qlist.end(), q + c, q + ce);
    it.q_end = (uint32_t)P.qlist.size();
    it.block_begin = P.nblocks;
    const uint32_t ntiles = (it.m_end - (it.m_begin & ~3u) + kFilterTile - 1) / kFilterTile;
    if ((uint64_t)P.nblocks + ntiles > 0x7fffffffull) {
      set_error("filter work list exceeds 2^31 blocks");
      return HS_ERR_UNSUPPORTED;
    }
    P.nblocks += ntiles;
    P.items.push_back(it);
    P.ncand += (uint64_t)(it.m_end - it.m_begin) * (ce - c);
  }
  return HS_OK;
}

// Appends the sub-plan `src` (built independently, indices relative to itself) to `dst`.
static int plan_merge(FilterPlan &dst, const FilterPlan &src) {
  if ((uint64_t)dst.nblocks + src.nblocks > 0x7fffffffull || (uint64_t)dst.nblocks_tc + src.nblocks_tc > 0x7fffffffull) {
    set_error("filter work list exceeds 2^31 blocks");
    return HS_ERR_UNSUPPORTED;
  }
  const uint32_t q_off = (uint32_t)dst.qlist.size(), qt_off = (uint32_t)dst.qlist_tc.size(),
                 qm_off = (uint32_t)dst.qlist_mma.size(), item_off = (uint32_t)dst.mma_items.size();
  for (WorkItem it : src.items) {
    it.q_begin += q_off;
    it.q_end += q_off;
    it.block_begin += dst.nblocks;
    dst.items.push_back(it);
  }
  for (WorkItem it : src.items_tc) {
    it.q_begin += qt_off;
    it.q_end += qt_off;
    it.block_begin += dst.nblocks_tc;
    dst.items_tc.push_back(it);
  }
  for (MmaItemHost it : src.mma_items) {
    it.q_begin += qm_off;
    it.q_end += qm_off;
    dst.mma_items.push_back(it);
  }
  for (MmaUnitHost un : src.mma_units) {
    un.item += item_off;
    dst.mma_units.push_back(un);
  }
  dst.qlist.insert(dst.qlist.end(), src.qlist.begin(), src.qlist.end());
  dst.qlist_tc.insert(dst.qlist_tc.end(), src.qlist_tc.begin(), src.qlist_tc.end());
  dst.qlist_mma.insert(dst.qlist_mma.end(), src.qlist_mma.begin(), src.qlist_mma.end());
  dst.nblocks += src.nblocks;
  dst.nblocks_tc += src.nblocks_tc;
  dst.ncand += src.ncand;
  dst.ncand_tc += src.ncand_tc;
  return HS_OK;
}

// Upload the plan and run both filter legs.  tq_base: tq row of query id x is x - tq_base.
static int plan_run(hs_ctx *ctx, FilterPlan &P, uint32_t tq_rows, uint32_t tq_base, int mode, uint64_t *nsurv) {
  *nsurv = 0;
  if (P.items.empty() && P.items_tc.empty() && P.mma_units.empty()) return HS_OK;
  MmaLaunch ml;
  if (!P.mma_units.empty()) {
    const uint32_t nunits = (uint32_t)P.mma_units.size();
    {
      if (ctx->plan_stats) {
        // width histogram of the tensor filter's (tile, query group) work: columns per MMA group
        uint64_t hist[9] = {0}, groups = 0, padded = 0;
        for (const MmaUnitHost &un : P.mma_units) {
          const MmaItemHost &it = P.mma_items[un.item];
          const uint64_t tiles = (un.m_end - (un.m_begin & ~15u) + 127) / 128;
          for (uint32_t q0 = it.q_begin; q0 < it.q_end; q0 += 256) {
            const uint32_t ng = std::min<uint32_t>(256, it.q_end - q0), ngp = (ng + 15u) & ~15u;
            groups += tiles;
            padded += tiles * 128 * ngp;
            hist[std::min<uint32_t>(8, ngp / 32)] += tiles;
          }
        }
        fprintf(stderr, "[plan] mma items %zu units %zu tile-groups %llu padded pairs %llu real pairs %llu | groups by width/32:",
                P.mma_items.size(), P.mma_units.size(), (unsigned long long)groups, (unsigned long long)padded,
                (unsigned long long)P.ncand_tc);
        for (int i = 0; i < 9; ++i) fprintf(stderr, " %llu", (unsigned long long)hist[i]);
        fprintf(stderr, "\n");
      }
    }
    const uint32_t grid = std::min<uint32_t>((uint32_t)ctx->num_sms, (nunits + 3) / 4);
    HS_TRY(ctx->d_mma_items.reserve(sizeof(MmaItemHost) * P.mma_items.size()));
    HS_TRY(ctx->d_mma_units.reserve(sizeof(MmaUnitHost) * P.mma_units.size()));
    HS_TRY(ctx->d_qlist_mma.reserve(sizeof(uint32_t) * P.qlist_mma.size()));
    HS_TRY(upload(ctx, ctx->d_mma_items.p, P.mma_items.data(), sizeof(MmaItemHost) * P.mma_items.size()));
    HS_TRY(upload(ctx, ctx->d_mma_units.p, P.mma_units.data(), sizeof(MmaUnitHost) * P.mma_units.size()));
    HS_TRY(upload(ctx, ctx->d_qlist_mma.p, P.qlist_mma.data(), sizeof(uint32_t) * P.qlist_mma.size()));
    MmaGeometry g;
    HS_TRY(mma_geometry(ctx, &g));
    HS_TRY(ctx->d_qb16.reserve(sizeof(uint16_t) * (size_t)std::max<uint32_t>(tq_rows, 1) * g.kp + 16));
    if (mode == kModeAllPairs) HS_TRY(launch_build_qb_codes(ctx, tq_base, tq_rows, ctx->d_qb16.p));
    else if (mode == kModeSelfJoin)  // queries = positions tq_base .. of the (single) table of the plan
      HS_TRY(launch_build_qb_store(ctx, P.mma_items[0].table, tq_base, tq_rows, ctx->d_qb16.p));
    else if (ctx->prm.metric == HS_METRIC_BLOSUM_INT)   // queries are residue codes (stage_queries checks)
      HS_TRY(launch_build_qb_qcodes(ctx, ctx->d_qcodes.as<uint8_t>(), tq_rows, ctx->d_qb16.p));
    else HS_TRY(launch_build_qb_points(ctx, ctx->d_q64.as<double>(), tq_rows, ctx->d_qb16.p));
    ml.grid = grid;
    ml.nunits = nunits;
    ml.qlist = ctx->d_qlist_mma.as<uint32_t>();
  }
  FilterArgs fa, ft;
  memset(&fa, 0, sizeof fa);
  fa.tq = ctx->d_tq.as<float>();
  fa.dsq32 = ctx->d_dsq32.as<float>();
  fa.stores = dev_stores(ctx);
  fa.npad = ctx->npad;
  fa.len = (int)ctx->prm.len;
  fa.thr = filter_threshold(ctx);
  fa.tq_base = tq_base;
  ft = fa;
  if (!P.items.empty()) {
    HS_TRY(ctx->d_work.reserve(sizeof(WorkItem) * P.items.size()));
    HS_TRY(ctx->d_qlist.reserve(sizeof(uint32_t) * P.qlist.size()));
    HS_TRY(upload(ctx, ctx->d_work.p, P.items.data(), sizeof(WorkItem) * P.items.size()));
    HS_TRY(upload(ctx, ctx->d_qlist.p, P.qlist.data(), sizeof(uint32_t) * P.qlist.size()));
    fa.items = ctx->d_work.as<WorkItem>();
    fa.nitems = (uint32_t)P.items.size();
    fa.qlist = ctx->d_qlist.as<uint32_t>();
  }
  if (!P.items_tc.empty()) {
    HS_TRY(ctx->d_work_tc.reserve(sizeof(WorkItem) * P.items_tc.size()));
    HS_TRY(ctx->d_qlist_tc.reserve(sizeof(uint32_t) * P.qlist_tc.size()));
    HS_TRY(upload(ctx, ctx->d_work_tc.p, P.items_tc.data(), sizeof(WorkItem) * P.items_tc.size()));
    HS_TRY(upload(ctx, ctx->d_qlist_tc.p, P.qlist_tc.data(), sizeof(uint32_t) * P.qlist_tc.size()));
    ft.items = ctx->d_work_tc.as<WorkItem>();
    ft.nitems = (uint32_t)P.items_tc.size();
    ft.qlist = ctx->d_qlist_tc.as<uint32_t>();
    HS_TRY(ctx->d_tq16.reserve(sizeof(uint16_t) * (size_t)tq_rows * tc_padded_k(ctx->prm.len) + 16));
    HS_TRY(launch_tq_to_half(ctx, ctx->d_tq.as<float>(), tq_rows, ctx->d_tq16.p));
  }
  return run_filter(ctx, fa, P.nblocks, mode, nsurv, &ft, P.nblocks_tc, &ml);
}

// Cluster: all pairs i < j inside the bucket [mb, me) of `table` through the tensor filter
// (queries = the bucket's own members).  Returns false in *used when the tensor path does not
// apply (integer metric, tiny bucket); survivors land in ctx->d_surv as usual.
// (one launch + synchronisation per bucket: below ~8 k members the tiled scalar self-join, which
// batches many buckets per launch, is faster -- measured at 1 M fragments)
constexpr uint32_t kSelfJoinMmaMin = 8192;
// Whether a bucket of n members is self-joined by the tensor filter (else: tiled scalar self-join).
bool selfjoin_uses_mma(const hs_ctx *ctx, uint32_t n) { return n >= kSelfJoinMmaMin && mma_filter_usable(ctx); }
// The pairs (q, m), q in [q_lo, q_hi), m in (q, me): the caller walks the bucket in query chunks so
// that the survivors of one call stay bounded (a whole 4 M-member bucket at once is 8e12 pairs).
int selfjoin_bucket_mma(hs_ctx *ctx, uint32_t table, uint32_t mb, uint32_t me, uint32_t q_lo, uint32_t q_hi,
                        uint64_t *nsurv, uint64_t *npairs, bool *used) {
  *used = false;
  *nsurv = 0;
  *npairs = 0;
  FilterPlan P;
  plan_init(ctx, P);
  if (!P.mma || me - mb < kSelfJoinMmaMin) return HS_OK;
  *used = true;
  if (q_hi <= q_lo) return HS_OK;
  std::vector<uint32_t> q(q_hi - q_lo);
  for (uint32_t i = 0; i < q_hi - q_lo; ++i) q[i] = q_lo + i;
  HS_TRY(plan_add(ctx, P, table, mb, me, q.data(), q.size(), true));
  if (!P.items.empty() || !P.items_tc.empty()) {
    set_error("selfjoin_bucket_mma: unexpected scalar work items");
    return HS_ERR_UNSUPPORTED;
  }
  if (P.mma_units.empty()) return HS_OK;
  // pairs with m > q inside the bucket
  for (uint32_t x = q_lo; x < q_hi; ++x) *npairs += (uint64_t)(me - 1 - x);
  return plan_run(ctx, P, q_hi - q_lo, q_lo, kModeSelfJoin, nsurv);
}

// ---- hits in the reference's output order -----------------------------------------
__global__ void hit_keys_kernel(const hs_hit *__restrict__ hits, uint64_t n, uint64_t *__restrict__ k0,
                                uint64_t *__restrict__ k1) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const hs_hit h = hits[i];
  k0[i] = h.db_id;
  k1[i] = ((uint64_t)h.query << 32) | (uint64_t)h.table_first;
}
__global__ void hit_gather_kernel(const hs_hit *__restrict__ in, const uint32_t *__restrict__ perm, uint64_t n,
                                  hs_hit *__restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[perm[i]];
}

// (query, first table, db id) as one 64-bit key: query in the top qbits, table in the next
// tbits, db id below.  *overflow is set when a field does not fit.
__global__ void hit_key1_kernel(const hs_hit *__restrict__ hits, uint64_t n, int tshift, int qshift,
                                uint64_t *__restrict__ k0, unsigned int *__restrict__ overflow) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const hs_hit h = hits[i];
  if ((h.db_id >> tshift) || ((uint64_t)h.table_first >> (qshift - tshift)) || ((uint64_t)h.query >> (64 - qshift)))
    *overflow = 1u;
  k0[i] = ((uint64_t)h.query << qshift) | ((uint64_t)h.table_first << tshift) | h.db_id;
}

// Compact (CSR) result: entry i of the sorted order as local id | table << id_bits, and dist2.
__global__ void hit_gather_compact_kernel(const hs_hit *__restrict__ in, const uint32_t *__restrict__ perm, uint64_t n,
                                          uint64_t id_base, int id_bits, uint32_t *__restrict__ idt,
                                          double *__restrict__ dist2) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const hs_hit h = in[perm[i]];
  idt[i] = (uint32_t)(h.db_id - id_base) | (h.table_first << id_bits);
  dist2[i] = h.dist2;
}
// offsets[q] = base + first sorted entry whose query is >= q, for the queries q in [qa, qb]
// (qb included: the end of the block's last segment).  keys: the sorted one-word hit keys.
__global__ void hit_offsets_kernel(const uint64_t *__restrict__ keys, uint64_t n, int qshift, uint32_t qa, uint32_t qb,
                                   uint64_t base, uint64_t *__restrict__ offsets) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  const uint32_t q_prev = i == 0 ? qa : (uint32_t)(keys[i - 1] >> qshift) + 1u;  // first query not yet closed
  const uint32_t q_here = i == n ? qb : (uint32_t)(keys[i] >> qshift);
  for (uint32_t q = q_prev; q <= q_here; ++q) offsets[q] = base + i;
}

static int bits_for(uint64_t nvalues) {  // bits that hold 0 .. nvalues-1
  int b = 1;
  while (b < 64 && (nvalues - 1) >> b) ++b;
  return b;
}

// Compact output of sort_hits: the block holds the queries [qa, qb), its entries start at `base`.
struct CompactBlock {
  uint32_t qa = 0, qb = 0;
  uint64_t base = 0;
  int id_bits = 0;
};
static int compact_id_bits(const hs_ctx *ctx, int *id_bits) {
  const int lb = bits_for(std::max<uint64_t>(ctx->N, 2)), tb = bits_for((uint64_t)ctx->prm.L + 1);
  if (lb + tb > 32) {
    set_error("compact hits: %d id bits + %d table bits exceed 32", lb, tb);
    return HS_ERR_UNSUPPORTED;
  }
  *id_bits = lb;
  return HS_OK;
}

// Sort d_hits[0..n) by (query, first table, db id); result in ctx->d_hits_sorted, or, with
// `cb`, as compact entries in ctx->d_cidt / d_cdist plus the block's offsets in ctx->d_coffsets.
struct SortedKeys {
  const uint64_t *keys = nullptr;  // one-word keys in sorted order (valid until the next sort)
  int tshift = 0;                  // the table field starts at this bit, the query field follows it
};
static int sort_hits(hs_ctx *ctx, hs_hit *d_hits, uint64_t n, const CompactBlock *cb = nullptr, SortedKeys *sk = nullptr) {
  if (cb) {
    HS_TRY(ctx->d_cidt.reserve(sizeof(uint32_t) * std::max<uint64_t>(n, 1)));
    HS_TRY(ctx->d_cdist.reserve(sizeof(double) * std::max<uint64_t>(n, 1)));
  }
  if (n == 0) {
    if (cb) {
      hit_offsets_kernel<<<1, 256, 0, ctx->stream>>>(nullptr, 0, 0, cb->qa, cb->qb, cb->base, ctx->d_coffsets.as<uint64_t>());
      // (n == 0: thread 0 alone writes offsets[qa .. qb] = base)
      ctx->stats.kernel_launches++;
      HS_CUDA(cudaGetLastError());
    }
    return HS_OK;
  }
  if (n >= (1ull << 32)) {
    set_error("sort_hits: more than 2^32 hits");
    return HS_ERR_UNSUPPORTED;
  }
  if (!cb) HS_TRY(ctx->d_hits_sorted.reserve(sizeof(hs_hit) * n));
  if (!sk) {
    // partition by the top key bits + per-bin sort in shared memory (hitsort.cu); it hands the
    // list back untouched when a field does not fit or a bin cannot be processed
    SegSortRequest rq;
    rq.compact = cb != nullptr;
    rq.hits_out = ctx->d_hits_sorted.as<hs_hit>();
    if (cb) {
      rq.idt = ctx->d_cidt.as<uint32_t>();
      rq.dist2 = ctx->d_cdist.as<double>();
      rq.offsets = ctx->d_coffsets.as<uint64_t>();
      rq.qa = cb->qa;
      rq.qb = cb->qb;
      rq.base = cb->base;
      rq.id_bits = cb->id_bits;
    }
    bool used = false;
    HS_TRY(sort_hits_segmented(ctx, d_hits, n, rq, &used));
    if (used) return HS_OK;
  }
  HS_TRY(ctx->d_hit_keys[0].reserve(sizeof(uint64_t) * n));
  HS_TRY(ctx->d_hit_perm.reserve(sizeof(uint32_t) * 2 * n));
  const unsigned grid = (unsigned)((n + 255) / 256);
  KeyPtrs in, sorted;
  memset(&in, 0, sizeof in);
  uint32_t *perm = ctx->d_hit_perm.as<uint32_t>();
  const hs_stats before = ctx->stats;
  // one-word key when the three fields fit 64 bits (they do unless ids are astronomically large)
  // fields packed tightly from bit 0 (db id, then table, then query): the radix sort walks 8-bit
  // windows over the varying bits, so a gap between the fields would cost an extra pass.  Ids
  // gathered from other ranks may exceed this rank's range: the kernel flags that (overflow) and
  // the two-word path is taken.
  const int qbits = bits_for(std::max<uint64_t>(ctx->hit_qmax, 1)), tbits = bits_for((uint64_t)ctx->prm.L + 1);
  const int ibits = bits_for(std::max<uint64_t>(ctx->hit_idmax ? ctx->hit_idmax : ctx->id_base + ctx->N, 2));
  const int tshift = ibits, qshift = ibits + tbits;
  bool one_word = qshift + qbits <= 64;
  if (one_word) {
    unsigned int *ovf = reinterpret_cast<unsigned int *>(ctx->d_counters.as<unsigned long long>() + 14);
    HS_CUDA(cudaMemsetAsync(ovf, 0, sizeof(unsigned int), ctx->stream));
    hit_key1_kernel<<<grid, 256, 0, ctx->stream>>>(d_hits, n, tshift, qshift, ctx->d_hit_keys[0].as<uint64_t>(), ovf);
    ctx->stats.kernel_launches++;
    unsigned int h_ovf = 0;
    HS_TRY(read_back(ctx, ovf, &h_ovf, sizeof h_ovf));
    one_word = h_ovf == 0;
  }
  if ((cb || sk) && !one_word) {
    set_error("compact / gathered hits need the one-word hit key (query, table and id within 64 bits)");
    return HS_ERR_UNSUPPORTED;
  }
  if (one_word) {
    in.w[0] = ctx->d_hit_keys[0].as<uint64_t>();
    const int kbits = qshift + qbits;
    HS_TRY(radix_sort_pairs(ctx, in, nullptr, n, 1, perm, perm + n, &sorted,
                            kbits >= 64 ? ~0ull : ((1ull << kbits) - 1ull)));
  } else {
    HS_TRY(ctx->d_hit_keys[1].reserve(sizeof(uint64_t) * n));
    hit_keys_kernel<<<grid, 256, 0, ctx->stream>>>(d_hits, n, ctx->d_hit_keys[0].as<uint64_t>(),
                                                  ctx->d_hit_keys[1].as<uint64_t>());
    ctx->stats.kernel_launches++;
    in.w[0] = ctx->d_hit_keys[0].as<uint64_t>();
    in.w[1] = ctx->d_hit_keys[1].as<uint64_t>();
    HS_TRY(radix_sort_pairs(ctx, in, nullptr, n, 2, perm, perm + n, &sorted));
  }
  // sort_passes / ms_sort_* describe the index build only
  ctx->stats.sort_passes = before.sort_passes;
  ctx->stats.ms_sort_upsweep = before.ms_sort_upsweep;
  ctx->stats.ms_sort_scan = before.ms_sort_scan;
  ctx->stats.ms_sort_downsweep = before.ms_sort_downsweep;
  if (sk) {
    sk->keys = sorted.w[0];
    sk->tshift = tshift;
  }
  if (cb) {
    hit_gather_compact_kernel<<<grid, 256, 0, ctx->stream>>>(d_hits, perm, n, ctx->id_base, cb->id_bits,
                                                            ctx->d_cidt.as<uint32_t>(), ctx->d_cdist.as<double>());
    hit_offsets_kernel<<<(unsigned)((n + 256) / 256), 256, 0, ctx->stream>>>(sorted.w[0], n, qshift, cb->qa, cb->qb, cb->base,
                                                                            ctx->d_coffsets.as<uint64_t>());
    ctx->stats.kernel_launches += 2;
  } else {
    hit_gather_kernel<<<grid, 256, 0, ctx->stream>>>(d_hits, perm, n, ctx->d_hits_sorted.as<hs_hit>());
    ctx->stats.kernel_launches++;
  }
  HS_CUDA(cudaGetLastError());
  return HS_OK;
}

// ---- search ----------------------------------------------------------------------------
constexpr uint32_t kSearchBlocks = 4;  // query blocks of the pipelined verify / sort / copy-out

// Query blocks grow geometrically (1/16, 3/16, 1/4, 1/2 of the queries): the first block is
// verified and sorted quickly, so that the PCIe copy chain -- the longer of the two pipelines --
// starts early; the later, larger blocks are always ready before the copy engine needs them.
struct SearchBlocks {
  uint32_t end[kSearchBlocks];  // block c holds the queries [end[c-1], end[c])
};
static SearchBlocks search_blocks(uint32_t Q) {
  SearchBlocks b;
  b.end[0] = Q / 16;
  b.end[1] = Q / 4;
  b.end[2] = Q / 2;
  b.end[3] = Q;
  return b;
}

// Regroups the survivor list by (query block, fragment-id block).  The query blocks are what the
// pipelined verify / sort / copy-out of a host-buffer search works through one after the other; the
// id blocks give the exact stage its locality: it gathers one 32-byte fragment record per survivor,
// and in filter order those gathers are random over the whole record array (128 bytes of DRAM
// traffic per record, 13 GB for 10^8 survivors), while inside an id block of 2^id_shift records
// (16 MB) they stay in L2 and every DRAM sector is fetched about once.  SCATTER = false: counts per
// bin; true: writes the survivors bin by bin (cursor[] preset to the bin offsets).  Counts are
// aggregated in shared memory: one global atomic per (thread block, bin) and tile of 4096 survivors.
constexpr int kSurvItems = 16;
constexpr uint32_t kSurvIdBins = 256, kSurvMaxBins = kSearchBlocks * kSurvIdBins;
struct SurvBinning {
  SearchBlocks qb;
  uint32_t nqblk;    // query blocks in use (1: no split by query)
  uint32_t nid;      // id blocks per query block
  int id_shift;      // id block of a fragment = id >> id_shift (clamped to nid - 1)
};
template <bool SCATTER>
__global__ void __launch_bounds__(256)
survivor_bin_kernel(const Survivor *__restrict__ surv, uint64_t n, const uint32_t *__restrict__ qlist_mma, SurvBinning sb,
                    unsigned long long *__restrict__ cursor, Survivor *__restrict__ out) {
  __shared__ unsigned int s_cnt[kSurvMaxBins];
  __shared__ unsigned long long s_base[kSurvMaxBins];
  const uint32_t nbins = sb.nqblk * sb.nid;
  const uint64_t tile = (uint64_t)blockDim.x * kSurvItems;
  for (uint64_t t0 = (uint64_t)blockIdx.x * tile; t0 < n; t0 += (uint64_t)gridDim.x * tile) {
    for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x) s_cnt[b] = 0u;
    __syncthreads();
    uint32_t bs[kSurvItems];   // bin | slot << 16 (a tile holds 4096 survivors: the slot fits 16 bits)
#pragma unroll
    for (int j = 0; j < kSurvItems; ++j) {
      const uint64_t i = t0 + (uint64_t)j * blockDim.x + threadIdx.x;
      bs[j] = 0xffffffffu;
      if (i < n) {
        const Survivor sv = surv[i];
        uint32_t bq = 0;
        if (sb.nqblk > 1) {
          const uint32_t q = (sv.pad & 1u) ? __ldg(qlist_mma + sv.query) : sv.query;
#pragma unroll
          for (int c = 0; c < (int)kSearchBlocks - 1; ++c) bq += q >= sb.qb.end[c] ? 1u : 0u;
        }
        // (survivors of the scalar filter carry a bucket position, not an id: first id block)
        const uint32_t ib = (sv.pad & 2u) ? min(sv.pos >> sb.id_shift, sb.nid - 1u) : 0u;
        const uint32_t bin = bq * sb.nid + ib;
        bs[j] = bin | (atomicAdd(&s_cnt[bin], 1u) << 16);
      }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x)
      if (s_cnt[b]) s_base[b] = atomicAdd(cursor + b, (unsigned long long)s_cnt[b]);
    __syncthreads();
    if (SCATTER) {
#pragma unroll
      for (int j = 0; j < kSurvItems; ++j)
        if (bs[j] != 0xffffffffu) {
          const uint64_t i = t0 + (uint64_t)j * blockDim.x + threadIdx.x;
          out[s_base[bs[j] & 0xffffu] + (bs[j] >> 16)] = surv[i];   // (re-read: L2-hot)
        }
    }
    __syncthreads();
  }
}

// d_surv[0 .. nsurv) -> d_surv_blk, by (query block, id block); cnt_qblk[c] = survivors of query block c.
static int bin_survivors(hs_ctx *ctx, uint64_t nsurv, uint32_t Q, uint32_t nqblk, unsigned long long *cnt_qblk,
                         bool by_id = true) {
  for (uint32_t c = 0; c < nqblk; ++c) cnt_qblk[c] = 0;
  if (nsurv == 0) return HS_OK;
  SurvBinning sb;
  sb.qb = search_blocks(Q);
  sb.nqblk = nqblk;
  sb.nid = by_id ? kSurvIdBins : 1u;
  int nbits = 1;
  while (nbits < 32 && (ctx->N >> nbits)) ++nbits;   // ids < 2^nbits
  sb.id_shift = std::max(0, nbits - 8);
  const uint32_t nbins = sb.nqblk * sb.nid;
  HS_TRY(ctx->d_surv_blk.reserve(sizeof(Survivor) * nsurv));
  HS_TRY(ctx->d_binctr.reserve(sizeof(unsigned long long) * 2 * kSurvMaxBins));
  unsigned long long *ctr = ctx->d_binctr.as<unsigned long long>();
  HS_CUDA(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long) * nbins, ctx->stream));
  const unsigned grid = (unsigned)std::min<uint64_t>((nsurv + 4095) / 4096, (uint64_t)ctx->num_sms * 8);
  survivor_bin_kernel<false><<<grid, 256, 0, ctx->stream>>>(ctx->d_surv.as<Survivor>(), nsurv, ctx->d_qlist_mma.as<uint32_t>(), sb, ctr, nullptr);
  HS_CUDA(cudaGetLastError());
  std::vector<unsigned long long> h(nbins), cur(nbins);
  HS_TRY(read_back(ctx, ctr, h.data(), sizeof(unsigned long long) * nbins));
  unsigned long long run = 0;
  for (uint32_t b = 0; b < nbins; ++b) {
    cur[b] = run;
    run += h[b];
    cnt_qblk[b / sb.nid] += h[b];
  }
  HS_TRY(upload(ctx, ctr + kSurvMaxBins, cur.data(), sizeof(unsigned long long) * nbins));
  survivor_bin_kernel<true><<<grid, 256, 0, ctx->stream>>>(ctx->d_surv.as<Survivor>(), nsurv, ctx->d_qlist_mma.as<uint32_t>(), sb, ctr + kSurvMaxBins, ctx->d_surv_blk.as<Survivor>());
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches += 2;
  return HS_OK;
}

struct QueryInput {
  const double *h_points = nullptr;  // host [Q][dim]
  const void *d_points = nullptr;    // device [Q][dim]
  const uint8_t *h_codes = nullptr;  // host [Q][len]
};

static int stage_queries(hs_ctx *ctx, const QueryInput &in, uint32_t Q) {
  const uint32_t dim = ctx->dim, len = ctx->prm.len;
  HS_TRY(ctx->d_q64.reserve(sizeof(double) * std::max<size_t>(1, (size_t)Q * dim)));
  ctx->have_qcodes = false;
  if (in.d_points) {
    HS_CUDA(cudaMemcpyAsync(ctx->d_q64.p, in.d_points, sizeof(double) * Q * dim, cudaMemcpyDeviceToDevice, ctx->stream));
  } else if (in.h_points) {
    HS_TRY(upload_user(ctx, ctx->d_q64.p, in.h_points, sizeof(double) * Q * dim));
  } else {
    // residue-code queries: embed with the ctx table (M1) on the host; Q is small
    std::vector<double> pts((size_t)Q * dim);
    for (uint32_t q = 0; q < Q; ++q)
      for (uint32_t p = 0; p < len; ++p) {
        const uint8_t c = in.h_codes[(size_t)q * len + p];
        if (c >= HS_AA) {
          set_error("query %u has residue code %u (must be 0..19)", q, (unsigned)c);
          return HS_ERR_INVALID;
        }
        memcpy(&pts[(size_t)q * dim + p * HS_CDIM], ctx->table64 + c * HS_CDIM, sizeof(double) * HS_CDIM);
      }
    HS_TRY(upload(ctx, ctx->d_q64.p, pts.data(), sizeof(double) * Q * dim));
    HS_TRY(ctx->d_qcodes.reserve((size_t)Q * len + 16));  // slack: word-wise reads (load_bytes)
    HS_TRY(upload(ctx, ctx->d_qcodes.p, in.h_codes, (size_t)Q * len));
    ctx->have_qcodes = true;
  }
  if (ctx->nranks > 1) {
    HS_TRY(comm_broadcast(ctx, ctx->d_q64.p, sizeof(double) * Q * dim));
    if (ctx->have_qcodes) HS_TRY(comm_broadcast(ctx, ctx->d_qcodes.p, (size_t)Q * len));
  }
  if (ctx->prm.metric == HS_METRIC_BLOSUM_INT && !ctx->have_qcodes) {
    set_error("the integer BLOSUM metric needs residue-code queries");
    return HS_ERR_INVALID;
  }
  if (ctx->prm.metric == HS_METRIC_EUCLID_FP64) {
    HS_TRY(ctx->d_qcodes_det.reserve((size_t)Q * len + 16));
    HS_TRY(ctx->d_qrow.reserve(std::max<size_t>(1, (size_t)Q)));
    HS_TRY(launch_detect_query_codes(ctx, ctx->d_q64.as<double>(), Q, ctx->d_qcodes_det.as<uint8_t>(),
                                     ctx->d_qrow.as<uint8_t>()));
  }
  return HS_OK;
}

static int build_tq(hs_ctx *ctx, uint32_t Q) {
  HS_TRY(ctx->d_tq.reserve(sizeof(float) * std::max<size_t>(1, (size_t)Q * ctx->prm.len * HS_AA)));
  if (ctx->prm.metric == HS_METRIC_BLOSUM_INT)
    return launch_build_tq_int(ctx, ctx->d_qcodes.as<uint8_t>(), Q, ctx->d_tq.as<float>());
  return launch_build_tq_points(ctx, ctx->d_q64.as<double>(), Q, ctx->d_tq.as<float>());
}

void fill_exact_common(hs_ctx *ctx, ExactArgs &ea, uint32_t Q) {
  memset(&ea, 0, sizeof ea);
  ea.metric = (int)ctx->prm.metric;
  ea.predicate = (int)ctx->prm.predicate;
  ea.len = (int)ctx->prm.len;
  ea.dim = (int)ctx->dim;
  ea.key_words = (int)ctx->key_words;
  ea.L = (int)ctx->prm.L;
  ea.R = ctx->prm.R;
  ea.sorted_ids = dev_sorted_ids(ctx);
  ea.codes = ctx->d_codes.as<uint8_t>();
  ea.rec = ctx->d_rec.as<uint8_t>();
  ea.rec_stride = ctx->rec_stride;
  ea.rec_rank_off = ctx->rec_rank_off;
  ea.N = ctx->N;
  ea.id_base = ctx->id_base;
  ea.table64 = ctx->d_table64.as<double>();
  ea.metric_tab = ctx->d_metric.as<int32_t>();
  ea.Q = Q;
  ea.keys = dev_keys(ctx);
  ea.qlist_mma = ctx->d_qlist_mma.as<uint32_t>();
}

// Finish a search / brute-force call: optional ordering, copy out, and -- on a context that
// joined a communicator -- the start of the merge of all ranks' lists into rank 0's memory
// (comm.cu; completed by hs_comm_result).  The caller's buffer always receives this rank's hits.
static int deliver_hits(hs_ctx *ctx, uint64_t nh, uint64_t dev_cap, hs_hit *hits_host, void *hits_dev,
                        uint64_t cap, uint64_t *nhits, cudaEvent_t ev_sort0, cudaEvent_t ev_sort1, uint32_t Q) {
  hs_hit *d_src = ctx->d_hits.as<hs_hit>();
  const uint64_t nvalid = std::min<uint64_t>(nh, dev_cap);
  const bool gather = ctx->nranks > 1 && cap > 0;   // cap == 0: count-only call, nothing to merge
  // the gather that used these buffers two searches ago must have read them
  if (gather && ctx->gather_seq >= 2) HS_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_gather[ctx->gather_seq & 1u], 0));
  HS_CUDA(cudaEventRecord(ev_sort0, ctx->stream));
  SortedKeys sk;
  if (((ctx->prm.flags & HS_FLAG_SORT_HITS) || gather) && nh <= dev_cap) {
    HS_TRY(sort_hits(ctx, d_src, nvalid, nullptr, gather ? &sk : nullptr));
    if (nvalid) d_src = ctx->d_hits_sorted.as<hs_hit>();
  }
  HS_CUDA(cudaEventRecord(ev_sort1, ctx->stream));
  if (gather) {
    const bool overflow = nh > dev_cap;
    HS_TRY(comm_gather_start(ctx, d_src, overflow ? 0 : nvalid, sk.keys, sk.tshift, Q, overflow));
    if (nvalid && !overflow) std::swap(ctx->d_hits_sorted, ctx->d_hits_sorted_alt);  // the gather reads d_src
  }
  *nhits = nh;
  ctx->stats.n_hits = nh;
  const uint64_t ncopy = std::min<uint64_t>(nvalid, cap);
  if (ncopy) {
    if (hits_dev)
      HS_CUDA(cudaMemcpyAsync(hits_dev, d_src, sizeof(hs_hit) * ncopy, cudaMemcpyDeviceToDevice, ctx->stream));
    else
      HS_CUDA(cudaMemcpyAsync(hits_host, d_src, sizeof(hs_hit) * ncopy, cudaMemcpyDeviceToHost, ctx->stream));
  }
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  if (nh > cap) {
    set_error("hit buffer too small: %llu hits, capacity %llu", (unsigned long long)nh, (unsigned long long)cap);
    return HS_ERR_CAPACITY;
  }
  return HS_OK;
}

static int search_impl(hs_ctx *ctx, const QueryInput &in, uint32_t Q, hs_hit *hits_host, void *hits_dev,
                       uint64_t cap, uint64_t *nhits, hs_compact_hits *compact = nullptr) {
  if (!ctx->indexed) {
    set_error("hs_search: call hs_build_index first");
    return HS_ERR_INVALID;
  }
  const uint32_t L = ctx->prm.L, KW = ctx->key_words;
  stats_begin(ctx);
  cudaEvent_t *ev = ctx->ev;
  HS_CUDA(cudaEventRecord(ev[0], ctx->stream));
  HS_TRY(stage_queries(ctx, in, Q));

  // query keys + probe
  HS_TRY(ctx->d_qkeys.reserve(sizeof(uint64_t) * std::max<size_t>(1, (size_t)L * Q * KW)));
  HS_TRY(ctx->d_qvalid.reserve(std::max<size_t>(1, (size_t)L * Q)));
  HS_TRY(ctx->d_qrange.reserve(sizeof(uint2) * std::max<size_t>(1, (size_t)L * Q)));
  HS_TRY(ctx->d_qrank.reserve(sizeof(uint32_t) * std::max<size_t>(1, (size_t)L * Q)));
  HS_TRY(launch_hash_queries(ctx, ctx->d_q64.as<double>(), Q, ctx->d_qkeys.as<uint64_t>(), ctx->d_qvalid.as<uint8_t>()));
  HS_CUDA(cudaEventRecord(ev[1], ctx->stream));
  for (uint32_t l = 0; l < L; ++l)
    HS_TRY(launch_probe(ctx, l, ctx->d_qkeys.as<uint64_t>(), ctx->d_qvalid.as<uint8_t>(), Q, ctx->d_qrange.as<uint2>(),
                         ctx->d_qrank.as<uint32_t>()));
  HS_TRY(build_tq(ctx, Q));
  std::vector<uint2> qrange((size_t)L * Q);
  HS_CUDA(cudaEventRecord(ev[2], ctx->stream));
  HS_TRY(read_back(ctx, ctx->d_qrange.p, qrange.data(), sizeof(uint2) * qrange.size()));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));

  // An index without code stores (tiny buckets, see ensure_code_stores): a small candidate set goes
  // to the exact stage unfiltered, a large one has the stores built now.
  bool bypass = false;
  if (!ctx->stores_built) {
    uint64_t ncand = 0;
    for (const uint2 &r : qrange) ncand += r.y - r.x;
    if (ncand <= ctx->bypass_max && ctx->prm.metric == HS_METRIC_EUCLID_FP64) bypass = true;
    else HS_TRY(ensure_code_stores(ctx));
  }
  auto run_bypass = [&](uint64_t *nsurv) -> int {
    const uint32_t nr = L * Q;
    HS_TRY(ctx->d_misc.reserve(sizeof(uint32_t) * 2 * ((size_t)nr + 1)));
    uint32_t *sz = ctx->d_misc.as<uint32_t>(), *off = sz + nr + 1;
    uint32_t *d_total = reinterpret_cast<uint32_t *>(ctx->d_counters.as<unsigned long long>() + 16);
    range_sizes_kernel<<<(nr + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_qrange.as<uint2>(), nr, sz);
    HS_TRY(exclusive_scan_u32(ctx, sz, off, nr, d_total));
    uint32_t total = 0;
    HS_TRY(read_back(ctx, d_total, &total, sizeof total));
    HS_TRY(ctx->d_surv.reserve(sizeof(Survivor) * std::max<size_t>(total, 1)));
    if (total)
      expand_candidates_kernel<<<(unsigned)(((uint64_t)nr * 32 + 255) / 256), 256, 0, ctx->stream>>>(
          ctx->d_qrange.as<uint2>(), off, Q, nr, ctx->d_surv.as<Survivor>());
    HS_CUDA(cudaGetLastError());
    ctx->stats.kernel_launches += 2;
    ctx->stats.n_candidates += total;
    *nsurv = total;
    return HS_OK;
  };

  // work list of the queries [qa, qb): queries grouped by bucket, chunked
  // (one sub-plan per table, built on its own host thread and merged in table order: the
  // planning of 10 k queries x 4 tables is ~0.7 ms per table of pure host time)
  auto make_plan = [&](FilterPlan &plan, uint32_t qa, uint32_t qb) -> int {
    const auto _tp0 = std::chrono::steady_clock::now();
    plan_init(ctx, plan);
    std::vector<FilterPlan> sub(L);
    std::vector<int> rcs(L, HS_OK);
    auto build = [&](uint32_t l) {
      FilterPlan &P = sub[l];
      plan_init(ctx, P);
      std::vector<uint64_t> order;
      std::vector<uint32_t> group;
      order.reserve(qb - qa);
      for (uint32_t q = qa; q < qb; ++q) {
        const uint2 r = qrange[(size_t)l * Q + q];
        if (r.y > r.x) order.push_back(((uint64_t)r.x << 32) | q);
      }
      std::sort(order.begin(), order.end());
      if (l == 0 && ctx->plan_stats)
        fprintf(stderr, "[plan] table 0 sorted at %.3f ms\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - _tp0).count());
      size_t i = 0;
      while (i < order.size() && rcs[l] == HS_OK) {
        const uint32_t mb = (uint32_t)(order[i] >> 32);
        size_t j = i;
        group.clear();
        while (j < order.size() && (uint32_t)(order[j] >> 32) == mb) group.push_back((uint32_t)order[j++]);
        const uint32_t me = qrange[(size_t)l * Q + group[0]].y;
        rcs[l] = plan_add(ctx, P, l, mb, me, group.data(), group.size(), false);
        i = j;
      }
    };
    if (L > 1 && (uint64_t)(qb - qa) * L >= 8192) {
      std::vector<std::thread> th;
      for (uint32_t l = 1; l < L; ++l) th.emplace_back(build, l);
      build(0);
      for (std::thread &t : th) t.join();
    } else {
      for (uint32_t l = 0; l < L; ++l) build(l);
    }
    if (ctx->plan_stats)
      fprintf(stderr, "[plan] sub-plans built at %.3f ms\n",
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - _tp0).count());
    {
      size_t ni = 0, nit = 0, nmi = 0, nmu = 0, nq = 0, nqt = 0, nqm = 0;
      for (const FilterPlan &P : sub) {
        ni += P.items.size(), nit += P.items_tc.size(), nmi += P.mma_items.size(), nmu += P.mma_units.size();
        nq += P.qlist.size(), nqt += P.qlist_tc.size(), nqm += P.qlist_mma.size();
      }
      plan.items.reserve(ni), plan.items_tc.reserve(nit), plan.mma_items.reserve(nmi), plan.mma_units.reserve(nmu);
      plan.qlist.reserve(nq), plan.qlist_tc.reserve(nqt), plan.qlist_mma.reserve(nqm);
    }
    for (uint32_t l = 0; l < L; ++l) {
      if (rcs[l] != HS_OK) {  // (the worker thread's message is thread-local)
        set_error("search: filter work list of table %u could not be built (status %d)", l, rcs[l]);
        return rcs[l];
      }
      HS_TRY(plan_merge(plan, sub[l]));
    }
    ctx->stats.n_candidates += plan.ncand;
    ctx->stats.n_candidates_tc += plan.ncand_tc;
    ctx->stats.n_work_items += plan.items.size() + plan.items_tc.size() + plan.mma_units.size();
    return HS_OK;
  };
  // exact + dedup + emit of the current survivor list into d_hits[0 .. hit_cap)
  unsigned long long *hit_count = ctx->d_counters.as<unsigned long long>() + 9;
  auto run_exact = [&](const Survivor *surv, uint64_t nsurv, uint64_t hit_cap) -> int {
    HS_CUDA(cudaMemsetAsync(hit_count, 0, sizeof(unsigned long long), ctx->stream));
    ExactArgs ea;
    fill_exact_common(ctx, ea, Q);
    ea.surv = surv;
    ea.nsurv = nsurv;
    ea.mode = kModeSearch;
    ea.q64 = ctx->prm.metric == HS_METRIC_EUCLID_FP64 ? ctx->d_q64.as<double>() : nullptr;
    ea.qcodes = ctx->have_qcodes ? ctx->d_qcodes.as<uint8_t>() : nullptr;
    if (ctx->prm.metric == HS_METRIC_EUCLID_FP64) {
      ea.qcodes = ctx->d_qcodes_det.as<uint8_t>();
      ea.qrow = ctx->d_qrow.as<uint8_t>();
    }
    ea.qkeys = ctx->d_qkeys.as<uint64_t>();
    ea.qvalid = ctx->d_qvalid.as<uint8_t>();
    ea.qrank = ctx->rank_mode ? ctx->d_qrank.as<uint32_t>() : nullptr;
    ea.hits = ctx->d_hits.as<hs_hit>();
    ea.hit_cap = hit_cap;
    ea.hit_count = hit_count;
    return launch_exact(ctx, ea);
  };
  ctx->hit_qmax = Q;
  const uint64_t dev_cap = std::max<uint64_t>(cap, 1);
  HS_TRY(ctx->d_hits.reserve(sizeof(hs_hit) * dev_cap));
  ctx->stats.ms_qhash = ev_ms(ev[0], ev[1]);
  ctx->stats.ms_probe = ev_ms(ev[1], ev[2]);

  // Host hit buffer + reference order: the filter runs once over all queries; its survivors are
  // then split by query block, and each block is verified, sorted and sent over PCIe on a
  // second stream while the next block is verified and sorted (hits are ordered by query first,
  // so the blocks concatenate into the final order).  HS_NO_PIPELINE=1 disables it.
  // (Splitting the *filter* by query block was measured too: it quarters the queries per
  // bucket and doubles the tensor filter's time, DESIGN.md.)
  int id_bits = 0;
  if (compact) {
    if (ctx->nranks > 1) {
      set_error("hs_search_points_compact: not available on a context that joined a communicator");
      return HS_ERR_UNSUPPORTED;
    }
    HS_TRY(compact_id_bits(ctx, &id_bits));
    compact->id_bits = (uint32_t)id_bits;
    HS_TRY(ctx->d_coffsets.reserve(sizeof(uint64_t) * ((size_t)Q + 1)));
  }
  const bool pipelined = compact || (hits_host && ctx->nranks == 1 && (ctx->prm.flags & HS_FLAG_SORT_HITS) && Q >= 2048 &&
                                     !ctx->no_pipeline);
  if (pipelined) {
    const uint32_t nblk = kSearchBlocks;
    const SearchBlocks sblk = search_blocks(Q);
    if (!ctx->copy_stream) HS_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    while (ctx->ev_chunk.size() < 4 * nblk + 8) {
      cudaEvent_t e;
      HS_CUDA(cudaEventCreate(&e));
      ctx->ev_chunk.push_back(e);
    }
    FilterPlan plan;
    if (!bypass) HS_TRY(make_plan(plan, 0, Q));
    HS_CUDA(cudaEventRecord(ev[12], ctx->stream));
    uint64_t nsurv = 0;
    if (bypass) HS_TRY(run_bypass(&nsurv));
    else HS_TRY(plan_run(ctx, plan, Q, 0, kModeSearch, &nsurv));
    ctx->stats.n_survivors = nsurv;
    // survivors by query block (and, inside a query block, by fragment-id block: bin_survivors)
    unsigned long long h_cnt[kSearchBlocks] = {0};
    HS_TRY(bin_survivors(ctx, nsurv, Q, nblk, h_cnt, ctx->surv_bins));
    HS_CUDA(cudaEventRecord(ev[3], ctx->stream));
    uint64_t done = 0, total = 0, soff = 0;
    for (uint32_t c = 0; c < nblk; ++c) {
      cudaEvent_t *ce = &ctx->ev_chunk[4 * c];  // [0] start [1] verified [2] sorted [3] copied
      HS_CUDA(cudaEventRecord(ce[0], ctx->stream));
      const uint64_t cap_left = cap - std::min(done, cap);
      HS_TRY(run_exact(ctx->d_surv_blk.as<Survivor>() + soff, h_cnt[c], cap_left));
      soff += h_cnt[c];
      unsigned long long nh = 0;
      HS_CUDA(cudaEventRecord(ce[1], ctx->stream));
      HS_TRY(read_back(ctx, hit_count, &nh, sizeof nh));
      const uint64_t nvalid = std::min<uint64_t>(nh, cap_left);
      // the sorted buffer written now was last read by the copy of block c-2
      if (c >= 2) HS_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[4 * (c - 2) + 3], 0));
      CompactBlock cb;
      cb.qa = c ? sblk.end[c - 1] : 0u;
      cb.qb = sblk.end[c];
      cb.base = done;
      cb.id_bits = id_bits;
      HS_TRY(sort_hits(ctx, ctx->d_hits.as<hs_hit>(), nvalid, compact ? &cb : nullptr));
      HS_CUDA(cudaEventRecord(ce[2], ctx->stream));
      if (nvalid || (compact && c + 1 == nblk)) HS_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ce[2], 0));
      if (nvalid && compact) {
        HS_CUDA(cudaMemcpyAsync(compact->idt + done, ctx->d_cidt.p, sizeof(uint32_t) * nvalid, cudaMemcpyDeviceToHost,
                                ctx->copy_stream));
        HS_CUDA(cudaMemcpyAsync(compact->dist2 + done, ctx->d_cdist.p, sizeof(double) * nvalid, cudaMemcpyDeviceToHost,
                                ctx->copy_stream));
      } else if (nvalid) {
        HS_CUDA(cudaMemcpyAsync(hits_host + done, ctx->d_hits_sorted.p, sizeof(hs_hit) * nvalid, cudaMemcpyDeviceToHost,
                                ctx->copy_stream));
      }
      if (compact && c + 1 == nblk)  // every block's offsets are written: one copy of the whole array
        HS_CUDA(cudaMemcpyAsync(compact->offsets, ctx->d_coffsets.p, sizeof(uint64_t) * ((size_t)Q + 1),
                                cudaMemcpyDeviceToHost, ctx->copy_stream));
      HS_CUDA(cudaEventRecord(ce[3], ctx->copy_stream));
      if (compact) {
        std::swap(ctx->d_cidt, ctx->d_cidt_alt);
        std::swap(ctx->d_cdist, ctx->d_cdist_alt);
      } else {
        std::swap(ctx->d_hits_sorted, ctx->d_hits_sorted_alt);
      }
      done += nvalid;
      total += nh;
    }
    HS_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    HS_CUDA(cudaEventRecord(ev[7], ctx->stream));
    HS_CUDA(cudaEventSynchronize(ev[7]));
    ctx->stats.ms_host = ev_ms(ev[2], ev[12]);
    ctx->stats.ms_filter = ev_ms(ev[12], ev[3]);
    for (uint32_t c = 0; c < nblk; ++c) {
      cudaEvent_t *ce = &ctx->ev_chunk[4 * c];
      ctx->stats.ms_exact += ev_ms(ce[0], ce[1]);
      ctx->stats.ms_hitsort += ev_ms(ce[1], ce[2]);
    }
    ctx->stats.n_hits = total;
    ctx->stats.ms_total = ev_ms(ev[0], ev[7]);
    *nhits = total;
    if (total > cap) {
      set_error("hit buffer too small: %llu hits, capacity %llu", (unsigned long long)total, (unsigned long long)cap);
      return HS_ERR_CAPACITY;
    }
    return HS_OK;
  }

  FilterPlan plan;
  const auto _t0 = std::chrono::steady_clock::now();
  if (!bypass) HS_TRY(make_plan(plan, 0, Q));
  if (ctx->plan_stats)
    fprintf(stderr, "[plan] host planning %.3f ms\n",
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - _t0).count());
  HS_CUDA(cudaEventRecord(ev[12], ctx->stream));
  uint64_t nsurv = 0;
  if (bypass) HS_TRY(run_bypass(&nsurv));
  else HS_TRY(plan_run(ctx, plan, Q, 0, kModeSearch, &nsurv));
  HS_CUDA(cudaEventRecord(ev[3], ctx->stream));
  ctx->stats.n_survivors = nsurv;
  // HS_SURV_BINS=1 regroups large lists by fragment-id block first so that the record gathers of the
  // exact stage stay in L2 (bin_survivors).  Off: measured at bench C2 the two regrouping passes cost
  // more (1.6 ms) than the exact stage gains (8.2 -> 7.6 ms).
  if (nsurv >= (1u << 20) && ctx->surv_bins) {
    unsigned long long one[kSearchBlocks];
    HS_TRY(bin_survivors(ctx, nsurv, Q, 1, one));
    HS_TRY(run_exact(ctx->d_surv_blk.as<Survivor>(), nsurv, dev_cap));
  } else {
    HS_TRY(run_exact(ctx->d_surv.as<Survivor>(), nsurv, dev_cap));
  }
  unsigned long long nh = 0;
  HS_CUDA(cudaEventRecord(ev[4], ctx->stream));
  HS_TRY(read_back(ctx, hit_count, &nh, sizeof nh));

  int rc = deliver_hits(ctx, nh, dev_cap, hits_host, hits_dev, cap, nhits, ev[5], ev[6], Q);
  HS_CUDA(cudaEventRecord(ev[7], ctx->stream));
  HS_CUDA(cudaEventSynchronize(ev[7]));
  ctx->stats.ms_host = ev_ms(ev[2], ev[12]);
  ctx->stats.ms_filter = ev_ms(ev[12], ev[3]);
  ctx->stats.ms_exact = ev_ms(ev[3], ev[4]);
  ctx->stats.ms_hitsort = ev_ms(ev[5], ev[6]);
  ctx->stats.ms_total = ev_ms(ev[0], ev[7]);
  return rc;
}

// ---- brute force ------------------------------------------------------------------------
__global__ void build_tq_from_db_kernel(const uint8_t *__restrict__ codes, uint64_t q0, uint32_t nq, int len,
                                        const float *__restrict__ dsq32, const int32_t *__restrict__ metric,
                                        int use_int, float *__restrict__ tq) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t n = (uint64_t)nq * len * HS_AA;
  if (i >= n) return;
  const int c = (int)(i % HS_AA);
  const uint64_t qp = i / HS_AA;
  const int cq = codes[q0 * len + qp];
  tq[i] = use_int ? (float)metric[c * HS_AA + cq] : dsq32[c * HS_AA + cq];
}

static int bruteforce_impl(hs_ctx *ctx, const QueryInput *in, uint32_t Q, hs_hit *hits, uint64_t cap,
                           uint64_t *nhits, void *hits_dev = nullptr) {
  if (ctx->N == 0) {
    set_error("hs_bruteforce: no fragments loaded");
    return HS_ERR_INVALID;
  }
  stats_begin(ctx);
  cudaEvent_t *ev = ctx->ev;
  HS_CUDA(cudaEventRecord(ev[0], ctx->stream));
  HS_TRY(ensure_identity_store(ctx));
  const uint32_t L = ctx->prm.L;
  const bool allpairs = (in == nullptr);
  const uint64_t N = ctx->N;
  const uint64_t dev_cap = std::max<uint64_t>(cap, 1);
  HS_TRY(ctx->d_hits.reserve(sizeof(hs_hit) * dev_cap));
  unsigned long long *hit_count = ctx->d_counters.as<unsigned long long>() + 9;
  HS_CUDA(cudaMemsetAsync(hit_count, 0, sizeof(unsigned long long), ctx->stream));
  if (!allpairs) HS_TRY(stage_queries(ctx, *in, Q));

  const uint64_t total_q = allpairs ? N : Q;
  ctx->hit_qmax = total_q;
  // all pairs: the DB is its own query set, taken in blocks whose filter tables
  // are resident at once; explicit queries: one block
  const uint64_t qblock = allpairs ? (1u << 20) : std::max<uint64_t>(total_q, 1);
  uint64_t ncand = 0, nsurv_total = 0;
  float ms_filter = 0.f, ms_exact = 0.f;
  for (uint64_t q0 = 0; q0 < total_q; q0 += qblock) {
    const uint32_t nq = (uint32_t)std::min<uint64_t>(qblock, total_q - q0);
    HS_TRY(ctx->d_tq.reserve(sizeof(float) * (size_t)nq * ctx->prm.len * HS_AA));
    if (allpairs) {
      const uint64_t n = (uint64_t)nq * ctx->prm.len * HS_AA;
      build_tq_from_db_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
          ctx->d_codes.as<uint8_t>(), q0, nq, (int)ctx->prm.len, ctx->d_dsq32.as<float>(),
          ctx->d_metric.as<int32_t>(), ctx->prm.metric == HS_METRIC_BLOSUM_INT, ctx->d_tq.as<float>());
      ctx->stats.kernel_launches++;
    } else {
      HS_TRY(build_tq(ctx, Q));
    }
    // work items: query chunks x member range
    FilterPlan plan;
    plan_init(ctx, plan);
    std::vector<uint32_t> qall(nq);
    for (uint32_t i = 0; i < nq; ++i) qall[i] = (uint32_t)(q0 + i);  // all pairs: DB ids
    HS_TRY(plan_add(ctx, plan, L, 0u, (uint32_t)N, qall.data(), nq, allpairs));
    ncand += plan.ncand;
    ctx->stats.n_candidates_tc += plan.ncand_tc;
    if (plan.items.empty() && plan.items_tc.empty() && plan.mma_units.empty()) continue;
    HS_CUDA(cudaEventRecord(ev[1], ctx->stream));
    uint64_t nsurv = 0;
    HS_TRY(plan_run(ctx, plan, nq, (uint32_t)q0, allpairs ? kModeAllPairs : kModeBrute, &nsurv));
    HS_CUDA(cudaEventRecord(ev[2], ctx->stream));
    nsurv_total += nsurv;
    ExactArgs ea;
    fill_exact_common(ctx, ea, allpairs ? 0 : Q);
    ea.surv = ctx->d_surv.as<Survivor>();
    ea.nsurv = nsurv;
    ea.mode = allpairs ? kModeAllPairs : kModeBrute;
    if (!allpairs) {
      ea.q64 = ctx->prm.metric == HS_METRIC_EUCLID_FP64 ? ctx->d_q64.as<double>() : nullptr;
      ea.qcodes = ctx->have_qcodes ? ctx->d_qcodes.as<uint8_t>() : nullptr;
      if (ctx->prm.metric == HS_METRIC_EUCLID_FP64) {
        ea.qcodes = ctx->d_qcodes_det.as<uint8_t>();
        ea.qrow = ctx->d_qrow.as<uint8_t>();
      }
    }
    ea.hits = ctx->d_hits.as<hs_hit>();
    ea.hit_cap = dev_cap;
    ea.hit_count = hit_count;
    HS_TRY(launch_exact(ctx, ea));
    HS_CUDA(cudaEventRecord(ev[3], ctx->stream));
    HS_CUDA(cudaEventSynchronize(ev[3]));
    ms_filter += ev_ms(ev[1], ev[2]);
    ms_exact += ev_ms(ev[2], ev[3]);
  }
  ctx->stats.n_candidates = ncand;
  ctx->stats.n_survivors = nsurv_total;
  unsigned long long nh = 0;
  HS_TRY(read_back(ctx, hit_count, &nh, sizeof nh));
  int rc = deliver_hits(ctx, nh, dev_cap, hits, hits_dev, cap, nhits, ev[5], ev[6], (uint32_t)total_q);
  HS_CUDA(cudaEventRecord(ev[7], ctx->stream));
  HS_CUDA(cudaEventSynchronize(ev[7]));
  ctx->stats.ms_filter = ms_filter;
  ctx->stats.ms_exact = ms_exact;
  ctx->stats.ms_hitsort = ev_ms(ev[5], ev[6]);
  ctx->stats.ms_total = ev_ms(ev[0], ev[7]);
  return rc;
}

// Residue codes index every table of the path (projection partial sums, embedding rows, residue-pair
// tables): a byte >= 20 in the DB would read past them.  Counted here, once per load, on the device.
__global__ void validate_codes_kernel(const uint8_t *__restrict__ codes, uint64_t b0, uint64_t b1,
                                      unsigned long long *__restrict__ bad) {
  // [b0, b1) with b0 a multiple of 16: whole 16-byte words, then the ragged tail byte by byte
  const uint64_t nvec = (b1 - b0) / 16;
  const uint4 *src = reinterpret_cast<const uint4 *>(codes + b0);
  unsigned int mine = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(src + i);
    mine |= __vcmpgeu4(v.x, 0x14141414u) | __vcmpgeu4(v.y, 0x14141414u) | __vcmpgeu4(v.z, 0x14141414u) |
            __vcmpgeu4(v.w, 0x14141414u);
  }
  if (blockIdx.x == 0)
    for (uint64_t i = b0 + nvec * 16 + threadIdx.x; i < b1; i += blockDim.x) mine |= codes[i] >= HS_AA ? 1u : 0u;
  if (__any_sync(0xffffffffu, mine != 0) && (threadIdx.x & 31) == 0) atomicAdd(bad, 1ull);
}
constexpr int kBadCodeSlot = 30;  // d_counters slot
// Enqueues the check of fragments [f0, f1) (f0 * len a multiple of 16) on the ctx stream.
static int validate_codes_enqueue(hs_ctx *ctx, uint64_t f0, uint64_t f1) {
  if (f1 <= f0) return HS_OK;
  const uint64_t b0 = f0 * ctx->prm.len, b1 = f1 * ctx->prm.len;
  const unsigned grid = (unsigned)std::min<uint64_t>(((b1 - b0) / 16 + 255) / 256 + 1, (uint64_t)ctx->num_sms * 8);
  validate_codes_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_codes.as<uint8_t>(), b0, b1,
                                                      ctx->d_counters.as<unsigned long long>() + kBadCodeSlot);
  HS_CUDA(cudaGetLastError());
  return HS_OK;
}
static int validate_codes_begin(hs_ctx *ctx) {
  HS_CUDA(cudaMemsetAsync(ctx->d_counters.as<unsigned long long>() + kBadCodeSlot, 0, sizeof(unsigned long long), ctx->stream));
  return HS_OK;
}
// Reads the verdict (synchronises the ctx stream); a bad DB is dropped.
static int validate_codes_end(hs_ctx *ctx, const char *who) {
  unsigned long long bad = 0;
  HS_TRY(read_back(ctx, ctx->d_counters.as<unsigned long long>() + kBadCodeSlot, &bad, sizeof bad));
  if (bad) {
    ctx->N = 0;
    ctx->npad = 0;
    ctx->hashed = ctx->indexed = ctx->have_codes_pm = ctx->have_rec = false;
    set_error("%s: the fragments hold residue codes outside 0..19 (hs_letter_to_code returns -1 for B, J, O, U, X, Z)", who);
    return HS_ERR_INVALID;
  }
  return HS_OK;
}
int protein_id_impl(hs_ctx *ctx, const uint32_t *start_index, uint32_t nstart, const uint32_t *pos, uint64_t n, uint32_t *out);
int validate_codes(hs_ctx *ctx, const char *who) {
  HS_TRY(validate_codes_begin(ctx));
  HS_TRY(validate_codes_enqueue(ctx, 0, ctx->N));
  return validate_codes_end(ctx, who);
}

}  // namespace hs

using namespace hs;

extern "C" {

int hs_device_available(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  for (int d = 0; d < n; ++d) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major == 10) return 1;
  }
  return 0;
}

int hs_create(hs_ctx_t **out, int device, const hs_params *params) {
  if (!out || !params) {
    set_error("hs_create: null argument");
    return HS_ERR_INVALID;
  }
  *out = nullptr;
  if (params->len == 0 || params->len > HS_MAX_LEN || params->K == 0 || params->K > HS_MAX_K || params->L == 0 ||
      params->L > HS_MAX_L) {
    set_error("hs_create: len must be 1..%d, K 1..%d, L 1..%d", HS_MAX_LEN, HS_MAX_K, HS_MAX_L);
    return HS_ERR_UNSUPPORTED;
  }
  if (!(params->W > 0.0) || !(params->R >= 0.0) || params->table_variant > HS_TABLE_PRINT6 ||
      params->metric > HS_METRIC_BLOSUM_INT || params->predicate > HS_PRED_SQRT_LE_R) {
    set_error("hs_create: bad W / R / enum value");
    return HS_ERR_INVALID;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("hs_create: no CUDA device (%s); this library has no CPU fallback",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return HS_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    set_error("hs_create: device %d out of range (0..%d)", device, ndev - 1);
    return HS_ERR_INVALID;
  }
  cudaDeviceProp prop;
  HS_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("hs_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
              prop.minor);
    return HS_ERR_CUDA;
  }
  HS_CUDA(cudaSetDevice(device));
  hs_ctx *ctx = new (std::nothrow) hs_ctx();
  if (!ctx) return HS_ERR_NOMEM;
  ctx->device = device;
  auto env_on = [](const char *name) {
    const char *e = getenv(name);
    return e && atoi(e) != 0;
  };
  ctx->no_pipeline = env_on("HS_NO_PIPELINE");
  ctx->no_load_overlap = env_on("HS_NO_LOAD_OVERLAP");
  ctx->plan_stats = env_on("HS_PLAN_STATS");
  ctx->no_hash_sort = env_on("HS_NO_HASH_SORT");
  ctx->no_lazy_stores = env_on("HS_NO_LAZY_STORES");
  ctx->bypass_max = kBypassMaxCandidates;
  if (const char *e = getenv("HS_BYPASS_MAX")) ctx->bypass_max = strtoull(e, nullptr, 10);
  if (const char *e = getenv("HS_EXACT_REP")) ctx->exact_rep = atoi(e) != 0;
  ctx->force_hash_collision = env_on("HS_FORCE_HASH_COLLISION");
  ctx->force_hash_sort = ctx->force_hash_collision || env_on("HS_FORCE_HASH_SORT");
  ctx->no_mma_filter = env_on("HS_NO_MMA_FILTER");
  ctx->no_mma_int = env_on("HS_NO_MMA_INT");
  ctx->surv_bins = env_on("HS_SURV_BINS");
  if (const char *e = getenv("HS_SEGSORT")) ctx->segsort = atoi(e) != 0;
  if (const char *e = getenv("HS_SEGSORT_MIN")) ctx->segsort_min = strtoull(e, nullptr, 10);
  if (const char *e = getenv("HS_SEGSORT_NBLK")) ctx->segsort_nblk = (uint32_t)std::max(0, atoi(e));
  ctx->segsort_prof = env_on("HS_SEGSORT_PROF");
  ctx->segsort_radix = env_on("HS_SEGSORT_RADIX");
  if (const char *e = getenv("HS_SEGSORT_BUF")) ctx->segsort_buf = (uint32_t)std::max(0, atoi(e));
  if (const char *e = getenv("HS_SELFJOIN_CHUNK"))
    if (atoi(e) >= 256) ctx->selfjoin_chunk = (uint32_t)atoi(e);
  ctx->num_sms = prop.multiProcessorCount;
  ctx->prm = *params;
  ctx->dim = params->len * HS_CDIM;
  ctx->rec_rank_off = (params->len + 1u) & ~1u;
  ctx->rec_stride = (params->len + 15u) & ~15u;
  coordinates_table(params->table_variant, ctx->table64);
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    set_error("hs_create: cudaStreamCreate failed");
    delete ctx;
    return HS_ERR_CUDA;
  }
  for (int i = 0; i < 16; ++i) cudaEventCreate(&ctx->ev[i]);
  int rc = ctx->d_counters.reserve(sizeof(unsigned long long) * 32);
  if (rc == HS_OK) rc = upload_tables(ctx);
  if (rc != HS_OK) {
    hs_destroy(ctx);
    return rc;
  }
  *out = ctx;
  return HS_OK;
}

void hs_destroy(hs_ctx_t *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  comm_destroy(ctx);
  DevBuf *bufs[] = {&ctx->d_table64, &ctx->d_dsq32, &ctx->d_metric, &ctx->d_a64, &ctx->d_a64t, &ctx->d_b64, &ctx->d_T32,
                    &ctx->d_b32, &ctx->d_eps32, &ctx->d_codes, &ctx->d_buckets, &ctx->d_codes_pm, &ctx->d_q64,
                    &ctx->d_qcodes, &ctx->d_qkeys, &ctx->d_qvalid, &ctx->d_qrange, &ctx->d_tq, &ctx->d_work,
                    &ctx->d_qlist, &ctx->d_surv, &ctx->d_hits, &ctx->d_counters, &ctx->d_hit_keys[0],
                    &ctx->d_hit_keys[1], &ctx->d_hit_keys[2], &ctx->d_hit_perm, &ctx->d_hits_sorted,
                    &ctx->d_hits_gathered, &ctx->d_misc, &ctx->d_parent, &ctx->d_tabptrs, &ctx->d_residues,
                    &ctx->d_starts, &ctx->d_metric32, &ctx->d_tq16, &ctx->d_work_tc, &ctx->d_qlist_tc, &ctx->d_large,
                    &ctx->d_qcodes_det, &ctx->d_qrow, &ctx->d_tab16, &ctx->d_qb16, &ctx->d_mma_items, &ctx->d_mma_units, &ctx->d_mma_cta, &ctx->d_qlist_mma, &ctx->d_surv_blk, &ctx->d_lut, &ctx->d_rinfo, &ctx->d_ranks, &ctx->d_rec, &ctx->d_qrank, &ctx->sort.vals_alt, &ctx->sort.tile_hist, &ctx->sort.digit_hist,
                    &ctx->sort.flags, &ctx->sort.block_sums, &ctx->sort.or_and};
  for (DevBuf *b : bufs) b->release();
  for (int w = 0; w < kMaxKeyWords; ++w) {
    ctx->sort.keys_alt[w].release();
    ctx->sort.keys_cur[w].release();
  }
  for (int l = 0; l < HS_MAX_L; ++l) {
    ctx->d_keys[l].release();
    ctx->tables[l].sorted_ids.release();
    ctx->tables[l].ukeys.release();
    ctx->tables[l].bstart.release();
    ctx->tables[l].codes_sorted.release();
  }
  for (int i = 0; i < 16; ++i) cudaEventDestroy(ctx->ev[i]);
  for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->ev_chunk) cudaEventDestroy(e);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  if (ctx->h_up) cudaFreeHost(ctx->h_up);
  ctx->d_hits_sorted_alt.release();
  ctx->d_events.release();
  ctx->d_binctr.release();
  DevBuf *cbufs[] = {&ctx->d_cidt, &ctx->d_cidt_alt, &ctx->d_cdist, &ctx->d_cdist_alt, &ctx->d_coffsets,
                     &ctx->d_seg_tab, &ctx->d_seg_key, &ctx->d_seg_dist, &ctx->d_seg_ctl};
  for (DevBuf *b : cbufs) b->release();
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int hs_get_stream(hs_ctx_t *ctx, void **stream_out) {
  if (!ctx || !stream_out) return HS_ERR_INVALID;
  *stream_out = (void *)ctx->stream;
  return HS_OK;
}

int hs_get_stats(hs_ctx_t *ctx, hs_stats *out) {
  if (!ctx || !out) return HS_ERR_INVALID;
  *out = ctx->stats;
  return HS_OK;
}

int hs_set_coordinates(hs_ctx_t *ctx, const double *table160) {
  if (!ctx || !table160) return HS_ERR_INVALID;
  HS_CUDA(cudaSetDevice(ctx->device));
  memcpy(ctx->table64, table160, sizeof ctx->table64);
  HS_TRY(upload_tables(ctx));
  if (ctx->have_projection) {
    std::vector<double> a = ctx->h_a, b = ctx->h_b;
    HS_TRY(setup_projection(ctx, a.data(), b.data()));
  }
  return HS_OK;
}

int hs_set_projection(hs_ctx_t *ctx, const double *a, const double *b) {
  if (!ctx || !a || !b) {
    set_error("hs_set_projection: null argument");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  return setup_projection(ctx, a, b);
}

static int load_common(hs_ctx *ctx, uint64_t N, uint64_t id_base) {
  if (N >= (1ull << 32) - 4096) {
    set_error("hs_load_fragments: at most 2^32-4096 fragments per GPU (ids are 32-bit)");
    return HS_ERR_UNSUPPORTED;
  }
  ctx->N = N;
  ctx->id_base = id_base;
  ctx->npad = (N + 15) & ~15ull;
  ctx->hashed = false;
  ctx->indexed = false;
  ctx->have_codes_pm = false;
  ctx->have_rec = false;
  return ctx->d_codes.reserve((size_t)N * ctx->prm.len + 64);
}

// hs_load_fragments with a projection already set: the host-to-device copy runs in blocks on a
// second stream and every block is hashed as soon as it has arrived, so the hash hides behind
// the PCIe transfer.  Returns HS_OK with *done = false when the plain path must be taken.
static int load_and_hash_overlapped(hs_ctx *ctx, const uint8_t *codes, uint64_t N, bool *done) {
  *done = false;
  if (!ctx->have_projection || !hash_single_launch_records(ctx) || N < 8 * kHashRangeAlign ||
      (ctx->prm.flags & (HS_FLAG_HASH_EXACT | HS_FLAG_HASH_AUDIT)) || ctx->no_load_overlap)
    return HS_OK;
  const uint32_t nblk = 8, len = ctx->prm.len;
  const uint64_t blk = ((N + nblk - 1) / nblk + kHashRangeAlign - 1) / kHashRangeAlign * kHashRangeAlign;
  if (!ctx->copy_stream) HS_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  while (ctx->ev_chunk.size() < nblk) {
    cudaEvent_t ev;
    HS_CUDA(cudaEventCreate(&ev));
    ctx->ev_chunk.push_back(ev);
  }
  stats_begin(ctx);
  unsigned long long *cnt = ctx->d_counters.as<unsigned long long>();
  HS_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * 8, ctx->stream));
  HS_TRY(validate_codes_begin(ctx));
  HS_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  HS_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev[0], 0));  // earlier work on the ctx stream is done with d_codes
  for (uint64_t f0 = 0, c = 0; f0 < N; f0 += blk, ++c) {
    const uint64_t f1 = std::min(N, f0 + blk);
    HS_CUDA(cudaMemcpyAsync(ctx->d_codes.as<uint8_t>() + f0 * len, codes + f0 * len, (f1 - f0) * len,
                            cudaMemcpyHostToDevice, ctx->copy_stream));
    HS_CUDA(cudaEventRecord(ctx->ev_chunk[c], ctx->copy_stream));
    HS_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[c], 0));
    HS_TRY(validate_codes_enqueue(ctx, f0, f1));
    HS_TRY(launch_hash_fast(ctx, false, f0, f1));
  }
  HS_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
  unsigned long long h[4];
  HS_TRY(read_back(ctx, cnt, h, sizeof h));
  HS_TRY(validate_codes_end(ctx, "hs_load_fragments"));
  ctx->stats.guard_hits = h[0];
  ctx->stats.guard_corrected = h[1];
  ctx->stats.ms_hash = ev_ms(ctx->ev[0], ctx->ev[1]);  // includes the transfer it overlaps
  ctx->stats.ms_total = ctx->stats.ms_hash;
  if (h[2]) {
    set_error("hs_load_fragments: %llu buckets outside the derived range (internal bound violated)", h[2]);
    return HS_ERR_UNSUPPORTED;
  }
  ctx->hashed = true;
  ctx->hash_stats = ctx->stats;
  *done = true;
  return HS_OK;
}

int hs_load_fragments(hs_ctx_t *ctx, const uint8_t *codes, uint64_t N, uint64_t id_base) {
  if (!ctx || (!codes && N)) {
    set_error("hs_load_fragments: null argument");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  HS_TRY(load_common(ctx, N, id_base));
  bool hashed = false;
  HS_TRY(load_and_hash_overlapped(ctx, codes, N, &hashed));
  if (hashed) return HS_OK;
  if (N) HS_CUDA(cudaMemcpyAsync(ctx->d_codes.p, codes, (size_t)N * ctx->prm.len, cudaMemcpyHostToDevice, ctx->stream));
  return validate_codes(ctx, "hs_load_fragments");
}

int hs_load_fragments_dev(hs_ctx_t *ctx, const void *codes_dev, uint64_t N, uint64_t id_base) {
  if (!ctx || (!codes_dev && N)) {
    set_error("hs_load_fragments_dev: null argument");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  HS_TRY(load_common(ctx, N, id_base));
  if (N) HS_CUDA(cudaMemcpyAsync(ctx->d_codes.p, codes_dev, (size_t)N * ctx->prm.len, cudaMemcpyDeviceToDevice, ctx->stream));
  return validate_codes(ctx, "hs_load_fragments_dev");
}

int hs_extract_windows(hs_ctx_t *ctx, const uint8_t *residues, const uint32_t *start_index, uint32_t nprot,
                       uint32_t stride, uint64_t id_base, uint32_t *pos_out, uint64_t pos_cap, uint64_t *nfrag) {
  if (!ctx || !residues || !start_index || !nfrag || stride == 0) {
    set_error("hs_extract_windows: bad argument");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  return extract_windows_impl(ctx, residues, start_index, nprot, stride, id_base, pos_out, pos_cap, nfrag);
}

int hs_protein_id(hs_ctx_t *ctx, const uint32_t *start_index, uint32_t nstart, const uint32_t *pos, uint64_t n,
                  uint32_t *protein_out) {
  if (!ctx || !start_index || nstart == 0 || (n && (!pos || !protein_out))) {
    set_error("hs_protein_id: bad argument");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  return protein_id_impl(ctx, start_index, nstart, pos, n, protein_out);
}

int hs_fragment_name(const char *protein_header, uint32_t protein_index, uint32_t offset, const char *kmer, uint32_t len,
                     uint64_t cnt, char *out, uint64_t out_cap) {
  if (!protein_header || (!kmer && len) || !out || out_cap == 0) {
    set_error("hs_fragment_name: bad argument");
    return HS_ERR_INVALID;
  }
  // `istringstream iss(name); iss >> name;` (protein2datapoints.cpp:61-63): leading white space skipped,
  // the token ends at the next white space
  const char *b = protein_header;
  while (*b == ' ' || *b == '\t' || *b == '\n' || *b == '\v' || *b == '\f' || *b == '\r') ++b;
  const char *e = b;
  while (*e && !(*e == ' ' || *e == '\t' || *e == '\n' || *e == '\v' || *e == '\f' || *e == '\r')) ++e;
  const int need = snprintf(out, (size_t)out_cap, "%.*s#%u$%u@%.*s*%llu", (int)(e - b), b, protein_index, offset, (int)len,
                            kmer ? kmer : "", (unsigned long long)cnt);
  if (need < 0 || (uint64_t)need >= out_cap) {
    set_error("hs_fragment_name: %d characters needed, capacity %llu", need + 1, (unsigned long long)out_cap);
    return HS_ERR_CAPACITY;
  }
  return HS_OK;
}

uint64_t hs_num_fragments(hs_ctx_t *ctx) { return ctx ? ctx->N : 0; }

int hs_hash(hs_ctx_t *ctx, int32_t *buckets_out) {
  if (!ctx) return HS_ERR_INVALID;
  if (!ctx->have_projection) {
    set_error("hs_hash: call hs_set_projection first");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  const uint32_t L = ctx->prm.L, K = ctx->prm.K;
  const uint64_t N = ctx->N;
  stats_begin(ctx);
  if (N == 0) {
    ctx->hashed = true;
    return HS_OK;
  }
  if (!ctx->rank_mode)
    for (uint32_t l = 0; l < L; ++l) HS_TRY(ctx->d_keys[l].reserve(sizeof(uint64_t) * ctx->key_words * N));
  if (buckets_out) HS_TRY(ctx->d_buckets.reserve(sizeof(int32_t) * N * L * K));
  unsigned long long *cnt = ctx->d_counters.as<unsigned long long>();
  HS_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * 8, ctx->stream));
  HS_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  if (ctx->prm.flags & HS_FLAG_HASH_EXACT) HS_TRY(launch_hash_exact(ctx, buckets_out != nullptr, false));
  else HS_TRY(launch_hash_fast(ctx, buckets_out != nullptr));
  HS_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
  if (ctx->prm.flags & HS_FLAG_HASH_AUDIT) HS_TRY(launch_hash_exact(ctx, false, true));
  unsigned long long h[4];
  if (buckets_out)
    HS_CUDA(cudaMemcpyAsync(buckets_out, ctx->d_buckets.p, sizeof(int32_t) * N * L * K, cudaMemcpyDeviceToHost, ctx->stream));
  HS_TRY(read_back(ctx, cnt, h, sizeof h));
  ctx->stats.guard_hits = h[0];
  ctx->stats.guard_corrected = h[1];
  ctx->stats.residual_flips = h[3];
  ctx->stats.ms_hash = ev_ms(ctx->ev[0], ctx->ev[1]);
  ctx->stats.ms_total = ctx->stats.ms_hash;
  if (h[2]) {
    set_error("hs_hash: %llu keys exceeded %u characters (internal bound violated)", h[2], 16 * ctx->key_words);
    return HS_ERR_UNSUPPORTED;
  }
  ctx->hashed = true;
  ctx->indexed = false;
  ctx->hash_stats = ctx->stats;
  return HS_OK;
}

int hs_hash_audit(hs_ctx_t *ctx, uint64_t *residual_flips) {
  if (!ctx || !residual_flips || !ctx->hashed) {
    set_error("hs_hash_audit: null argument or hs_hash not run");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  *residual_flips = 0;
  if (ctx->N == 0) return HS_OK;
  unsigned long long *cnt = ctx->d_counters.as<unsigned long long>();
  HS_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * 8, ctx->stream));
  HS_TRY(launch_hash_exact(ctx, false, true));
  unsigned long long h[4];
  HS_CUDA(cudaMemcpyAsync(h, cnt, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  *residual_flips = h[3];
  ctx->stats.residual_flips = h[3];
  ctx->hash_stats.residual_flips = h[3];
  return HS_OK;
}

int hs_get_keys(hs_ctx_t *ctx, uint32_t table, uint64_t *keys_out) {
  if (!ctx || !keys_out || table >= ctx->prm.L || !ctx->hashed) {
    set_error("hs_get_keys: bad argument or hs_hash not run");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  const uint64_t N = ctx->N;
  const uint32_t KW = ctx->key_words;
  if (ctx->rank_mode) {
    // rank path: the key of a fragment is the key string of its bucket rank
    std::vector<uint16_t> rk(N);
    if (N)
      HS_CUDA(cudaMemcpy(rk.data(), ctx->d_ranks.as<uint16_t>() + (size_t)table * ctx->npad, sizeof(uint16_t) * N,
                         cudaMemcpyDeviceToHost));
    const uint32_t nr = ctx->rank_nr[table];
    const std::vector<uint64_t> &keys = ctx->h_rkeys[table];
    for (uint64_t i = 0; i < N; ++i)
      for (uint32_t w = 0; w < KW; ++w) keys_out[i * KW + w] = keys[(size_t)w * nr + rk[i]];
    return HS_OK;
  }
  std::vector<uint64_t> tmp((size_t)KW * N);
  if (N) HS_CUDA(cudaMemcpy(tmp.data(), ctx->d_keys[table].p, sizeof(uint64_t) * KW * N, cudaMemcpyDeviceToHost));
  for (uint64_t i = 0; i < N; ++i)
    for (uint32_t w = 0; w < KW; ++w) keys_out[i * KW + w] = tmp[(size_t)w * N + i];
  return HS_OK;
}

int hs_build_index(hs_ctx_t *ctx) {
  if (!ctx) return HS_ERR_INVALID;
  HS_CUDA(cudaSetDevice(ctx->device));
  const bool ran_hash = !ctx->hashed;
  if (ran_hash) HS_TRY(hs_hash(ctx, nullptr));
  hs_stats hash_stats = ctx->hash_stats;  // of hs_hash, or of the overlapped hash of hs_load_fragments
  if (!ran_hash) {  // an earlier call already reported them
    hash_stats.kernel_launches = 0;
    hash_stats.ms_hash = 0.f;
  }
  stats_begin(ctx);
  ctx->stats.guard_hits = hash_stats.guard_hits;
  ctx->stats.guard_corrected = hash_stats.guard_corrected;
  ctx->stats.residual_flips = hash_stats.residual_flips;
  ctx->stats.kernel_launches = hash_stats.kernel_launches;
  const uint32_t L = ctx->prm.L;
  float ms_sort = 0.f, ms_group = 0.f, ms_permute = 0.f;
  cudaEvent_t *ev = ctx->ev;
  HS_CUDA(cudaEventRecord(ev[8], ctx->stream));
  ctx->hash_sort_choice = -1;
  if (ctx->N) {
    for (uint32_t l = 0; l < L; ++l) {
      HS_CUDA(cudaEventRecord(ev[2], ctx->stream));
      HS_TRY(build_table_index(ctx, l, ev[3], ev[4], false));
      HS_CUDA(cudaEventSynchronize(ev[4]));
      ms_sort += ev_ms(ev[2], ev[3]);
      ms_group += ev_ms(ev[3], ev[4]);
    }
    // bucket-order code stores: one L2-blocked gather for all tables, else table by table;
    // left to the first user when the buckets are tiny (ensure_code_stores)
    HS_CUDA(cudaEventRecord(ev[2], ctx->stream));
    uint64_t nb_sum = 0;
    for (uint32_t l = 0; l < L; ++l) nb_sum += ctx->tables[l].nb;
    ctx->stores_built = false;
    const bool lazy = !ctx->no_lazy_stores && ctx->N >= (1u << 20) && nb_sum * 4 > (uint64_t)L * ctx->N &&
                      ctx->prm.metric == HS_METRIC_EUCLID_FP64;
    if (!lazy) HS_TRY(ensure_code_stores(ctx));
    else HS_TRY(ensure_records(ctx));   // (the exact stage reads the fragment records; the store build makes them otherwise)
    HS_CUDA(cudaEventRecord(ev[5], ctx->stream));
    HS_CUDA(cudaEventSynchronize(ev[5]));
    ms_permute = ev_ms(ev[2], ev[5]);
  } else {
    for (uint32_t l = 0; l < L; ++l) ctx->tables[l].nb = ctx->tables[l].nslots = 0;
    ctx->stores_built = true;
  }
  HS_CUDA(cudaEventRecord(ev[9], ctx->stream));
  HS_CUDA(cudaEventSynchronize(ev[9]));
  HS_TRY(upload_table_pointers(ctx));
  ctx->stats.ms_hash = hash_stats.ms_hash;
  ctx->stats.ms_sort = ms_sort;
  ctx->stats.ms_group = ms_group;
  ctx->stats.ms_permute = ms_permute;
  ctx->stats.ms_total = hash_stats.ms_hash + ev_ms(ev[8], ev[9]);
  ctx->indexed = true;
  return HS_OK;
}

int hs_table_sizes(hs_ctx_t *ctx, uint64_t *sizes_out) {
  if (!ctx || !sizes_out || !ctx->indexed) {
    set_error("hs_table_sizes: index not built");
    return HS_ERR_INVALID;
  }
  for (uint32_t l = 0; l < ctx->prm.L; ++l) sizes_out[l] = ctx->tables[l].nb;
  return HS_OK;
}

int hs_get_table(hs_ctx_t *ctx, uint32_t table, uint32_t *ids_out, uint32_t *starts_out) {
  if (!ctx || !ctx->indexed || table >= ctx->prm.L) {
    set_error("hs_get_table: index not built or bad table");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  const TableIndex &T = ctx->tables[table];
  if (ids_out && ctx->N) HS_CUDA(cudaMemcpy(ids_out, T.sorted_ids.p, sizeof(uint32_t) * ctx->N, cudaMemcpyDeviceToHost));
  if (starts_out && ctx->N) {
    if (T.nslots == T.nb) {
      HS_CUDA(cudaMemcpy(starts_out, T.bstart.p, sizeof(uint32_t) * (T.nb + 1), cudaMemcpyDeviceToHost));
    } else {  // rank path: drop the empty bucket slots
      std::vector<uint32_t> all(T.nslots + 1);
      HS_CUDA(cudaMemcpy(all.data(), T.bstart.p, sizeof(uint32_t) * (T.nslots + 1), cudaMemcpyDeviceToHost));
      uint64_t o = 0;
      for (uint64_t i = 0; i < T.nslots; ++i)
        if (all[i + 1] != all[i]) starts_out[o++] = all[i];
      starts_out[o] = all[T.nslots];
    }
  }
  return HS_OK;
}

int hs_search_points(hs_ctx_t *ctx, const double *qpoints, uint32_t Q, hs_hit *hits, uint64_t cap, uint64_t *nhits) {
  if (!ctx || (!qpoints && Q) || (!hits && cap) || !nhits) {
    set_error("hs_search_points: null argument");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  QueryInput in;
  in.h_points = qpoints;
  return search_impl(ctx, in, Q, hits, nullptr, cap, nhits);
}

int hs_search_codes(hs_ctx_t *ctx, const uint8_t *qcodes, uint32_t Q, hs_hit *hits, uint64_t cap, uint64_t *nhits) {
  if (!ctx || (!qcodes && Q) || (!hits && cap) || !nhits) {
    set_error("hs_search_codes: null argument");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  QueryInput in;
  in.h_codes = qcodes;
  return search_impl(ctx, in, Q, hits, nullptr, cap, nhits);
}

int hs_search_points_compact(hs_ctx_t *ctx, const double *qpoints, uint32_t Q, hs_compact_hits *out, uint64_t *nhits) {
  if (!ctx || (!qpoints && Q) || !out || !out->offsets || (out->cap && (!out->idt || !out->dist2)) || !nhits) {
    set_error("hs_search_points_compact: null argument");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  QueryInput in;
  in.h_points = qpoints;
  return search_impl(ctx, in, Q, nullptr, nullptr, out->cap, nhits, out);
}

int hs_expand_hits(const hs_compact_hits *in, uint32_t Q, uint64_t id_base, hs_hit *hits_out) {
  if (!in || !in->offsets || (in->offsets[Q] && (!in->idt || !in->dist2 || !hits_out)) || in->id_bits == 0 || in->id_bits > 31) {
    set_error("hs_expand_hits: bad argument");
    return HS_ERR_INVALID;
  }
  const uint32_t id_mask = (1u << in->id_bits) - 1u;
  for (uint32_t q = 0; q < Q; ++q)
    for (uint64_t i = in->offsets[q]; i < in->offsets[q + 1]; ++i) {
      hs_hit h;
      h.query = q;
      h.table_first = in->idt[i] >> in->id_bits;
      h.db_id = id_base + (in->idt[i] & id_mask);
      h.dist2 = in->dist2[i];
      hits_out[i] = h;
    }
  return HS_OK;
}

int hs_search_points_dev(hs_ctx_t *ctx, const void *qpoints_dev, uint32_t Q, void *hits_dev, uint64_t cap,
                         uint64_t *nhits) {
  if (!ctx || (!qpoints_dev && Q) || (!hits_dev && cap) || !nhits) {
    set_error("hs_search_points_dev: null argument");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  QueryInput in;
  in.d_points = qpoints_dev;
  return search_impl(ctx, in, Q, nullptr, hits_dev, cap, nhits);
}

int hs_bruteforce_codes(hs_ctx_t *ctx, const uint8_t *qcodes, uint32_t Q, hs_hit *hits, uint64_t cap,
                        uint64_t *nhits) {
  if (!ctx || (!hits && cap) || !nhits) {
    set_error("hs_bruteforce_codes: null argument");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  if (!qcodes) return bruteforce_impl(ctx, nullptr, 0, hits, cap, nhits);
  QueryInput in;
  in.h_codes = qcodes;
  return bruteforce_impl(ctx, &in, Q, hits, cap, nhits);
}

int hs_bruteforce_points(hs_ctx_t *ctx, const double *qpoints, uint32_t Q, hs_hit *hits, uint64_t cap,
                         uint64_t *nhits) {
  if (!ctx || (!qpoints && Q) || (!hits && cap) || !nhits) {
    set_error("hs_bruteforce_points: null argument");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  QueryInput in;
  in.h_points = qpoints;
  return bruteforce_impl(ctx, &in, Q, hits, cap, nhits);
}

int hs_bruteforce_points_dev(hs_ctx_t *ctx, const void *qpoints_dev, uint32_t Q, void *hits_dev, uint64_t cap,
                             uint64_t *nhits) {
  if (!ctx || (!qpoints_dev && Q) || (!hits_dev && cap) || !nhits) {
    set_error("hs_bruteforce_points_dev: null argument");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  QueryInput in;
  in.d_points = qpoints_dev;
  return bruteforce_impl(ctx, &in, Q, nullptr, cap, nhits, hits_dev);
}

int hs_cluster(hs_ctx_t *ctx, uint32_t *label_out) {
  if (!ctx || !label_out) {
    set_error("hs_cluster: null argument");
    return HS_ERR_INVALID;
  }
  if (!ctx->indexed) {
    set_error("hs_cluster: call hs_build_index first");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  HS_TRY(ensure_code_stores(ctx));
  return cluster_impl(ctx, label_out);
}

int hs_greedy_cluster(hs_ctx_t *ctx, uint32_t *center_out, uint32_t *round_out, uint8_t *state_out) {
  if (!ctx || !center_out) {
    set_error("hs_greedy_cluster: null argument");
    return HS_ERR_INVALID;
  }
  if (!ctx->indexed) {
    set_error("hs_greedy_cluster: call hs_build_index first");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  return greedy_cluster_impl(ctx, center_out, round_out, state_out);
}

int hs_union_find(hs_ctx_t *ctx, uint32_t n, const uint32_t *eu, const uint32_t *ev, uint64_t ne, uint32_t *label_out) {
  if (!ctx || (n && !label_out) || (ne && (!eu || !ev))) {
    set_error("hs_union_find: null argument");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  stats_begin(ctx);
  return union_find_impl(ctx, n, eu, ev, ne, label_out);
}
}
