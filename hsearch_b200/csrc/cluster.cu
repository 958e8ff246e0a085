// K5: near-pair union-find clustering (U1 of SURVEY.md 8a).
//
// Reference pieces: UnionFind::FindRoot / JoinUnion, pcluster/src/pcluster/
// union_find.cpp:16-33 (never reached by the reference's main, so the cluster
// is the composition SURVEY.md 8c defines): for every table, every pair of
// fragments sharing a bucket whose distance is within R (PairwiseDistance with
// sqrt, hclust2.cpp:64-71,119-120, or the integer DistanceScore) is an edge;
// the clusters are the connected components, labelled by their smallest id.
// Root identity in the reference depends on edge order; the partition does
// not, and min-id hooking makes the device result order-independent.
//
// Small buckets (<= kSmallBucket members) are enumerated pair by pair by one
// warp each; large buckets go through the tiled filter of verify.cu.
#include <algorithm>
#include <vector>

#include "hash.cuh"
#include "internal.cuh"

namespace hs {

constexpr uint32_t kSmallBucket = 64;
constexpr uint32_t kSelfJoinTail = 512;  // members at the end of a tensor-joined bucket left to the scalar self-join

__global__ void iota32_kernel(uint32_t *p, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (uint32_t)i;
}

// warp per bucket: all pairs (i<j) of a small bucket through the filter bound
__global__ void small_bucket_pairs_kernel(const uint32_t *__restrict__ bstart, uint64_t nb, uint32_t table,
                                          const uint8_t *__restrict__ store, uint64_t npad, int len,
                                          const float *__restrict__ pair32 /* [20][20] */, float thr,
                                          Survivor *__restrict__ surv, unsigned long long surv_cap,
                                          unsigned long long *__restrict__ surv_count,
                                          uint2 *__restrict__ large, unsigned int *__restrict__ nlarge,
                                          unsigned long long *__restrict__ npairs, uint32_t part, uint32_t nparts) {
  const uint64_t b = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= nb) return;
  const uint32_t s = bstart[b], e = bstart[b + 1];
  const uint32_t n = e - s;
  if (n < 2) return;
  if (n > kSmallBucket) {   // (every part collects all large buckets and takes its share of them later)
    if (lane == 0) {
      const unsigned int i = atomicAdd(nlarge, 1u);
      large[i] = make_uint2(s, e);
    }
    return;
  }
  if (b % nparts != part) return;  // pair work split over the ranks of a communicator (replicated DB)
  const uint32_t np = n * (n - 1) / 2;
  if (lane == 0) atomicAdd(npairs, (unsigned long long)np);
  for (uint32_t p = lane; p < np; p += 32) {
    // p -> (i, j), i < j, row-major over the strict upper triangle
    uint32_t i = 0, rem = p, row = n - 1;
    while (rem >= row) {
      rem -= row;
      --row;
      ++i;
    }
    const uint32_t j = i + 1 + rem;
    float d = 0.f;
    for (int q = 0; q < len; ++q) {
      const int ci = store[(uint64_t)q * npad + s + i] / kCodeScale;
      const int cj = store[(uint64_t)q * npad + s + j] / kCodeScale;
      d += __ldg(pair32 + ci * HS_AA + cj);
    }
    if (d <= thr) {
      const unsigned long long idx = atomicAdd(surv_count, 1ull);
      if (idx < surv_cap) {
        Survivor sv;
        sv.query = s + i;
        sv.table = table;
        sv.pos = s + j;
        sv.pad = 0;
        surv[idx] = sv;
      }
    }
  }
}

__global__ void uf_flatten_kernel(const uint32_t *__restrict__ parent, uint64_t n, uint32_t *__restrict__ label) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x = (uint32_t)i;
  while (parent[x] != x) x = parent[x];
  label[i] = x;
}

// Labels of the other ranks' partial clusterings (each: smallest id of the fragment's component under
// that rank's share of the pairs) united into this rank's forest.
__global__ void uf_merge_labels_kernel(uint32_t *__restrict__ parent, const uint32_t *__restrict__ labels, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t l = labels[i];
  if (l != (uint32_t)i) uf_union(parent, (uint32_t)i, l);
}

int comm_allgather_u32(hs_ctx *ctx, const uint32_t *d_send, uint32_t *d_recv, uint64_t n);

// On a context that joined a communicator every rank holds the SAME database and index, and the
// pair work -- what the run time consists of: 2.4e14 pairs at 50 M fragments -- is split: small and
// medium buckets by bucket number, the query chunks of the large buckets round-robin.  Each rank
// unites the edges of its share; the flattened labels (4 bytes per fragment and rank) are
// all-gathered with NCCL and every rank unites the others' labels into its forest: all ranks end
// with the labels of the complete edge set.  No bucket, key or code leaves its GPU.
int cluster_impl(hs_ctx *ctx, uint32_t *label_out) {
  const uint64_t N = ctx->N;
  const uint32_t L = ctx->prm.L;
  const uint32_t nparts = ctx->nranks > 1 ? (uint32_t)ctx->nranks : 1u, part = ctx->nranks > 1 ? (uint32_t)ctx->rank : 0u;
  uint64_t work_no = 0;   // large-bucket chunks and medium buckets, numbered alike on every rank
  stats_begin(ctx);
  if (N == 0) return HS_OK;
  cudaEvent_t *ev = ctx->ev;
  HS_CUDA(cudaEventRecord(ev[0], ctx->stream));
  HS_TRY(ctx->d_parent.reserve(sizeof(uint32_t) * 2 * N));
  uint32_t *parent = ctx->d_parent.as<uint32_t>();
  iota32_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(parent, N);
  ctx->stats.kernel_launches++;
  unsigned long long *counters = ctx->d_counters.as<unsigned long long>();
  unsigned long long *edge_count = counters + 10, *npairs = counters + 11;
  unsigned int *nlarge = reinterpret_cast<unsigned int *>(counters + 12);
  HS_CUDA(cudaMemsetAsync(edge_count, 0, sizeof(unsigned long long) * 2, ctx->stream));
  const float thr = filter_threshold(ctx);
  const float *pair32 = ctx->prm.metric == HS_METRIC_BLOSUM_INT ? ctx->d_metric32.as<float>() : ctx->d_dsq32.as<float>();
  uint64_t ncand = 0, nsurv_total = 0;
  float ms_filter = 0.f, ms_exact = 0.f;

  for (uint32_t l = 0; l < L; ++l) {
    const TableIndex &T = ctx->tables[l];
    if (T.nslots == 0) continue;
    HS_CUDA(cudaEventRecord(ev[1], ctx->stream));
    // --- small buckets on the device, large ones collected ---
    HS_TRY(ctx->d_large.reserve(sizeof(uint2) * (N / kSmallBucket + 2)));
    unsigned long long h_cnt = 0;
    unsigned int h_nlarge = 0;
    unsigned long long *scnt = counters + 8;
    for (int attempt = 0;; ++attempt) {
      if (ctx->d_surv.cap < sizeof(Survivor) * (1u << 20)) HS_TRY(ctx->d_surv.reserve(sizeof(Survivor) * (1u << 22)));
      HS_CUDA(cudaMemsetAsync(scnt, 0, sizeof(unsigned long long), ctx->stream));
      HS_CUDA(cudaMemsetAsync(nlarge, 0, sizeof(unsigned int), ctx->stream));
      HS_CUDA(cudaMemsetAsync(npairs, 0, sizeof(unsigned long long), ctx->stream));
      const uint64_t nthreads = T.nslots * 32;
      small_bucket_pairs_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, ctx->stream>>>(
          T.bstart.as<uint32_t>(), T.nslots, l, T.codes_sorted.as<uint8_t>(), ctx->npad, (int)ctx->prm.len, pair32, thr,
          ctx->d_surv.as<Survivor>(), ctx->d_surv.cap / sizeof(Survivor), scnt, ctx->d_large.as<uint2>(), nlarge,
          npairs, part, nparts);
      HS_CUDA(cudaGetLastError());
      ctx->stats.kernel_launches++;
      HS_CUDA(cudaMemcpyAsync(&h_cnt, scnt, sizeof h_cnt, cudaMemcpyDeviceToHost, ctx->stream));
      HS_CUDA(cudaMemcpyAsync(&h_nlarge, nlarge, sizeof h_nlarge, cudaMemcpyDeviceToHost, ctx->stream));
      HS_CUDA(cudaStreamSynchronize(ctx->stream));
      if (h_cnt <= ctx->d_surv.cap / sizeof(Survivor)) break;
      if (attempt >= 2) {
        set_error("hs_cluster: survivor buffer kept overflowing");
        return HS_ERR_NOMEM;
      }
      HS_TRY(ctx->d_surv.reserve(sizeof(Survivor) * (size_t)(h_cnt + h_cnt / 8 + 1024)));
    }
    unsigned long long h_np = 0;
    HS_CUDA(cudaMemcpy(&h_np, npairs, sizeof h_np, cudaMemcpyDeviceToHost));
    ncand += h_np;
    HS_CUDA(cudaEventRecord(ev[2], ctx->stream));

    ExactArgs ea;
    fill_exact_common(ctx, ea, 0);
    ea.mode = kModeSelfJoin;
    ea.parent = parent;
    ea.edge_count = edge_count;
    ea.surv = ctx->d_surv.as<Survivor>();
    ea.nsurv = h_cnt;
    HS_TRY(launch_exact(ctx, ea));
    nsurv_total += h_cnt;
    HS_CUDA(cudaEventRecord(ev[3], ctx->stream));

    // --- large buckets through the tiled filter ---
    float ms_f2 = 0.f, ms_e2 = 0.f;
    if (h_nlarge) {
      std::vector<uint2> large(h_nlarge);
      HS_CUDA(cudaMemcpy(large.data(), ctx->d_large.p, sizeof(uint2) * h_nlarge, cudaMemcpyDeviceToHost));
      std::sort(large.begin(), large.end(), [](const uint2 &a, const uint2 &b) { return a.x < b.x; });
      // buckets of >= 8192 members: all pairs through the tensor filter (queries = the bucket's
      // own members), one bucket at a time; the rest through the tiled scalar self-join below
      {
        std::vector<uint2> rest;
        for (const uint2 &bk : large) {
          // in chunks of ctx->selfjoin_chunk query members: the survivors of a chunk are verified and
          // united before the next chunk is filtered, so the survivor buffer stays bounded
          // whatever the bucket size (50 M fragments: buckets of millions of members)
          // The last kSelfJoinTail members pair among themselves on the scalar self-join (too few
          // members are left after them for a tensor tile); every tensor chunk keeps >= 1024 queries.
          bool used = bk.y - bk.x >= 4 * kSelfJoinTail && selfjoin_uses_mma(ctx, bk.y - bk.x);
          const uint32_t chunk = std::max<uint32_t>(ctx->selfjoin_chunk, 1024u);
          const uint32_t tail_lo = bk.y - kSelfJoinTail;
          for (uint32_t q_lo = bk.x; used && q_lo < tail_lo;) {
            uint32_t q_hi = (uint32_t)std::min<uint64_t>((uint64_t)q_lo + chunk, tail_lo);
            if (tail_lo - q_hi < 1024u) q_hi = tail_lo;
            if ((work_no++ % nparts) != part) {   // another rank's chunk
              q_lo = q_hi;
              continue;
            }
            uint64_t nsurv = 0, np = 0;
            HS_CUDA(cudaEventRecord(ev[4], ctx->stream));
            HS_TRY(selfjoin_bucket_mma(ctx, l, bk.x, bk.y, q_lo, q_hi, &nsurv, &np, &used));
            if (!used) {
              set_error("hs_cluster: the tensor self-join refused a bucket it had accepted");
              return HS_ERR_UNSUPPORTED;
            }
            HS_CUDA(cudaEventRecord(ev[5], ctx->stream));
            ncand += np;
            ctx->stats.n_candidates_tc += np;
            ea.surv = ctx->d_surv.as<Survivor>();
            ea.nsurv = nsurv;
            ea.qlist_mma = ctx->d_qlist_mma.as<uint32_t>();
            HS_TRY(launch_exact(ctx, ea));
            HS_CUDA(cudaEventRecord(ev[6], ctx->stream));
            HS_CUDA(cudaEventSynchronize(ev[6]));
            nsurv_total += nsurv;
            ms_f2 += ev_ms(ev[4], ev[5]);
            ms_e2 += ev_ms(ev[5], ev[6]);
            q_lo = q_hi;
          }
          rest.push_back(used ? make_uint2(tail_lo, bk.y) : bk);
        }
        large.swap(rest);
      }
      std::vector<WorkItem> items;
      uint32_t nblocks = 0;
      size_t bi = 0;
      while (bi < large.size()) {
        items.clear();
        nblocks = 0;
        // batch buckets until the block budget is reached
        for (; bi < large.size(); ++bi) {
          const uint32_t ms = large[bi].x, me = large[bi].y;
          if (((work_no + bi) % nparts) != part) continue;   // another rank's bucket
          uint64_t need = 0;
          for (uint32_t qb = ms; qb + 1 < me; qb += kQueriesPerItem)
            need += ((uint64_t)me - ((qb + 1) & ~3u) + kFilterTile - 1) / kFilterTile;
          if (!items.empty() && (uint64_t)nblocks + need > 0x3fffffffull) break;
          if (need > 0x7fffffffull) {
            set_error("hs_cluster: bucket of %u members is too large to join", me - ms);
            return HS_ERR_UNSUPPORTED;
          }
          for (uint32_t qb = ms; qb + 1 < me; qb += kQueriesPerItem) {
            WorkItem it;
            it.table = l;
            it.m_begin = qb + 1;
            it.m_end = me;
            it.q_begin = qb;
            it.q_end = std::min<uint32_t>(qb + kQueriesPerItem, me - 1);
            it.block_begin = nblocks;
            nblocks += (it.m_end - (it.m_begin & ~3u) + kFilterTile - 1) / kFilterTile;
            items.push_back(it);
          }
          ncand += (uint64_t)(me - ms) * (me - ms - 1) / 2;
        }
        HS_TRY(ctx->d_work.reserve(sizeof(WorkItem) * items.size()));
        HS_CUDA(cudaMemcpyAsync(ctx->d_work.p, items.data(), sizeof(WorkItem) * items.size(), cudaMemcpyHostToDevice,
                                ctx->stream));
        HS_CUDA(cudaEventRecord(ev[4], ctx->stream));
        FilterArgs fa;
        memset(&fa, 0, sizeof fa);
        fa.items = ctx->d_work.as<WorkItem>();
        fa.nitems = (uint32_t)items.size();
        fa.dsq32 = pair32;
        fa.stores = dev_stores(ctx);
        fa.npad = ctx->npad;
        fa.len = (int)ctx->prm.len;
        fa.thr = thr;
        uint64_t nsurv = 0;
        HS_TRY(run_filter(ctx, fa, nblocks, kModeSelfJoin, &nsurv));
        HS_CUDA(cudaEventRecord(ev[5], ctx->stream));
        ea.surv = ctx->d_surv.as<Survivor>();
        ea.nsurv = nsurv;
        HS_TRY(launch_exact(ctx, ea));
        HS_CUDA(cudaEventRecord(ev[6], ctx->stream));
        HS_CUDA(cudaEventSynchronize(ev[6]));
        nsurv_total += nsurv;
        ms_f2 += ev_ms(ev[4], ev[5]);
        ms_e2 += ev_ms(ev[5], ev[6]);
      }
      work_no += large.size();
    }
    HS_CUDA(cudaEventSynchronize(ev[3]));
    ms_filter += ev_ms(ev[1], ev[2]) + ms_f2;
    ms_exact += ev_ms(ev[2], ev[3]) + ms_e2;
  }
  uint32_t *label = parent + N;
  uf_flatten_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(parent, N, label);
  ctx->stats.kernel_launches++;
  HS_CUDA(cudaGetLastError());
  if (nparts > 1) {
    // the exchange step: every rank's labels to every rank, united into the local forest
    HS_TRY(ctx->d_misc.reserve(sizeof(uint32_t) * (size_t)N * nparts));
    uint32_t *all = ctx->d_misc.as<uint32_t>();
    HS_TRY(comm_allgather_u32(ctx, label, all, N));
    for (uint32_t r = 0; r < nparts; ++r) {
      if (r == part) continue;
      uf_merge_labels_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(parent, all + (size_t)r * N, N);
      ctx->stats.kernel_launches++;
    }
    uf_flatten_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(parent, N, label);
    ctx->stats.kernel_launches++;
    HS_CUDA(cudaGetLastError());
  }
  unsigned long long h_edges = 0;
  HS_CUDA(cudaMemcpyAsync(&h_edges, edge_count, sizeof h_edges, cudaMemcpyDeviceToHost, ctx->stream));
  HS_CUDA(cudaMemcpyAsync(label_out, label, sizeof(uint32_t) * N, cudaMemcpyDeviceToHost, ctx->stream));
  HS_CUDA(cudaEventRecord(ev[7], ctx->stream));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->stats.n_candidates = ncand;
  ctx->stats.n_survivors = nsurv_total;
  ctx->stats.n_edges = h_edges;
  ctx->stats.ms_filter = ms_filter;
  ctx->stats.ms_exact = ms_exact;
  ctx->stats.ms_total = ev_ms(ev[0], ev[7]);
  return HS_OK;
}


// ---- union-find over an explicit edge list (cross-shard cluster) ---------
__global__ void uf_edges_kernel(const uint32_t *__restrict__ eu, const uint32_t *__restrict__ ev, uint64_t ne,
                                uint32_t n, uint32_t *__restrict__ parent, unsigned int *__restrict__ bad) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ne) return;
  const uint32_t u = eu[i], v = ev[i];
  if (u >= n || v >= n) {
    *bad = 1u;
    return;
  }
  if (u != v) uf_union(parent, u, v);
}

// FindRoot / JoinUnion (union_find.cpp:16-33) over ne edges between n ids; label_out[i] =
// smallest id of i's component (the partition does not depend on the edge order).
int union_find_impl(hs_ctx *ctx, uint32_t n, const uint32_t *eu, const uint32_t *ev, uint64_t ne, uint32_t *label_out) {
  if (n == 0) return HS_OK;
  HS_TRY(ctx->d_parent.reserve(sizeof(uint32_t) * 2 * (size_t)n));
  HS_TRY(ctx->d_misc.reserve(sizeof(uint32_t) * 2 * (size_t)std::max<uint64_t>(ne, 1) + 16));
  uint32_t *parent = ctx->d_parent.as<uint32_t>(), *label = parent + n;
  uint32_t *d_eu = ctx->d_misc.as<uint32_t>(), *d_ev = d_eu + ne;
  unsigned int *bad = reinterpret_cast<unsigned int *>(ctx->d_counters.as<unsigned long long>() + 12);
  HS_CUDA(cudaMemsetAsync(bad, 0, sizeof(unsigned int), ctx->stream));
  iota32_kernel<<<(unsigned)(((uint64_t)n + 255) / 256), 256, 0, ctx->stream>>>(parent, n);
  if (ne) {
    HS_CUDA(cudaMemcpyAsync(d_eu, eu, sizeof(uint32_t) * ne, cudaMemcpyHostToDevice, ctx->stream));
    HS_CUDA(cudaMemcpyAsync(d_ev, ev, sizeof(uint32_t) * ne, cudaMemcpyHostToDevice, ctx->stream));
    uf_edges_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, ctx->stream>>>(d_eu, d_ev, ne, n, parent, bad);
  }
  uf_flatten_kernel<<<(unsigned)(((uint64_t)n + 255) / 256), 256, 0, ctx->stream>>>(parent, n, label);
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches += 3;
  unsigned int h_bad = 0;
  HS_CUDA(cudaMemcpyAsync(&h_bad, bad, sizeof h_bad, cudaMemcpyDeviceToHost, ctx->stream));
  HS_CUDA(cudaMemcpyAsync(label_out, label, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h_bad) {
    set_error("hs_union_find: an edge endpoint is >= n");
    return HS_ERR_INVALID;
  }
  return HS_OK;
}

// ---- CL1: greedy centre clustering (hclust2.cpp:86-151, hclust3.cpp:87-152) -----------
// L rounds; round l walks the buckets of table l.  Inside a bucket (members in ascending
// id = insertion order) the centres are the members already marked 1, then every
// unprocessed member (0), in order, joins the first centre within R (sqrt predicate,
// hclust2.cpp:64-71,119-120) or becomes a candidate centre itself.  Points that joined (2)
// leave all later rounds.  A live point sits in exactly one bucket per round and a bucket
// touches only its own members, so buckets are independent: one warp per bucket, the
// reference's sequential order kept inside the bucket (32 members at a time: each lane
// scans the existing centres, then the still-unmatched lanes are resolved in lane order).
constexpr int kGreedyThreads = 128;

template <int NV>
__device__ __forceinline__ void greedy_load(const uint8_t *rec, uint32_t RS, uint32_t id, uint32_t (&w)[4 * NV]) {
  const uint4 *src = reinterpret_cast<const uint4 *>(rec + (uint64_t)id * RS);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const uint4 r = __ldg(src + v);
    w[4 * v + 0] = r.x;
    w[4 * v + 1] = r.y;
    w[4 * v + 2] = r.z;
    w[4 * v + 3] = r.w;
  }
}

// PairwiseDistance(a, b) <= R: FP32 bound first (provably never rejects a pair the FP64
// computation accepts, filter_threshold()), then the reference's FP64 sum and sqrt.
template <int NV>
__device__ __forceinline__ bool greedy_near(const uint32_t (&x)[4 * NV], const uint32_t (&c)[4 * NV], int len,
                                            const float *s_d32, const double *s_sq, float thr, double R) {
  float f = 0.f;
#pragma unroll
  for (int p = 0; p < 16 * NV; ++p)
    if (p < len) f += s_d32[((c[p >> 2] >> (8 * (p & 3))) & 0xff) * HS_AA + ((x[p >> 2] >> (8 * (p & 3))) & 0xff)];
  if (f > thr) return false;
  double dis = 0.0;
#pragma unroll
  for (int p = 0; p < 16 * NV; ++p)
    if (p < len) {
      const int xc = (x[p >> 2] >> (8 * (p & 3))) & 0xff, cc = (c[p >> 2] >> (8 * (p & 3))) & 0xff;
      const double *row = s_sq + (cc * HS_AA + xc) * HS_CDIM;
#pragma unroll
      for (int j = 0; j < HS_CDIM; ++j) dis = __dadd_rn(dis, row[j]);
    }
  return !(sqrt(dis) > R);
}

template <int NV>
__global__ void __launch_bounds__(kGreedyThreads)
greedy_round_kernel(const uint32_t *__restrict__ bstart, uint64_t nslots, const uint32_t *__restrict__ ids,
                    const uint8_t *__restrict__ rec, uint32_t RS, int len, const double *__restrict__ table64,
                    const float *__restrict__ dsq32, float thr, double R, uint32_t round, uint8_t *__restrict__ merged,
                    uint32_t *__restrict__ center_of, uint32_t *__restrict__ round_of,
                    uint32_t *__restrict__ centers /* [N] scratch, one slice per bucket */) {
  __shared__ double s_sq[HS_AA * HS_AA * HS_CDIM];  // (table[x][j] - table[c][j])^2, r * r as in hclust2.cpp:67-68
  __shared__ float s_d32[HS_AA * HS_AA];
  for (int i = threadIdx.x; i < HS_AA * HS_AA * HS_CDIM; i += blockDim.x) {
    const int j = i % HS_CDIM, pair = i / HS_CDIM;
    const int cc = pair / HS_AA, xc = pair - cc * HS_AA;
    const double r = __dsub_rn(table64[xc * HS_CDIM + j], table64[cc * HS_CDIM + j]);
    s_sq[i] = __dmul_rn(r, r);
  }
  for (int i = threadIdx.x; i < HS_AA * HS_AA; i += blockDim.x) s_d32[i] = dsq32[i];
  __syncthreads();
  const uint64_t b = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= nslots) return;
  const uint32_t s = bstart[b], e = bstart[b + 1];
  if (e <= s) return;
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t *cen = centers + s;
  uint32_t nc = 0;
  // centres so far: the members already marked 1, in member order
  for (uint32_t base = s; base < e; base += 32) {
    const uint32_t i = base + lane;
    const uint32_t id = i < e ? ids[i] : 0u;
    const bool isc = i < e && merged[id] == 1;
    const uint32_t m = __ballot_sync(0xffffffffu, isc);
    if (isc) cen[nc + __popc(m & lt)] = id;
    nc += __popc(m);
  }
  __syncwarp();
  for (uint32_t base = s; base < e; base += 32) {
    const uint32_t i = base + lane;
    const uint32_t id = i < e ? ids[i] : 0u;
    const bool live = i < e && merged[id] == 0;
    uint32_t x[4 * NV];
    if (live) greedy_load<NV>(rec, RS, id, x);
    int found = -1;
    if (live) {
      for (uint32_t j = 0; j < nc; ++j) {
        uint32_t c[4 * NV];
        greedy_load<NV>(rec, RS, cen[j], c);
        if (greedy_near<NV>(x, c, len, s_d32, s_sq, thr, R)) {
          found = (int)j;
          break;
        }
      }
    }
    // members that matched no existing centre, in order: the first becomes a candidate
    // centre, the later ones are tested against it (first centre within R wins)
    uint32_t rem = __ballot_sync(0xffffffffu, live && found < 0);
    while (rem) {
      const int k = __ffs((int)rem) - 1;
      uint32_t c[4 * NV];
#pragma unroll
      for (int v = 0; v < 4 * NV; ++v) c[v] = __shfl_sync(0xffffffffu, x[v], k);
      if (lane == k) cen[nc] = id;
      rem &= ~(1u << k);
      const bool join = ((rem >> lane) & 1u) && greedy_near<NV>(x, c, len, s_d32, s_sq, thr, R);
      if (join) found = (int)nc;
      rem &= ~__ballot_sync(0xffffffffu, join);
      ++nc;
    }
    __syncwarp();
    if (live && found >= 0) {
      const uint32_t c = cen[found];
      merged[id] = 2;          // has been added to another cluster
      merged[c] = 1;           // to be the real centre (hclust2.cpp:122)
      center_of[id] = c;
      round_of[id] = round;
    }
    __syncwarp();
  }
}

__global__ void greedy_init_kernel(uint64_t n, uint8_t *merged, uint32_t *center_of, uint32_t *round_of) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  merged[i] = 0;
  center_of[i] = (uint32_t)i;
  round_of[i] = 0xffffffffu;
}

int greedy_cluster_impl(hs_ctx *ctx, uint32_t *center_out, uint32_t *round_out, uint8_t *state_out) {
  const uint64_t N = ctx->N;
  const uint32_t L = ctx->prm.L;
  stats_begin(ctx);
  if (N == 0) return HS_OK;
  if (ctx->prm.metric != HS_METRIC_EUCLID_FP64) {
    set_error("hs_greedy_cluster: Euclidean metric only (hclust2.cpp:64-71)");
    return HS_ERR_UNSUPPORTED;
  }
  cudaEvent_t *ev = ctx->ev;
  HS_CUDA(cudaEventRecord(ev[0], ctx->stream));
  HS_TRY(ensure_records(ctx));
  HS_TRY(ctx->d_parent.reserve(sizeof(uint32_t) * 3 * N + N + 64));
  uint32_t *center_of = ctx->d_parent.as<uint32_t>();
  uint32_t *round_of = center_of + N;
  uint32_t *centers = round_of + N;
  uint8_t *merged = reinterpret_cast<uint8_t *>(centers + N);
  greedy_init_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(N, merged, center_of, round_of);
  ctx->stats.kernel_launches++;
  const float thr = filter_threshold(ctx);
  for (uint32_t l = 0; l < L; ++l) {
    const TableIndex &T = ctx->tables[l];
    if (T.nslots == 0) continue;
    const uint64_t nthreads = T.nslots * 32;
    const unsigned grid = (unsigned)((nthreads + kGreedyThreads - 1) / kGreedyThreads);
    if (ctx->prm.len <= 16)
      greedy_round_kernel<1><<<grid, kGreedyThreads, 0, ctx->stream>>>(
          T.bstart.as<uint32_t>(), T.nslots, T.sorted_ids.as<uint32_t>(), ctx->d_rec.as<uint8_t>(), ctx->rec_stride,
          (int)ctx->prm.len, ctx->d_table64.as<double>(), ctx->d_dsq32.as<float>(), thr, ctx->prm.R, l, merged,
          center_of, round_of, centers);
    else
      greedy_round_kernel<2><<<grid, kGreedyThreads, 0, ctx->stream>>>(
          T.bstart.as<uint32_t>(), T.nslots, T.sorted_ids.as<uint32_t>(), ctx->d_rec.as<uint8_t>(), ctx->rec_stride,
          (int)ctx->prm.len, ctx->d_table64.as<double>(), ctx->d_dsq32.as<float>(), thr, ctx->prm.R, l, merged,
          center_of, round_of, centers);
    HS_CUDA(cudaGetLastError());
    ctx->stats.kernel_launches++;
  }
  if (center_out) HS_CUDA(cudaMemcpyAsync(center_out, center_of, sizeof(uint32_t) * N, cudaMemcpyDeviceToHost, ctx->stream));
  if (round_out) HS_CUDA(cudaMemcpyAsync(round_out, round_of, sizeof(uint32_t) * N, cudaMemcpyDeviceToHost, ctx->stream));
  if (state_out) HS_CUDA(cudaMemcpyAsync(state_out, merged, N, cudaMemcpyDeviceToHost, ctx->stream));
  HS_CUDA(cudaEventRecord(ev[7], ctx->stream));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->stats.ms_total = ev_ms(ev[0], ev[7]);
  return HS_OK;
}

}  // namespace hs
