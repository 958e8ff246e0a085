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

#include "internal.cuh"

namespace hs {

constexpr uint32_t kSmallBucket = 64;

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
                                          unsigned long long *__restrict__ npairs) {
  const uint64_t b = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= nb) return;
  const uint32_t s = bstart[b], e = bstart[b + 1];
  const uint32_t n = e - s;
  if (n < 2) return;
  if (n > kSmallBucket) {
    if (lane == 0) {
      const unsigned int i = atomicAdd(nlarge, 1u);
      large[i] = make_uint2(s, e);
    }
    return;
  }
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

int cluster_impl(hs_ctx *ctx, uint32_t *label_out) {
  const uint64_t N = ctx->N;
  const uint32_t L = ctx->prm.L;
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
          npairs);
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
      std::vector<WorkItem> items;
      uint32_t nblocks = 0;
      size_t bi = 0;
      while (bi < large.size()) {
        items.clear();
        nblocks = 0;
        // batch buckets until the block budget is reached
        for (; bi < large.size(); ++bi) {
          const uint32_t ms = large[bi].x, me = large[bi].y;
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
    }
    HS_CUDA(cudaEventSynchronize(ev[3]));
    ms_filter += ev_ms(ev[1], ev[2]) + ms_f2;
    ms_exact += ev_ms(ev[2], ev[3]) + ms_e2;
  }
  uint32_t *label = parent + N;
  uf_flatten_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(parent, N, label);
  ctx->stats.kernel_launches++;
  HS_CUDA(cudaGetLastError());
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

}  // namespace hs
