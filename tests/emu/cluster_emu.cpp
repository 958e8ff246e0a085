// CPU emulation run of the near-pair clustering path (U1 as the composition SURVEY.md 8c defines: every in-bucket
// pair within R is an edge, UnionFind of pcluster/src/pcluster/union_find.cpp:3-33 over the edges, labels = smallest
// id of the component) against the oracle's orc_cluster (oracle/hs_oracle.c, linked in).  The library's kernels run
// unchanged, in the order cluster_impl (csrc/cluster.cu) launches them: per table small_bucket_pairs_kernel (warp per
// bucket of <= 64 members, larger buckets collected), the tiled scalar self-join filter_kernel<SelfJoin> over the
// larger buckets with the work items cluster_impl builds, exact_kernel in self-join mode after each (FP64 distance,
// sqrt predicate, lock-free union), uf_flatten at the end; also split over two emulated ranks (buckets dealt by
// number) whose forests are merged through their labels (uf_merge_labels_kernel), as on a communicator.
// Compile with -ffp-contract=off.  cluster_kernels.inc is cut out of the sources by tests/test_emu_cluster.py.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <random>
#include <string>
#include <vector>

#include "cuda_emu.h"
#include "../../hsearch_b200/csrc/common.cuh"

extern "C" {
void orc_get_coordinates_print6(double *out160);
void orc_lsh_generate(uint64_t seed, uint32_t dim, uint32_t K, double W, double *a, double *b);
void orc_hash_points(const double *pts, uint64_t N, uint32_t dim, const double *a, const double *b, uint32_t K, uint32_t L,
                     double W, int *out);
void orc_embed(const uint8_t *codes, uint32_t len, const double *table160, double *point);
uint64_t orc_cluster(const uint8_t *codes, uint64_t N, uint32_t len, const double *table160, const double *a, const double *b,
                     uint32_t K, uint32_t L, double W, double R, int metric, uint32_t *label_out);
}

namespace hs {
void set_error(const char *, ...) {}
#include "cluster_kernels.inc"
}  // namespace hs

using namespace hs;

struct Index {
  std::vector<uint32_t> ids, bstart;
  std::vector<uint8_t> store;
};

static bool run_rank(uint32_t part, uint32_t nparts, const hs_ctx *ctx, const std::vector<Index> &tabs, uint64_t N, int len, double R,
                     const std::vector<uint8_t> &rec, uint32_t RS, const float *dsq32, uint64_t npad, std::vector<uint32_t> &parent,
                     unsigned long long *edges_out) {
  const uint32_t L = (uint32_t)tabs.size();
  parent.resize(N);
  bool ok = emu_launch((unsigned)((N + 255) / 256), 256, [&]() { iota32_kernel(parent.data(), N); });
  const float thr = filter_threshold(ctx);
  std::vector<const uint32_t *> id_ptrs(L);
  std::vector<const uint8_t *> store_ptrs(L);
  for (uint32_t l = 0; l < L; ++l) {
    id_ptrs[l] = tabs[l].ids.data();
    store_ptrs[l] = tabs[l].store.data();
  }
  unsigned long long edges = 0, hit_count = 0;
  std::vector<int32_t> metric(400, 0);
  uint64_t work_no = 0;
  for (uint32_t l = 0; l < L && ok; ++l) {
    const Index &T = tabs[l];
    const uint64_t nslots = T.bstart.size() - 1;
    std::vector<Survivor> surv((size_t)4 << 20);
    std::vector<uint2> large(N / kSmallBucket + 2);
    unsigned long long scnt = 0, npairs = 0;
    unsigned int nlarge = 0;
    ok = ok && emu_launch((unsigned)((nslots * 32 + 255) / 256), 256, [&]() {
      small_bucket_pairs_kernel(T.bstart.data(), nslots, l, T.store.data(), npad, len, dsq32, thr, surv.data(), surv.size(), &scnt,
                                large.data(), &nlarge, &npairs, part, nparts);
    });
    if (!ok || scnt > surv.size()) return false;
    ExactArgs ea;
    memset(&ea, 0, sizeof ea);
    ea.mode = kModeSelfJoin;
    ea.metric = HS_METRIC_EUCLID_FP64;
    ea.predicate = HS_PRED_SQRT_LE_R;
    ea.len = len; ea.dim = len * HS_CDIM; ea.key_words = 1; ea.L = (int)L;
    ea.R = R;
    ea.sorted_ids = id_ptrs.data();
    ea.rec = rec.data();
    ea.rec_stride = RS;
    ea.N = N;
    ea.table64 = ctx->table64;
    ea.metric_tab = metric.data();
    ea.parent = parent.data();
    ea.edge_count = &edges;
    ea.hit_count = &hit_count;
    ea.surv = surv.data();
    ea.nsurv = scnt;
    if (scnt) ok = ok && emu_launch(2, kExactThreads, [&]() { exact_kernel<1, 1>(ea); });
    // the larger buckets through the tiled scalar self-join, work items as cluster_impl builds them
    std::sort(large.begin(), large.begin() + nlarge, [](const uint2 &x, const uint2 &y) { return x.x < y.x; });
    std::vector<WorkItem> items;
    uint32_t nblocks = 0;
    for (size_t bi = 0; bi < nlarge; ++bi) {
      const uint32_t ms = large[bi].x, me = large[bi].y;
      if (((work_no + bi) % nparts) != part) continue;
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
    }
    work_no += nlarge;
    if (!items.empty()) {
      unsigned long long cnt2 = 0;
      FilterArgs fa;
      memset(&fa, 0, sizeof fa);
      fa.items = items.data();
      fa.nitems = (uint32_t)items.size();
      fa.dsq32 = dsq32;
      fa.stores = store_ptrs.data();
      fa.npad = npad;
      fa.len = len;
      fa.thr = thr;
      fa.surv = surv.data();
      fa.surv_cap = surv.size();
      fa.surv_count = &cnt2;
      ok = ok && emu_launch(nblocks, kFilterThreads, [&]() { filter_kernel<kModeSelfJoin, 10>(fa); });
      if (!ok || cnt2 > surv.size()) return false;
      ea.surv = surv.data();
      ea.nsurv = cnt2;
      if (cnt2) ok = ok && emu_launch(2, kExactThreads, [&]() { exact_kernel<1, 1>(ea); });
    }
  }
  *edges_out = edges;
  return ok;
}

static bool test_cluster(uint64_t N, uint32_t K, uint32_t L, double W, double R, int family, unsigned seed) {
  const int len = 10, dim = len * HS_CDIM;
  hs_ctx ctx_storage;
  hs_ctx *ctx = &ctx_storage;
  memset(&ctx->prm, 0, sizeof ctx->prm);
  ctx->prm.len = len; ctx->prm.K = K; ctx->prm.L = L; ctx->prm.W = W; ctx->prm.R = R;
  ctx->prm.metric = HS_METRIC_EUCLID_FP64;
  orc_get_coordinates_print6(ctx->table64);
  std::mt19937 rng(seed);
  const uint32_t nfam = std::max<uint32_t>(1, (uint32_t)(N / family));
  std::vector<uint8_t> roots((size_t)nfam * len), codes(N * len);
  for (auto &c : roots) c = (uint8_t)(rng() % 20);
  for (uint64_t i = 0; i < N; ++i) {
    memcpy(&codes[i * len], &roots[(size_t)(i % nfam) * len], len);
    for (int s = 0; s < (int)(rng() % 3); ++s) codes[i * len + rng() % len] = (uint8_t)(rng() % 20);
  }
  std::vector<double> a((size_t)L * K * dim), b((size_t)L * K), pts(N * dim);
  for (uint32_t l = 0; l < L; ++l) orc_lsh_generate(99 + seed + l, dim, K, W, &a[(size_t)l * K * dim], &b[(size_t)l * K]);
  for (uint64_t i = 0; i < N; ++i) orc_embed(&codes[i * len], len, ctx->table64, &pts[i * dim]);
  std::vector<uint32_t> want(N);
  const uint64_t want_edges = orc_cluster(codes.data(), N, len, ctx->table64, a.data(), b.data(), K, L, W, R, 0, want.data());
  // the index of every table: fragments grouped by key string, ascending id inside a bucket; position-major code * 4 stores
  std::vector<int> bk(N * L * K);
  orc_hash_points(pts.data(), N, dim, a.data(), b.data(), K, L, W, bk.data());
  const uint64_t npad = (N + 15) & ~15ull;
  std::vector<Index> tabs(L);
  for (uint32_t l = 0; l < L; ++l) {
    std::map<std::string, std::vector<uint32_t>> groups;
    for (uint64_t i = 0; i < N; ++i) {
      std::string s;
      for (uint32_t k = 0; k < K; ++k) s += std::to_string(bk[(i * L + l) * K + k]);
      groups[s].push_back((uint32_t)i);
    }
    tabs[l].bstart.push_back(0);
    for (auto &kv : groups) {
      tabs[l].ids.insert(tabs[l].ids.end(), kv.second.begin(), kv.second.end());
      tabs[l].bstart.push_back((uint32_t)tabs[l].ids.size());
    }
    tabs[l].store.assign((size_t)len * npad + 256, 0);
    for (uint64_t j = 0; j < N; ++j)
      for (int p = 0; p < len; ++p) tabs[l].store[(uint64_t)p * npad + j] = (uint8_t)(codes[(uint64_t)tabs[l].ids[j] * len + p] * kCodeScale);
  }
  const uint32_t RS = 16;
  std::vector<uint8_t> rec(N * RS + 64, 0);
  for (uint64_t i = 0; i < N; ++i) memcpy(&rec[i * RS], &codes[i * len], len);
  float dsq32[HS_AA * HS_AA];   // upload_tables (api.cu): squared residue distances in FP32
  for (int c = 0; c < HS_AA; ++c)
    for (int d = 0; d < HS_AA; ++d) {
      double s = 0.0;
      for (int j = 0; j < HS_CDIM; ++j) {
        const double r = ctx->table64[c * HS_CDIM + j] - ctx->table64[d * HS_CDIM + j];
        s += r * r;
      }
      dsq32[c * HS_AA + d] = (float)s;
    }
  // one rank
  std::vector<uint32_t> parent, label(N);
  unsigned long long edges = 0;
  if (!run_rank(0, 1, ctx, tabs, N, len, R, rec, RS, dsq32, npad, parent, &edges)) return false;
  if (!emu_launch((unsigned)((N + 255) / 256), 256, [&]() { uf_flatten_kernel(parent.data(), N, label.data()); })) return false;
  if (label != want || edges != want_edges) {
    printf("  one rank: labels or edge count (%llu, oracle %llu) differ\n", edges, (unsigned long long)want_edges);
    return false;
  }
  // two ranks: the pair work split by bucket number, forests merged through their labels
  std::vector<uint32_t> p0, p1, l0(N), l1(N);
  unsigned long long e0 = 0, e1 = 0;
  if (!run_rank(0, 2, ctx, tabs, N, len, R, rec, RS, dsq32, npad, p0, &e0) || !run_rank(1, 2, ctx, tabs, N, len, R, rec, RS, dsq32, npad, p1, &e1))
    return false;
  bool ok = emu_launch((unsigned)((N + 255) / 256), 256, [&]() { uf_flatten_kernel(p1.data(), N, l1.data()); });
  ok = ok && emu_launch((unsigned)((N + 255) / 256), 256, [&]() { uf_merge_labels_kernel(p0.data(), l1.data(), N); });
  ok = ok && emu_launch((unsigned)((N + 255) / 256), 256, [&]() { uf_flatten_kernel(p0.data(), N, l0.data()); });
  if (!ok || l0 != want || e0 + e1 != want_edges) {
    printf("  two ranks: labels or edge count (%llu + %llu, oracle %llu) differ\n", e0, e1, (unsigned long long)want_edges);
    return false;
  }
  uint32_t ncomp = 0;
  for (uint64_t i = 0; i < N; ++i) ncomp += want[i] == i;
  printf("  (%llu fragments, %llu edges, %u components; two ranks: %llu + %llu edges)\n", (unsigned long long)N,
         (unsigned long long)want_edges, ncomp, e0, e1);
  return want_edges > 0 && ncomp < N;
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  int nbad = 0;
  auto report = [&](const char *what, bool ok) {
    printf("%s -> %s\n", what, ok ? "ok" : "FAILED");
    if (!ok) ++nbad;
  };
  report("4000 fragments in families of 10, K 4 L 4 W 50 R 25 (small buckets)", test_cluster(4000, 4, 4, 50.0, 25.0, 10, 1));
  report("3000 fragments in families of 30, K 2 L 3 W 80 R 28 (buckets of hundreds: the tiled self-join)", test_cluster(3000, 2, 3, 80.0, 28.0, 30, 2));
  return nbad ? 1 : 0;
}
