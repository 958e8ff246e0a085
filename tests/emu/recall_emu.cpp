// CPU emulation run of the recall join (hsearch_b200/csrc/evaluate.cu: recall_order / recall_segments /
// recall_join kernels -- evaulate() and weight() of motif_both_points.cpp:67-87,100-165 on binary hit lists)
// against the oracle's sequential restatement (oracle/hs_oracle.c: orc_evaluate, itself pinned bit-exact against
// the reference's own function).  Counts and distance bins must be equal, the weighted sums within 1e-12
// relative (the device adds them in a fixed parallel order).  recall_kernels.inc is cut out of evaluate.cu by
// tests/test_emu_recall.py.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <random>
#include <vector>

#include "../../include/hsearch_b200.h"
#include "cuda_emu.h"

extern "C" {
typedef struct {
  uint32_t query;
  uint32_t table_first;
  uint64_t db_id;
  double dist2;
} orc_hit;
int orc_evaluate(const orc_hit *truth, const double *tdis, uint64_t nt, const orc_hit *found, uint64_t nf, double R,
                 uint32_t nbins, double *tp_out, double *fn_out, uint64_t *n_tp, uint64_t *n_fn, uint64_t *n_extra,
                 uint64_t *tp_bin, uint64_t *fn_bin);
}

namespace hs {
#include "recall_kernels.inc"
}  // namespace hs

using namespace hs;

static bool test_recall(uint32_t Q, uint32_t ntab, uint64_t nt, double keep, double R, unsigned seed) {
  std::mt19937_64 rng(seed);
  // ground truth: unique (query, id) pairs with distances up to R (some beyond 24, where weight() < 1),
  // in (query, id) order as the brute force returns them
  std::vector<hs_hit> truth;
  for (uint64_t i = 0; i < nt; ++i) {
    hs_hit h;
    h.query = (uint32_t)(rng() % Q);
    h.table_first = 0;
    h.db_id = rng() % 100000;
    const double dis = (double)(rng() % 100000) / 100000.0 * R;
    h.dist2 = dis * dis;
    truth.push_back(h);
  }
  std::sort(truth.begin(), truth.end(), [](const hs_hit &a, const hs_hit &b) { return a.query != b.query ? a.query < b.query : a.db_id < b.db_id; });
  truth.erase(std::unique(truth.begin(), truth.end(), [](const hs_hit &a, const hs_hit &b) { return a.query == b.query && a.db_id == b.db_id; }),
              truth.end());
  nt = truth.size();
  // found: a fraction of the truth (first table drawn at random) plus pairs that are not in the truth
  std::vector<hs_hit> found;
  for (const hs_hit &g : truth)
    if ((double)(rng() % 1000) / 1000.0 < keep) {
      hs_hit h = g;
      h.table_first = (uint32_t)(rng() % ntab);
      found.push_back(h);
    }
  for (uint64_t i = 0; i < nt / 10 + 3; ++i) {
    hs_hit h;
    h.query = (uint32_t)(rng() % Q);
    h.table_first = (uint32_t)(rng() % ntab);
    h.db_id = 200000 + rng() % 1000;   // ids the truth never uses
    h.dist2 = 1.0;
    found.push_back(h);
  }
  std::sort(found.begin(), found.end(), [](const hs_hit &a, const hs_hit &b) {
    if (a.query != b.query) return a.query < b.query;
    if (a.table_first != b.table_first) return a.table_first < b.table_first;
    return a.db_id < b.db_id;
  });
  found.erase(std::unique(found.begin(), found.end(), [](const hs_hit &a, const hs_hit &b) {
                return a.query == b.query && a.table_first == b.table_first && a.db_id == b.db_id;
              }), found.end());
  const uint64_t nf = found.size();
  // the oracle joins lists in (query, id) order
  std::vector<orc_hit> ot(nt), of(nf);
  std::vector<double> tdis(nt);
  for (uint64_t i = 0; i < nt; ++i) {
    ot[i] = orc_hit{truth[i].query, 0, truth[i].db_id, truth[i].dist2};
    tdis[i] = sqrt(truth[i].dist2);
  }
  for (uint64_t i = 0; i < nf; ++i) of[i] = orc_hit{found[i].query, found[i].table_first, found[i].db_id, found[i].dist2};
  std::sort(of.begin(), of.end(), [](const orc_hit &a, const orc_hit &b) { return a.query != b.query ? a.query < b.query : a.db_id < b.db_id; });
  double wtp = 0, wfn = 0;
  uint64_t n_tp = 0, n_fn = 0, n_extra = 0;
  std::vector<uint64_t> tpb(HS_RECALL_BINS), fnb(HS_RECALL_BINS);
  if (orc_evaluate(ot.data(), tdis.data(), nt, of.data(), nf, R, HS_RECALL_BINS, &wtp, &wfn, &n_tp, &n_fn, &n_extra, tpb.data(), fnb.data()))
    return false;
  // the device path (evaluate_recall_dev)
  RecallCounters c;
  memset(&c, 0, sizeof c);
  std::vector<uint64_t> seg((size_t)Q + 2, 0);
  const unsigned grid = 5;
  std::vector<double> part(2 * grid, 0.0);
  bool ok = true;
  if (nf > 1) ok = emu_launch((unsigned)((nf - 1 + 255) / 256), 256, [&]() { recall_order_kernel(found.data(), nf, &c); });
  ok = ok && emu_launch((Q + 1 + 255) / 256, 256, [&]() { recall_segments_kernel(found.data(), nf, Q, seg.data()); });
  ok = ok && emu_launch(grid, kRecallThreads, [&]() {
    recall_join_kernel(truth.data(), nt, found.data(), seg.data(), Q, ntab, 1, R, &c, part.data(), part.data() + grid);
  });
  if (!ok || c.unsorted || c.n_badq || c.n_over) return false;
  double ftp = 0, ffn = 0;
  for (unsigned b = 0; b < grid; ++b) {
    ftp += part[b];
    ffn += part[grid + b];
  }
  const double tp = (double)c.ones_tp + ftp, fn = (double)c.ones_fn + ffn;
  if (c.n_tp != n_tp || c.n_fn != n_fn || nf - c.n_tp != n_extra) {
    printf("  counts: tp %llu fn %llu extra %llu, oracle %llu %llu %llu\n", c.n_tp, c.n_fn, (unsigned long long)(nf - c.n_tp),
           (unsigned long long)n_tp, (unsigned long long)n_fn, (unsigned long long)n_extra);
    return false;
  }
  for (int i = 0; i < HS_RECALL_BINS; ++i)
    if (c.tp_bin[i] != tpb[i] || c.fn_bin[i] != fnb[i]) {
      printf("  bin %d differs\n", i);
      return false;
    }
  if (fabs(tp - wtp) > 1e-12 * fabs(wtp) || fabs(fn - wfn) > 1e-12 * fabs(wfn)) {
    printf("  weighted sums: %.17g %.17g, oracle %.17g %.17g\n", tp, fn, wtp, wfn);
    return false;
  }
  // an out-of-order found list is reported
  if (nf > 2) {
    std::swap(found[0], found[nf - 1]);
    memset(&c, 0, sizeof c);
    if (!emu_launch((unsigned)((nf - 1 + 255) / 256), 256, [&]() { recall_order_kernel(found.data(), nf, &c); }) || c.unsorted == 0) return false;
  }
  return n_tp > 0 && n_fn > 0;
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  int nbad = 0;
  auto report = [&](const char *what, bool ok) {
    printf("%s -> %s\n", what, ok ? "ok" : "FAILED");
    if (!ok) ++nbad;
  };
  report("recall: 300 queries, 4 tables, 20000 true pairs, a third found, R 30", test_recall(300, 4, 20000, 0.33, 30.0, 1));
  report("recall: 7 queries, 1 table, 3000 true pairs, most found, R 45", test_recall(7, 1, 3000, 0.9, 45.0, 2));
  report("recall: 2000 queries, 8 tables, 5000 true pairs (many queries without a hit), R 20", test_recall(2000, 8, 5000, 0.5, 20.0, 3));
  return nbad ? 1 : 0;
}
