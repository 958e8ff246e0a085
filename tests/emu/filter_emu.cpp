// CPU emulation run of the scalar candidate filter (hsearch_b200/csrc/verify.cu: build_tq_points_kernel and
// filter_kernel; api.cu: filter_threshold) against the oracle's brute force (oracle/hs_oracle.c;
// motif_both_points.cpp:176-183): the filter may let extra pairs through -- the exact stage decides -- but it
// must never drop a pair whose FP64 distance is within R, and it should be tight.  Checked over all
// (query, fragment) pairs of a seeded set with planted neighbours, for the Euclidean tables (FP32 sums of
// per-position squared distances against a threshold rounded up) and the integer metric (exact in FP32).
// Compile with -ffp-contract=off.  filter_kernels.inc is cut out of the sources by tests/test_emu_filter.py.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <random>
#include <set>
#include <vector>

#include "cuda_emu.h"
#include "../../hsearch_b200/csrc/common.cuh"

extern "C" {
typedef struct {
  uint32_t query;
  uint32_t table_first;
  uint64_t db_id;
  double dist2;
} orc_hit;
void orc_get_coordinates_print6(double *out160);
void orc_blosum_metric(int *out400);
void orc_embed(const uint8_t *codes, uint32_t len, const double *table160, double *point);
uint64_t orc_bruteforce(const double *db, uint64_t N, const double *queries, uint32_t Q, uint32_t dim, double R, int pred,
                        orc_hit *hits, uint64_t cap);
uint64_t orc_bruteforce_int(const uint8_t *db, uint64_t N, const uint8_t *qcodes, uint32_t Q, uint32_t len, int R,
                            orc_hit *hits, uint64_t cap);
}

namespace hs {
void set_error(const char *, ...) {}
#include "filter_kernels.inc"
}  // namespace hs

using namespace hs;

template <int LENB>
static bool test_filter(int len, double R, uint64_t N, uint32_t Q, bool integer, unsigned seed) {
  const int dim = len * HS_CDIM;
  hs_ctx ctx_storage;
  hs_ctx *ctx = &ctx_storage;
  memset(&ctx->prm, 0, sizeof ctx->prm);
  ctx->prm.len = len; ctx->prm.R = R;
  ctx->prm.metric = integer ? HS_METRIC_BLOSUM_INT : HS_METRIC_EUCLID_FP64;
  orc_get_coordinates_print6(ctx->table64);
  std::mt19937 rng(seed);
  std::vector<uint8_t> codes(N * len), qcodes((size_t)Q * len);
  for (auto &c : codes) c = (uint8_t)(rng() % 20);
  for (uint32_t q = 0; q < Q; ++q) {   // queries: mutants of DB fragments, so that pairs near the threshold exist
    const uint64_t src = rng() % N;
    memcpy(&qcodes[(size_t)q * len], &codes[src * len], len);
    for (int s = 0; s < (int)(rng() % 5); ++s) qcodes[(size_t)q * len + rng() % len] = (uint8_t)(rng() % 20);
  }
  std::vector<double> db(N * dim), qp((size_t)Q * dim);
  for (uint64_t i = 0; i < N; ++i) orc_embed(&codes[i * len], len, ctx->table64, &db[i * dim]);
  for (uint32_t q = 0; q < Q; ++q) orc_embed(&qcodes[(size_t)q * len], len, ctx->table64, &qp[(size_t)q * dim]);
  std::vector<orc_hit> want(N * Q);
  const uint64_t nw = integer ? orc_bruteforce_int(codes.data(), N, qcodes.data(), Q, len, (int)R, want.data(), want.size())
                              : orc_bruteforce(db.data(), N, qp.data(), Q, dim, R, 0, want.data(), want.size());
  // per-query filter tables
  const uint64_t ntq = (uint64_t)Q * len * HS_AA;
  std::vector<float> tq(ntq);
  int metric[400];
  orc_blosum_metric(metric);
  std::vector<int32_t> metric32(metric, metric + 400);
  bool ok = integer ? emu_launch((unsigned)((ntq + 255) / 256), 256, [&]() { build_tq_int_kernel(qcodes.data(), Q, len, metric32.data(), tq.data()); })
                    : emu_launch((unsigned)((ntq + 255) / 256), 256, [&]() { build_tq_points_kernel(qp.data(), Q, len, ctx->table64, tq.data()); });
  // identity-order, position-major store of code * 4
  const uint64_t npad = (N + 15) & ~15ull;
  std::vector<uint8_t> store((size_t)len * npad + 64, 0);
  for (uint64_t i = 0; i < N; ++i)
    for (int p = 0; p < len; ++p) store[(uint64_t)p * npad + i] = (uint8_t)(codes[i * len + p] * kCodeScale);
  const uint8_t *stores[1] = {store.data()};
  std::vector<uint32_t> qlist(Q);
  for (uint32_t q = 0; q < Q; ++q) qlist[q] = q;
  // two work items: the members split at an odd position (tiles start at m_begin rounded down to 4)
  const uint32_t cutpos = (uint32_t)(N / 2) | 1u;
  WorkItem items[2];
  items[0] = WorkItem{0, 0, cutpos, 0, Q, 0};
  const uint32_t nb0 = (cutpos + kFilterTile - 1) / kFilterTile;
  items[1] = WorkItem{0, cutpos, (uint32_t)N, 0, Q, nb0};
  const uint32_t nb1 = ((uint32_t)N - (cutpos & ~3u) + kFilterTile - 1) / kFilterTile;
  std::vector<Survivor> surv(N * Q);
  unsigned long long count = 0;
  FilterArgs fa;
  memset(&fa, 0, sizeof fa);
  fa.items = items;
  fa.nitems = 2;
  fa.qlist = qlist.data();
  fa.tq = tq.data();
  fa.stores = stores;
  fa.npad = npad;
  fa.len = len;
  fa.thr = filter_threshold(ctx);   // the library's own threshold
  fa.surv = surv.data();
  fa.surv_cap = surv.size();
  fa.surv_count = &count;
  ok = ok && emu_launch(nb0 + nb1, kFilterThreads, [&]() { filter_kernel<kModeSearch, LENB>(fa); });
  if (!ok) return false;
  std::set<uint64_t> passed;
  for (unsigned long long i = 0; i < count; ++i) {
    if (!passed.insert((uint64_t)surv[i].query * N + surv[i].pos).second) {
      printf("  pair reported twice\n");
      return false;
    }
  }
  for (uint64_t i = 0; i < nw; ++i)
    if (!passed.count((uint64_t)want[i].query * N + want[i].db_id)) {
      printf("  the filter dropped a pair within R (query %u, fragment %llu, d2 %.17g)\n", want[i].query,
             (unsigned long long)want[i].db_id, want[i].dist2);
      return false;
    }
  // tightness: whatever passed is within R up to the rounding slack
  uint64_t extra = count - nw;
  printf("  (%llu pairs within R, %llu extra survivors among %llu pairs)\n", (unsigned long long)nw, (unsigned long long)extra,
         (unsigned long long)(N * Q));
  return nw > 15 && (integer ? extra == 0 : extra <= nw / 100 + 2);
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  int nbad = 0;
  auto report = [&](const char *what, bool ok) {
    printf("%s -> %s\n", what, ok ? "ok" : "FAILED");
    if (!ok) ++nbad;
  };
  report("Euclidean, len 10, R 34", test_filter<10>(10, 34.0, 6001, 24, false, 1));
  report("Euclidean, len 25, R 60", test_filter<25>(25, 60.0, 3003, 12, false, 2));
  report("integer metric, len 10, R 60", test_filter<10>(10, 60.0, 5000, 24, true, 3));
  return nbad ? 1 : 0;
}
