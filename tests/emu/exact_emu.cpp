// CPU emulation run of the exact stage (hsearch_b200/csrc/verify.cu: exact_kernel -- the FP64 distance in the
// reference's operation order, the predicate, the first-table-wins rule, hit emission) against the oracle's
// Search() / brute force (oracle/hs_oracle.c, linked in; motif_both_points.cpp:176-183,224-245,
// motif_both_points_noLSH.cpp:36-56, evaluate_correlation.cpp:26-41).  The survivors handed to the kernel are
// ALL members of every query's buckets, so what is checked is exactly what the stage decides: hit set, first
// table and the FP64 distance, bit for bit.  Compile with -ffp-contract=off.  exact_kernels.inc is cut out of
// the sources by tests/test_emu_exact.py.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <random>
#include <string>
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
void orc_get_coordinates_print6(double *out160);
void orc_blosum_metric(int *out400);
void orc_lsh_generate(uint64_t seed, uint32_t dim, uint32_t K, double W, double *a, double *b);
void orc_hash_points(const double *pts, uint64_t N, uint32_t dim, const double *a, const double *b, uint32_t K, uint32_t L,
                     double W, int *out);
void orc_embed(const uint8_t *codes, uint32_t len, const double *table160, double *point);
uint64_t orc_search(const double *db, uint64_t N, const double *queries, uint32_t Q, uint32_t dim, const double *a,
                    const double *b, uint32_t K, uint32_t L, double W, double R, int pred, orc_hit *hits, uint64_t cap,
                    uint64_t *table_sizes, uint64_t *ncandidates);
uint64_t orc_bruteforce(const double *db, uint64_t N, const double *queries, uint32_t Q, uint32_t dim, double R, int pred,
                        orc_hit *hits, uint64_t cap);
uint64_t orc_bruteforce_int(const uint8_t *db, uint64_t N, const uint8_t *qcodes, uint32_t Q, uint32_t len, int R,
                            orc_hit *hits, uint64_t cap);
}

namespace hs {
#include "exact_kernels.inc"
}  // namespace hs

using namespace hs;

struct HitKey {
  uint32_t q, t;
  uint64_t id, dbits;
  bool operator<(const HitKey &o) const {
    if (q != o.q) return q < o.q;
    if (t != o.t) return t < o.t;
    return id < o.id;
  }
  bool operator==(const HitKey &o) const { return q == o.q && t == o.t && id == o.id && dbits == o.dbits; }
};
static uint64_t bits_of(double d) {
  uint64_t u;
  memcpy(&u, &d, 8);
  return u;
}

static uint64_t key_of(const int *buckets, int K) {   // one-word packed key of a bucket tuple (hash.cuh)
  std::string s;
  for (int k = 0; k < K; ++k) s += std::to_string(buckets[k]);
  uint64_t w = 0;
  for (char c : s) w = (w << 4) | (c == '-' ? 11u : (uint64_t)(c - '0' + 1));
  return w;
}

struct Setup {
  int len, dim, K, L;
  double W, R;
  uint64_t N;
  uint32_t Q;
  double table[HS_AA * HS_CDIM];
  int32_t metric[400];
  std::vector<uint8_t> codes, qcodes, rec;
  std::vector<double> db, q, a, b;
};

static void make_setup(Setup &S, unsigned seed) {
  S.dim = S.len * HS_CDIM;
  orc_get_coordinates_print6(S.table);
  int m[400];
  orc_blosum_metric(m);
  for (int i = 0; i < 400; ++i) S.metric[i] = m[i];
  std::mt19937 rng(seed);
  S.codes.resize(S.N * S.len);
  for (auto &c : S.codes) c = (uint8_t)(rng() % 20);
  S.qcodes.resize((size_t)S.Q * S.len);
  for (uint32_t i = 0; i < S.Q; ++i) {
    if (i % 2 == 0) {   // planted: a DB fragment with up to two substitutions
      const uint64_t src = rng() % S.N;
      memcpy(&S.qcodes[(size_t)i * S.len], &S.codes[src * S.len], S.len);
      for (int s = 0; s < (int)(rng() % 3); ++s) S.qcodes[(size_t)i * S.len + rng() % S.len] = (uint8_t)(rng() % 20);
    } else {
      for (int p = 0; p < S.len; ++p) S.qcodes[(size_t)i * S.len + p] = (uint8_t)(rng() % 20);
    }
  }
  S.db.resize(S.N * S.dim);
  for (uint64_t i = 0; i < S.N; ++i) orc_embed(&S.codes[i * S.len], S.len, S.table, &S.db[i * S.dim]);
  S.q.resize((size_t)S.Q * S.dim);
  for (uint32_t i = 0; i < S.Q; ++i) orc_embed(&S.qcodes[(size_t)i * S.len], S.len, S.table, &S.q[(size_t)i * S.dim]);
  S.a.resize((size_t)S.L * S.K * S.dim);
  S.b.resize((size_t)S.L * S.K);
  for (int l = 0; l < S.L; ++l) orc_lsh_generate(777 + seed + l, S.dim, S.K, S.W, &S.a[(size_t)l * S.K * S.dim], &S.b[(size_t)l * S.K]);
  const uint32_t RS = (S.len + 15u) & ~15u;
  S.rec.assign(S.N * RS + 64, 0);
  for (uint64_t i = 0; i < S.N; ++i) memcpy(&S.rec[i * RS], &S.codes[i * S.len], S.len);
}

template <int NV, int REP>
static bool run_exact(const Setup &S, ExactArgs ea, std::vector<HitKey> &got) {
  std::vector<hs_hit> hits(ea.hit_cap);
  unsigned long long count = 0, edges = 0;
  ea.hits = hits.data();
  ea.hit_count = &count;
  ea.edge_count = &edges;
  const int threads = REP == 8 ? kExactThreadsRep : kExactThreads;
  if (!emu_launch(3, threads, [&]() { exact_kernel<NV, REP>(ea); })) return false;
  if (count > ea.hit_cap) return false;
  got.clear();
  for (unsigned long long i = 0; i < count; ++i) got.push_back({hits[i].query, hits[i].table_first, hits[i].db_id, bits_of(hits[i].dist2)});
  std::sort(got.begin(), got.end());
  return true;
}

static std::vector<HitKey> as_keys(const std::vector<orc_hit> &h, uint64_t n) {
  std::vector<HitKey> v;
  for (uint64_t i = 0; i < n; ++i) v.push_back({h[i].query, h[i].table_first, h[i].db_id, bits_of(h[i].dist2)});
  std::sort(v.begin(), v.end());
  return v;
}

// LSH search: survivors = every member of every query's bucket in every table
template <int NV, int REP>
static bool test_search(int len, double W, double R, uint64_t N, uint32_t Q, bool string_queries, unsigned seed, bool rank_path = false) {
  Setup S;
  S.len = len; S.K = 4; S.L = 4; S.W = W; S.R = R; S.N = N; S.Q = Q;
  make_setup(S, seed);
  std::vector<orc_hit> want(N * Q);
  const uint64_t nw = orc_search(S.db.data(), N, S.q.data(), Q, S.dim, S.a.data(), S.b.data(), S.K, S.L, W, R, 0, want.data(), want.size(),
                                 nullptr, nullptr);
  std::vector<int> bk(N * S.L * S.K), qbk((size_t)Q * S.L * S.K);
  orc_hash_points(S.db.data(), N, S.dim, S.a.data(), S.b.data(), S.K, S.L, W, bk.data());
  orc_hash_points(S.q.data(), Q, S.dim, S.a.data(), S.b.data(), S.K, S.L, W, qbk.data());
  std::vector<std::vector<uint64_t>> keys(S.L, std::vector<uint64_t>(N));
  std::vector<uint64_t> qkeys((size_t)S.L * Q);
  std::vector<uint8_t> qvalid((size_t)S.L * Q, 1);
  for (int l = 0; l < S.L; ++l) {
    for (uint64_t i = 0; i < N; ++i) keys[l][i] = key_of(&bk[(i * S.L + l) * S.K], S.K);
    for (uint32_t q = 0; q < Q; ++q) qkeys[(size_t)l * Q + q] = key_of(&qbk[((size_t)q * S.L + l) * S.K], S.K);
  }
  std::vector<Survivor> surv;
  for (int l = 0; l < S.L; ++l)
    for (uint32_t q = 0; q < Q; ++q)
      for (uint64_t i = 0; i < N; ++i)
        if (keys[l][i] == qkeys[(size_t)l * Q + q]) surv.push_back(Survivor{q, (uint32_t)l, (uint32_t)i, 2u});
  std::mt19937 rng(seed + 99);
  std::shuffle(surv.begin(), surv.end(), rng);
  std::vector<const uint64_t *> kptr(S.L);
  for (int l = 0; l < S.L; ++l) kptr[l] = keys[l].data();
  std::vector<const uint32_t *> ids(S.L, nullptr);
  std::vector<uint8_t> qrow(Q, 1);
  ExactArgs ea;
  memset(&ea, 0, sizeof ea);
  ea.surv = surv.data();
  ea.nsurv = surv.size();
  ea.mode = kModeSearch;
  ea.metric = HS_METRIC_EUCLID_FP64;
  ea.predicate = HS_PRED_D2_LE_R2;
  ea.len = len; ea.dim = S.dim; ea.key_words = 1; ea.L = S.L;
  ea.R = R;
  ea.sorted_ids = ids.data();
  ea.codes = S.codes.data();
  ea.rec = S.rec.data();
  ea.rec_stride = (len + 15u) & ~15u;
  ea.N = N; ea.id_base = 1000000;
  ea.table64 = S.table;
  ea.metric_tab = S.metric;
  ea.q64 = S.q.data();
  if (string_queries) {   // dense queries recognised as embedded residue strings: the shared residue-pair table
    ea.qcodes = S.qcodes.data();
    ea.qrow = qrow.data();
  }
  ea.Q = Q;
  ea.keys = kptr.data();
  ea.qkeys = qkeys.data();
  ea.qvalid = qvalid.data();
  // rank path (the headline configuration): a fragment's bucket in table l is the u16 rank of its key among the
  // table's keys, kept in its 32-byte record behind the codes; a query's is qrank[l][q] (0xffffffff: no bucket)
  std::vector<uint8_t> rec32;
  std::vector<uint32_t> qrank;
  if (rank_path) {
    const uint32_t RS = 32, off = (uint32_t)((len + 1) & ~1);
    rec32.assign(N * RS + 64, 0);
    qrank.assign((size_t)S.L * Q, 0xffffffffu);
    for (uint64_t i = 0; i < N; ++i) memcpy(&rec32[i * RS], &S.codes[i * len], len);
    for (int l = 0; l < S.L; ++l) {
      std::vector<uint64_t> distinct(keys[l]);
      std::sort(distinct.begin(), distinct.end());
      distinct.erase(std::unique(distinct.begin(), distinct.end()), distinct.end());
      if (distinct.size() > 65535) return false;
      for (uint64_t i = 0; i < N; ++i) {
        const uint16_t r = (uint16_t)(std::lower_bound(distinct.begin(), distinct.end(), keys[l][i]) - distinct.begin());
        memcpy(&rec32[i * RS + off + 2 * l], &r, 2);
      }
      for (uint32_t q = 0; q < Q; ++q) {
        auto it = std::lower_bound(distinct.begin(), distinct.end(), qkeys[(size_t)l * Q + q]);
        if (it != distinct.end() && *it == qkeys[(size_t)l * Q + q]) qrank[(size_t)l * Q + q] = (uint32_t)(it - distinct.begin());
      }
    }
    ea.rec = rec32.data();
    ea.rec_stride = RS;
    ea.rec_rank_off = off;
    ea.qrank = qrank.data();
    ea.keys = nullptr;
    ea.qkeys = nullptr;
    ea.qvalid = nullptr;
  }
  ea.hit_cap = surv.size() + 1;
  std::vector<HitKey> got;
  if (!run_exact<NV, REP>(S, ea, got)) return false;
  std::vector<HitKey> w = as_keys(want, nw);
  for (auto &h : w) h.id += 1000000;
  if (nw == 0 || got.size() != w.size() || !std::equal(got.begin(), got.end(), w.begin())) {
    printf("  search: %zu hits, oracle %llu (or a field differs) among %zu survivors\n", got.size(), (unsigned long long)nw, surv.size());
    return false;
  }
  return true;
}

// brute force: every (query, fragment) pair is a survivor; sqrt predicate; Euclidean or the integer metric
static bool test_brute(int len, double R, uint64_t N, uint32_t Q, bool integer, unsigned seed) {
  Setup S;
  S.len = len; S.K = 1; S.L = 1; S.W = 50; S.R = R; S.N = N; S.Q = Q;
  make_setup(S, seed);
  std::vector<orc_hit> want(N * Q);
  const uint64_t nw = integer ? orc_bruteforce_int(S.codes.data(), N, S.qcodes.data(), Q, len, (int)R, want.data(), want.size())
                              : orc_bruteforce(S.db.data(), N, S.q.data(), Q, S.dim, R, 1, want.data(), want.size());
  std::vector<Survivor> surv;
  for (uint32_t q = 0; q < Q; ++q)
    for (uint64_t i = 0; i < N; ++i) surv.push_back(Survivor{q, 0u, (uint32_t)i, 2u});
  std::vector<const uint32_t *> ids(1, nullptr);
  ExactArgs ea;
  memset(&ea, 0, sizeof ea);
  ea.surv = surv.data();
  ea.nsurv = surv.size();
  ea.mode = kModeBrute;
  ea.metric = integer ? HS_METRIC_BLOSUM_INT : HS_METRIC_EUCLID_FP64;
  ea.predicate = HS_PRED_SQRT_LE_R;
  ea.len = len; ea.dim = S.dim; ea.key_words = 1; ea.L = 1;
  ea.R = R;
  ea.sorted_ids = ids.data();
  ea.codes = S.codes.data();
  ea.rec = S.rec.data();
  ea.rec_stride = (len + 15u) & ~15u;
  ea.N = N;
  ea.table64 = S.table;
  ea.metric_tab = S.metric;
  if (integer) ea.qcodes = S.qcodes.data();
  else ea.q64 = S.q.data();
  ea.Q = Q;
  ea.hit_cap = surv.size() + 1;
  std::vector<HitKey> got;
  const bool ok = len <= 16 ? run_exact<1, 1>(S, ea, got) : run_exact<2, 1>(S, ea, got);
  if (!ok) return false;
  std::vector<HitKey> w = as_keys(want, nw);
  if (nw == 0 || got.size() != w.size() || !std::equal(got.begin(), got.end(), w.begin())) {
    printf("  brute force: %zu hits, oracle %llu (or a field differs)\n", got.size(), (unsigned long long)nw);
    return false;
  }
  return true;
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  int nbad = 0;
  auto report = [&](const char *what, bool ok) {
    printf("%s -> %s\n", what, ok ? "ok" : "FAILED");
    if (!ok) ++nbad;
  };
  report("search len 10 W 50 R 30, residue-string queries, plain pair table", test_search<1, 1>(10, 50.0, 30.0, 4000, 40, true, 1));
  report("search len 10 W 50 R 30, residue-string queries, replicated pair table", test_search<1, 8>(10, 50.0, 30.0, 4000, 40, true, 2));
  report("search len 10 W 50 R 36, dense queries", test_search<1, 1>(10, 50.0, 36.0, 3000, 30, false, 3));
  report("search len 10 W 50 R 30, rank path (u16 ranks in 32-byte records), replicated pair table",
         test_search<1, 8>(10, 50.0, 30.0, 4000, 40, true, 8, true));
  report("search len 10 W 50 R 34, rank path, dense queries", test_search<1, 1>(10, 50.0, 34.0, 3000, 30, false, 9, true));
  report("search len 25 W 120 R 60, residue-string queries", test_search<2, 1>(25, 120.0, 60.0, 2000, 30, true, 4));
  report("brute force len 10 R 38 (sqrt predicate)", test_brute(10, 38.0, 1500, 20, false, 5));
  report("brute force len 10 R 40, integer metric", test_brute(10, 40.0, 1500, 20, true, 6));
  report("brute force len 20 R 90, integer metric", test_brute(20, 90.0, 800, 12, true, 7));
  return nbad ? 1 : 0;
}
