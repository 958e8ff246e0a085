// oracle/_ref harness, TU 1: the reference's own motif_both_points.cpp compiled
// in place (main renamed), exposing its LSH class and its Search() through a
// small C ABI.  TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Built by
// oracle/Makefile only where /root/reference exists; no reference source is
// copied into this repository.
#include "fixed_rd.hpp"
#include <sstream>
#include <fstream>
#include <iostream>
#include <algorithm>
#include <cstdio>
#include <ctime>
#define private public  // reach LSH::a / LSH::b (lsh.hpp:61-68) without editing the file
#define main hs_ref_motif_both_points_main
#include "hclust/src/hclust/motif_both_points.cpp"
#undef main
#undef private

struct ref_hit {
  uint32_t query;
  uint32_t table_first;  // not observable from the reference's output; left 0xffffffff
  uint64_t db_id;
  double dist2;          // PairwiseDistance_square of the reference
};

extern "C" {

// LSH::LSH (lsh.hpp:10-31) with the engine seeded `seed`.
void ref_lsh_generate(uint64_t seed, uint32_t dim, uint32_t K, double W, double *a, double *b) {
  hs_fixed_rd::reset(seed);
  LSH lsh(dim, K, W);
  for (uint32_t k = 0; k < K; ++k) {
    for (uint32_t i = 0; i < dim; ++i) a[(size_t)k * dim + i] = lsh.a[k][i];
    b[k] = lsh.b[k];
  }
}

// LSH::HashBucketIndex / HashKey (lsh.hpp:44-59) over N dense points, L tables
// seeded seed_base + l.  buckets [N][L][K]; keys: N*L strings of stride keystride.
void ref_hash_points(const double *pts, uint64_t N, uint32_t dim, uint32_t K, uint32_t L, double W,
                     uint64_t seed_base, int *buckets, char *keys, uint32_t keystride) {
  hs_fixed_rd::reset(seed_base);
  std::vector<LSH> funs;
  for (uint32_t l = 0; l < L; ++l) funs.push_back(LSH(dim, K, W));
  std::vector<double> p(dim);
  for (uint64_t i = 0; i < N; ++i) {
    p.assign(pts + i * dim, pts + (i + 1) * dim);
    for (uint32_t l = 0; l < L; ++l) {
      for (uint32_t k = 0; k < K; ++k) buckets[(i * L + l) * K + k] = funs[l].HashBucketIndex(p, k);
      if (keys) {
        std::string s = funs[l].HashKey(p);
        snprintf(keys + (i * L + l) * keystride, keystride, "%s", s.c_str());
      }
    }
  }
}

// Search() (motif_both_points.cpp:195-250) as shipped.  Names are "q<i>" and
// "k<j>"; the text output is parsed back into (query, db id).  Returns #hits in
// the reference's output order; table_sizes from the "table size" lines (:217);
// printed[i] = the distance as printed (6 significant digits, :240-241);
// seconds = clock() around Search() exactly like main (:373,384).
uint64_t ref_search(const double *db, uint64_t N, const double *queries, uint32_t Q, uint32_t dim,
                    uint32_t K, uint32_t L, double W, double R, uint64_t seed_base,
                    const char *tmp_path, ref_hit *hits, double *printed, uint64_t cap,
                    uint64_t *table_sizes, double *seconds) {
  DIMENSION = dim;
  KMERLENGTH = dim / AACoordinateSize;
  std::vector<Point> kmers(N), centers(Q);
  std::vector<std::string> kn(N), cn(Q);
  for (uint64_t i = 0; i < N; ++i) {
    kmers[i].data.assign(db + i * dim, db + (i + 1) * dim);
    kn[i] = "k" + std::to_string(i);
  }
  for (uint32_t i = 0; i < Q; ++i) {
    centers[i].data.assign(queries + (size_t)i * dim, queries + (size_t)(i + 1) * dim);
    cn[i] = "q" + std::to_string(i);
  }
  hs_fixed_rd::reset(seed_base);
  std::ostringstream captured;
  std::streambuf *old = std::cout.rdbuf(captured.rdbuf());
  clock_t t0 = clock();
  Search(kmers, centers, kn, cn, K, L, W, R, tmp_path);
  clock_t t1 = clock();
  std::cout.rdbuf(old);
  if (seconds) *seconds = (t1 - t0) / (double)CLOCKS_PER_SEC;
  if (table_sizes) {
    std::istringstream iss(captured.str());
    std::string w1, w2;
    uint64_t v, l = 0;
    while (iss >> w1 >> w2 >> v)
      if (w1 == "table" && w2 == "size" && l < L) table_sizes[l++] = v;
  }
  std::ifstream fin(tmp_path);
  std::string qn, dn;
  double dis;
  uint64_t nh = 0;
  while (fin >> qn >> dn >> dis) {
    if (nh < cap) {
      uint32_t q = (uint32_t)strtoul(qn.c_str() + 1, nullptr, 10);
      uint64_t j = strtoull(dn.c_str() + 1, nullptr, 10);
      hits[nh].query = q;
      hits[nh].table_first = 0xffffffffu;
      hits[nh].db_id = j;
      hits[nh].dist2 = PairwiseDistance_square(kmers[j], centers[q]);
      if (printed) printed[nh] = dis;
    }
    nh++;
  }
  std::remove(tmp_path);
  return nh;
}

// The index-build half only (lsh.hpp HashKey + unordered_map insert,
// motif_both_points.cpp:206-218), for stage-level CPU timing.
double ref_build_tables_seconds(const double *db, uint64_t N, uint32_t dim, uint32_t K, uint32_t L,
                                double W, uint64_t seed_base, uint64_t *table_sizes) {
  DIMENSION = dim;
  std::vector<Point> kmers(N);
  for (uint64_t i = 0; i < N; ++i) kmers[i].data.assign(db + i * dim, db + (i + 1) * dim);
  hs_fixed_rd::reset(seed_base);
  clock_t t0 = clock();
  std::vector<HashTable> lsh_tables(L);
  std::vector<LSH> lsh_funs;
  for (uint32_t l = 0; l < L; ++l) lsh_funs.push_back(LSH(dim, K, W));
  for (uint32_t l = 0; l < L; ++l) BuildLSHTalbe(kmers, lsh_funs[l], lsh_tables[l]);
  clock_t t1 = clock();
  if (table_sizes)
    for (uint32_t l = 0; l < L; ++l) table_sizes[l] = lsh_tables[l].size();
  return (t1 - t0) / (double)CLOCKS_PER_SEC;
}

// weight() (motif_both_points.cpp:67-87)
double ref_weight(double dis, double R) { return weight(dis, R); }

// evaulate() (motif_both_points.cpp:100-165) on the reference's two text files; also writes
// <out>.accuracy.txt and prints the reference's own lines to stdout
double ref_evaluate(const char *ground_truth, const char *output_file, double R) {
  return evaulate(ground_truth, output_file, R);
}
}
