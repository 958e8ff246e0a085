// oracle/_ref harness, TU 2: the reference's motif_both_points_noLSH.cpp
// compiled in place (main renamed).  TEST INFRASTRUCTURE ONLY.
#include <sstream>
#include <fstream>
#include <iostream>
#include <cstdio>
#include <ctime>
#include <unistd.h>
#define main hs_ref_nolsh_main
#include "hclust/src/hclust/motif_both_points_noLSH.cpp"
#undef main

struct ref_hit {
  uint32_t query;
  uint32_t table_first;
  uint64_t db_id;
  double dist2;  // here: the reference's PairwiseDistance (WITH sqrt), squared is not available
};

extern "C" {
// Search() (motif_both_points_noLSH.cpp:36-56).  The non-hit dump
// (<out>notlessthan.txt, :40-42) is pre-symlinked to /dev/null.  hits[i].dist2
// holds PairwiseDistance() (sqrt applied) recomputed with the reference's own
// function at full precision; printed[i] is the 6-digit text.
uint64_t ref_bruteforce(const double *db, uint64_t N, const double *queries, uint32_t Q, uint32_t dim,
                        double R, const char *tmp_path, ref_hit *hits, double *printed, uint64_t cap,
                        double *seconds) {
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
  std::string nxot = std::string(tmp_path) + "notlessthan.txt";
  std::remove(nxot.c_str());
  if (symlink("/dev/null", nxot.c_str()) != 0) return (uint64_t)-1;
  clock_t t0 = clock();
  Search(kmers, centers, kn, cn, R, tmp_path);
  clock_t t1 = clock();
  if (seconds) *seconds = (t1 - t0) / (double)CLOCKS_PER_SEC;
  std::remove(nxot.c_str());
  std::ifstream fin(tmp_path);
  std::string qn, dn;
  double dis;
  uint64_t nh = 0;
  while (fin >> qn >> dn >> dis) {
    if (nh < cap) {
      uint32_t q = (uint32_t)strtoul(qn.c_str() + 1, nullptr, 10);
      uint64_t j = strtoull(dn.c_str() + 1, nullptr, 10);
      hits[nh].query = q;
      hits[nh].table_first = 0;
      hits[nh].db_id = j;
      hits[nh].dist2 = PairwiseDistance(kmers[j], centers[q]);
      if (printed) printed[nh] = dis;
    }
    nh++;
  }
  std::remove(tmp_path);
  return nh;
}
}
