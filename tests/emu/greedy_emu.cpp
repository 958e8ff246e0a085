// CPU emulation run of the greedy centre clustering kernel (hsearch_b200/csrc/cluster.cu: greedy_round_kernel,
// one warp per bucket) against a sequential restatement of Clustering() of hclust2.cpp:86-151 (= hclust3.cpp:
// 87-152): L rounds; in round l every point not yet merged into a cluster (state != 2) sits in the bucket of
// its key in table l; inside a bucket, members in ascending id order, the centres are the members already in
// state 1, then every state-0 member joins the first centre within R (sqrt(d2) <= R) -- which becomes a real
// centre (state 1) -- or becomes a candidate centre itself.  Bucket keys and distances come from the oracle
// (oracle/hs_oracle.c, linked in).  Compile with -ffp-contract=off.  greedy_kernels.inc is cut out of cluster.cu
// by tests/test_emu_greedy.py.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <limits>
#include <map>
#include <random>
#include <string>
#include <vector>

#include "../../include/hsearch_b200.h"
#include "cuda_emu.h"

extern "C" {
void orc_get_coordinates_print6(double *out160);
void orc_lsh_generate(uint64_t seed, uint32_t dim, uint32_t K, double W, double *a, double *b);
void orc_hash_points(const double *pts, uint64_t N, uint32_t dim, const double *a, const double *b, uint32_t K, uint32_t L,
                     double W, int *out);
void orc_embed(const uint8_t *codes, uint32_t len, const double *table160, double *point);
double orc_dist2(const double *x, const double *y, uint32_t dim);
}

namespace hs {
#include "greedy_kernels.inc"
}  // namespace hs

using namespace hs;

template <int NV>
static bool test_greedy(int len, int K, int L, double W, double R, uint32_t N, int family, unsigned seed) {
  const int dim = len * HS_CDIM;
  double table[HS_AA * HS_CDIM];
  orc_get_coordinates_print6(table);
  std::mt19937 rng(seed);
  // families of near-duplicates (up to two substitutions) so that clusters form
  const uint32_t nfam = std::max<uint32_t>(1, N / family);
  std::vector<uint8_t> roots((size_t)nfam * len), codes((size_t)N * len);
  for (auto &c : roots) c = (uint8_t)(rng() % 20);
  for (uint32_t i = 0; i < N; ++i) {
    memcpy(&codes[(size_t)i * len], &roots[(size_t)(i % nfam) * len], len);
    for (int s = 0; s < (int)(rng() % 3); ++s) codes[(size_t)i * len + rng() % len] = (uint8_t)(rng() % 20);
  }
  std::vector<double> pts((size_t)N * dim);
  for (uint32_t i = 0; i < N; ++i) orc_embed(&codes[(size_t)i * len], len, table, &pts[(size_t)i * dim]);
  std::vector<double> a((size_t)L * K * dim), b((size_t)L * K);
  for (int l = 0; l < L; ++l) orc_lsh_generate(4242 + seed + l, dim, K, W, &a[(size_t)l * K * dim], &b[(size_t)l * K]);
  std::vector<int> bk((size_t)N * L * K);
  orc_hash_points(pts.data(), N, dim, a.data(), b.data(), K, L, W, bk.data());
  auto key_of = [&](uint32_t i, int l) {
    std::string s;
    for (int k = 0; k < K; ++k) s += std::to_string(bk[((size_t)i * L + l) * K + k]);   // HashKey, lsh.hpp:51-59
    return s;
  };
  // ---- the reference's procedure, sequentially
  std::vector<uint8_t> want_state(N, 0);
  std::vector<uint32_t> want_center(N);
  for (uint32_t i = 0; i < N; ++i) want_center[i] = i;
  for (int l = 0; l < L; ++l) {
    std::map<std::string, std::vector<uint32_t>> tab;
    for (uint32_t i = 0; i < N; ++i)
      if (want_state[i] != 2) tab[key_of(i, l)].push_back(i);
    for (auto &kv : tab) {
      const std::vector<uint32_t> &ids = kv.second;
      std::vector<uint32_t> centers;
      for (uint32_t id : ids)
        if (want_state[id] == 1) centers.push_back(id);
      for (uint32_t id : ids) {
        if (want_state[id] == 0) {
          for (uint32_t c : centers)
            if (sqrt(orc_dist2(&pts[(size_t)id * dim], &pts[(size_t)c * dim], dim)) <= R) {
              want_center[id] = c;
              want_state[c] = 1;
              want_state[id] = 2;
              break;
            }
        }
        if (want_state[id] == 0) centers.push_back(id);
      }
    }
  }
  // ---- the kernel: one index per table over ALL points (as hs_build_index leaves it), ids ascending in a bucket
  const uint32_t RS = (len + 15u) & ~15u;
  std::vector<uint8_t> rec((size_t)N * RS + 64, 0);
  for (uint32_t i = 0; i < N; ++i) memcpy(&rec[(size_t)i * RS], &codes[(size_t)i * len], len);
  std::vector<uint8_t> state(N, 0);
  std::vector<uint32_t> center(N), round_of(N, 0xffffffffu), scratch(N, 0);
  for (uint32_t i = 0; i < N; ++i) center[i] = i;
  std::vector<float> dsq32(HS_AA * HS_AA, 0.f);
  const float thr = std::numeric_limits<float>::infinity();   // (the FP32 bound is an optimisation: switched off here)
  for (int l = 0; l < L; ++l) {
    std::map<std::string, std::vector<uint32_t>> tab;
    for (uint32_t i = 0; i < N; ++i) tab[key_of(i, l)].push_back(i);
    std::vector<uint32_t> bstart(1, 0), ids;
    for (auto &kv : tab) {
      ids.insert(ids.end(), kv.second.begin(), kv.second.end());
      bstart.push_back((uint32_t)ids.size());
      if (bstart.size() % 7 == 0) bstart.push_back((uint32_t)ids.size());   // empty slots, as on the rank path
    }
    const uint64_t nslots = bstart.size() - 1;
    const unsigned grid = (unsigned)((nslots * 32 + kGreedyThreads - 1) / kGreedyThreads);
    if (!emu_launch(grid, kGreedyThreads, [&]() {
          greedy_round_kernel<NV>(bstart.data(), nslots, ids.data(), rec.data(), RS, len, table, dsq32.data(), thr, R, (uint32_t)l,
                                  state.data(), center.data(), round_of.data(), scratch.data());
        }))
      return false;
  }
  uint32_t joined = 0, centres = 0;
  for (uint32_t i = 0; i < N; ++i) {
    if (state[i] != want_state[i] || center[i] != want_center[i]) {
      printf("  point %u: state %d centre %u, expected state %d centre %u\n", i, state[i], center[i], want_state[i], want_center[i]);
      return false;
    }
    if ((state[i] == 2) != (round_of[i] != 0xffffffffu)) return false;
    joined += state[i] == 2;
    centres += state[i] == 1;
  }
  printf("  (%u points: %u joined a cluster, %u real centres)\n", N, joined, centres);
  return joined > N / 4 && centres > 0;
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  int nbad = 0;
  auto report = [&](const char *what, bool ok) {
    printf("%s -> %s\n", what, ok ? "ok" : "FAILED");
    if (!ok) ++nbad;
  };
  report("len 10 K 4 L 8 W 50 R 25, families of 8", test_greedy<1>(10, 4, 8, 50.0, 25.0, 3000, 8, 1));
  report("len 10 K 2 L 3 W 80 R 30, families of 40 (buckets of hundreds of members)", test_greedy<1>(10, 2, 3, 80.0, 30.0, 2000, 40, 2));
  report("len 25 K 4 L 4 W 120 R 45, families of 6", test_greedy<2>(25, 4, 4, 120.0, 45.0, 1500, 6, 3));
  return nbad ? 1 : 0;
}
