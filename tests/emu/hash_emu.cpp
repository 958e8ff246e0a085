// CPU emulation run of the FP64 hash kernels (hsearch_b200/csrc/hash.cu: hash_exact_kernel, the all-FP64 hash
// and audit in the reference's operation order, lsh.hpp:33-59; hash_queries_kernel, motif_both_points.cpp:227)
// and the packed digit-string key (hash.cuh: KeyBuilder) against the oracle (oracle/hs_oracle.c, linked in):
// bucket ints bit-equal, keys equal to the nibble packing of the concatenated std::to_string strings.
// Compile with -ffp-contract=off.  hash_kernels.inc is cut out of the sources by tests/test_emu_hash.py.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <random>
#include <string>
#include <vector>

#include "../../include/hsearch_b200.h"
#include "cuda_emu.h"

extern "C" {
void orc_get_coordinates_print6(double *out160);
void orc_get_coordinates(double *out160);
void orc_lsh_generate(uint64_t seed, uint32_t dim, uint32_t K, double W, double *a, double *b);
void orc_hash_codes(const uint8_t *codes, uint64_t N, uint32_t len, const double *table160, const double *a, const double *b,
                    uint32_t K, uint32_t L, double W, int *out);
void orc_hash_points(const double *pts, uint64_t N, uint32_t dim, const double *a, const double *b, uint32_t K, uint32_t L,
                     double W, int *out);
void orc_embed(const uint8_t *codes, uint32_t len, const double *table160, double *point);
}

namespace hs {
constexpr int kHashThreads = 256;
#include "hash_kernels.inc"
}  // namespace hs

using namespace hs;

// HashKey (lsh.hpp:51-59): std::to_string of every bucket, no separator; packed one nibble per character
// ('0'..'9' -> 1..10, '-' -> 11), appended left to right, word 0 least significant
template <int KW>
static void pack_expected(const int *buckets, int K, uint64_t (&w)[KW], int *nchars) {
  std::string s;
  for (int k = 0; k < K; ++k) s += std::to_string(buckets[k]);
  for (int i = 0; i < KW; ++i) w[i] = 0;
  for (char c : s) {
    const uint64_t nib = c == '-' ? 11 : (uint64_t)(c - '0' + 1);
    for (int j = KW - 1; j > 0; --j) w[j] = (w[j] << 4) | (w[j - 1] >> 60);
    w[0] = (w[0] << 4) | nib;
  }
  *nchars = (int)s.size();
}

template <int KW>
static bool test_hash(int len, int K, int L, double W, bool print6, uint64_t N, uint32_t Q, unsigned seed) {
  const int dim = len * HS_CDIM;
  double table[HS_AA * HS_CDIM];
  if (print6) orc_get_coordinates_print6(table);
  else orc_get_coordinates(table);
  std::vector<double> a((size_t)L * K * dim), b((size_t)L * K);
  for (int l = 0; l < L; ++l) orc_lsh_generate(12345 + seed + l, dim, K, W, a.data() + (size_t)l * K * dim, b.data() + (size_t)l * K);
  std::mt19937 rng(seed);
  std::vector<uint8_t> codes(N * len);
  for (auto &c : codes) c = (uint8_t)(rng() % 20);
  std::vector<int> want(N * L * K), got(N * L * K, 12345678);
  orc_hash_codes(codes.data(), N, len, table, a.data(), b.data(), K, L, W, want.data());
  std::vector<std::vector<uint64_t>> keys(L, std::vector<uint64_t>((size_t)KW * N, 0));
  std::vector<uint64_t *> kptr(L);
  for (int l = 0; l < L; ++l) kptr[l] = keys[l].data();
  unsigned long long counters[4] = {0, 0, 0, 0};
  const unsigned grid = (unsigned)((N + kHashThreads - 1) / kHashThreads);
  auto launch = [&](int audit) {
    return emu_launch(grid, kHashThreads, [&]() {
      hash_exact_kernel<KW>(codes.data(), N, len, table, a.data(), b.data(), W, K, L, dim, 0, L, kptr.data(), got.data(), audit,
                            nullptr, nullptr, nullptr, 0, counters);
    });
  };
  if (!launch(0)) return false;
  if (got != want) {
    printf("  bucket ints differ from the oracle\n");
    return false;
  }
  int overlong = 0;
  for (uint64_t i = 0; i < N; ++i)
    for (int l = 0; l < L; ++l) {
      uint64_t w[KW];
      int nc;
      pack_expected<KW>(&want[(i * L + l) * K], K, w, &nc);
      if (nc > 16 * KW) {
        ++overlong;
        continue;
      }
      for (int j = 0; j < KW; ++j)
        if (keys[l][(uint64_t)j * N + i] != w[j]) {
          printf("  packed key of fragment %llu, table %d differs\n", (unsigned long long)i, l);
          return false;
        }
    }
  if (counters[2] != (unsigned long long)overlong || counters[3] != 0) {
    printf("  counters: overflow %llu (expected %d), flips %llu\n", counters[2], overlong, counters[3]);
    return false;
  }
  // the audit finds nothing, then exactly the one key that was tampered with
  counters[2] = 0;
  if (!launch(1) || counters[3] != 0) return false;
  keys[L - 1][N / 2] ^= 0x10;
  if (!launch(1) || counters[3] != 1) {
    printf("  audit: %llu flips reported after one key was changed\n", counters[3]);
    return false;
  }
  // queries as dense points
  std::vector<uint8_t> qcodes((size_t)Q * len);
  for (auto &c : qcodes) c = (uint8_t)(rng() % 20);
  std::vector<double> q((size_t)Q * dim);
  for (uint32_t i = 0; i < Q; ++i) orc_embed(qcodes.data() + (size_t)i * len, len, table, q.data() + (size_t)i * dim);
  std::vector<int> qwant((size_t)Q * L * K);
  orc_hash_points(q.data(), Q, dim, a.data(), b.data(), K, L, W, qwant.data());
  int GP = 1;
  while (GP < K) GP <<= 1;
  std::vector<double> a64t((size_t)L * dim * GP, 0.0);
  for (int l = 0; l < L; ++l)
    for (int k = 0; k < K; ++k)
      for (int i = 0; i < dim; ++i) a64t[((size_t)l * dim + i) * GP + k] = a[((size_t)l * K + k) * dim + i];
  std::vector<uint64_t> qkeys((size_t)L * Q * KW, 0);
  std::vector<uint8_t> qvalid((size_t)L * Q, 0xff);
  const uint64_t nthreads = (uint64_t)Q * L * GP;
  if (!emu_launch((unsigned)((nthreads + 127) / 128), 128, [&]() {
        hash_queries_kernel<KW>(q.data(), Q, dim, a64t.data(), b.data(), W, K, GP, L, qkeys.data(), qvalid.data());
      }))
    return false;
  for (int l = 0; l < L; ++l)
    for (uint32_t i = 0; i < Q; ++i) {
      uint64_t w[KW];
      int nc;
      pack_expected<KW>(&qwant[((size_t)i * L + l) * K], K, w, &nc);
      const bool valid = nc <= 16 * KW;
      if ((qvalid[(size_t)l * Q + i] != 0) != valid) {
        printf("  query %u table %d: validity differs\n", i, l);
        return false;
      }
      if (valid)
        for (int j = 0; j < KW; ++j)
          if (qkeys[((size_t)l * Q + i) * KW + j] != w[j]) {
            printf("  query %u table %d: key differs\n", i, l);
            return false;
          }
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
  report("len 10 K 4 L 4 W 50 print6, one key word", test_hash<1>(10, 4, 4, 50.0, true, 3000, 70, 1));
  report("len 10 K 4 L 4 W 20 full table, one key word", test_hash<1>(10, 4, 4, 20.0, false, 2000, 33, 2));
  report("len 25 K 4 L 2 W 4 (negative, two-digit buckets), two key words", test_hash<2>(25, 4, 2, 4.0, true, 1500, 40, 3));
  report("len 10 K 16 L 2 W 10, four key words", test_hash<4>(10, 16, 2, 10.0, true, 1200, 50, 4));
  report("len 8 K 3 L 5 W 0.05, one key word (keys longer than 16 characters are counted, not stored)", test_hash<1>(8, 3, 5, 0.05, true, 600, 20, 5));
  report("len 1 K 1 L 1 W 1", test_hash<1>(1, 1, 1, 1.0, true, 300, 10, 6));
  return nbad ? 1 : 0;
}
