// Numerical check of the tensor filter's margins (hsearch_b200/csrc/filter_mma.cu; DESIGN.md "Filter margins"):
// the filter keeps a (query, member) pair iff  D = <x~_m, q~> + c_q >= rowthr_m  with FP16 operands and FP32
// accumulation, and must never drop a pair whose FP64 distance is within R (motif_both_points.cpp:176-183).
// The tcgen05 pipeline itself cannot run on a CPU; what runs here, unchanged, is everything that DETERMINES the
// decision: the library's FP16 embedding rows and rounded-down row norms (mma_upload_tables), its geometry and
// error bound beta (mma_geometry), its query rows with the split constant c_q (build_qb_points_kernel /
// build_qb_codes_kernel, write_cq) -- cut out of the source -- and the row-threshold and threshold formulas of
// the kernel (two statements each, restated below; tests/test_emu_margin.py fails if the source's change).  D is
// accumulated in FP32 in four different orders (the hardware's is unspecified) and the SMALLEST must pass for
// every pair the oracle's brute force finds within R.  margin_kernels.inc is cut out by the test.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <random>
#include <vector>

#include "cuda_emu.h"
#include "../../hsearch_b200/csrc/common.cuh"

typedef _Float16 __half;
static inline __half __double2half(double d) { return (__half)d; }
static inline __half __float2half(float f) { return (__half)f; }
static inline float __half2float(__half h) { return (float)h; }

extern "C" {
typedef struct {
  uint32_t query;
  uint32_t table_first;
  uint64_t db_id;
  double dist2;
} orc_hit;
void orc_get_coordinates_print6(double *out160);
void orc_get_coordinates(double *out160);
void orc_embed(const uint8_t *codes, uint32_t len, const double *table160, double *point);
uint64_t orc_bruteforce(const double *db, uint64_t N, const double *queries, uint32_t Q, uint32_t dim, double R, int pred,
                        orc_hit *hits, uint64_t cap);
}

namespace hs {
void set_error(const char *, ...) {}
constexpr int kMmaMaxCodeRing = 8;
constexpr int kMmaAGroupBytes = 2048;
#include "margin_kernels.inc"
}  // namespace hs

using namespace hs;

// FP32 accumulation of kp exact products in several orders; returns the smallest result
static float accumulate_min(const std::vector<float> &t) {
  const int n = (int)t.size();
  float fwd = 0.f, bwd = 0.f;
  for (int i = 0; i < n; ++i) fwd += t[i];
  for (int i = n - 1; i >= 0; --i) bwd += t[i];
  std::vector<float> v(t);
  for (int m = n; m > 1; m = (m + 1) / 2)             // pairwise tree
    for (int i = 0; i < m / 2; ++i) v[i] = v[2 * i] + v[2 * i + 1], v[(m + 1) / 2 - 1] = (m & 1) ? v[m - 1] : v[(m + 1) / 2 - 1];
  float tree = 0.f;
  {
    std::vector<float> w(t);
    int m = n;
    while (m > 1) {
      int o = 0;
      for (int i = 0; i + 1 < m; i += 2) w[o++] = w[i] + w[i + 1];
      if (m & 1) w[o++] = w[m - 1];
      m = o;
    }
    tree = w[0];
  }
  float chunks = 0.f;                                  // K = 16 per MMA instruction: chunk sums, then a running FP32 sum
  for (int c = 0; c < n; c += 16) {
    double s = 0.0;
    for (int i = c; i < std::min(n, c + 16); ++i) s += (double)t[i];
    chunks += (float)s;
  }
  (void)v;
  return std::min(std::min(fwd, bwd), std::min(tree, chunks));
}

static bool test_margin(int len, double R, bool print6, uint64_t N, uint32_t Q, bool code_queries, unsigned seed) {
  const int dim = len * HS_CDIM;
  hs_ctx ctx_storage;
  hs_ctx *ctx = &ctx_storage;
  memset(&ctx->prm, 0, sizeof ctx->prm);
  ctx->prm.len = len; ctx->prm.R = R; ctx->prm.metric = HS_METRIC_EUCLID_FP64;
  ctx->dim = dim;
  if (print6) orc_get_coordinates_print6(ctx->table64);
  else orc_get_coordinates(ctx->table64);
  memcpy(ctx->ftable64, ctx->table64, sizeof ctx->ftable64);   // Euclidean metric: the filter embeds with the coordinates
  ctx->have_ftable = true;
  MmaGeometry g;
  if (mma_geometry(ctx, &g) != HS_OK || mma_upload_tables(ctx) != HS_OK) return false;   // the library's host code
  const __half *tab16 = ctx->d_tab16.as<__half>();
  const float *nx32 = reinterpret_cast<const float *>(ctx->d_tab16.as<char>() + sizeof(__half) * HS_AA * HS_CDIM);
  std::mt19937 rng(seed);
  std::vector<uint8_t> codes(N * len), qcodes((size_t)Q * len);
  for (auto &c : codes) c = (uint8_t)(rng() % 20);
  for (uint32_t q = 0; q < Q; ++q) {
    const uint64_t src = rng() % N;
    memcpy(&qcodes[(size_t)q * len], &codes[src * len], len);
    for (int s = 0; s < (int)(rng() % 5); ++s) qcodes[(size_t)q * len + rng() % len] = (uint8_t)(rng() % 20);
  }
  std::vector<double> db(N * dim), qp((size_t)Q * dim);
  for (uint64_t i = 0; i < N; ++i) orc_embed(&codes[i * len], len, ctx->table64, &db[i * dim]);
  for (uint32_t q = 0; q < Q; ++q) orc_embed(&qcodes[(size_t)q * len], len, ctx->table64, &qp[(size_t)q * dim]);
  if (!code_queries)   // dense centres off the lattice of residue strings (Pfam-style centres)
    for (auto &v : qp) v += ((double)(rng() % 2001) - 1000.0) * 1e-3;
  std::vector<orc_hit> want(N * Q);
  const uint64_t nw = orc_bruteforce(db.data(), N, qp.data(), Q, dim, R, 0, want.data(), want.size());
  // the B operand: one FP16 row per query, coordinates + c_q split hi / lo
  std::vector<__half> qb16((size_t)Q * g.kp);
  bool ok;
  if (code_queries)
    ok = emu_launch((Q + 127) / 128, 128, [&]() { build_qb_codes_kernel(qcodes.data(), 0, Q, len, g.kp, ctx->table64, g.beta, qb16.data()); });
  else
    ok = emu_launch((Q + 127) / 128, 128, [&]() { build_qb_points_kernel(qp.data(), Q, dim, g.kp, g.beta, qb16.data()); });
  if (!ok) return false;
  // launch_filter_mma: threshold and beta as the kernel receives them
  const double rr = R * R;
  const double r2 = rr * (1.0 + 1e-12) + 1e-30;
  float thr = (float)r2;
  if ((double)thr < r2) thr = nextafterf(thr, INFINITY);
  const float beta = (float)(g.beta * 1.0001);
  std::vector<uint8_t> is_hit(N * Q, 0);
  for (uint64_t i = 0; i < nw; ++i) is_hit[(uint64_t)want[i].query * N + want[i].db_id] = 1;
  uint64_t passed = 0, dropped = 0;
  std::vector<float> t((size_t)g.kp);
  for (uint64_t m = 0; m < N; ++m) {
    // the A row of member m and its threshold, as the producer warps build them
    float nx = 0.f;
    for (int p = 0; p < len; ++p) nx += nx32[codes[m * len + p]];
    float rt = 0.5f * (nx * (1.0f - 4e-6f) * (1.0f - beta) - thr);
    rt -= (nx + thr) * 2.4e-7f + 1e-6f;
    for (uint32_t q = 0; q < Q; ++q) {
      const __half *brow = &qb16[(size_t)q * g.kp];
      for (int k = 0; k < g.kp; ++k) {
        float av = 0.f;
        if (k < dim) av = (float)tab16[codes[m * len + k / HS_CDIM] * HS_CDIM + k % HS_CDIM];
        else if (k < dim + 2) av = 1.0f;
        t[k] = av * (float)brow[k];      // exact: two 11-bit significands
      }
      const float D = accumulate_min(t);
      const bool pass = D >= rt;
      passed += pass;
      if (is_hit[(uint64_t)q * N + m] && !pass) ++dropped;
    }
  }
  printf("  (kp %d, beta %.3e: %llu pairs within R, %llu pass the filter among %llu)\n", g.kp, g.beta, (unsigned long long)nw,
         (unsigned long long)passed, (unsigned long long)(N * Q));
  if (dropped) {
    printf("  %llu pairs within R would be DROPPED\n", (unsigned long long)dropped);
    return false;
  }
  return nw > 20 && passed >= nw && passed <= nw + nw / 2 + 20;   // safe, and tight (the exact stage sees few extra pairs)
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  int nbad = 0;
  auto report = [&](const char *what, bool ok) {
    printf("%s -> %s\n", what, ok ? "ok" : "FAILED");
    if (!ok) ++nbad;
  };
  report("len 10, R 30, residue-string queries, print6 table", test_margin(10, 30.0, true, 4000, 40, true, 1));
  report("len 10, R 34, dense queries off the lattice, full table", test_margin(10, 34.0, false, 4000, 40, false, 2));
  report("len 25, R 60, residue-string queries", test_margin(25, 60.0, true, 1500, 24, true, 3));
  report("len 8, R 24, dense queries", test_margin(8, 24.0, true, 3000, 30, false, 4));
  return nbad ? 1 : 0;
}
