// CPU emulation run of the PRODUCTION hash (hsearch_b200/csrc/hash.cu: hash_fast_kernel -- FP32 residue-projection
// partial sums with the FP64 re-evaluation inside a guard band, dense u16 bucket ranks, fragment records) in the
// configurations the library launches, with the projection data prepared by the library's own host code:
// setup_projection / setup_ranks run unchanged on the real hs_ctx over a stand-in CUDA runtime
// (tests/emu/stub/cuda_runtime.h: device memory = host memory), so the guard bands, the bucket ranges and the
// tuple -> rank tables under test are the product's.  Checked against the oracle (oracle/hs_oracle.c, lsh.hpp:33-59):
// bucket ints bit-equal for every fragment, table and projection; ranks = position of the fragment's key string
// among the table's possible strings; records = codes + ranks; packed keys on the packed-key path.  The FP32 path
// only adds and multiplies without contraction, so SSE arithmetic (-ffp-contract=off) reproduces the device's.
// hashfast_kernels.inc / hashfast_host.inc are cut out of hash.cuh / hash.cu by tests/test_emu_hashfast.py.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <random>
#include <string>
#include <vector>

#include "cuda_emu.h"
#include "../../hsearch_b200/csrc/common.cuh"

extern "C" {
void orc_get_coordinates_print6(double *out160);
void orc_lsh_generate(uint64_t seed, uint32_t dim, uint32_t K, double W, double *a, double *b);
void orc_hash_codes(const uint8_t *codes, uint64_t N, uint32_t len, const double *table160, const double *a, const double *b,
                    uint32_t K, uint32_t L, double W, int *out);
}

namespace hs {
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fputc('\n', stderr);
}
constexpr int kHashThreads = 256;
constexpr int kHashRepThreads = 1024;
#include "hashfast_kernels.inc"
#include "hashfast_host.inc"
}  // namespace hs

using namespace hs;

static uint64_t pack1(const int *buckets, int K, int *nchars, uint64_t *hi) {   // two-word packed key of a bucket tuple
  std::string s;
  for (int k = 0; k < K; ++k) s += std::to_string(buckets[k]);
  uint64_t w0 = 0, w1 = 0;
  for (char c : s) {
    w1 = (w1 << 4) | (w0 >> 60);
    w0 = (w0 << 4) | (c == '-' ? 11u : (uint64_t)(c - '0' + 1));
  }
  *nchars = (int)s.size();
  *hi = w1;
  return w0;
}

template <int NQ, int KW, bool RANK, int REP, int NT, bool K4, int G>
static bool launch(hs_ctx *ctx, const HashChunkArgs &args, int chunk, const uint8_t *codes, uint64_t N, int32_t *buckets,
                   unsigned long long *counters, unsigned grid) {
  const int len = (int)ctx->prm.len, P = 4 * NQ;
  unsigned int tile_counter = 2u * grid * G;
  const float *T = ctx->d_T32.as<float>() + (size_t)chunk * len * HS_AA * P;
  return emu_launch(grid, NT, [&]() {
    hash_fast_kernel<NQ, KW, RANK, REP, NT, K4, G>(codes, N, len, T, ctx->d_b32.as<float>() + (size_t)chunk * P,
                                                    ctx->d_eps32.as<float>() + (size_t)chunk * P, (float)(1.0 / ctx->prm.W),
                                                    ctx->table64, ctx->d_a64.as<double>(), ctx->d_b64.as<double>(), ctx->prm.W,
                                                    (int)ctx->prm.K, (int)ctx->Kp, (int)ctx->prm.L, (int)ctx->dim, args, buckets,
                                                    counters, 0, N, &tile_counter);
  });
}

static bool test_hash_fast(uint32_t len, uint32_t K, uint32_t L, double W, uint64_t N, bool expect_rank, unsigned seed,
                           unsigned long long *guard_hits) {
  hs_ctx ctx_storage;
  hs_ctx *ctx = &ctx_storage;
  memset(&ctx->prm, 0, sizeof ctx->prm);
  ctx->prm.len = len; ctx->prm.K = K; ctx->prm.L = L; ctx->prm.W = W; ctx->prm.R = 30.0;
  ctx->dim = len * HS_CDIM;
  orc_get_coordinates_print6(ctx->table64);
  std::vector<double> a((size_t)L * K * ctx->dim), b((size_t)L * K);
  for (uint32_t l = 0; l < L; ++l) orc_lsh_generate(12345 + seed + l, ctx->dim, K, W, &a[(size_t)l * K * ctx->dim], &b[(size_t)l * K]);
  if (setup_projection(ctx, a.data(), b.data()) != HS_OK) return false;   // the library's host code, unchanged
  if (ctx->rank_mode != expect_rank) {
    printf("  rank path %d, expected %d\n", (int)ctx->rank_mode, (int)expect_rank);
    return false;
  }
  std::mt19937 rng(seed);
  std::vector<uint8_t> codes(N * len + 64, 0);
  for (uint64_t i = 0; i < N * len; ++i) codes[i] = (uint8_t)(rng() % 20);
  std::vector<int> want(N * L * K);
  orc_hash_codes(codes.data(), N, len, ctx->table64, a.data(), b.data(), K, L, W, want.data());
  ctx->N = N;
  ctx->npad = (N + 15) & ~15ull;
  std::vector<int32_t> buckets(N * L * K, 12345678);
  std::vector<uint16_t> ranks((size_t)L * ctx->npad + 64, 0xffff);
  std::vector<uint8_t> rec((size_t)N * ctx->rec_stride + 64, 0xee);
  std::vector<std::vector<uint64_t>> keys(L, std::vector<uint64_t>((size_t)ctx->key_words * N, 0));
  unsigned long long counters[32];
  memset(counters, 0, sizeof counters);
  // launch_hash_fast, chunk by chunk
  const bool full_rec = ctx->rank_mode && ctx->nchunks == 1;
  for (uint32_t chunk = 0; chunk < ctx->nchunks; ++chunk) {
    HashChunkArgs args;
    memset(&args, 0, sizeof args);
    args.l0 = (int)(chunk * ctx->tpc);
    args.ntab = (int)std::min<uint32_t>(ctx->tpc, L - args.l0);
    for (int t = 0; t < args.ntab; ++t) {
      const uint32_t l = (uint32_t)args.l0 + t;
      if (ctx->rank_mode) {
        args.ranks[t] = ranks.data() + (size_t)l * ctx->npad;
        args.lut[t] = ctx->d_lut.as<uint16_t>() + ctx->rank_lut_off[l];
        for (uint32_t k = 0; k < K; ++k) {
          args.lo[t * ctx->Kp + k] = ctx->rank_lo[(size_t)l * K + k];
          args.rng[t * ctx->Kp + k] = ctx->rank_rng[(size_t)l * K + k];
        }
      } else {
        args.keys[t] = keys[l].data();
      }
    }
    for (uint32_t sl = 0; sl < 4 * ctx->nq; ++sl) {
      args.b32[sl] = ctx->h_b32[(size_t)chunk * 4 * ctx->nq + sl];
      args.eps32[sl] = ctx->h_eps32[(size_t)chunk * 4 * ctx->nq + sl];
    }
    args.rec = rec.data();
    args.rec_stride = ctx->rec_stride;
    args.rec_rank_off = ctx->rec_rank_off;
    args.full_rec = full_rec ? 1 : 0;
    args.k4_full = (K == 4 && (uint32_t)args.ntab * 4 == 4 * ctx->nq) ? 1 : 0;
    bool ok;
    const unsigned nq = ctx->nq, kw = ctx->key_words;
    if (ctx->rank_mode && nq == 4 && args.k4_full)        // the headline configuration's launch
      ok = launch<4, 1, true, 8, kHashRepThreads, true, 8>(ctx, args, chunk, codes.data(), N, buckets.data(), counters, 3);
    else if (ctx->rank_mode && nq == 2)
      ok = launch<2, 1, true, 8, kHashRepThreads, false, 8>(ctx, args, chunk, codes.data(), N, buckets.data(), counters, 2);
    else if (ctx->rank_mode && nq == 1)
      ok = launch<1, 1, true, 1, kHashThreads, false, 1>(ctx, args, chunk, codes.data(), N, buckets.data(), counters, 5);
    else if (!ctx->rank_mode && nq == 4 && kw == 1)
      ok = launch<4, 1, false, 1, kHashThreads, false, 1>(ctx, args, chunk, codes.data(), N, buckets.data(), counters, 4);
    else if (!ctx->rank_mode && nq == 4 && kw == 2)
      ok = launch<4, 2, false, 1, kHashThreads, false, 1>(ctx, args, chunk, codes.data(), N, buckets.data(), counters, 4);
    else if (!ctx->rank_mode && nq == 2 && kw == 1)
      ok = launch<2, 1, false, 1, kHashThreads, false, 1>(ctx, args, chunk, codes.data(), N, buckets.data(), counters, 4);
    else {
      printf("  configuration nq=%u key_words=%u rank=%d is not instantiated in this harness\n", nq, kw, (int)ctx->rank_mode);
      return false;
    }
    if (!ok) return false;
  }
  *guard_hits = counters[0];
  if (memcmp(buckets.data(), want.data(), sizeof(int) * want.size()) != 0) {
    printf("  bucket ints differ from the oracle\n");
    return false;
  }
  if (counters[2] != 0) {
    printf("  %llu buckets outside the derived range / over-long keys\n", counters[2]);
    return false;
  }
  for (uint64_t i = 0; i < N; ++i)
    for (uint32_t l = 0; l < L; ++l) {
      int nc;
      uint64_t hi;
      const uint64_t w0 = pack1(&want[(i * L + l) * K], (int)K, &nc, &hi);
      if (ctx->rank_mode) {
        const uint16_t r = ranks[(size_t)l * ctx->npad + i];
        const std::vector<uint64_t> &rk = ctx->h_rkeys[l];   // [KW][nr]: the packed key of every rank, ascending
        const uint32_t nr = ctx->rank_nr[l];
        bool ok = r < nr && rk[r] == w0 && (ctx->key_words < 2 || rk[(size_t)nr + r] == hi);
        if (ok && r > 0) ok = ctx->key_words >= 2 ? true : rk[r - 1] < rk[r];
        uint16_t in_rec;
        memcpy(&in_rec, &rec[i * ctx->rec_stride + ctx->rec_rank_off + 2 * l], 2);
        if (!ok || in_rec != r || memcmp(&rec[i * ctx->rec_stride], &codes[i * len], len) != 0) {
          printf("  fragment %llu table %u: rank / record differs\n", (unsigned long long)i, l);
          return false;
        }
      } else if (keys[l][i] != w0 || (ctx->key_words >= 2 && keys[l][N + i] != hi)) {
        printf("  fragment %llu table %u: packed key differs\n", (unsigned long long)i, l);
        return false;
      }
    }
  return true;
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  int nbad = 0;
  unsigned long long guard_total = 0;
  auto report = [&](const char *what, bool ok, unsigned long long g) {
    printf("%s -> %s  (%llu projections re-evaluated in FP64)\n", what, ok ? "ok" : "FAILED", g);
    guard_total += g;
    if (!ok) ++nbad;
  };
  unsigned long long g;
  bool ok;
  ok = test_hash_fast(10, 4, 4, 50.0, 20000, true, 1, &g);
  report("len 10 K 4 L 4 W 50: rank path, replicated tables, 1024 threads in 8 groups (the headline launch)", ok, g);
  ok = test_hash_fast(10, 4, 2, 50.0, 6000, true, 2, &g);
  report("len 10 K 4 L 2 W 50: rank path, two quads", ok, g);
  ok = test_hash_fast(10, 4, 4, 20.0, 6000, false, 6, &g);
  report("len 10 K 4 L 4 W 20: packed keys", ok, g);
  ok = test_hash_fast(16, 2, 1, 40.0, 3000, true, 3, &g);
  report("len 16 K 2 L 1 W 40: rank path, plain table layout", ok, g);
  ok = test_hash_fast(10, 4, 4, 1.0, 5000, false, 4, &g);
  report("len 10 K 4 L 4 W 1: packed keys (too many possible buckets for ranks)", ok, g);
  ok = test_hash_fast(10, 4, 4, 0.001, 1500, false, 5, &g);
  report("len 10 K 4 L 4 W 0.001: nearly every projection inside the guard band", ok, g);
  printf("guard band exercised: %s\n", guard_total > 1000 ? "yes -> ok" : "no -> FAILED");
  if (guard_total <= 1000) ++nbad;
  return nbad ? 1 : 0;
}
