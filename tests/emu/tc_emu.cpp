// CPU emulation run of the one-hot tensor filter (hsearch_b200/csrc/filter_tc.cu: tq_to_half_kernel, filter_tc_kernel
// -- the tcgen05 fallback of the integer metric for len > 30 / HS_NO_MMA_INT, and the first tensor filter of the
// Euclidean metric) over the same emulated hardware as tests/emu/mma_emu.cpp (mbarrier, tensor memory, tcgen05.mma on
// the kernel's own descriptors, commit, tcgen05.ld).  With A one-hot the accumulator is the sum of len FP16 table
// entries in ascending position order, so the kernel's survivor set must equal a direct evaluation, and it must keep
// every pair the oracle's brute force finds within R (Euclidean: tables rounded DOWN to FP16; integer metric: exact).
// tc_kernels.inc is cut out of the sources by tests/test_emu_tc.py.
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

typedef _Float16 __half;
static inline __half __float2half(float f) { return (__half)f; }
static inline __half __float2half_rd(float f) {   // round toward minus infinity
  __half h = (__half)f;
  if ((float)h > f) {
    uint16_t b;
    memcpy(&b, &h, 2);
    if (b == 0x0000) b = 0x8001;                    // +0 -> smallest negative subnormal
    else if (b & 0x8000) ++b;                       // negative: larger magnitude
    else --b;                                       // positive: smaller magnitude
    memcpy(&h, &b, 2);
  }
  return h;
}

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

// ---- the hardware the kernel talks to (as in mma_emu.cpp) ---------------------------------------
static float emu_tmem[128][512];
static inline uint32_t smem_u32(const void *p) { return (uint32_t)(int32_t)((const char *)p - (const char *)emu_dyn_smem); }
static inline void *smem_ptr(uint32_t a) { return (char *)emu_dyn_smem + (int32_t)a; }
struct EmuMbar {
  uint16_t pending, init;
  int32_t tx : 31;
  uint32_t phase : 1;
};
static inline void mbar_init(uint32_t bar, uint32_t count) {
  EmuMbar *b = (EmuMbar *)smem_ptr(bar);
  b->pending = b->init = (uint16_t)count;
  b->tx = 0;
  b->phase = 0;
}
static inline void mbar_wait(uint32_t bar, uint32_t parity) {
  while (((const EmuMbar *)smem_ptr(bar))->phase == parity) emu_yield();
}
static inline void tc_fence_before() {}
static inline void tc_fence_after() {}
static inline void fence_async_smem() {}
static inline void emu_tmem_alloc(uint32_t *slot) { *slot = 0u; }
static inline void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t N = ((idesc >> 17) & 0x3fu) << 3;
  const uint32_t a0 = (uint32_t)(adesc & 0x3fffu) << 4, alb = (uint32_t)((adesc >> 16) & 0x3fffu) << 4, asb = (uint32_t)((adesc >> 32) & 0x3fffu) << 4;
  const uint32_t b0 = (uint32_t)(bdesc & 0x3fffu) << 4, blb = (uint32_t)((bdesc >> 16) & 0x3fffu) << 4, bsb = (uint32_t)((bdesc >> 32) & 0x3fffu) << 4;
  const uint32_t col0 = tmem_d & 0xffffu;
  for (uint32_t r = 0; r < 128; ++r)
    for (uint32_t n = 0; n < N; ++n) {
      float acc = accumulate ? emu_tmem[r][col0 + n] : 0.f;
      for (uint32_t k = 0; k < 16; ++k) {
        const __half av = *(const __half *)smem_ptr(a0 + (k >> 3) * alb + (r >> 3) * asb + (r & 7) * 16 + (k & 7) * 2);
        const __half bv = *(const __half *)smem_ptr(b0 + (k >> 3) * blb + (n >> 3) * bsb + (n & 7) * 16 + (k & 7) * 2);
        acc += (float)av * (float)bv;
      }
      emu_tmem[r][col0 + n] = acc;
    }
}
static inline void umma_commit(uint32_t bar) {
  EmuMbar *b = (EmuMbar *)smem_ptr(bar);
  if (--b->pending == 0) {
    b->phase ^= 1u;
    b->pending = b->init;
  }
}
static inline void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  const uint32_t lane = (taddr >> 16) + (threadIdx.x & 31), col = taddr & 0xffffu;
  for (int i = 0; i < 16; ++i) v[i] = emu_tmem[lane][col + i];
}

#include "tc_kernels.inc"
}  // namespace hs

using namespace hs;

static bool test_tc(int len, double R, uint64_t N, uint32_t Q, bool integer, unsigned seed) {
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
  for (uint32_t q = 0; q < Q; ++q) {
    const uint64_t src = rng() % N;
    memcpy(&qcodes[(size_t)q * len], &codes[src * len], len);
    for (int s = 0; s < (int)(rng() % 4); ++s) qcodes[(size_t)q * len + rng() % len] = (uint8_t)(rng() % 20);
  }
  std::vector<double> db(N * dim), qp((size_t)Q * dim);
  for (uint64_t i = 0; i < N; ++i) orc_embed(&codes[i * len], len, ctx->table64, &db[i * dim]);
  for (uint32_t q = 0; q < Q; ++q) orc_embed(&qcodes[(size_t)q * len], len, ctx->table64, &qp[(size_t)q * dim]);
  std::vector<orc_hit> hits(N * Q);
  const uint64_t nh = integer ? orc_bruteforce_int(codes.data(), N, qcodes.data(), Q, len, (int)R, hits.data(), hits.size())
                              : orc_bruteforce(db.data(), N, qp.data(), Q, dim, R, 0, hits.data(), hits.size());
  // query tables, FP32 then FP16 (rounded down)
  const int klen = len * HS_AA;
  int kp, kc;
  size_t smem;
  if (tc_geometry(len, &kp, &kc, &smem) != HS_OK || smem > sizeof emu_dyn_smem) return false;
  const uint64_t ntq = (uint64_t)Q * klen;
  std::vector<float> tq(ntq);
  int metric[400];
  orc_blosum_metric(metric);
  std::vector<int32_t> metric32(metric, metric + 400);
  bool ok = integer ? emu_launch((unsigned)((ntq + 255) / 256), 256, [&]() { build_tq_int_kernel(qcodes.data(), Q, len, metric32.data(), tq.data()); })
                    : emu_launch((unsigned)((ntq + 255) / 256), 256, [&]() { build_tq_points_kernel(qp.data(), Q, len, ctx->table64, tq.data()); });
  std::vector<__half> tq16((size_t)Q * kp);
  ok = ok && emu_launch((unsigned)(((uint64_t)Q * kp + 255) / 256), 256, [&]() { tq_to_half_kernel(tq.data(), Q, klen, kp, tq16.data()); });
  if (!ok) return false;
  const uint64_t npad = (N + 15) & ~15ull;
  std::vector<uint8_t> store((size_t)len * npad + 256, 0);
  for (uint64_t i = 0; i < N; ++i)
    for (int p = 0; p < len; ++p) store[(uint64_t)p * npad + i] = (uint8_t)(codes[i * len + p] * kCodeScale);
  const uint8_t *stores[1] = {store.data()};
  // work items of <= 128 queries over ragged member ranges, blocks of `tpb` 128-member tiles
  const uint32_t tpb = 2;
  std::vector<uint32_t> qlist;
  std::vector<WorkItem> items;
  uint32_t nblocks = 0, nextm = 5;
  const uint32_t widths[] = {128, 7, 33, 100};
  for (uint32_t w : widths) {
    w = std::min<uint32_t>(w, Q);
    WorkItem it;
    it.table = 0;
    it.q_begin = (uint32_t)qlist.size();
    for (uint32_t i = 0; i < w; ++i) qlist.push_back((uint32_t)(rng() % Q));
    it.q_end = (uint32_t)qlist.size();
    it.m_begin = nextm;
    it.m_end = std::min<uint32_t>((uint32_t)N, nextm + 200 + (uint32_t)(rng() % 500));
    it.block_begin = nblocks;
    nblocks += (it.m_end - it.m_begin + tpb * kTcM - 1) / (tpb * kTcM);
    items.push_back(it);
    nextm = it.m_end % (uint32_t)(N - 800);
  }
  std::vector<Survivor> surv((size_t)1 << 20);
  unsigned long long count = 0;
  TcArgs a;
  memset(&a, 0, sizeof a);
  a.items = items.data();
  a.nitems = (uint32_t)items.size();
  a.qlist = qlist.data();
  a.tq16 = tq16.data();
  a.stores = stores;
  a.npad = npad;
  a.len = len; a.kp = kp; a.kc = kc;
  a.tiles_per_block = tpb;
  a.lbo = kTcColGroupBytes;
  a.sbo = 128;
  a.thr = filter_tc_threshold(ctx);   // the library's own threshold
  a.surv = surv.data();
  a.surv_cap = surv.size();
  a.surv_count = &count;
  if (!emu_launch(nblocks, kTcThreads, [&]() { filter_tc_kernel<kModeSearch>(a); }) || count > surv.size()) return false;
  std::multiset<std::pair<uint32_t, uint32_t>> got, want;   // (query id, member position) per work-list pair
  for (unsigned long long i = 0; i < count; ++i) got.insert({surv[i].query, surv[i].pos});
  std::set<uint64_t> near;
  for (uint64_t i = 0; i < nh; ++i) near.insert((uint64_t)hits[i].query * N + hits[i].db_id);
  uint64_t within = 0;
  for (const WorkItem &it : items)
    for (uint32_t m = it.m_begin; m < it.m_end; ++m)
      for (uint32_t qi = it.q_begin; qi < it.q_end; ++qi) {
        const uint32_t q = qlist[qi];
        float s = 0.f;
        for (int p = 0; p < len; ++p) s += (float)tq16[(size_t)q * kp + p * HS_AA + codes[(uint64_t)m * len + p]];
        if (s <= a.thr) want.insert({q, m});
        if (near.count((uint64_t)q * N + m)) {
          ++within;
          if (!(s <= a.thr)) {
            printf("  a pair within R would be dropped\n");
            return false;
          }
        }
      }
  if (got != want) {
    printf("  %zu survivors, the direct evaluation gives %zu\n", got.size(), want.size());
    return false;
  }
  printf("  (kp %d, %u blocks: %zu survivors, %llu of the work list's pairs within R)\n", kp, nblocks, got.size(), (unsigned long long)within);
  return within > 0 && got.size() >= within;
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  int nbad = 0;
  auto report = [&](const char *what, bool ok) {
    printf("%s -> %s\n", what, ok ? "ok" : "FAILED");
    if (!ok) ++nbad;
  };
  report("Euclidean, len 10, R 34", test_tc(10, 34.0, 3000, 200, false, 1));
  report("integer metric, len 10, R 60", test_tc(10, 60.0, 3000, 200, true, 2));
  report("integer metric, len 32, R 150 (the path's reason to exist: len > 30)", test_tc(32, 150.0, 2000, 150, true, 3));
  return nbad ? 1 : 0;
}
