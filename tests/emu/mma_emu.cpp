// CPU emulation run of the pipelined tensor filter (hsearch_b200/csrc/filter_mma.cu: filter_mma_kernel -- 22 warps
// in four roles handing work through mbarriers: scheduler + bulk-copy loader, A-tile producers, the single MMA
// issuer, epilogue warps reading the accumulators from tensor memory).  The kernel body runs unchanged; what the
// hardware provides is emulated below, synchronously: mbarriers (arrival count + transaction bytes + phase),
// cp.async.bulk (a copy that completes its bytes on the barrier), tensor memory (128 lanes x 512 FP32 columns),
// tcgen05.mma on the no-swizzle K-major shared-memory descriptors the kernel builds (M = 128, N from the
// instruction descriptor, K = 16, FP16 x FP16 -> FP32), tcgen05.commit, tcgen05.ld 32x32b.x16.  Every (query,
// member) pair of the work list is also evaluated directly, with the same accumulation order as the emulated
// MMA, and the kernel's survivor set must equal that evaluation -- which in turn keeps every pair the oracle's
// brute force finds within R.  This checks the pipeline's LOGIC (unit and tile walk, row / column masks, stage and
// ring reuse, id hand-over, survivor slots and flushes); the GPU tests check it on the real tensor cores.
// mma_kernels.inc is cut out of the source by tests/test_emu_mma.py.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <random>
#include <set>
#include <tuple>
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
void orc_embed(const uint8_t *codes, uint32_t len, const double *table160, double *point);
uint64_t orc_bruteforce(const double *db, uint64_t N, const double *queries, uint32_t Q, uint32_t dim, double R, int pred,
                        orc_hit *hits, uint64_t cap);
}

namespace hs {
void set_error(const char *, ...) {}

// ---- the hardware the kernel talks to ------------------------------------------------------------
static float emu_tmem[128][512];
// shared-space addresses: byte offsets from the dynamic shared memory arena (operands live in it; the barriers
// of the static MmaShared lie elsewhere in the image, which a 32-bit signed offset still reaches)
static inline uint32_t smem_addr(const void *p) { return (uint32_t)(int32_t)((const char *)p - (const char *)emu_dyn_smem); }
static inline void *smem_ptr(uint32_t a) { return (char *)emu_dyn_smem + (int32_t)a; }
struct EmuMbar {   // the 64-bit mbarrier object
  uint16_t pending, init;
  int32_t tx : 31;
  uint32_t phase : 1;
};
static_assert(sizeof(EmuMbar) == 8, "mbarrier objects are 64-bit");
static inline void mbar_check(EmuMbar *b) {
  if (b->pending == 0 && b->tx == 0) {
    b->phase ^= 1u;
    b->pending = b->init;
  }
}
static inline void mbar_init(uint32_t bar, uint32_t count) {
  EmuMbar *b = (EmuMbar *)smem_ptr(bar);
  b->pending = b->init = (uint16_t)count;
  b->tx = 0;
  b->phase = 0;
}
static inline bool mbar_try_wait(uint32_t bar, uint32_t parity) {   // true once the phase of that parity has completed
  const EmuMbar *b = (const EmuMbar *)smem_ptr(bar);
  if (b->phase != parity) return true;
  emu_yield();
  return false;
}
static inline bool mbar_test_wait(uint32_t bar, uint32_t parity) { return mbar_try_wait(bar, parity); }
static inline void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
static inline void mbar_wait_spin(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
static inline void mbar_arrive(uint32_t bar) {
  EmuMbar *b = (EmuMbar *)smem_ptr(bar);
  if (b->pending == 0) {
    fprintf(stderr, "emu: arrival on a completed mbarrier phase\n");
    abort();
  }
  --b->pending;
  mbar_check(b);
}
static inline void mbar_arrive_warp(uint32_t bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
static inline void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  EmuMbar *b = (EmuMbar *)smem_ptr(bar);
  b->tx += (int32_t)bytes;
  --b->pending;
  mbar_check(b);
}
static inline void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  memcpy(smem_ptr(dst), src, bytes);
  EmuMbar *b = (EmuMbar *)smem_ptr(bar);
  b->tx -= (int32_t)bytes;
  mbar_check(b);
}
static inline void fence_async_shared() {}
static inline void tc_before() {}
static inline void tc_after() {}
static inline void emu_tmem_alloc(uint32_t *slot) { *slot = 0u; }
// D[128][N] (+)= A[128][16] * B[N][16]^T, operands by their shared-memory matrix descriptors
static inline void mma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
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
static inline void mma_commit(uint32_t bar) { mbar_arrive(bar); }   // (the emulated MMAs have completed when issued)
static inline void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  const uint32_t lane = (taddr >> 16) + (threadIdx.x & 31), col = taddr & 0xffffu;
  for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(emu_tmem[lane][col + i]);
}
static inline void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  const uint32_t lane = (taddr >> 16) + (threadIdx.x & 31), col = taddr & 0xffffu;
  for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(emu_tmem[lane][col + i]);
}
static inline void tmem_ld_wait() {}
static inline float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
static inline float fmaxf3_unused() { return 0.f; }

#include "mma_kernels.inc"
}  // namespace hs

using namespace hs;

typedef std::tuple<uint32_t, uint32_t, uint32_t> Triple;   // (index into the query list, table, fragment id)

static bool test_mma(int len, double R, uint64_t N, uint32_t Q, unsigned seed) {
  const int dim = len * HS_CDIM;
  hs_ctx ctx_storage;
  hs_ctx *ctx = &ctx_storage;
  memset(&ctx->prm, 0, sizeof ctx->prm);
  ctx->prm.len = len; ctx->prm.R = R; ctx->prm.metric = HS_METRIC_EUCLID_FP64;
  ctx->dim = dim;
  orc_get_coordinates_print6(ctx->table64);
  memcpy(ctx->ftable64, ctx->table64, sizeof ctx->ftable64);
  ctx->have_ftable = true;
  MmaGeometry g;
  if (mma_geometry(ctx, &g) != HS_OK || mma_upload_tables(ctx) != HS_OK) return false;
  if (g.smem > sizeof emu_dyn_smem) return false;
  const __half *tab16 = ctx->d_tab16.as<__half>();
  const float *nx32 = reinterpret_cast<const float *>(ctx->d_tab16.as<char>() + sizeof(__half) * HS_AA * HS_CDIM);
  std::mt19937 rng(seed);
  std::vector<uint8_t> codes(N * len), qcodes((size_t)Q * len);
  for (auto &c : codes) c = (uint8_t)(rng() % 20);
  for (uint32_t q = 0; q < Q; ++q) {
    const uint64_t src = rng() % N;
    memcpy(&qcodes[(size_t)q * len], &codes[src * len], len);
    for (int s = 0; s < (int)(rng() % 4); ++s) qcodes[(size_t)q * len + rng() % len] = (uint8_t)(rng() % 20);
  }
  std::vector<__half> qb16((size_t)Q * g.kp);
  if (!emu_launch((Q + 127) / 128, 128, [&]() { build_qb_codes_kernel(qcodes.data(), 0, Q, len, g.kp, ctx->table64, g.beta, qb16.data()); }))
    return false;
  // two "tables": table 0 in identity order, table 1 a permutation of the fragments with its id list
  const uint64_t npad = (N + 15) & ~15ull;
  std::vector<uint8_t> store0((size_t)len * npad + 512, 0), store1((size_t)len * npad + 512, 0);
  std::vector<uint32_t> ids1(N);
  for (uint64_t i = 0; i < N; ++i) ids1[i] = (uint32_t)i;
  std::shuffle(ids1.begin(), ids1.end(), rng);
  for (uint64_t i = 0; i < N; ++i)
    for (int p = 0; p < len; ++p) {
      store0[(uint64_t)p * npad + i] = (uint8_t)(codes[i * len + p] * kCodeScale);
      store1[(uint64_t)p * npad + i] = (uint8_t)(codes[(uint64_t)ids1[i] * len + p] * kCodeScale);
    }
  const uint8_t *stores[2] = {store0.data(), store1.data()};
  const uint32_t *sorted_ids[2] = {nullptr, ids1.data()};
  // work list: items (a bucket's query set) of various widths, units (member ranges) of various sizes and alignments
  std::vector<uint32_t> qlist;
  std::vector<MmaItem> items;
  std::vector<MmaUnit> units;
  const uint32_t widths[] = {9, 130, 256, 300, (uint32_t)g.qmax, 16, 1};
  uint32_t nextm = 3;
  for (uint32_t w : widths) {
    w = std::min<uint32_t>(w, std::min<uint32_t>(Q, (uint32_t)g.qmax));
    MmaItem it;
    it.table = (uint32_t)(items.size() & 1);
    it.q_begin = (uint32_t)qlist.size();
    for (uint32_t i = 0; i < w; ++i) qlist.push_back((uint32_t)(rng() % Q));
    it.q_end = (uint32_t)qlist.size();
    it.pad = 0;
    items.push_back(it);
    const int nunits = 1 + (int)(rng() % 3);
    for (int u = 0; u < nunits; ++u) {
      MmaUnit un;
      un.item = (uint32_t)items.size() - 1;
      un.m_begin = nextm % (uint32_t)N;
      const uint32_t span = 1 + (uint32_t)(rng() % 700);
      un.m_end = std::min<uint32_t>((uint32_t)N, un.m_begin + span);
      un.pad = 0;
      units.push_back(un);
      nextm = un.m_end + (uint32_t)(rng() % 50);
      if (nextm >= N) nextm = (uint32_t)(rng() % 100);
    }
  }
  // launch_filter_mma: arguments as the kernel receives them
  const double rr = R * R;
  const double r2 = rr * (1.0 + 1e-12) + 1e-30;
  float thr = (float)r2;
  if ((double)thr < r2) thr = nextafterf(thr, INFINITY);
  const float beta = (float)(g.beta * 1.0001);
  std::vector<Survivor> surv((size_t)8 << 20);
  unsigned long long surv_count = 0;
  uint32_t unit_counter = 0;
  uint4 tab16v[HS_AA];
  memcpy(tab16v, tab16, sizeof tab16v);
  MmaArgs a;
  memset(&a, 0, sizeof a);
  a.items = items.data();
  a.units = units.data();
  a.nunits = (uint32_t)units.size();
  a.unit_counter = &unit_counter;
  a.qlist = qlist.data();
  a.qb16 = qb16.data();
  a.tq_base = 0;
  a.stores = stores;
  a.sorted_ids = sorted_ids;
  a.npad = npad;
  a.len = len; a.kp = g.kp; a.nstages = g.nstages; a.qmax = g.qmax; a.cring = g.cring;
  a.thr = thr; a.beta = beta;
  a.tab16 = tab16v;
  a.nx32 = nx32;
  a.surv = surv.data();
  a.surv_cap = surv.size();
  a.surv_count = &surv_count;
  bool ok;
  if (len <= 10) ok = emu_launch(3, kMmaThreads, [&]() { filter_mma_kernel<10>(a); });
  else ok = emu_launch(3, kMmaThreads, [&]() { filter_mma_kernel<16>(a); });
  if (!ok || surv_count > surv.size()) return false;
  std::multiset<Triple> got;
  for (unsigned long long i = 0; i < surv_count; ++i) {
    if (surv[i].pad != 3) return false;
    got.insert(Triple(surv[i].query, surv[i].table, surv[i].pos));
  }
  // direct evaluation of every (query, member) pair of the work list, same accumulation order as the emulated MMA
  std::multiset<Triple> want;
  for (const MmaUnit &un : units) {
    const MmaItem &it = items[un.item];
    for (uint32_t pos = un.m_begin; pos < un.m_end; ++pos) {
      const uint32_t id = it.table ? ids1[pos] : pos;
      float nx = 0.f;
      for (int p = 0; p < len; ++p) nx += nx32[codes[(uint64_t)id * len + p]];
      float rt = 0.5f * (nx * (1.0f - 4e-6f) * (1.0f - beta) - thr);
      rt -= (nx + thr) * 2.4e-7f + 1e-6f;
      for (uint32_t qi = it.q_begin; qi < it.q_end; ++qi) {
        const __half *brow = &qb16[(size_t)qlist[qi] * g.kp];
        float D = 0.f;
        for (int k = 0; k < g.kp; ++k) {
          float av = 0.f;
          if (k < dim) av = (float)tab16[codes[(uint64_t)id * len + k / HS_CDIM] * HS_CDIM + k % HS_CDIM];
          else if (k < dim + 2) av = 1.0f;
          D += av * (float)brow[k];
        }
        if (D >= rt) want.insert(Triple(qi, it.table, id));
      }
    }
  }
  if (got != want) {
    printf("  %zu survivors, the direct evaluation gives %zu\n", got.size(), want.size());
    return false;
  }
  // ... and nothing within R is lost: the oracle's brute force over the same pairs
  std::vector<double> db(N * dim), qp((size_t)Q * dim);
  for (uint64_t i = 0; i < N; ++i) orc_embed(&codes[i * len], len, ctx->table64, &db[i * dim]);
  for (uint32_t q = 0; q < Q; ++q) orc_embed(&qcodes[(size_t)q * len], len, ctx->table64, &qp[(size_t)q * dim]);
  std::vector<orc_hit> hits(N * Q);
  const uint64_t nh = orc_bruteforce(db.data(), N, qp.data(), Q, dim, R, 0, hits.data(), hits.size());
  std::set<uint64_t> near;
  for (uint64_t i = 0; i < nh; ++i) near.insert((uint64_t)hits[i].query * N + hits[i].db_id);
  uint64_t pairs_within = 0;
  for (const MmaUnit &un : units) {
    const MmaItem &it = items[un.item];
    for (uint32_t pos = un.m_begin; pos < un.m_end; ++pos) {
      const uint32_t id = it.table ? ids1[pos] : pos;
      for (uint32_t qi = it.q_begin; qi < it.q_end; ++qi)
        if (near.count((uint64_t)qlist[qi] * N + id)) {
          ++pairs_within;
          if (!got.count(Triple(qi, it.table, id))) {
            printf("  a pair within R was dropped\n");
            return false;
          }
        }
    }
  }
  printf("  (%zu units, %zu items, %zu survivors, %llu of the work list's pairs within R; %llu polling yields)\n", units.size(),
         items.size(), got.size(), (unsigned long long)pairs_within, (unsigned long long)emu_yields);
  return got.size() > 0 && pairs_within > 0;
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  int nbad = 0;
  auto report = [&](const char *what, bool ok) {
    printf("%s -> %s\n", what, ok ? "ok" : "FAILED");
    if (!ok) ++nbad;
  };
  report("len 10, R 30: 7 items of 1 .. 512 queries, ragged units, two tables", test_mma(10, 30.0, 3000, 600, 1));
  report("len 10, R 36: another work list", test_mma(10, 36.0, 2000, 520, 2));
  report("len 16, R 44", test_mma(16, 44.0, 1500, 300, 3));
  return nbad ? 1 : 0;
}
