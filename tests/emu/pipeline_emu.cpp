// The whole hot path on the CPU, stage after stage, against the oracle's Search() (oracle/hs_oracle.c;
// motif_both_points.cpp:195-250): every stage below is the library's kernel, unchanged, cut out of its source by
// tests/test_emu_pipeline.py (pipeline_kernels.inc, ~2,700 lines) and run over tests/emu/cuda_emu.h, and every stage
// consumes what the previous one produced:
//   setup_projection (host, over the stand-in CUDA runtime)  ->  hash_fast_kernel (headline launch configuration:
//   1024 threads in 8 groups, replicated tables, dense u16 ranks, 32-byte records)  ->  rank upsweep / scan /
//   downsweep / bounds per table  ->  gather_runs + gather_blocked (bucket-ordered code stores)  ->
//   hash_queries_kernel + probe_kernel  ->  a work list of one item per probed bucket  ->  filter_mma_kernel (over the
//   emulated mbarriers, bulk copies, tensor memory and MMA of tests/emu/mma_emu.cpp)  ->  exact_kernel (FP64, rank-path
//   first-table rule)  ->  seg_hist / scan / seg_scatter / seg_sort (the hit list's order).
// The final list must equal the oracle's: order, first table, ids and FP64 distances, bit for bit.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <numeric>
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
void orc_lsh_generate(uint64_t seed, uint32_t dim, uint32_t K, double W, double *a, double *b);
void orc_embed(const uint8_t *codes, uint32_t len, const double *table160, double *point);
uint64_t orc_search(const double *db, uint64_t N, const double *queries, uint32_t Q, uint32_t dim, const double *a,
                    const double *b, uint32_t K, uint32_t L, double W, double R, int pred, orc_hit *hits, uint64_t cap,
                    uint64_t *table_sizes, uint64_t *ncandidates);
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


#include "pipeline_kernels.inc"
}  // namespace hs

using namespace hs;

static bool scan_u32(std::vector<uint32_t> &v) {   // exclusive_scan_u32 with the library's three kernels
  const uint64_t n = v.size();
  if (n == 0) return true;
  const uint32_t nblocks = (uint32_t)((n + kScanTile - 1) / kScanTile);
  std::vector<uint32_t> sums(nblocks + 1, 0);
  bool ok = emu_launch(nblocks, kScanThreads, [&]() { scan_reduce_kernel(v.data(), n, sums.data()); });
  ok = ok && emu_launch(1, kScanThreads, [&]() { scan_sums_kernel(sums.data(), nblocks, nullptr); });
  ok = ok && emu_launch(nblocks, kScanThreads, [&]() { scan_downsweep_kernel(v.data(), v.data(), n, sums.data()); });
  return ok;
}

static int bits_for_(uint64_t nvalues) {
  int b = 1;
  while (b < 64 && (nvalues - 1) >> b) ++b;
  return b;
}

static bool test_pipeline(uint64_t N, uint32_t Q, double W, double R, unsigned seed) {
  const uint32_t len = 10, K = 4, L = 4;
  hs_ctx ctx_storage;
  hs_ctx *ctx = &ctx_storage;
  memset(&ctx->prm, 0, sizeof ctx->prm);
  ctx->prm.len = len; ctx->prm.K = K; ctx->prm.L = L; ctx->prm.W = W; ctx->prm.R = R;
  ctx->prm.metric = HS_METRIC_EUCLID_FP64;
  ctx->dim = len * HS_CDIM;
  const int dim = (int)ctx->dim;
  orc_get_coordinates_print6(ctx->table64);
  memcpy(ctx->ftable64, ctx->table64, sizeof ctx->ftable64);
  ctx->have_ftable = true;
  std::vector<double> a((size_t)L * K * dim), b((size_t)L * K);
  for (uint32_t l = 0; l < L; ++l) orc_lsh_generate(12345 + l, dim, K, W, &a[(size_t)l * K * dim], &b[(size_t)l * K]);
  if (setup_projection(ctx, a.data(), b.data()) != HS_OK || !ctx->rank_mode || ctx->nq != 4 || ctx->nchunks != 1) return false;
  std::mt19937 rng(seed);
  std::vector<uint8_t> codes(N * len + 64, 0), qcodes((size_t)Q * len);
  for (uint64_t i = 0; i < N * len; ++i) codes[i] = (uint8_t)(rng() % 20);
  for (uint32_t q = 0; q < Q; ++q) {
    if (q % 2 == 0) {
      const uint64_t src = rng() % N;
      memcpy(&qcodes[(size_t)q * len], &codes[src * len], len);
      for (int s = 0; s < (int)(rng() % 3); ++s) qcodes[(size_t)q * len + rng() % len] = (uint8_t)(rng() % 20);
    } else {
      for (uint32_t p = 0; p < len; ++p) qcodes[(size_t)q * len + p] = (uint8_t)(rng() % 20);
    }
  }
  std::vector<double> db(N * dim), qp((size_t)Q * dim);
  for (uint64_t i = 0; i < N; ++i) orc_embed(&codes[i * len], len, ctx->table64, &db[i * dim]);
  for (uint32_t q = 0; q < Q; ++q) orc_embed(&qcodes[(size_t)q * len], len, ctx->table64, &qp[(size_t)q * dim]);
  std::vector<orc_hit> want(N * Q);
  const uint64_t nw = orc_search(db.data(), N, qp.data(), Q, dim, a.data(), b.data(), K, L, W, R, 0, want.data(), want.size(), nullptr, nullptr);

  // ---- K1: the production hash -> ranks [L][npad], records [N][32]
  ctx->N = N;
  ctx->npad = (N + 15) & ~15ull;
  const uint64_t npad = ctx->npad;
  const uint32_t RS = ctx->rec_stride;
  std::vector<uint16_t> ranks((size_t)L * npad + 64, 0xffff);
  std::vector<uint8_t> rec((size_t)N * RS + 64, 0xee);
  unsigned long long counters[32];
  memset(counters, 0, sizeof counters);
  {
    HashChunkArgs args;
    memset(&args, 0, sizeof args);
    args.l0 = 0;
    args.ntab = (int)L;
    for (uint32_t t = 0; t < L; ++t) {
      args.ranks[t] = ranks.data() + (size_t)t * npad;
      args.lut[t] = ctx->d_lut.as<uint16_t>() + ctx->rank_lut_off[t];
      for (uint32_t k = 0; k < K; ++k) {
        args.lo[t * ctx->Kp + k] = ctx->rank_lo[(size_t)t * K + k];
        args.rng[t * ctx->Kp + k] = ctx->rank_rng[(size_t)t * K + k];
      }
    }
    for (uint32_t sl = 0; sl < 16; ++sl) {
      args.b32[sl] = ctx->h_b32[sl];
      args.eps32[sl] = ctx->h_eps32[sl];
    }
    args.rec = rec.data();
    args.rec_stride = RS;
    args.rec_rank_off = ctx->rec_rank_off;
    args.full_rec = 1;
    args.k4_full = 1;
    unsigned int tile_counter = 2u * 3 * 8;
    if (!emu_launch(3, kHashRepThreads, [&]() {
          hash_fast_kernel<4, 1, true, 8, kHashRepThreads, true, 8>(codes.data(), N, (int)len, ctx->d_T32.as<float>(), ctx->d_b32.as<float>(),
                                                                    ctx->d_eps32.as<float>(), (float)(1.0 / W), ctx->table64,
                                                                    ctx->d_a64.as<double>(), ctx->d_b64.as<double>(), W, (int)K, (int)ctx->Kp,
                                                                    (int)L, dim, args, nullptr, counters, 0, N, &tile_counter);
        }))
      return false;
    if (counters[2]) return false;
  }
  // ---- K2: per table the rank sort, the slot boundaries; then the code stores of all tables
  std::vector<std::vector<uint32_t>> ids(L, std::vector<uint32_t>(N)), bstart(L);
  for (uint32_t l = 0; l < L; ++l) {
    const uint32_t nr = ctx->rank_nr[l];
    const uint16_t *rk = ranks.data() + (size_t)l * npad;
    const uint32_t ntiles = (uint32_t)((N + kRkTile - 1) / kRkTile);
    std::vector<uint32_t> tile_hist((size_t)ntiles * 256), vtmp(N);
    std::vector<uint16_t> ktmp(N + 16), ksorted(N + 16);
    int hibits = 0;
    while (((uint64_t)256 << hibits) < nr) ++hibits;
    const int npass = nr > 256 ? 2 : 1;
    bool ok = true;
    for (int pi = 0; pi < npass; ++pi) {
      const int shift = 8 * pi;
      const uint32_t mask = pi == 0 ? 0xffu : ((1u << hibits) - 1u);
      const uint16_t *kin = pi == 0 ? rk : ktmp.data();
      tile_hist.assign((size_t)ntiles * 256 + 1, 0);
      ok = ok && emu_launch(ntiles, kRkThreads, [&]() { rank_upsweep_kernel(kin, N, shift, mask, tile_hist.data(), ntiles); });
      tile_hist.resize((size_t)ntiles * 256);
      ok = ok && scan_u32(tile_hist);
      tile_hist.resize((size_t)ntiles * 256 + 1);
      if (npass == 1)
        ok = ok && emu_launch(ntiles, kRkThreads, [&]() { rank_downsweep_kernel<true, true>(kin, nullptr, ksorted.data(), ids[l].data(), N, shift, mask, tile_hist.data(), ntiles); });
      else if (pi == 0)
        ok = ok && emu_launch(ntiles, kRkThreads, [&]() { rank_downsweep_kernel<true, false>(kin, nullptr, ktmp.data(), vtmp.data(), N, shift, mask, tile_hist.data(), ntiles); });
      else
        ok = ok && emu_launch(ntiles, kRkThreads, [&]() { rank_downsweep_kernel<false, true>(kin, vtmp.data(), ksorted.data(), ids[l].data(), N, shift, mask, tile_hist.data(), ntiles); });
    }
    bstart[l].assign(nr + 1, 0);
    unsigned int nb = 0;
    ok = ok && emu_launch((unsigned)(((N + 7) / 8 + 255) / 256), 256, [&]() { rank_bounds_kernel(ksorted.data(), N, nr, bstart[l].data(), &nb); });
    if (!ok) return false;
  }
  std::vector<std::vector<uint8_t>> stores(L, std::vector<uint8_t>((size_t)len * npad + 512, 0));
  {
    const uint32_t chunk = 1500, nchunks = (uint32_t)((N + chunk - 1) / chunk);
    std::vector<GatherTab> tabs(L);
    std::vector<std::vector<uint32_t>> slots(L);
    uint64_t run_total = 0;
    uint32_t groups = 0;
    for (uint32_t l = 0; l < L; ++l) {
      for (uint32_t s = 0; s + 1 < bstart[l].size(); ++s)
        for (uint32_t p = bstart[l][s]; p < bstart[l][s + 1]; p += kGatherPart) slots[l].push_back(p);
      slots[l].push_back((uint32_t)N);
      const uint64_t nv = slots[l].size() - 1;
      tabs[l].ids = ids[l].data();
      tabs[l].bstart = slots[l].data();
      tabs[l].out = stores[l].data();
      tabs[l].nslots = (uint32_t)nv;
      tabs[l].ngroups = (uint32_t)((nv + kGatherSlots - 1) / kGatherSlots);
      tabs[l].groups_before = groups;
      tabs[l].run_off = run_total;
      groups += tabs[l].ngroups;
      run_total += (uint64_t)(nchunks + 1) * nv;
    }
    std::vector<uint32_t> runs(run_total + 16);
    bool ok = emu_launch(groups, kGatherSlots, [&]() { gather_runs_kernel(tabs.data(), L, nchunks, chunk, runs.data()); });
    ok = ok && emu_launch(groups * nchunks, kGatherThreads, [&]() { gather_blocked_kernel<1>(tabs.data(), L, groups, runs.data(), rec.data(), RS, npad, (int)len); });
    if (!ok) return false;
  }
  // ---- queries: hash, probe
  std::vector<uint64_t> qkeys((size_t)L * Q);
  std::vector<uint8_t> qvalid((size_t)L * Q);
  {
    const int GP = (int)ctx->qh_group;
    const uint64_t nthreads = (uint64_t)Q * L * GP;
    if (!emu_launch((unsigned)((nthreads + 127) / 128), 128, [&]() {
          hash_queries_kernel<1>(qp.data(), Q, dim, ctx->d_a64t.as<double>(), ctx->d_b64.as<double>(), W, (int)K, GP, (int)L, qkeys.data(), qvalid.data());
        }))
      return false;
  }
  std::vector<uint2> qrange((size_t)L * Q);
  std::vector<uint32_t> qrank((size_t)L * Q);
  for (uint32_t l = 0; l < L; ++l)
    if (!emu_launch((Q + 127) / 128, 128, [&]() {
          probe_kernel<1>(qkeys.data() + (size_t)l * Q, qvalid.data() + (size_t)l * Q, Q, ctx->h_rkeys[l].data(), ctx->rank_nr[l], bstart[l].data(),
                          qrange.data() + (size_t)l * Q, qrank.data() + (size_t)l * Q);
        }))
      return false;
  // ---- the filter's work list: one item per probed bucket (its queries), one unit per bucket
  MmaGeometry g;
  if (mma_geometry(ctx, &g) != HS_OK || mma_upload_tables(ctx) != HS_OK) return false;
  std::vector<uint32_t> qlist;
  std::vector<MmaItem> items;
  std::vector<MmaUnit> units;
  uint64_t ncand = 0;
  for (uint32_t l = 0; l < L; ++l) {
    std::map<uint32_t, std::vector<uint32_t>> by_slot;
    for (uint32_t q = 0; q < Q; ++q)
      if (qrank[(size_t)l * Q + q] != 0xffffffffu && qrange[(size_t)l * Q + q].y > qrange[(size_t)l * Q + q].x) by_slot[qrank[(size_t)l * Q + q]].push_back(q);
    for (auto &kv : by_slot) {
      MmaItem it;
      it.table = l;
      it.q_begin = (uint32_t)qlist.size();
      qlist.insert(qlist.end(), kv.second.begin(), kv.second.end());
      it.q_end = (uint32_t)qlist.size();
      it.pad = 0;
      if (it.q_end - it.q_begin > (uint32_t)g.qmax) return false;
      items.push_back(it);
      MmaUnit un;
      un.item = (uint32_t)items.size() - 1;
      un.m_begin = bstart[l][kv.first];
      un.m_end = bstart[l][kv.first + 1];
      un.pad = 0;
      units.push_back(un);
      ncand += (uint64_t)(un.m_end - un.m_begin) * (it.q_end - it.q_begin);
    }
  }
  std::vector<__half> qb16((size_t)Q * g.kp);
  if (!emu_launch((Q + 127) / 128, 128, [&]() { build_qb_points_kernel(qp.data(), Q, dim, g.kp, g.beta, qb16.data()); })) return false;
  const double r2 = R * R * (1.0 + 1e-12) + 1e-30;
  float thr = (float)r2;
  if ((double)thr < r2) thr = nextafterf(thr, INFINITY);
  std::vector<Survivor> surv(ncand + 16);
  unsigned long long surv_count = 0;
  {
    const uint8_t *store_ptrs[HS_MAX_L] = {nullptr};
    const uint32_t *id_ptrs[HS_MAX_L] = {nullptr};
    for (uint32_t l = 0; l < L; ++l) {
      store_ptrs[l] = stores[l].data();
      id_ptrs[l] = ids[l].data();
    }
    uint32_t unit_counter = 0;
    uint4 tab16v[HS_AA];
    memcpy(tab16v, ctx->d_tab16.p, sizeof tab16v);
    MmaArgs ma;
    memset(&ma, 0, sizeof ma);
    ma.items = items.data();
    ma.units = units.data();
    ma.nunits = (uint32_t)units.size();
    ma.unit_counter = &unit_counter;
    ma.qlist = qlist.data();
    ma.qb16 = qb16.data();
    ma.stores = store_ptrs;
    ma.sorted_ids = id_ptrs;
    ma.npad = npad;
    ma.len = (int)len; ma.kp = g.kp; ma.nstages = g.nstages; ma.qmax = g.qmax; ma.cring = g.cring;
    ma.thr = thr;
    ma.beta = (float)(g.beta * 1.0001);
    ma.tab16 = tab16v;
    ma.nx32 = reinterpret_cast<const float *>(ctx->d_tab16.as<char>() + sizeof(__half) * HS_AA * HS_CDIM);
    ma.surv = surv.data();
    ma.surv_cap = surv.size();
    ma.surv_count = &surv_count;
    if (!emu_launch(4, kMmaThreads, [&]() { filter_mma_kernel<10>(ma); }) || surv_count > surv.size()) return false;
  }
  // ---- K4: exact stage (rank-path first-table rule), then the hit list's order
  std::vector<hs_hit> hits(surv_count + 16);
  unsigned long long hit_count = 0, edges = 0;
  {
    std::vector<const uint32_t *> id_ptrs(L);
    for (uint32_t l = 0; l < L; ++l) id_ptrs[l] = ids[l].data();
    std::vector<uint8_t> qrow(Q, 1);
    ExactArgs ea;
    memset(&ea, 0, sizeof ea);
    ea.surv = surv.data();
    ea.nsurv = surv_count;
    ea.mode = kModeSearch;
    ea.metric = HS_METRIC_EUCLID_FP64;
    ea.predicate = HS_PRED_D2_LE_R2;
    ea.len = (int)len; ea.dim = dim; ea.key_words = 1; ea.L = (int)L;
    ea.R = R;
    ea.sorted_ids = id_ptrs.data();
    ea.codes = codes.data();
    ea.rec = rec.data();
    ea.rec_stride = RS;
    ea.rec_rank_off = ctx->rec_rank_off;
    ea.qrank = qrank.data();
    ea.N = N;
    ea.table64 = ctx->table64;
    std::vector<int32_t> metric(400, 0);
    ea.metric_tab = metric.data();
    ea.q64 = qp.data();
    ea.qcodes = qcodes.data();     // (what detect_query_codes_kernel recovers from dense queries that embed residue strings)
    ea.qrow = qrow.data();
    ea.Q = Q;
    ea.qlist_mma = qlist.data();
    ea.hits = hits.data();
    ea.hit_cap = hits.size();
    ea.hit_count = &hit_count;
    ea.edge_count = &edges;
    if (!emu_launch(3, kExactThreadsRep, [&]() { exact_kernel<1, 8>(ea); }) || hit_count > hits.size()) return false;
  }
  const uint64_t n = hit_count;
  std::vector<hs_hit> sorted(n + 8);
  if (n) {
    SegFields f;
    const int tbits = bits_for_((uint64_t)L + 1), ibits = bits_for_(std::max<uint64_t>(N, 2));
    f.qbits = bits_for_(std::max<uint64_t>(Q, 1));
    f.tshift = ibits;
    f.qshift = ibits + tbits;
    const int kbits = f.qshift + f.qbits, pb = std::min(kbits, kSegMaxBinBits);
    f.shift = kbits - pb;
    f.rb = f.shift;
    f.nbins = (uint32_t)(((((uint64_t)Q << f.qshift) - 1ull) >> f.shift) + 1ull);
    f.nblk = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((n + 8191) / 8192, 4));
    f.chunk = (n + f.nblk - 1) / f.nblk;
    std::vector<uint32_t> tab((size_t)f.nbins * f.nblk), pkey(n);
    std::vector<double> pdist(n);
    unsigned int ctl[4] = {0, 0, 0, 0};
    bool ok = emu_launch(f.nblk, kSegThreads, [&]() { seg_hist_kernel(hits.data(), n, f, tab.data(), ctl); });
    ok = ok && scan_u32(tab);
    ok = ok && emu_launch(f.nblk, kSegThreads, [&]() { seg_scatter_kernel(hits.data(), n, f, tab.data(), pkey.data(), pdist.data()); });
    SegOut o;
    o.hits = sorted.data();
    o.idt = nullptr;
    o.dist2 = nullptr;
    o.id_base = 0;
    o.id_bits = 0;
    ok = ok && emu_launch(3, 1024, [&]() { seg_sort_kernel<1024>(pkey.data(), pdist.data(), n, f, tab.data(), o, kSegBufMax, kSegBufMax, 0u, ctl + 2, ctl); });
    if (!ok || ctl[0] || ctl[1]) return false;
  }
  printf("  (%llu candidates, %llu survivors, %llu hits; the oracle: %llu hits; %llu projections re-evaluated in FP64)\n",
         (unsigned long long)ncand, surv_count, (unsigned long long)n, (unsigned long long)nw, counters[0]);
  if (n != nw || nw == 0) return false;
  for (uint64_t i = 0; i < n; ++i)
    if (sorted[i].query != want[i].query || sorted[i].table_first != want[i].table_first || sorted[i].db_id != want[i].db_id ||
        memcmp(&sorted[i].dist2, &want[i].dist2, 8) != 0) {
      printf("  hit %llu differs from the oracle's\n", (unsigned long long)i);
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
  report("20000 fragments x 300 queries, len 10, K = L = 4, W 50, R 30", test_pipeline(20000, 300, 50.0, 30.0, 1));
  report("9000 fragments x 500 queries, W 50, R 34", test_pipeline(9000, 500, 50.0, 34.0, 2));
  return nbad ? 1 : 0;
}
