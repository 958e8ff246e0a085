// CPU emulation run of the index-build sort kernels of hsearch_b200/csrc/radix_sort.cu (B1: the
// unordered_map insert of motif_both_points.cpp:212-218 as a stable sort): the rank path (u16 bucket
// ranks, two 8-bit passes, bucket boundaries), the device-wide exclusive scan, and one pass of the
// general (key word, index) radix sort.  The kernel text is compiled unchanged over cuda_emu.h
// (radixsort_kernels.inc is cut out of the .cu file by tests/test_emu_ranksort.py); the host
// orchestration below mirrors build_table_index_ranks / exclusive_scan_u32 / radix_sort_pairs.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <numeric>
#include <random>
#include <vector>

#include "cuda_emu.h"

namespace hs {
constexpr int kCodeScale = 4;   // (csrc/common.cuh) the bucket-ordered code store keeps code * 4
constexpr int kMaxKeyWords = 4;
struct KeyPtrs {
  uint64_t *w[kMaxKeyWords];
};
#include "radixsort_kernels.inc"
#include "probe_kernels.inc"
#include "gather_kernels.inc"
}  // namespace hs

using namespace hs;

// exclusive_scan_u32 with the library's three kernels
static bool scan_u32(std::vector<uint32_t> &v, uint32_t *total) {
  const uint64_t n = v.size();
  if (n == 0) return true;
  const uint32_t nblocks = (uint32_t)((n + kScanTile - 1) / kScanTile);
  std::vector<uint32_t> sums(nblocks + 1, 0);
  bool ok = emu_launch(nblocks, kScanThreads, [&]() { scan_reduce_kernel(v.data(), n, sums.data()); });
  ok = ok && emu_launch(1, kScanThreads, [&]() { scan_sums_kernel(sums.data(), nblocks, total); });
  ok = ok && emu_launch(nblocks, kScanThreads, [&]() { scan_downsweep_kernel(v.data(), v.data(), n, sums.data()); });
  return ok;
}

static bool test_scan(uint64_t n, unsigned seed) {
  std::mt19937 rng(seed);
  std::vector<uint32_t> v(n), want(n);
  for (auto &x : v) x = rng() % 1000;
  uint32_t run = 0;
  for (uint64_t i = 0; i < n; ++i) {
    want[i] = run;
    run += v[i];
  }
  uint32_t total = 0;
  if (!scan_u32(v, &total)) return false;
  return v == want && total == run;
}

// build_table_index_ranks: ids in bucket order, sorted ranks, slot boundaries
static bool test_rank_index(uint64_t n, uint32_t nr, unsigned seed) {
  std::mt19937 rng(seed);
  std::vector<uint16_t> ranks(n + 16, 0);
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t r = rng() % nr;
    if (rng() % 4 == 0) r = (rng() % 3) * (nr / 3) % nr;   // a few heavy buckets
    ranks[i] = (uint16_t)r;
  }
  const uint32_t ntiles = (uint32_t)((n + kRkTile - 1) / kRkTile);
  std::vector<uint32_t> tile_hist((uint64_t)ntiles * 256 + 1, 0), vtmp(n), ids(n);
  std::vector<uint16_t> ktmp(n + 16, 0), ksorted(n + 16, 0);
  int hibits = 0;
  while (((uint64_t)256 << hibits) < nr) ++hibits;
  const int npass = nr > 256 ? 2 : 1;
  bool ok = true;
  for (int pi = 0; pi < npass; ++pi) {
    const int shift = 8 * pi;
    const uint32_t mask = pi == 0 ? 0xffu : ((1u << hibits) - 1u);
    const uint16_t *kin = pi == 0 ? ranks.data() : ktmp.data();
    ok = ok && emu_launch(ntiles, kRkThreads, [&]() { rank_upsweep_kernel(kin, n, shift, mask, tile_hist.data(), ntiles); });
    tile_hist.resize((uint64_t)ntiles * 256);
    ok = ok && scan_u32(tile_hist, nullptr);
    tile_hist.resize((uint64_t)ntiles * 256 + 1);
    if (npass == 1)
      ok = ok && emu_launch(ntiles, kRkThreads, [&]() {
        rank_downsweep_kernel<true, true>(kin, nullptr, ksorted.data(), ids.data(), n, shift, mask, tile_hist.data(), ntiles);
      });
    else if (pi == 0)
      ok = ok && emu_launch(ntiles, kRkThreads, [&]() {
        rank_downsweep_kernel<true, false>(kin, nullptr, ktmp.data(), vtmp.data(), n, shift, mask, tile_hist.data(), ntiles);
      });
    else
      ok = ok && emu_launch(ntiles, kRkThreads, [&]() {
        rank_downsweep_kernel<false, true>(kin, vtmp.data(), ksorted.data(), ids.data(), n, shift, mask, tile_hist.data(), ntiles);
      });
  }
  std::vector<uint32_t> bstart(nr + 1, 0xdeadbeefu);
  unsigned int nb = 0;
  ok = ok && emu_launch((unsigned)(((n + 7) / 8 + 255) / 256), 256, [&]() { rank_bounds_kernel(ksorted.data(), n, nr, bstart.data(), &nb); });
  if (!ok) return false;
  // expectation: stable sort of the ids by rank
  std::vector<uint32_t> want(n);
  std::iota(want.begin(), want.end(), 0u);
  std::stable_sort(want.begin(), want.end(), [&](uint32_t a, uint32_t b) { return ranks[a] < ranks[b]; });
  unsigned int distinct = 0;
  for (uint64_t i = 0; i < n; ++i) {
    if (ids[i] != want[i] || ksorted[i] != ranks[want[i]]) {
      printf("  rank index: mismatch at %llu\n", (unsigned long long)i);
      return false;
    }
    if (i == 0 || ksorted[i] != ksorted[i - 1]) ++distinct;
  }
  for (uint32_t r = 0; r <= nr; ++r) {
    const uint32_t lb = (uint32_t)(std::lower_bound(ksorted.begin(), ksorted.begin() + n, (uint16_t)std::min<uint32_t>(r, 65535)) - ksorted.begin());
    const uint32_t expect = r == nr ? (uint32_t)n : lb;
    if (bstart[r] != expect) {
      printf("  rank index: slot %u starts at %u, expected %u\n", r, bstart[r], expect);
      return false;
    }
  }
  if (nb != distinct) {
    printf("  rank index: %u non-empty slots, expected %u\n", nb, distinct);
    return false;
  }
  return true;
}

// one pass of the general sort: stable partition of (key word, index) by an 8-bit digit
static bool test_radix_pass(uint64_t n, int shift, bool with_vals, unsigned seed) {
  std::mt19937_64 rng(seed);
  std::vector<uint64_t> kin(n), kout(n, 0);
  std::vector<uint32_t> vin(n), vout(n, 0);
  for (uint64_t i = 0; i < n; ++i) {
    kin[i] = rng();
    vin[i] = (uint32_t)(rng() % 1000000);
  }
  const uint32_t ntiles = (uint32_t)((n + kSortTile - 1) / kSortTile);
  std::vector<uint32_t> tile_hist((uint64_t)ntiles * 256, 0);
  bool ok = emu_launch(ntiles, kSortThreads, [&]() { radix_upsweep_kernel(kin.data(), n, shift, 0xffu, tile_hist.data(), ntiles); });
  ok = ok && scan_u32(tile_hist, nullptr);
  KeyPtrs in{}, out{};
  in.w[0] = kin.data();
  out.w[0] = kout.data();
  ok = ok && emu_launch(ntiles, kSortThreads, [&]() {
    radix_downsweep_kernel<1>(in, with_vals ? vin.data() : nullptr, out, vout.data(), n, 0, shift, 0xffu, tile_hist.data(), ntiles);
  });
  if (!ok) return false;
  std::vector<uint32_t> order(n);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return ((kin[a] >> shift) & 0xff) < ((kin[b] >> shift) & 0xff); });
  for (uint64_t i = 0; i < n; ++i)
    if (kout[i] != kin[order[i]] || vout[i] != (with_vals ? vin[order[i]] : order[i])) {
      printf("  radix pass: mismatch at %llu\n", (unsigned long long)i);
      return false;
    }
  return true;
}

// probe: a query's key -> the slot holding it and the slot's member range (HashTable::find,
// motif_both_points.cpp:228-231), on ascending multi-word keys and on the hashed-key index
template <int KW>
static bool test_probe(uint64_t nb, uint32_t Q, unsigned seed) {
  std::mt19937_64 rng(seed);
  // nb distinct keys in ascending order (word KW-1 most significant), bucket sizes 1..5
  std::vector<std::vector<uint64_t>> ks(nb, std::vector<uint64_t>(KW));
  for (auto &k : ks)
    for (int w = 0; w < KW; ++w) k[w] = rng() % 7;   // few values per word: ties in the high words are common
  auto less = [](const std::vector<uint64_t> &x, const std::vector<uint64_t> &y) {
    for (int w = KW - 1; w >= 0; --w)
      if (x[w] != y[w]) return x[w] < y[w];
    return false;
  };
  std::sort(ks.begin(), ks.end(), less);
  ks.erase(std::unique(ks.begin(), ks.end()), ks.end());
  nb = ks.size();
  std::vector<uint64_t> ukeys((size_t)KW * nb);
  std::vector<uint32_t> bstart(nb + 1, 0);
  for (uint64_t i = 0; i < nb; ++i) {
    for (int w = 0; w < KW; ++w) ukeys[(uint64_t)w * nb + i] = ks[i][w];
    bstart[i + 1] = bstart[i] + 1 + (uint32_t)(rng() % 5);
  }
  std::vector<uint64_t> qkeys((size_t)Q * KW);
  std::vector<uint8_t> qvalid(Q, 1);
  for (uint32_t q = 0; q < Q; ++q) {
    if (q % 3 == 0) {
      const auto &k = ks[rng() % nb];
      for (int w = 0; w < KW; ++w) qkeys[(size_t)q * KW + w] = k[w];
    } else {
      for (int w = 0; w < KW; ++w) qkeys[(size_t)q * KW + w] = rng() % 8;
    }
    if (q % 17 == 0) qvalid[q] = 0;
  }
  std::vector<uint2> qrange(Q, uint2{7, 7});
  std::vector<uint32_t> qrank(Q, 7);
  if (!emu_launch((Q + 127) / 128, 128, [&]() {
        probe_kernel<KW>(qkeys.data(), qvalid.data(), Q, ukeys.data(), nb, bstart.data(), qrange.data(), qrank.data());
      }))
    return false;
  // the hashed-key index over the same buckets: slots in ascending order of the 64-bit key hash, a slot's key
  // is that of its first member
  std::vector<uint64_t> order(nb);
  std::iota(order.begin(), order.end(), 0ull);
  std::vector<uint64_t> hs_(nb);
  for (uint64_t i = 0; i < nb; ++i) hs_[i] = key_hash<KW>(ks[i].data());
  std::sort(order.begin(), order.end(), [&](uint64_t x, uint64_t y) { return hs_[x] < hs_[y]; });
  const uint64_t N = bstart[nb];
  std::vector<uint64_t> uhash(nb), dbkeys((size_t)KW * N, ~0ull);
  std::vector<uint32_t> hstart(nb + 1, 0), ids(N);
  std::iota(ids.begin(), ids.end(), 0u);
  std::shuffle(ids.begin(), ids.end(), rng);
  for (uint64_t s = 0; s < nb; ++s) {
    const uint64_t i = order[s];
    uhash[s] = hs_[i];
    hstart[s + 1] = hstart[s] + (bstart[i + 1] - bstart[i]);
    for (uint32_t m = hstart[s]; m < hstart[s + 1]; ++m)
      for (int w = 0; w < KW; ++w) dbkeys[(uint64_t)w * N + ids[m]] = ks[i][w];
  }
  std::vector<uint2> hrange(Q, uint2{7, 7});
  std::vector<uint32_t> hrank(Q, 7);
  if (!emu_launch((Q + 127) / 128, 128, [&]() {
        probe_hashed_kernel<KW>(qkeys.data(), qvalid.data(), Q, uhash.data(), dbkeys.data(), N, ids.data(), nb, hstart.data(),
                                hrange.data(), hrank.data());
      }))
    return false;
  uint32_t found = 0;
  for (uint32_t q = 0; q < Q; ++q) {
    std::vector<uint64_t> k(qkeys.begin() + (size_t)q * KW, qkeys.begin() + (size_t)(q + 1) * KW);
    auto it = std::lower_bound(ks.begin(), ks.end(), k, less);
    const bool hit = qvalid[q] && it != ks.end() && *it == k;
    const uint64_t i = hit ? (uint64_t)(it - ks.begin()) : 0;
    const uint32_t want_slot = hit ? (uint32_t)i : 0xffffffffu;
    const uint2 want_r = hit ? uint2{bstart[i], bstart[i + 1]} : uint2{0, 0};
    if (qrank[q] != want_slot || qrange[q].x != want_r.x || qrange[q].y != want_r.y) {
      printf("  probe: query %u\n", q);
      return false;
    }
    const uint64_t s = hit ? (uint64_t)(std::find(order.begin(), order.end(), i) - order.begin()) : 0;
    const uint2 want_h = hit ? uint2{hstart[s], hstart[s + 1]} : uint2{0, 0};
    if (hrank[q] != (hit ? (uint32_t)s : 0xffffffffu) || hrange[q].x != want_h.x || hrange[q].y != want_h.y) {
      printf("  hashed probe: query %u\n", q);
      return false;
    }
    found += hit;
  }
  return found > 0 && found < Q;
}

// The bucket-ordered, position-major code stores of all tables by the L2-blocked gather (build_code_stores_blocked):
// the record array is walked block by block, every bucket part's run inside a block found by gather_runs_kernel.
// (Built with -DHS_GATHER_PART=64 so that buckets are cut into several parts at this size.)
static bool test_gather(uint64_t n, uint32_t L, uint32_t nr, int len, uint32_t chunk, unsigned seed) {
  std::mt19937 rng(seed);
  const uint32_t RS = (uint32_t)((len + 15) & ~15);
  const uint64_t npad = (n + 15) & ~15ull;
  std::vector<uint8_t> rec(n * RS + 64, 0);
  for (uint64_t i = 0; i < n; ++i)
    for (int p = 0; p < len; ++p) rec[i * RS + p] = (uint8_t)(rng() % 20);
  const uint32_t nchunks = (uint32_t)((n + chunk - 1) / chunk);
  std::vector<std::vector<uint32_t>> ids(L), slots(L);
  std::vector<std::vector<uint8_t>> out(L, std::vector<uint8_t>((size_t)len * npad + 256, 0xee));
  std::vector<GatherTab> tabs(L);
  uint64_t run_total = 0;
  uint32_t groups = 0;
  for (uint32_t l = 0; l < L; ++l) {
    std::vector<uint16_t> rank(n);
    for (auto &r : rank) r = (uint16_t)(rng() % 5 == 0 ? 3 : rng() % nr);   // one large bucket among small ones
    ids[l].resize(n);
    std::iota(ids[l].begin(), ids[l].end(), 0u);
    std::stable_sort(ids[l].begin(), ids[l].end(), [&](uint32_t a, uint32_t b) { return rank[a] < rank[b]; });
    std::vector<uint32_t> bstart(nr + 1, 0);
    for (uint64_t i = 0; i < n; ++i) bstart[rank[i] + 1]++;
    for (uint32_t r = 0; r < nr; ++r) bstart[r + 1] += bstart[r];
    for (uint32_t b = 0; b < nr; ++b)
      for (uint32_t p = bstart[b]; p < bstart[b + 1]; p += kGatherPart) slots[l].push_back(p);
    slots[l].push_back((uint32_t)n);
    const uint64_t nv = slots[l].size() - 1;
    tabs[l].ids = ids[l].data();
    tabs[l].bstart = slots[l].data();
    tabs[l].out = out[l].data();
    tabs[l].nslots = (uint32_t)nv;
    tabs[l].ngroups = (uint32_t)((nv + kGatherSlots - 1) / kGatherSlots);
    tabs[l].groups_before = groups;
    tabs[l].run_off = run_total;
    groups += tabs[l].ngroups;
    run_total += (uint64_t)(nchunks + 1) * nv;
  }
  std::vector<uint32_t> runs(run_total + 16, 0xdeadbeefu);
  bool ok = emu_launch(groups, kGatherSlots, [&]() { gather_runs_kernel(tabs.data(), L, nchunks, chunk, runs.data()); });
  ok = ok && emu_launch(groups * nchunks, kGatherThreads, [&]() {
    gather_blocked_kernel<1>(tabs.data(), L, groups, runs.data(), rec.data(), RS, npad, len);
  });
  if (!ok) return false;
  for (uint32_t l = 0; l < L; ++l)
    for (uint64_t j = 0; j < n; ++j)
      for (int p = 0; p < len; ++p)
        if (out[l][(uint64_t)p * npad + j] != (uint8_t)(rec[(uint64_t)ids[l][j] * RS + p] * kCodeScale)) {
          printf("  table %u position %llu residue %d differs\n", l, (unsigned long long)j, p);
          return false;
        }
  return true;
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  int bad = 0;
  auto report = [&](const char *what, bool ok) {
    printf("%s -> %s\n", what, ok ? "ok" : "FAILED");
    if (!ok) ++bad;
  };
  report("scan n=1", test_scan(1, 1));
  report("scan n=4096", test_scan(4096, 2));
  report("scan n=70001", test_scan(70001, 3));
  report("rank index n=20011 nr=14641 (two passes)", test_rank_index(20011, 14641, 4));
  report("rank index n=4096 nr=14641 (one whole tile)", test_rank_index(4096, 14641, 5));
  report("rank index n=9000 nr=200 (one pass)", test_rank_index(9000, 200, 6));
  report("rank index n=5 nr=65536", test_rank_index(5, 65536, 7));
  report("rank index n=12289 nr=257", test_rank_index(12289, 257, 8));
  report("radix pass n=9000 shift=8 implicit index", test_radix_pass(9000, 8, false, 9));
  report("radix pass n=4097 shift=56 values", test_radix_pass(4097, 56, true, 10));
  report("probe, one-word keys, 300 slots", test_probe<1>(300, 500, 11));
  report("probe, three-word keys, 200 slots", test_probe<3>(200, 400, 12));
  report("probe, one slot", test_probe<2>(1, 40, 13));
  report("blocked gather: 5000 fragments, 2 tables, record blocks of 300", test_gather(5000, 2, 40, 10, 300, 14));
  report("blocked gather: 3001 fragments, 4 tables, record blocks of 1000, len 16", test_gather(3001, 4, 700, 16, 1000, 15));
  return bad ? 1 : 0;
}
