// CPU emulation run of the multi-GPU hit merge (hsearch_b200/csrc/comm.cu: seg_offsets / seg_counts / seg_dest /
// scatter_merged kernels): every rank holds its hits in the reference's order (query, first table, ascending db id;
// motif_both_points.cpp:224-245) over its own id block, counts its (query, table) segments, and -- after the counts
// of all ranks were exchanged (an NCCL all-gather on the device; a copy here) -- writes every hit to its final
// position of the merged list in rank 0's memory.  The result must be the sorted union of all ranks' lists.  The
// GPU suite's 2-GPU test needs two devices; this runs the same kernels for 2, 3 and 8 ranks on every CPU run.
// merge_kernels.inc is cut out of comm.cu by tests/test_emu_merge.py.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <random>
#include <vector>

#include "../../include/hsearch_b200.h"
#include "cuda_emu.h"

namespace hs {
#include "merge_kernels.inc"
}  // namespace hs

using namespace hs;

static int bits_for_(uint64_t nvalues) {
  int b = 1;
  while (b < 64 && (nvalues - 1) >> b) ++b;
  return b;
}

static bool hit_less(const hs_hit &a, const hs_hit &b) {
  if (a.query != b.query) return a.query < b.query;
  if (a.table_first != b.table_first) return a.table_first < b.table_first;
  return a.db_id < b.db_id;
}

static bool test_merge(int G, uint32_t Q, uint32_t L, uint64_t per_rank, uint64_t nshard, bool overflow, unsigned seed) {
  std::mt19937_64 rng(seed);
  const int tbits = bits_for_((uint64_t)L + 1);
  const uint32_t S = Q << tbits;
  const int ibits = bits_for_(nshard * G);
  const int tshift = ibits;
  std::vector<std::vector<hs_hit>> hits(G);
  std::vector<std::vector<uint64_t>> keys(G);
  std::vector<hs_hit> all;
  for (int r = 0; r < G; ++r) {
    const uint64_t n = r == 1 ? 0 : per_rank / 2 + rng() % per_rank;     // one rank without any hit
    for (uint64_t i = 0; i < n; ++i) {
      hs_hit h;
      h.query = (uint32_t)(rng() % 7 == 0 ? rng() % 3 : rng() % Q);        // a few crowded queries
      h.table_first = (uint32_t)(rng() % L);
      h.db_id = (uint64_t)r * nshard + rng() % nshard;                      // ids ascend with the rank
      h.dist2 = (double)(rng() % 1000000) / 3.0;
      hits[r].push_back(h);
    }
    std::sort(hits[r].begin(), hits[r].end(), hit_less);
    hits[r].erase(std::unique(hits[r].begin(), hits[r].end(), [](const hs_hit &a, const hs_hit &b) { return !hit_less(a, b) && !hit_less(b, a); }),
                  hits[r].end());
    for (const hs_hit &h : hits[r]) {
      keys[r].push_back(((uint64_t)h.query << (tshift + tbits)) | ((uint64_t)h.table_first << tshift) | h.db_id);
      all.push_back(h);
    }
  }
  std::sort(all.begin(), all.end(), hit_less);
  const uint64_t total = all.size();
  const uint64_t cap = overflow ? total - 1 : total + 5;
  // per rank: segment offsets and counts
  std::vector<std::vector<uint64_t>> off(G, std::vector<uint64_t>((size_t)S + 1, ~0ull)), dst(G, std::vector<uint64_t>(S, ~0ull));
  std::vector<uint32_t> cnt_all((size_t)G * S, 0);
  bool ok = true;
  for (int r = 0; r < G; ++r) {
    const uint64_t n = hits[r].size();
    ok = ok && emu_launch((unsigned)((n + 256) / 256), 256, [&]() { seg_offsets_kernel(keys[r].data(), n, tshift, S, off[r].data()); });
    ok = ok && emu_launch((S + 255) / 256, 256, [&]() { seg_counts_kernel(off[r].data(), S, &cnt_all[(size_t)r * S]); });   // (+ all-gather)
  }
  std::vector<hs_hit> out(cap + 8);
  memset(out.data(), 0xff, sizeof(hs_hit) * out.size());
  for (int r = 0; r < G; ++r) {
    unsigned long long info[2] = {0, 0};
    const uint64_t n = hits[r].size();
    ok = ok && emu_launch(1, kDestThreads, [&]() { seg_dest_kernel(cnt_all.data(), S, G, r, cap, dst[r].data(), info); });
    if (info[0] != total || info[1] != (overflow ? 1ull : 0ull)) {
      printf("  rank %d: total %llu (expected %llu), overflow flag %llu\n", r, info[0], (unsigned long long)total, info[1]);
      return false;
    }
    if (n) ok = ok && emu_launch(3, kScatterWarps * 32, [&]() { scatter_merged_kernel(hits[r].data(), n, tbits, off[r].data(), dst[r].data(), info, out.data()); });
  }
  if (!ok) return false;
  if (overflow) {   // nothing may have been written
    for (size_t i = 0; i < sizeof(hs_hit) * out.size(); ++i)
      if (reinterpret_cast<unsigned char *>(out.data())[i] != 0xff) return false;
    return true;
  }
  for (uint64_t i = 0; i < total; ++i)
    if (memcmp(&out[i], &all[i], sizeof(hs_hit)) != 0) {
      printf("  merged list differs at %llu of %llu\n", (unsigned long long)i, (unsigned long long)total);
      return false;
    }
  for (size_t i = sizeof(hs_hit) * total; i < sizeof(hs_hit) * out.size(); ++i)
    if (reinterpret_cast<unsigned char *>(out.data())[i] != 0xff) return false;   // nothing past the end
  return total > 0;
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  int nbad = 0;
  auto report = [&](const char *what, bool ok) {
    printf("%s -> %s\n", what, ok ? "ok" : "FAILED");
    if (!ok) ++nbad;
  };
  report("2 ranks, 300 queries, 4 tables", test_merge(2, 300, 4, 20000, 1000000, false, 1));
  report("3 ranks, 50 queries, 8 tables (long segments: whole warps inside one segment)", test_merge(3, 50, 8, 30000, 500000, false, 2));
  report("8 ranks, 1200 queries, 4 tables (more than 1024 x 4 segments: several scan rounds)", test_merge(8, 1200, 4, 6000, 125000, false, 3));
  report("4 ranks, receive buffer one hit too small", test_merge(4, 100, 4, 5000, 100000, true, 4));
  return nbad ? 1 : 0;
}
