// CPU emulation run of the segmented hit sort kernels (hsearch_b200/csrc/hitsort.cu) against
// std::sort: the kernel text between the markers is compiled unchanged over tests/emu/cuda_emu.h;
// the host orchestration below mirrors sort_hits_segmented.  Usage: segsort_emu <kernels.inc>
// is produced by the Makefile-free recipe in tests/test_emu_segsort.py.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <random>
#include <vector>

#include "../../include/hsearch_b200.h"
#include "cuda_emu.h"

namespace hs {
#include "segsort_kernels.inc"
}  // namespace hs

using namespace hs;

static int bits_for(uint64_t nvalues) {
  int b = 1;
  while (b < 64 && (nvalues - 1) >> b) ++b;
  return b;
}

struct Case {
  uint32_t Q, L;
  uint64_t N, id_base;
  uint64_t n;
  uint32_t bufcap;
  bool compact;
  unsigned skew;  // 0: uniform queries; k: most hits in k queries
  int radix = 0;       // 1: the per-bin sort is forced onto the radix passes
};

static bool run_case(const Case &c, unsigned seed) {
  std::mt19937_64 rng(seed);
  // unique (query, id) pairs
  std::vector<hs_hit> hits;
  {
    std::vector<uint64_t> seen;
    while (hits.size() < c.n) {
      hs_hit h;
      uint32_t q = (uint32_t)(rng() % c.Q);
      if (c.skew && (rng() % 10) != 0) q = (uint32_t)((rng() % c.skew) * 37 % c.Q);
      h.query = q;
      h.table_first = (uint32_t)(rng() % c.L);
      h.db_id = c.id_base + rng() % c.N;
      h.dist2 = (double)(rng() % 100000) / 7.0;
      seen.push_back(((uint64_t)h.query << 40) ^ h.db_id);
      hits.push_back(h);
    }
    // drop duplicates of (query, id)
    std::vector<size_t> idx(hits.size());
    for (size_t i = 0; i < idx.size(); ++i) idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&](size_t a, size_t b) {
      if (hits[a].query != hits[b].query) return hits[a].query < hits[b].query;
      return hits[a].db_id < hits[b].db_id;
    });
    std::vector<hs_hit> u;
    for (size_t i = 0; i < idx.size(); ++i)
      if (i == 0 || hits[idx[i]].query != hits[idx[i - 1]].query || hits[idx[i]].db_id != hits[idx[i - 1]].db_id)
        u.push_back(hits[idx[i]]);
    std::shuffle(u.begin(), u.end(), rng);
    hits.swap(u);
  }
  const uint64_t n = hits.size();
  SegFields f;
  const int tbits = bits_for((uint64_t)c.L + 1), ibits = bits_for(std::max<uint64_t>(c.id_base + c.N, 2));
  f.qbits = bits_for(std::max<uint64_t>(c.Q, 1));
  f.tshift = ibits;
  f.qshift = ibits + tbits;
  const int kbits = f.qshift + f.qbits;
  if (kbits > 64 || f.qbits > kSegMaxBinBits) {
    printf("  (not applicable: key layout)\n");
    return true;
  }
  const int pb = std::min(kbits, kSegMaxBinBits);
  f.shift = kbits - pb;
  f.rb = f.shift;
  if (f.rb > 32) {
    printf("  (not applicable: rb > 32)\n");
    return true;
  }
  f.nbins = (uint32_t)(((((uint64_t)c.Q << f.qshift) - 1ull) >> f.shift) + 1ull);
  const uint64_t want_blk = (n + 8191) / 8192;
  f.nblk = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(want_blk, 6));
  f.chunk = (n + f.nblk - 1) / f.nblk;
  const uint64_t ntab = (uint64_t)f.nbins * f.nblk;
  std::vector<uint32_t> tab(ntab + 1, 0), pkey(n);
  std::vector<double> pdist(n);
  unsigned int ctl[4] = {0, 0, 0, 0};
  bool launched = emu_launch(f.nblk, kSegThreads, [&]() { seg_hist_kernel(hits.data(), n, f, tab.data(), ctl); });
  {
    uint32_t run = 0;
    for (uint64_t i = 0; i < ntab; ++i) {
      const uint32_t v = tab[i];
      tab[i] = run;
      run += v;
    }
  }
  launched = launched && emu_launch(f.nblk, kSegThreads, [&]() { seg_scatter_kernel(hits.data(), n, f, tab.data(), pkey.data(), pdist.data()); });
  std::vector<hs_hit> out_hits(n);
  std::vector<uint32_t> idt(n);
  std::vector<double> cd(n);
  std::vector<uint64_t> offsets((size_t)c.Q + 1, ~0ull);
  SegOut o;
  o.hits = c.compact ? nullptr : out_hits.data();
  o.idt = idt.data();
  o.dist2 = cd.data();
  o.id_base = c.id_base;
  o.id_bits = bits_for(std::max<uint64_t>(c.N, 2));
  const unsigned grid = std::min<uint32_t>(f.nbins, 3);
  launched = launched && emu_launch(grid, 1024, [&]() {
    seg_sort_kernel<1024>(pkey.data(), pdist.data(), n, f, tab.data(), o, c.bufcap, kSegBufMax, c.radix ? 1u : 0u, ctl + 2, ctl);
  });
  if (c.compact) {
    const uint32_t nq = c.Q + 1;
    launched = launched && emu_launch((nq + 255) / 256, 256, [&]() { seg_query_offsets_kernel(tab.data(), n, f, 0, c.Q, 1000, offsets.data()); });
  }
  if (!launched) return false;
  // expectation
  std::vector<hs_hit> want = hits;
  std::sort(want.begin(), want.end(), [](const hs_hit &a, const hs_hit &b) {
    if (a.query != b.query) return a.query < b.query;
    if (a.table_first != b.table_first) return a.table_first < b.table_first;
    return a.db_id < b.db_id;
  });
  if (ctl[0]) {
    printf("  field overflow flagged (unexpected)\n");
    return false;
  }
  if (ctl[1]) {
    printf("  handed back (range overflow) -- n=%llu bufcap=%u\n", (unsigned long long)n, c.bufcap);
    return true;
  }
  for (uint64_t i = 0; i < n; ++i) {
    bool ok;
    if (c.compact) {
      ok = idt[i] == ((uint32_t)(want[i].db_id - c.id_base) | (want[i].table_first << o.id_bits)) && cd[i] == want[i].dist2;
    } else {
      ok = out_hits[i].query == want[i].query && out_hits[i].table_first == want[i].table_first &&
           out_hits[i].db_id == want[i].db_id && out_hits[i].dist2 == want[i].dist2;
    }
    if (!ok) {
      printf("  MISMATCH at %llu of %llu\n", (unsigned long long)i, (unsigned long long)n);
      return false;
    }
  }
  if (c.compact) {
    uint64_t p = 0;
    for (uint32_t q = 0; q <= c.Q; ++q) {
      while (p < n && want[p].query < q) ++p;
      if (offsets[q] != 1000 + p) {
        printf("  OFFSET mismatch at query %u: %llu want %llu\n", q, (unsigned long long)offsets[q], (unsigned long long)(1000 + p));
        return false;
      }
    }
  }
  return true;
}

int main(int argc, char **argv) {
  setvbuf(stdout, nullptr, _IONBF, 0);
  const int only = argc > 1 ? atoi(argv[1]) : -1;
#if HS_SEG_BIN_BITS >= 15
  const Case cases[] = {
      {300, 4, 60000, 0, 3000, 22528, false, 0},            // small bins, one ranking step
      {300, 4, 60000, 0, 3000, 22528, true, 0},
      {10000, 4, 100000000ull, 0, 40000, 22528, true, 3},   // the bench's key layout; three queries hold most hits
      {10000, 4, 100000000ull, 0, 40000, 22528, false, 3},
      {9000, 4, 200000, 0, 60000, 22528, false, 2},         // bins beyond the buffer: the range path
      {9000, 4, 200000, 0, 30000, 2000, true, 2},           // ... forced by a small buffer
      {300, 4, 60000, 5000000000ull, 3000, 48, false, 0},   // id_base beyond 32 bits
      {4, 4, 3000, 0, 6, 22528, false, 0},                  // rb = 2
      {1, 1, 5, 0, 5, 22528, true, 0},                      // every bit in the bin
      {2600, 4, 60000, 0, 5000, 0, false, 0},               // buffer 0: handed back
      {32768, 32, 1000, 125000000ull, 20000, 22528, true, 5},
      {10000, 4, 100000000ull, 0, 40000, 22528, true, 3, 1},
  };
#else
  // (a build with few bins, -DHS_SEG_BIN_BITS=6: the same kernels, the per-bin loop is short)
  const Case cases[] = {
      {50, 4, 60000, 0, 3000, 22528, false, 0},
      {50, 4, 60000, 0, 3000, 22528, true, 0},
      {64, 4, 100000000ull, 0, 40000, 22528, true, 3},      // three bins of ~12 k keys: three ranking steps, four passes
      {64, 4, 100000000ull, 0, 40000, 22528, false, 3},
      {60, 4, 200000, 0, 60000, 22528, false, 2},           // two bins beyond the buffer: the range path
      {60, 4, 200000, 0, 30000, 2000, true, 2},             // ... forced by a small buffer
      {64, 4, 1000, 125000000ull, 20000, 22528, true, 5},
      {33, 4, 60000, 0, 5000, 0, false, 0},                 // buffer 0: handed back
      {1, 1, 5, 0, 5, 22528, true, 0},                      // every bit in the bin
      {7, 2, 100, 0, 600, 9, true, 0},                      // tiny key, tiny buffer
      {64, 4, 100000000ull, 0, 40000, 22528, false, 3, 1},  // the same lists with the per-bin sort forced onto the radix passes
      {50, 4, 60000, 0, 3000, 22528, true, 0, 1},
      {60, 4, 200000, 0, 30000, 2000, true, 2, 1},
      {60, 4, 200000, 0, 60000, 22528, false, 2, 1},
      {64, 4, 3000, 8000000ull, 30000, 22528, true, 1},     // ids in a narrow band far from 0: keys pile up in a few buckets -> radix passes by themselves
  };
#endif
  int bad = 0, i = 0;
  for (const Case &c : cases) {
    if (only >= 0 && only != i) {
      ++i;
      continue;
    }
    const bool ok = run_case(c, 1234 + i);
    printf("case %d: Q=%u L=%u N=%llu id_base=%llu n=%llu buf=%u radix=%d %s -> %s\n", i, c.Q, c.L, (unsigned long long)c.N,
           (unsigned long long)c.id_base, (unsigned long long)c.n, c.bufcap, c.radix, c.compact ? "compact" : "plain", ok ? "ok" : "FAILED");
    if (!ok) ++bad;
    ++i;
  }
  return bad ? 1 : 0;
}
