// CPU emulation run of the front-end and union-find kernels against the oracle (oracle/hs_oracle.c,
// linked in: test infrastructure): window extraction and ProteinID (extract.cu; protein.hpp:28-39,
// kmer_search.cpp:64-83), six-frame translation (sequence.cu; orf.cc:39-74), the lock-free union-find
// (verify.cuh, cluster.cu; union_find.cpp:16-33).  The kernel text is compiled unchanged over
// cuda_emu.h (frontend_kernels.inc is cut out of the sources by tests/test_emu_frontend.py).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <math.h>

#include <random>
#include <string>
#include <vector>

#include "cuda_emu.h"

extern "C" {
uint64_t orc_extract_windows(const uint8_t *residues, const uint32_t *start_index, uint32_t nprot, uint32_t L, uint32_t stride,
                             uint8_t *out_codes, uint32_t *out_pos);
uint32_t orc_protein_id(const uint32_t *start_index, uint32_t nstart, uint32_t pos);
int orc_orf6(const char *dna, int n, char *out, int *kept);
void orc_union_find_labels(uint32_t n, const uint32_t *eu, const uint32_t *ev, uint64_t ne, uint32_t *label_out);
void orc_klsh_generate(uint32_t feat, uint32_t bits, double sigma, double *w, double *t, double *b);
uint64_t orc_klsh_hash(const double *p, uint32_t feat, uint32_t bits, const double *w, const double *t, const double *b);
void orc_kmer3_features(const char *seq, uint32_t n, double *feat512);
}

namespace hs {
#include "frontend_kernels.inc"
}  // namespace hs

using namespace hs;

static bool test_extract(uint32_t nprot, int len, uint32_t stride, unsigned seed) {
  std::mt19937 rng(seed);
  std::vector<uint32_t> start(nprot + 1, 0);
  for (uint32_t p = 0; p < nprot; ++p) {
    uint32_t plen = rng() % 60;                         // ragged: empty, shorter than the window, exactly the window, longer
    if (p % 7 == 0) plen = (uint32_t)len;
    if (p % 11 == 0) plen = (uint32_t)len - 1;
    start[p + 1] = start[p] + plen;
  }
  const uint32_t total = start[nprot];
  std::vector<uint8_t> res(total + 16);
  for (auto &c : res) c = (uint8_t)(rng() % 20);
  // host part of extract_windows_impl: first fragment of every protein
  std::vector<uint32_t> fstart(nprot + 1);
  uint64_t n = 0;
  for (uint32_t p = 0; p < nprot; ++p) {
    fstart[p] = (uint32_t)n;
    const uint32_t plen = start[p + 1] - start[p];
    if (plen >= (uint32_t)len) n += (uint64_t)(plen - len) / stride + 1;
  }
  fstart[nprot] = (uint32_t)n;
  std::vector<uint8_t> want_codes((n + 1) * len), got_codes((n + 1) * len, 0xff);
  std::vector<uint32_t> want_pos(n + 1), got_pos(n + 1, 0xffffffffu);
  const uint64_t nw = orc_extract_windows(res.data(), start.data(), nprot, (uint32_t)len, stride, want_codes.data(), want_pos.data());
  if (nw != n) {
    printf("  extract: %llu fragments, oracle %llu\n", (unsigned long long)n, (unsigned long long)nw);
    return false;
  }
  if (n) {
    if (!emu_launch((unsigned)((n + 255) / 256), 256, [&]() {
          extract_windows_kernel(res.data(), start.data(), fstart.data(), nprot, stride, len, n, got_codes.data(), got_pos.data());
        }))
      return false;
    if (memcmp(want_codes.data(), got_codes.data(), n * len) || memcmp(want_pos.data(), got_pos.data(), n * 4)) {
      printf("  extract: windows differ from the oracle\n");
      return false;
    }
  }
  // ProteinID of every position, the sentinel and beyond (searched over all nprot + 1 entries)
  std::vector<uint32_t> pos, got(0);
  for (uint32_t x = 0; x <= total + 3; ++x) pos.push_back(x);
  got.assign(pos.size(), 0xffffffffu);
  if (!emu_launch((unsigned)((pos.size() + 255) / 256), 256,
                  [&]() { protein_id_kernel(start.data(), nprot + 1, pos.data(), pos.size(), got.data()); }))
    return false;
  for (size_t i = 0; i < pos.size(); ++i)
    if (got[i] != orc_protein_id(start.data(), nprot + 1, pos[i])) {
      printf("  protein id of position %u: %u, oracle %u\n", pos[i], got[i], orc_protein_id(start.data(), nprot + 1, pos[i]));
      return false;
    }
  return true;
}

static bool test_orf6(uint32_t nseq, unsigned seed) {
  std::mt19937 rng(seed);
  std::string dna;
  std::vector<uint64_t> start(nseq + 1, 0);
  for (uint32_t s = 0; s < nseq; ++s) {
    const uint32_t n = s < 8 ? s : rng() % 200;           // lengths 0..7 too
    for (uint32_t i = 0; i < n; ++i) dna.push_back("ACGT"[rng() % 4]);
    start[s + 1] = dna.size();
  }
  std::vector<char> aa(2 * dna.size() + 6ull * nseq + 64, 0);
  std::vector<int32_t> aa_len(6ull * nseq, -1);
  std::vector<uint8_t> bad(nseq, 0);
  if (!emu_launch((unsigned)((6ull * nseq + 127) / 128), 128,
                  [&]() { orf6_kernel(dna.data(), start.data(), nseq, aa.data(), aa_len.data(), bad.data()); }))
    return false;
  for (uint32_t s = 0; s < nseq; ++s) {
    const uint64_t s0 = start[s];
    const int n = (int)(start[s + 1] - s0);
    std::vector<char> want((size_t)6 * (n / 3 + 2) + 8, 0);
    int kept[6];
    orc_orf6(dna.data() + s0, n, want.data(), kept);
    for (int f = 0; f < 6; ++f) {
      const char *w = want.data() + (size_t)f * (n / 3 + 2);
      const char *g = aa.data() + 2 * s0 + 6ull * s + (uint64_t)f * (uint64_t)(n / 3 + 1);
      const int m = aa_len[(size_t)s * 6 + f];
      if (bad[s] || m != (int)strlen(w) || memcmp(w, g, m) || (m >= 6) != (kept[f] != 0)) {
        printf("  orf6: sequence %u frame %d differs from the oracle\n", s, f);
        return false;
      }
    }
  }
  return true;
}

// E5 + KL1: 3-mer histogram over the reduced alphabet and the 16 random-Fourier bits (pcluster.cpp:19-32,
// lsh.cpp:8-15,40-49).  Under the emulation cos() is libm's, the function the reference calls, so every bit must
// equal the oracle's (on the GPU a bit inside the 1e-9 guard band is recomputed on the host).
static bool test_kmer3_klsh(uint32_t nprot, unsigned seed) {
  std::mt19937 rng(seed);
  const uint32_t bits = 16;
  std::vector<double> w((size_t)bits * kFeat), t(bits), b(bits), wT((size_t)kFeat * bits);
  orc_klsh_generate(kFeat, bits, 0.2, w.data(), t.data(), b.data());
  for (uint32_t i = 0; i < bits; ++i)
    for (int j = 0; j < kFeat; ++j) wT[(size_t)j * bits + i] = w[(size_t)i * kFeat + j];
  const char *aa = "ARNDCQEGHILKMFPSTWYV";
  std::string res;
  std::vector<uint64_t> start(nprot + 1, 0);
  for (uint32_t p = 0; p < nprot; ++p) {
    const uint32_t n = p < 5 ? p : 20 + rng() % 400;      // proteins shorter than a 3-mer too
    for (uint32_t i = 0; i < n; ++i) res.push_back(aa[rng() % 20]);
    start[p + 1] = res.size();
  }
  std::vector<uint32_t> feat((size_t)nprot * kFeat, 7);
  std::vector<uint64_t> hash(nprot, 7);
  std::vector<uint8_t> flags(nprot, 7);
  if (!emu_launch(4, 128, [&]() {
        kmer3_klsh_kernel(res.data(), start.data(), nprot, wT.data(), t.data(), b.data(), bits, feat.data(), hash.data(), flags.data());
      }))
    return false;
  for (uint32_t p = 0; p < nprot; ++p) {
    const uint32_t n = (uint32_t)(start[p + 1] - start[p]);
    double want[512];
    orc_kmer3_features(res.data() + start[p], n, want);
    for (int i = 0; i < kFeat; ++i)
      if ((double)feat[(size_t)p * kFeat + i] != want[i]) {
        printf("  3-mer histogram of protein %u differs\n", p);
        return false;
      }
    if (n < 3) {
      if (flags[p] != 0) return false;
      continue;
    }
    if ((flags[p] & 1) == 0 || (flags[p] & 4)) return false;
    if (hash[p] != orc_klsh_hash(want, kFeat, bits, w.data(), t.data(), b.data())) {
      printf("  KLSH value of protein %u differs\n", p);
      return false;
    }
  }
  return true;
}

static bool test_union_find(uint32_t n, uint64_t ne, unsigned seed) {
  std::mt19937 rng(seed);
  std::vector<uint32_t> eu(ne), ev(ne);
  for (uint64_t e = 0; e < ne; ++e) {
    eu[e] = rng() % n;
    ev[e] = e % 5 == 0 ? (eu[e] + 1) % n : rng() % n;   // chains among random edges; self loops occur
    if (e % 97 == 0) ev[e] = eu[e];
  }
  std::vector<uint32_t> want(n), parent(n), label(n), parent2(n), label2(n);
  orc_union_find_labels(n, eu.data(), ev.data(), ne, want.data());
  unsigned int bad = 0;
  bool ok = emu_launch((n + 255) / 256, 256, [&]() { iota32_kernel(parent.data(), n); });
  ok = ok && emu_launch((unsigned)((ne + 255) / 256), 256, [&]() { uf_edges_kernel(eu.data(), ev.data(), ne, n, parent.data(), &bad); });
  ok = ok && emu_launch((n + 255) / 256, 256, [&]() { uf_flatten_kernel(parent.data(), n, label.data()); });
  if (!ok || bad || label != want) {
    printf("  union-find: labels differ from the oracle\n");
    return false;
  }
  // two forests over the two halves of the edge list, the second's labels united into the first (the multi-GPU merge)
  const uint64_t h = ne / 2;
  ok = emu_launch((n + 255) / 256, 256, [&]() { iota32_kernel(parent.data(), n); });
  ok = ok && emu_launch((n + 255) / 256, 256, [&]() { iota32_kernel(parent2.data(), n); });
  ok = ok && emu_launch((unsigned)((h + 255) / 256), 256, [&]() { uf_edges_kernel(eu.data(), ev.data(), h, n, parent.data(), &bad); });
  ok = ok && emu_launch((unsigned)((ne - h + 255) / 256), 256,
                        [&]() { uf_edges_kernel(eu.data() + h, ev.data() + h, ne - h, n, parent2.data(), &bad); });
  ok = ok && emu_launch((n + 255) / 256, 256, [&]() { uf_flatten_kernel(parent2.data(), n, label2.data()); });
  ok = ok && emu_launch((n + 255) / 256, 256, [&]() { uf_merge_labels_kernel(parent.data(), label2.data(), n); });
  ok = ok && emu_launch((n + 255) / 256, 256, [&]() { uf_flatten_kernel(parent.data(), n, label.data()); });
  if (!ok || bad || label != want) {
    printf("  union-find: merged labels differ from the oracle\n");
    return false;
  }
  // an endpoint out of range is flagged
  std::vector<uint32_t> bu{1, n}, bv{2, 3};
  ok = emu_launch(1, 256, [&]() { uf_edges_kernel(bu.data(), bv.data(), 2, n, parent.data(), &bad); });
  return ok && bad == 1;
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  int nbad = 0;
  auto report = [&](const char *what, bool ok) {
    printf("%s -> %s\n", what, ok ? "ok" : "FAILED");
    if (!ok) ++nbad;
  };
  report("windows + protein id: 300 proteins, len 10, stride 1", test_extract(300, 10, 1, 1));
  report("windows + protein id: 200 proteins, len 25, stride 3", test_extract(200, 25, 3, 2));
  report("windows + protein id: 1 protein, len 1, stride 1", test_extract(1, 1, 1, 3));
  report("six-frame translation: 120 sequences", test_orf6(120, 4));
  report("3-mer histograms + KLSH bits: 80 proteins", test_kmer3_klsh(80, 8));
  report("union-find: 5000 ids, 3000 edges", test_union_find(5000, 3000, 5));
  report("union-find: 400 ids, 4000 edges", test_union_find(400, 4000, 6));
  report("union-find: 1000 ids, no edge", test_union_find(1000, 0, 7));
  return nbad ? 1 : 0;
}
