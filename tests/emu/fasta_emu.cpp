// CPU emulation run of the device FASTA parser (hsearch_b200/csrc/fasta.cu: fasta_pass_kernel<count / emit>,
// patch and boundary kernels) against a sequential restatement of ProteinDB::ReadFASTAFile
// (pcluster/src/pcluster/read_proteins.cpp:6-41): a name per header line (up to the first space), a sequence
// only when non-empty, letters of AA20 kept, every other letter replaced by AA20[rand() % 20] with one rand()
// per replaced letter in file order, everything else dropped.  The host orchestration below mirrors
// parse_fasta_gpu_impl.  fasta_kernels.inc is cut out of fasta.cu by tests/test_emu_fasta.py.
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <random>
#include <string>
#include <vector>

#include "cuda_emu.h"

namespace hs {
#include "fasta_kernels.inc"
}  // namespace hs

using namespace hs;

static const char AA20[] = "ARNDCEQGHILKMFPSTWYV";   // pcluster/src/pcluster/util.hpp:97

struct Parsed {
  std::vector<std::string> names, seqs;
};

static Parsed reference_parse(const std::string &text) {
  Parsed P;
  std::string seq;
  size_t pos = 0;
  while (pos < text.size()) {                      // getline: the last line needs no newline
    size_t e = text.find('\n', pos);
    if (e == std::string::npos) e = text.size();
    const std::string line = text.substr(pos, e - pos);
    pos = e + 1;
    if (!line.empty() && line[0] == '>') {
      if (!seq.empty()) {
        P.seqs.push_back(seq);
        seq.clear();
      }
      const size_t sp = line.find(' ');
      P.names.push_back(sp == std::string::npos ? line.substr(1) : line.substr(1, sp - 1));
      continue;
    }
    for (char c : line) {
      if (c != '\0' && strchr(AA20, c)) seq.push_back(c);
      else if (isalpha((unsigned char)c)) seq.push_back(AA20[rand() % 20]);
    }
  }
  if (!seq.empty()) P.seqs.push_back(seq);
  return P;
}

static std::string make_text(unsigned seed, size_t target, bool long_lines) {
  std::mt19937 rng(seed);
  std::string t;
  const char *junk = "BJOUXZbjouxz*-0123456789 .\r\t";
  while (t.size() < target) {
    const int kind = rng() % 10;
    if (kind < 3) {   // header, with or without a description
      t += ">sp|P" + std::to_string(rng() % 100000);
      if (rng() % 2) t += " some description > here";
      t += "\n";
      if (rng() % 6 == 0) continue;                 // (the next line may be another header: empty sequence)
    }
    if (kind == 9 && rng() % 3 == 0) {
      t += "\n";                                    // empty line
      continue;
    }
    const size_t len = long_lines && rng() % 4 == 0 ? 3000 + rng() % 9000 : rng() % 90;
    for (size_t i = 0; i < len; ++i) {
      const unsigned r = rng() % 100;
      if (r < 80) t.push_back(AA20[rng() % 20]);
      else if (r < 90) t.push_back((char)tolower(AA20[rng() % 20]));
      else t.push_back(junk[rng() % 28]);
    }
    t += "\n";
  }
  if (seed % 2) t.pop_back();                        // no newline at the end of the file
  return t;
}

static bool test_fasta(unsigned seed, size_t target, bool long_lines, bool leading_sequence) {
  std::string text = make_text(seed, target, long_lines);
  if (leading_sequence) text = "MKV\nLLA\n" + text;   // sequence lines before the first header
  const uint64_t n = text.size();
  srand(1234 + seed);
  const Parsed want = reference_parse(text);
  const int after_ref = rand();
  // ---- device path (parse_fasta_gpu_impl)
  srand(1234 + seed);
  std::vector<unsigned char> buf(n + 64, '\n');
  memcpy(buf.data(), text.data(), n);
  const uint32_t nchunks = (uint32_t)((n + kFaChunk - 1) / kFaChunk);
  std::vector<uint32_t> cnt(3 * (size_t)nchunks + 8, 0);
  FaOut none{};
  if (!emu_launch(nchunks, kFaThreads, [&]() { fasta_pass_kernel<false>(buf.data(), n, cnt.data(), nchunks, none); })) return false;
  uint32_t tot[3];
  for (int q = 0; q < 3; ++q) {
    uint32_t run = 0;
    for (uint32_t c = 0; c < nchunks; ++c) {
      const uint32_t v = cnt[(size_t)q * nchunks + c];
      cnt[(size_t)q * nchunks + c] = run;
      run += v;
    }
    tot[q] = run;
  }
  const uint32_t n_res = tot[0], n_rnd = tot[1], n_hdr = tot[2];
  std::vector<char> letters(n_rnd);
  for (uint32_t i = 0; i < n_rnd; ++i) letters[i] = AA20[rand() % 20];
  const int after_dev = rand();
  std::vector<char> residues(n_res + 16, 0);
  std::vector<uint32_t> rand_pos(n_rnd + 1), name_len(n_hdr + 1), hdr_res(n_hdr + 1), flag(n_hdr + 2), scan(n_hdr + 2);
  std::vector<uint64_t> name_begin(n_hdr + 1), start(n_hdr + 3, 0);
  FaOut fo;
  fo.residues = residues.data();
  fo.rand_pos = rand_pos.data();
  fo.name_begin = name_begin.data();
  fo.name_len = name_len.data();
  fo.hdr_res = hdr_res.data();
  if (!emu_launch(nchunks, kFaThreads, [&]() { fasta_pass_kernel<true>(buf.data(), n, cnt.data(), nchunks, fo); })) return false;
  if (n_rnd && !emu_launch((n_rnd + 255) / 256, 256, [&]() { fasta_patch_kernel(residues.data(), rand_pos.data(), letters.data(), n_rnd); }))
    return false;
  const unsigned gb = (n_hdr + 1 + 255) / 256;
  if (!emu_launch(gb, 256, [&]() { fasta_bound_flags_kernel(hdr_res.data(), n_hdr, n_res, flag.data()); })) return false;
  uint32_t n_seq = 0;
  for (uint32_t j = 0; j <= n_hdr; ++j) {
    scan[j] = n_seq;
    n_seq += flag[j];
  }
  if (!emu_launch(gb, 256, [&]() { fasta_bound_scatter_kernel(hdr_res.data(), n_hdr, n_res, flag.data(), scan.data(), start.data()); }))
    return false;
  // ---- compare
  if (n_hdr != want.names.size() || n_seq != want.seqs.size() || after_dev != after_ref) {
    printf("  %u names (expected %zu), %u sequences (expected %zu), rand() state %s\n", n_hdr, want.names.size(), n_seq,
           want.seqs.size(), after_dev == after_ref ? "equal" : "DIFFERS");
    return false;
  }
  for (uint32_t h = 0; h < n_hdr; ++h)
    if (text.substr(name_begin[h], name_len[h]) != want.names[h]) {
      printf("  name %u: '%s', expected '%s'\n", h, text.substr(name_begin[h], name_len[h]).c_str(), want.names[h].c_str());
      return false;
    }
  if (n_seq && start[n_seq] != n_res) return false;
  for (uint32_t s = 0; s < n_seq; ++s)
    if (std::string(residues.data() + start[s], residues.data() + start[s + 1]) != want.seqs[s]) {
      printf("  sequence %u differs\n", s);
      return false;
    }
  return n_hdr > 0 && n_seq > 0 && n_rnd > 0;
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  int nbad = 0;
  auto report = [&](const char *what, bool ok) {
    printf("%s -> %s\n", what, ok ? "ok" : "FAILED");
    if (!ok) ++nbad;
  };
  report("20 KB, short lines", test_fasta(1, 20000, false, false));
  report("30 KB, short lines, text ends without a newline... or with one", test_fasta(2, 30000, false, false));
  report("60 KB, lines of up to 12 KB (line starts several chunks back)", test_fasta(3, 60000, true, false));
  report("45 KB, long lines, sequence lines before the first header", test_fasta(4, 45000, true, true));
  report("5 KB (two chunks)", test_fasta(5, 5000, false, true));
  report("300 bytes (one partial chunk)", test_fasta(7, 300, false, false));
  return nbad ? 1 : 0;
}
