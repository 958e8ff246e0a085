// orf -- drop-in for the reference's six-frame translator (orf/orf_main.cc:8-20 over
// ORF::orf6, orf/orf.cc:39-74): `orf -q <fasta>` writes <fasta>_translatedAA.fasta with, for
// every DNA sequence, the translated frames of at least six residues (frames 0-2 forward, then
// 0-2 of the reverse complement, each up to its first stop codon) as
//     <name>_<j>        (j counts the frames kept for this sequence)
//     <amino acids>
// The translation runs on the GPU (hs_orf6).  The reference's own program does not compile in
// its tree (util/option.h, fasta_file.h and bio_util.h are not vendored), so there is
// no binary to compare with: the FASTA reader here takes the header line after '>' as the name
// and joins the sequence lines; the debug lines orf.cc prints to stdout (:40,47) are not produced.
#include "common.hpp"

using namespace hscli;

int main(int argc, const char **argv) {
  try {
    std::string query_file;
    for (int i = 1; i + 1 < argc; ++i)
      if (!strcmp(argv[i], "-q")) query_file = argv[i + 1];
    if (query_file.empty()) {
      fprintf(stderr, "usage: %s -q <DNA fasta file>\n", strip_path(argv[0]).c_str());
      return EXIT_FAILURE;
    }
    std::ifstream fin(query_file.c_str());
    if (!fin) throw CliError("cannot open " + query_file);
    std::vector<std::string> names;
    std::string dna, line;
    std::vector<uint64_t> start;
    while (std::getline(fin, line)) {
      if (!line.empty() && line.back() == '\r') line.pop_back();
      if (line.empty()) continue;
      if (line[0] == '>') {
        names.push_back(line.substr(1));
        start.push_back(dna.size());
      } else if (!names.empty()) {
        dna += line;
      }
    }
    start.push_back(dna.size());
    const uint32_t nseq = (uint32_t)names.size();
    std::ofstream fout((query_file + "_translatedAA.fasta").c_str());
    if (nseq) {
      hs_params prm;
      memset(&prm, 0, sizeof prm);
      prm.len = 1;
      prm.K = prm.L = 1;
      prm.W = 1.0;
      Ctx ctx(device_from_env(), prm);
      std::vector<char> aa(2 * dna.size() + 6 * (size_t)nseq + 1);
      std::vector<int32_t> aa_len(6 * (size_t)nseq);
      check(hs_orf6(ctx.h, dna.data(), start.data(), nseq, aa.data(), aa.size(), aa_len.data()), "hs_orf6");
      for (uint32_t s = 0; s < nseq; ++s) {
        const uint64_t ls = start[s + 1] - start[s];
        const uint64_t base = 2 * start[s] + 6 * (uint64_t)s;
        uint32_t j = 0;
        for (int f = 0; f < 6; ++f) {
          const int32_t n = aa_len[6 * (size_t)s + f];
          if (n < 6) continue;  // orf.cc:58,70
          fout << names[s] << "_" << j++ << "\n";
          fout.write(aa.data() + base + (uint64_t)f * (ls / 3 + 1), n);
          fout << "\n";
        }
      }
    }
    fout.close();
  } catch (const CliError &e) {
    fprintf(stderr, "%s\n", e.what());
    return EXIT_FAILURE;
  } catch (std::bad_alloc &) {
    fprintf(stderr, "ERROR: could not allocate memory\n");
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
