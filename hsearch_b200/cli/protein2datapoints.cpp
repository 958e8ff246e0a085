// protein2datapoints -- drop-in for the reference's extractor + embedder
// (hclust/src/hclust/protein2datapoints.cpp:33-152): FASTA with one sequence per
// line (ProteinDB, protein.hpp:41-71) -> data-point text file: for each of the
// first -n proteins, windows of -l residues at a random stride of 30..49
// (rand(), :57,70), k-mers already seen skipped (:52-56), one name line
// `name#i$j@KMER*cnt` and one line of 8*len coordinates at 6 significant digits.
//
// Host-only program (text formatting; no kernel is worth launching for it).
// HS_SEED pins srand(); proteins shorter than the window are skipped (the
// source underflows an unsigned there, :45).
#include <unordered_set>

#include "common.hpp"

using namespace hscli;

int main(int argc, const char **argv) {
  srand((unsigned)time(NULL));
  try {
    banner(argc, argv);
    std::string protein_file, output_file;
    unsigned kmer_length = 25, num_out = 0;
    Options opt(strip_path(argv[0]), "protein sequences to data points");
    opt.add("db", 'd', "protein database file", true, protein_file);
    opt.add("len", 'l', "kmer length", true, kmer_length);
    opt.add("nnn", 'n', "num of proteins out", true, num_out);
    opt.add("output", 'o', "output file name", true, output_file);
    std::vector<std::string> rest;
    opt.parse(argc, argv, rest);
    if (handled_help(argc, opt)) return EXIT_SUCCESS;

    const clock_t start = clock();
    if (const char *e = getenv("HS_SEED")) srand((unsigned)strtoul(e, nullptr, 10));
    const ProteinStore db = read_protein_db(protein_file);
    if (const char *e = getenv("HS_SEED")) srand((unsigned)strtoul(e, nullptr, 10));
    double table[HS_AA * HS_CDIM];
    hs_get_coordinates(HS_TABLE_FULL, table);  // coordinates[base[c]][p], util.hpp:21-42

    std::cout << "protein to data points... " << std::endl;
    std::ofstream fout(output_file.c_str());
    std::unordered_set<std::string> seen;
    uint32_t cnt = 0;
    for (uint32_t i = 0; i < db.nprot(); ++i) {
      std::cout << i << " " << db.nprot() << std::endl;
      if (i >= num_out) break;
      const uint32_t plen = db.start[i + 1] - db.start[i];
      if (plen < kmer_length) continue;
      for (uint32_t j = 0; j <= plen - kmer_length;) {
        const uint32_t pos = db.start[i] + j;
        const std::string kmer = db.letters.substr(pos, kmer_length);
        if (!seen.insert(kmer).second) {
          j += 30 + rand() % 20;
          continue;
        }
        // name#i$j@KMER*cnt (protein2datapoints.cpp:61-65) through the C ABI
        const std::string &header = i < db.names.size() ? db.names[i] : std::string();
        std::vector<char> line(header.size() + kmer.size() + 64);
        if (hs_fragment_name(header.c_str(), i, j, kmer.c_str(), kmer_length, cnt, line.data(), line.size()) != HS_OK)
          throw CliError(hs_last_error());
        fout << line.data() << "\n";
        for (uint32_t l = 0; l < kmer_length; ++l) {
          const double *row = table + db.codes[pos + l] * HS_CDIM;
          for (int p = 0; p < HS_CDIM; ++p) {
            if (l || p) fout << " ";
            fout << fmt_g(row[p]);
          }
        }
        fout << "\n";
        ++cnt;
        j += 30 + rand() % 20;
      }
    }
    fout.close();
    printf("It takes %lf seconds\n", (clock() - start) / (double)CLOCKS_PER_SEC);
  } catch (const OptionError &e) {
    fprintf(stderr, "%s\n", e.what());
    return EXIT_FAILURE;
  } catch (const CliError &e) {
    fprintf(stderr, "%s\n", e.what());
    return EXIT_FAILURE;
  } catch (std::bad_alloc &) {
    fprintf(stderr, "ERROR: could not allocate memory\n");
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
