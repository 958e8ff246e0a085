// pcluster -- drop-in for the data-parallel part of the reference's protein clustering program
// (pcluster/src/pcluster/pcluster.cpp:84-181): `pcluster -d <protein fasta> -o <output> [-t n]`.
// ProteinDB::ReadFASTAFile (read_proteins.cpp:6-41) and PreClustering (pcluster.cpp:11-35: 512-bin
// histogram of reduced-alphabet 3-mers per protein, 16-bit KLSH value, proteins grouped by that
// value) run on the GPU (hs_parse_fasta_gpu, hs_kmer3_klsh); the progress lines on stderr are the
// reference's.  The per-group stage behind it -- CHashSearch, the RAPSearch2-derived aligner
// (hash_search.cpp, blast_stat.cpp) -- is outside the hot path (DESIGN.md "Out of
// scope"), and the reference's own program does not compile in its tree, so there
// is no binary to compare with.  The output file therefore holds the pre-groups, the input of that
// stage: one line per group in ascending KLSH value,
//     <klsh value>\t<number of proteins>\t<protein index> <protein index> ...
// (indices into the non-empty sequences of the file, in file order, as pro_seqs numbers them;
// the reference walks its unordered_map in an unspecified order).  The reference seeds rand()
// with time(NULL) for the letters it replaces (read_proteins.cpp:31); HS_SEED pins it.
#include <ctime>
#include <map>

#include "common.hpp"

using namespace hscli;

int main(int argc, const char **argv) {
  srand((unsigned)time(nullptr));
  if (const char *e = getenv("HS_SEED")) srand((unsigned)strtoul(e, nullptr, 10));
  try {
    bool help = false;
    for (int i = 1; i < argc; ++i)
      if (!strcmp(argv[i], "-help") || !strcmp(argv[i], "-about") || !strcmp(argv[i], "-?")) help = true;
    if (argc > 1 && !help) {
      fprintf(stderr, "[WELCOME TO PCLUSTER v%s]\n", "1.0");   // pcluster_version, util.hpp:66
      fprintf(stderr, "[%s", argv[0]);
      for (int i = 1; i < argc; ++i) fprintf(stderr, " %s", argv[i]);
      fprintf(stderr, "]\n");
    }
    std::string protein_file, output_file;
    int num_of_threads = 1;
    Options opt(strip_path(argv[0]), "cluster protein sequences");
    opt.add("database", 'd', "protein database file", true, protein_file);
    opt.add("output", 'o', "output file name", true, output_file);
    opt.add("thread", 't', "number of threads for mapping", false, num_of_threads);
    std::vector<std::string> rest;
    opt.parse(argc, argv, rest);
    if (handled_help(argc, opt)) return EXIT_SUCCESS;

    std::ifstream fin(protein_file.c_str(), std::ios::binary);
    if (!fin) throw CliError("cannot open input file " + protein_file);
    std::string text((std::istreambuf_iterator<char>(fin)), std::istreambuf_iterator<char>());
    fin.close();

    hs_params prm;
    memset(&prm, 0, sizeof prm);
    prm.len = 1;
    prm.K = prm.L = 1;
    prm.W = 1.0;
    Ctx ctx(device_from_env(), prm);

    // ---- ReadFASTAFile on the device (every byte of the file is at most one residue)
    std::vector<char> residues(text.size() + 1);
    size_t nlines = 1;
    for (char c : text) nlines += c == '\n';
    std::vector<uint64_t> start(nlines + 2), name_begin(nlines + 1);
    std::vector<uint32_t> name_len(nlines + 1);
    uint32_t nseq = 0, nnames = 0;
    uint64_t nres = 0;
    check(hs_parse_fasta_gpu(ctx.h, text.data(), text.size(), residues.data(), residues.size(), start.data(), start.size(),
                             name_begin.data(), name_len.data(), name_begin.size(), &nseq, &nnames, &nres),
          "hs_parse_fasta_gpu");
    fprintf(stderr, "[THE TOTAL NUMBER OF PROTEINS IN THE DATABASE IS %u]\n", nseq);

    // ---- PreClustering (pcluster.cpp:11-35)
    const clock_t t0 = clock();
    const uint32_t feature_size = 512, bit_num = 16;   // pow(8, HASHLEN), pcluster.cpp:13-14
    const double sigma = 0.2;
    std::vector<double> w((size_t)bit_num * feature_size), t(bit_num), b(bit_num);
    check(hs_klsh_generate(feature_size, bit_num, sigma, w.data(), t.data(), b.data()), "hs_klsh_generate");
    std::vector<uint64_t> hv(nseq);
    std::vector<uint8_t> valid(nseq);
    uint64_t fixed = 0;
    if (nseq)
      check(hs_kmer3_klsh(ctx.h, residues.data(), start.data(), nseq, w.data(), t.data(), b.data(), bit_num, nullptr,
                          hv.data(), valid.data(), &fixed),
            "hs_kmer3_klsh");
    std::map<uint64_t, std::vector<uint32_t>> groups;
    for (uint32_t i = 0; i < nseq; ++i)
      if (valid[i]) groups[hv[i]].push_back(i);   // (shorter than HASHLEN: skipped, pcluster.cpp:22-24)
    fprintf(stderr, "[NUMBER OF PRE-GROUPS %lu]\n", (unsigned long)groups.size());
    fprintf(stderr, "[Locality-Sensitive Hashing Pre-Clustering TAKES %lf SECONDS]\n", (clock() - t0) / (double)CLOCKS_PER_SEC);

    std::ofstream fout(output_file.c_str());
    if (!fout) throw CliError("cannot open output file " + output_file);
    uint32_t group_id = 0;
    for (const auto &g : groups) {
      fprintf(stderr, "[CLUSTERING GROUP %u of %lu]\n", group_id++, (unsigned long)groups.size());
      fprintf(stderr, "[THE NUMBER OF SEQUENCES IN THIS GROUP IS %lu]\n", (unsigned long)g.second.size());
      fout << g.first << "\t" << g.second.size() << "\t";
      for (size_t k = 0; k < g.second.size(); ++k) fout << (k ? " " : "") << g.second[k];
      fout << "\n";
    }
    fout.close();
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
