// motif_both_points_noLSH -- drop-in for the reference's brute-force program
// (hclust/src/hclust/motif_both_points_noLSH.cpp:58-172): all centers x all
// k-mers, hits (sqrt(d2) <= R, the predicate of :46-47) to <output> in center
// order then k-mer order, through hs_bruteforce_points.
//
// <output>notlessthan.txt (the dump of every non-hit, :47-49, 6 GB per 10^8
// pairs) is created but left empty unless HS_NOLSH_NONHITS=1, in which case a
// second all-distances pass fills it (small inputs only).
#include "common.hpp"

using namespace hscli;

int main(int argc, const char **argv) {
  srand((unsigned)time(NULL));
  try {
    banner(argc, argv);
    std::string kmer_file, center_file, output_file;
    unsigned kmer_length = 25;
    double hash_R = 200;
    Options opt(strip_path(argv[0]), "cluster kmers to motifs");
    opt.add("db", 'd', "protein database file", true, kmer_file);
    opt.add("center", 'c', "centers from Pfam database", true, center_file);
    opt.add("len", 'l', "kmer length", true, kmer_length);
    opt.add("threshold", 'T', "kmer threshold", true, hash_R);
    opt.add("output", 'o', "output file name", true, output_file);
    std::vector<std::string> rest;
    opt.parse(argc, argv, rest);
    if (handled_help(argc, opt)) return EXIT_SUCCESS;

    const uint32_t dim = HS_CDIM * kmer_length;
    std::cout << "Read Kmers..." << std::endl;
    const PointFile kmers = read_points(kmer_file, dim);
    std::cout << "Read Centers..." << std::endl;
    const PointFile centers = read_points(center_file, dim);
    std::cout << "number of kmers " << kmers.size() << std::endl;
    std::cout << "number of centers " << centers.size() << std::endl;
    const clock_t start = clock();

    std::vector<uint8_t> codes;
    const int variant = points_to_codes(kmers, kmer_length, codes);
    if (variant < 0)
      throw CliError("the -db point file holds vectors that are not residue embeddings; this build stores DB "
                     "fragments as residue codes (write the file with protein2datapoints)");
    hs_params prm;
    memset(&prm, 0, sizeof prm);
    prm.len = kmer_length;
    prm.K = 1;
    prm.L = 1;
    prm.W = 1.0;
    prm.R = hash_R;
    prm.table_variant = (uint32_t)variant;
    prm.metric = HS_METRIC_EUCLID_FP64;
    prm.predicate = HS_PRED_SQRT_LE_R;
    prm.flags = HS_FLAG_SORT_HITS;
    std::ofstream fnot((output_file + "notlessthan.txt").c_str());
    std::ofstream fout(output_file.c_str());
    {
      // HS_DEVICES=0,1,...: the kmers are sharded over several GPUs (common.hpp, sharded_search)
      const ShardedResult res = sharded_search(
          devices_from_env(), prm, codes.data(), kmers.size(), 0,
          [&](hs_ctx_t *h, const uint8_t *shard, uint64_t n, uint64_t id_base) {
            check(hs_load_fragments(h, shard, n, id_base), "hs_load_fragments");
          },
          [&](hs_ctx_t *h, hs_hit *buf, uint64_t cap, uint64_t *n) {
            return hs_bruteforce_points(h, centers.data.data(), (uint32_t)centers.size(), buf, cap, n);
          });
      for (const hs_hit &h : res.hits)
        fout << centers.names[h.query] << " " << kmers.names[h.db_id] << " " << fmt_g(sqrt(h.dist2)) << "\n";
    }
    if (const char *e = getenv("HS_NOLSH_NONHITS")) {
      if (atoi(e)) {
        prm.R = 1e300;
        Ctx ctx(device_from_env(), prm);
        check(hs_load_fragments(ctx.h, codes.data(), kmers.size(), 0), "hs_load_fragments");
        const std::vector<hs_hit> all = collect_hits(
            [&](hs_hit *buf, uint64_t cap, uint64_t *n) {
              return hs_bruteforce_points(ctx.h, centers.data.data(), (uint32_t)centers.size(), buf, cap, n);
            },
            (uint64_t)centers.size() * kmers.size() + 1);
        for (const hs_hit &h : all)
          if (sqrt(h.dist2) > hash_R)
            fnot << centers.names[h.query] << " " << kmers.names[h.db_id] << " " << fmt_g(sqrt(h.dist2)) << "\n";
      }
    }
    fnot.close();
    fout.close();
    printf("Searching takes %lf seconds\n", (clock() - start) / (double)CLOCKS_PER_SEC);
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
