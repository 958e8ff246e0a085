// motif_both_points -- drop-in for the reference's HSEARCH motif search program
// (hclust/src/hclust/motif_both_points.cpp:252-396): same flags, same input and
// output files, same stdout lines; Search() (:195-250) runs on the GPU through
// the C ABI (hs_build_index + hs_search_points).
//
// Differences a user can see: K and L are still forced to 4 (:380-381) unless
// HS_HASH_K / HS_HASH_L are set; the DB point file must hold residue embeddings
// (what protein2datapoints writes) because the device stores fragments as
// 1-byte residue codes; HS_REF_SEED pins the projection seed.
#include "common.hpp"

using namespace hscli;

int main(int argc, const char **argv) {
  srand((unsigned)time(NULL));
  try {
    banner(argc, argv);
    std::string kmer_file, center_file, output_file, ground_truth;
    unsigned kmer_length = 25;
    double hash_W = 50, hash_R = 200;
    Options opt(strip_path(argv[0]), "cluster kmers to motifs");
    opt.add("db", 'd', "protein database file", true, kmer_file);
    opt.add("center", 'c', "centers from Pfam database", true, center_file);
    opt.add("len", 'l', "kmer length", true, kmer_length);
    opt.add("window", 'W', "bucket width", true, hash_W);
    opt.add("threshold", 'T', "kmer threshold", true, hash_R);
    opt.add("groundtruth", 'g', "groundtruth", true, ground_truth);
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

    const clock_t start_s = clock();
    const double p1 = 0.9, p2 = 0.2, roh = log(1 / p1) / log(1 / p2);
    unsigned hash_K = 4, hash_L = 4;
    if (const char *e = getenv("HS_HASH_K")) hash_K = (unsigned)atoi(e);
    if (const char *e = getenv("HS_HASH_L")) hash_L = (unsigned)atoi(e);
    printf("p1 = %lf p2 = %lf roh = %lf hash_K = %u hash_L = %u\n", p1, p2, roh, hash_K, hash_L);
    fflush(stdout);

    std::vector<uint8_t> codes;
    const int variant = points_to_codes(kmers, kmer_length, codes);
    if (variant < 0)
      throw CliError("the -db point file holds vectors that are not residue embeddings; this build stores DB "
                     "fragments as residue codes (write the file with protein2datapoints)");

    hs_params prm;
    memset(&prm, 0, sizeof prm);
    prm.len = kmer_length;
    prm.K = hash_K;
    prm.L = hash_L;
    prm.W = hash_W;
    prm.R = hash_R;
    prm.table_variant = (uint32_t)variant;
    prm.metric = HS_METRIC_EUCLID_FP64;
    prm.predicate = HS_PRED_D2_LE_R2;  // dis_square <= hash_R_square (:204,239)
    prm.flags = HS_FLAG_SORT_HITS;     // center order, table order, ascending kmer id (:224-245)
    // one LSH object per table, each seeded by its own random_device draw (:206-211)
    const uint64_t seed = projection_seed_base();
    std::vector<double> a((size_t)hash_L * hash_K * dim), b((size_t)hash_L * hash_K);
    for (unsigned l = 0; l < hash_L; ++l)
      check(hs_generate_projection(seed + l, dim, hash_K, hash_W, &a[(size_t)l * hash_K * dim], &b[(size_t)l * hash_K]),
            "hs_generate_projection");
    // HS_DEVICES=0,1,...: the kmers are sharded over several GPUs (common.hpp, sharded_search)
    const ShardedResult res = sharded_search(
        devices_from_env(), prm, codes.data(), kmers.size(), hash_L,
        [&](hs_ctx_t *h, const uint8_t *shard, uint64_t n, uint64_t id_base) {
          check(hs_set_projection(h, a.data(), b.data()), "hs_set_projection");
          check(hs_load_fragments(h, shard, n, id_base), "hs_load_fragments");
          check(hs_build_index(h), "hs_build_index");
        },
        [&](hs_ctx_t *h, hs_hit *buf, uint64_t cap, uint64_t *n) {
          return hs_search_points(h, centers.data.data(), (uint32_t)centers.size(), buf, cap, n);
        });
    for (unsigned l = 0; l < hash_L; ++l) std::cout << "table size " << res.table_sizes[l] << std::endl;
    const std::vector<hs_hit> &hits = res.hits;
    {
      std::ofstream fout(output_file.c_str());
      for (const hs_hit &h : hits)
        fout << centers.names[h.query] << " " << kmers.names[h.db_id] << " " << fmt_g(sqrt(h.dist2)) << "\n";
    }
    const clock_t end_s = clock();
    std::cout << "evaulate ..." << std::endl;
    const double recall = evaluate_recall(ground_truth, output_file, hash_R);
    printf("ACCURACY: %lf %lf\n", recall, (end_s - start_s) / (double)CLOCKS_PER_SEC);
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
