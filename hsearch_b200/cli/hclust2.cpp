// hclust2 / hclust3 -- drop-ins for the reference's k-mer clustering programs
// (hclust/src/hclust/hclust2.cpp:153-252, hclust3.cpp:154-266; the two differ only in
// when the embedding is computed and in one progress line): same flags, same input and
// output files, same stdout lines; Clustering() (hclust2.cpp:86-151) runs on the GPU
// through the C ABI (hs_build_index + hs_greedy_cluster).
//
// Differences a user can see: k-mers must consist of the 20 amino-acid letters (the
// reference replaces any other letter by a random residue, hclust2.cpp:53-55);
// HS_REF_SEED pins the projection seeds (one std::random_device draw per round,
// hclust2.cpp:104, lsh.hpp:15-16).
#include <algorithm>

#include "common.hpp"

using namespace hscli;

int main(int argc, const char **argv) {
  srand((unsigned)time(NULL));
  try {
    bool help = false;
    for (int i = 1; i < argc; ++i)
      if (!strcmp(argv[i], "-help") || !strcmp(argv[i], "-about") || !strcmp(argv[i], "-?")) help = true;
    if (argc > 1 && !help) {  // hclust2.cpp:169-177: this program greets as PMF
      fprintf(stdout, "[WELCOME TO PMF v%s]\n", kVersion);
      fprintf(stdout, "[%s", argv[0]);
      for (int i = 1; i < argc; ++i) fprintf(stdout, " %s", argv[i]);
      fprintf(stdout, "]\n");
    }
    std::string kmers_file, output_file;
    unsigned len = 25, hash_K = 16, hash_L = 32;
    double hash_W = 50, hash_R = 200;
    Options opt(strip_path(argv[0]), "cluster kmers to motifs");
    opt.add("kmers", 'k', "kmers file", true, kmers_file);
    opt.add("len", 'l', "kmer length", true, len);
    opt.add("hash_K", 'K', "number of random lines", true, hash_K);
    opt.add("hash_L", 'L', "number of hash tables", true, hash_L);
    opt.add("window", 'W', "bucket width", true, hash_W);
    opt.add("threshold", 'T', "clustering threshold", true, hash_R);
    opt.add("output", 'o', "output file name", true, output_file);
    std::vector<std::string> rest;
    opt.parse(argc, argv, rest);
    if (handled_help(argc, opt)) return EXIT_SUCCESS;

    // tokens: ">name" then the k-mer (hclust2.cpp:232-240)
    std::vector<std::string> names;
    std::vector<uint8_t> codes;
    {
      std::ifstream fin(kmers_file.c_str());
      std::string tok, name;
      while (fin >> tok) {
        if (tok[0] == '>') {
          name = tok.substr(1);
          if (!(fin >> tok)) break;
          if (tok.size() < len)
            throw CliError("k-mer '" + tok + "' is shorter than -l (the reference reads past its end)");
          for (unsigned i = 0; i < len; ++i) {
            const int c = (tok[i] >= 'A' && tok[i] <= 'Z') ? hs_letter_to_code(tok[i]) : -1;
            if (c < 0) throw CliError("k-mer '" + tok + "' holds a letter that is not one of the 20 amino acids");
            codes.push_back((uint8_t)c);
          }
          names.push_back(name);
        }
      }
    }
    const size_t n = names.size();
    printf("The number of kmers is %u\n", (unsigned)n);
    const clock_t start = clock();
    std::cout << "Clustering... " << std::endl;

    hs_params prm;
    memset(&prm, 0, sizeof prm);
    prm.len = len;
    prm.K = hash_K;
    prm.L = hash_L;
    prm.W = hash_W;
    prm.R = hash_R;
    prm.table_variant = HS_TABLE_FULL;      // KmerToCoordinates embeds with util.hpp:21-42 as written
    prm.metric = HS_METRIC_EUCLID_FP64;
    prm.predicate = HS_PRED_SQRT_LE_R;      // PairwiseDistance(...) <= hash_R (hclust2.cpp:64-71,119-120)
    std::vector<uint32_t> center(n), round(n);
    std::vector<uint8_t> merged(n);
    const clock_t cstart = clock();
    if (n) {
      Ctx ctx(device_from_env(), prm);
      const uint32_t dim = HS_CDIM * len;
      const uint64_t seed = projection_seed_base();
      std::vector<double> a((size_t)hash_L * hash_K * dim), b((size_t)hash_L * hash_K);
      for (unsigned l = 0; l < hash_L; ++l)
        check(hs_generate_projection(seed + l, dim, hash_K, hash_W, &a[(size_t)l * hash_K * dim], &b[(size_t)l * hash_K]),
              "hs_generate_projection");
      check(hs_set_projection(ctx.h, a.data(), b.data()), "hs_set_projection");
      check(hs_load_fragments(ctx.h, codes.data(), n, 0), "hs_load_fragments");
      check(hs_build_index(ctx.h), "hs_build_index");
      check(hs_greedy_cluster(ctx.h, center.data(), round.data(), merged.data()), "hs_greedy_cluster");
    }
#ifdef HCLUST3
    for (unsigned l = 0; l < hash_L; ++l) std::cout << "BuildLSHTalbes... " << std::endl;  // hclust3.cpp:77, once per round
#endif
    printf("ClusteringTime takes %lf seconds\n", (clock() - cstart) / (double)CLOCKS_PER_SEC);

    // clusters[i].ids (hclust2.cpp:96-99,121): the head, then its members in joining order
    // = by round, inside a round by ascending id (members of one bucket, in id order)
    std::vector<uint32_t> order(n);
    for (size_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
      if (center[x] != center[y]) return center[x] < center[y];
      const bool hx = center[x] == x && merged[x] != 2, hy = center[y] == y && merged[y] != 2;
      if (hx != hy) return hx;
      if (round[x] != round[y]) return round[x] < round[y];
      return x < y;
    });
    std::vector<uint32_t> size(n, 0), first(n, 0);
    for (size_t i = 0; i < n; ++i) size[center[order[i]]]++;
    for (size_t i = n; i-- > 0;) first[center[order[i]]] = (uint32_t)i;
    std::ofstream fout(output_file.c_str());
    uint32_t cluster_id = 0, num_of_kmers = 0;
    for (size_t i = 0; i < n; ++i) {
      if (merged[i] == 1 || merged[i] == 0) {
        fout << "#clusterid:" << cluster_id++ << ":size" << size[i] << std::endl;
        num_of_kmers += size[i];
        for (uint32_t j = 0; j < size[i]; ++j) fout << names[order[first[i] + j]] << std::endl;
      }
    }
    std::cout << "num_of_kmers = " << num_of_kmers << std::endl;
    fout.close();
    printf("Clustering takes %lf seconds\n", (clock() - start) / (double)CLOCKS_PER_SEC);
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
