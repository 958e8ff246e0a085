// Shared pieces of the drop-in command-line programs: the reference's banner,
// its text file formats, and thin RAII over the C ABI.  Host code only; every
// compute step goes through libhsearch_b200.so (include/hsearch_b200.h).
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <fstream>
#include <iostream>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/hsearch_b200.h"
#include "options.hpp"

namespace hscli {

static const char *const kVersion = "1.0";  // hclust_version, hclust/src/hclust/util.hpp

struct CliError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

inline void check(int rc, const char *what) {
  if (rc != HS_OK) throw CliError(std::string(what) + ": " + hs_last_error());
}

// "[WELCOME TO HSEARCH v1.0]" + echoed argv (motif_both_points.cpp:266-274).
// Returns true when one of -help / -about / -? is on the command line.
inline bool banner(int argc, const char **argv) {
  bool help = false;
  for (int i = 1; i < argc; ++i)
    if (!strcmp(argv[i], "-help") || !strcmp(argv[i], "-about") || !strcmp(argv[i], "-?")) help = true;
  if (argc > 1 && !help) {
    fprintf(stdout, "[WELCOME TO HSEARCH v%s]\n", kVersion);
    fprintf(stdout, "[%s", argv[0]);
    for (int i = 1; i < argc; ++i) fprintf(stdout, " %s", argv[i]);
    fprintf(stdout, "]\n");
  }
  return help;
}

// The reference's exit protocol for help / about / missing options
// (motif_both_points.cpp:324-335): message on stderr, status 0.
inline bool handled_help(int argc, const Options &opt) {
  if (argc == 1 || opt.help_requested()) {
    fprintf(stderr, "%s\n", opt.help_message().c_str());
    return true;
  }
  if (opt.about_requested()) {
    fprintf(stderr, "%s\n", opt.about_message().c_str());
    return true;
  }
  if (opt.option_missing()) {
    fprintf(stderr, "%s\n", opt.option_missing_message().c_str());
    return true;
  }
  return false;
}

// Seed of the LSH projection.  The reference draws it from std::random_device
// (lsh.hpp:15-16), one draw per table; HS_REF_SEED pins the sequence the same
// way oracle/fixed_rd.hpp pins the reference binaries (n-th draw = seed + n).
inline uint64_t projection_seed_base() {
  if (const char *e = getenv("HS_REF_SEED")) return strtoull(e, nullptr, 10);
  std::random_device rd;
  return rd();
}

// ---- point files (motif_both_points.cpp:343-370) ------------------------------
// Repeated: one name line, one line of DIM whitespace-separated decimals.
struct PointFile {
  std::vector<std::string> names;
  std::vector<double> data;  // [n][dim]
  size_t size() const { return names.size(); }
};

inline PointFile read_points(const std::string &path, uint32_t dim) {
  PointFile pf;
  std::ifstream fin(path.c_str());
  std::string line;
  while (std::getline(fin, line)) {
    pf.names.push_back(line);
    std::getline(fin, line);
    const size_t base = pf.data.size();
    pf.data.resize(base + dim, 0.0);
    const char *p = line.c_str();
    for (uint32_t i = 0; i < dim; ++i) {
      char *end = nullptr;
      const double v = strtod(p, &end);
      if (end == p) break;  // short line: remaining coordinates stay 0, as with a failed operator>>
      pf.data[base + i] = v;
      p = end;
    }
  }
  return pf;
}

// DB points written by protein2datapoints are embeddings of residue strings:
// every 8-vector is a row of `coordinates` after the 6-significant-digit print
// (HS_TABLE_PRINT6) or at full precision (HS_TABLE_FULL).  Recover the 1-byte
// codes the device stores.  Returns the table variant, or -1 when some vector
// is not a table row (arbitrary DB points are outside this build's scope).
inline int points_to_codes(const PointFile &pf, uint32_t len, std::vector<uint8_t> &codes) {
  for (uint32_t variant : {(uint32_t)HS_TABLE_PRINT6, (uint32_t)HS_TABLE_FULL}) {
    double tab[HS_AA * HS_CDIM];
    hs_get_coordinates(variant, tab);
    codes.assign(pf.size() * len, 0);
    bool ok = true;
    for (size_t i = 0; ok && i < pf.size(); ++i)
      for (uint32_t p = 0; ok && p < len; ++p) {
        const double *v = &pf.data[(i * len + p) * HS_CDIM];
        int found = -1;
        for (int c = 0; c < HS_AA && found < 0; ++c)
          if (!memcmp(v, tab + c * HS_CDIM, sizeof(double) * HS_CDIM)) found = c;
        if (found < 0) ok = false;
        else codes[i * len + p] = (uint8_t)found;
      }
    if (ok) return (int)variant;
  }
  return -1;
}

// ostream << double at the default precision == "%g" (6 significant digits).
inline std::string fmt_g(double v) {
  char buf[64];
  snprintf(buf, sizeof buf, "%g", v);
  return buf;
}

struct Ctx {
  hs_ctx_t *h = nullptr;
  Ctx(int device, const hs_params &p) { check(hs_create(&h, device, &p), "hs_create"); }
  ~Ctx() { hs_destroy(h); }
  Ctx(const Ctx &) = delete;
  Ctx &operator=(const Ctx &) = delete;
};

inline int device_from_env() {
  const char *e = getenv("HS_DEVICE");
  return e ? atoi(e) : 0;
}

// Calls a hit-producing entry point, growing the buffer on HS_ERR_CAPACITY.
template <class F>
std::vector<hs_hit> collect_hits(F call, uint64_t first_cap = 1u << 20) {
  std::vector<hs_hit> hits(first_cap);
  for (;;) {
    uint64_t n = 0;
    const int rc = call(hits.data(), (uint64_t)hits.size(), &n);
    if (rc == HS_ERR_CAPACITY) {
      hits.resize(n);
      continue;
    }
    check(rc, "search");
    hits.resize(n);
    return hits;
  }
}

// ---- multi-GPU mode of the search programs -------------------------------------------
// HS_DEVICES=0,1,...: the database is cut into contiguous id blocks, one per device; every
// device builds its own tables with the same projection and answers all centres (one host
// thread and one context per device); the lists are merged on the host into the reference's
// output order -- centre, table, ascending kmer id (motif_both_points.cpp:224-245): ids ascend
// with the device index, so inside a (centre, table) segment the devices' runs follow each other.
inline std::vector<int> devices_from_env() {
  std::vector<int> dev;
  if (const char *e = getenv("HS_DEVICES")) {
    std::stringstream ss(e);
    std::string tok;
    while (std::getline(ss, tok, ','))
      if (!tok.empty()) dev.push_back(atoi(tok.c_str()));
  }
  if (dev.empty()) dev.push_back(device_from_env());
  return dev;
}

struct ShardedResult {
  std::vector<hs_hit> hits;             // merged, reference order
  std::vector<uint64_t> table_sizes;    // distinct buckets per table over the whole database
};

// run(ctx, shard_codes, shard_n, id_base) prepares one device (projection, load, index) ;
// search(ctx, buf, cap, n) is the hit-producing call.  want_sizes: count the buckets (LSH tables).
template <class Prep, class Search>
ShardedResult sharded_search(const std::vector<int> &devices, const hs_params &prm, const uint8_t *codes, uint64_t n,
                             uint32_t n_tables, Prep prep, Search search) {
  const size_t G = devices.size();
  std::vector<std::vector<hs_hit>> lists(G);
  std::vector<std::vector<std::vector<uint64_t>>> keys(G);   // [device][table] -> packed keys of the shard
  std::vector<std::string> errors(G);
  std::vector<std::thread> threads;
  for (size_t g = 0; g < G; ++g)
    threads.emplace_back([&, g]() {
      try {
        const uint64_t lo = n * g / G, hi = n * (g + 1) / G;
        Ctx ctx(devices[g], prm);
        prep(ctx.h, codes + lo * prm.len, hi - lo, lo);
        if (n_tables) {
          hs_stats st;
          check(hs_get_stats(ctx.h, &st), "hs_get_stats");
          const uint32_t kw = st.key_words ? st.key_words : 1;
          keys[g].resize(n_tables);
          for (uint32_t l = 0; l < n_tables && hi > lo; ++l) {
            std::vector<uint64_t> k((size_t)(hi - lo) * kw);
            check(hs_get_keys(ctx.h, l, k.data()), "hs_get_keys");
            // one 64-bit digest per key is enough to count distinct buckets (FNV over the words)
            keys[g][l].resize(hi - lo);
            for (uint64_t i = 0; i < hi - lo; ++i) {
              uint64_t hsh = 1469598103934665603ull;
              for (uint32_t w = 0; w < kw; ++w) hsh = (hsh ^ k[i * kw + w]) * 1099511628211ull;
              keys[g][l][i] = kw == 1 ? k[i] : hsh;
            }
            std::sort(keys[g][l].begin(), keys[g][l].end());
            keys[g][l].erase(std::unique(keys[g][l].begin(), keys[g][l].end()), keys[g][l].end());
          }
        }
        lists[g] = collect_hits([&](hs_hit *buf, uint64_t cap, uint64_t *nh) { return search(ctx.h, buf, cap, nh); });
      } catch (const std::exception &e) {
        errors[g] = e.what();
      }
    });
  for (std::thread &t : threads) t.join();
  for (const std::string &e : errors)
    if (!e.empty()) throw CliError(e);
  ShardedResult r;
  for (uint32_t l = 0; l < n_tables; ++l) {
    std::vector<uint64_t> all;
    for (size_t g = 0; g < G; ++g) all.insert(all.end(), keys[g][l].begin(), keys[g][l].end());
    std::sort(all.begin(), all.end());
    r.table_sizes.push_back((uint64_t)(std::unique(all.begin(), all.end()) - all.begin()));
  }
  // merge: every list is in (query, table, id) order
  size_t total = 0;
  for (const auto &v : lists) total += v.size();
  r.hits.reserve(total);
  std::vector<size_t> cur(G, 0);
  while (r.hits.size() < total) {
    // the smallest (query, table) at the cursors, then each device's run of it in device order
    uint64_t best = ~0ull;
    for (size_t g = 0; g < G; ++g)
      if (cur[g] < lists[g].size())
        best = std::min(best, ((uint64_t)lists[g][cur[g]].query << 32) | lists[g][cur[g]].table_first);
    for (size_t g = 0; g < G; ++g)
      while (cur[g] < lists[g].size() &&
             (((uint64_t)lists[g][cur[g]].query << 32) | lists[g][cur[g]].table_first) == best)
        r.hits.push_back(lists[g][cur[g]++]);
  }
  return r;
}

// ---- recall evaluation (motif_both_points.cpp:27-165) -----------------------------
struct MotifHit {
  std::string motif, protein;
  double dis;
};

// weight(): 1 below distance 24, then 1/(d-24) clipped to [0,1]; a ground-truth
// distance above R + 0.1 aborts the run with status 0 (:67-71).
inline double recall_weight(double dis, double R) {
  if (dis > R + 0.1) {
    std::cout << "err " << dis << std::endl;
    exit(0);
  }
  if (dis < 24) return 1;
  const double w = 1 / (dis - 24);
  if (w > 1 || w < 0) return 1;
  return w;
}

inline std::vector<MotifHit> read_hit_file(const std::string &path) {
  std::vector<MotifHit> v;
  std::ifstream fin(path.c_str());
  MotifHit h;
  while (fin >> h.motif >> h.protein >> h.dis) v.push_back(h);
  return v;
}

// Merge-join of the (already sorted) ground truth with the sorted LSH hits;
// prints the reference's diagnostic lines and writes <output>.accuracy.txt.
inline double evaluate_recall(const std::string &ground_truth, const std::string &output_file, double R) {
  const std::vector<MotifHit> truth = read_hit_file(ground_truth);
  std::vector<MotifHit> found = read_hit_file(output_file);
  std::sort(found.begin(), found.end(), [](const MotifHit &a, const MotifHit &b) {
    return a.motif == b.motif ? a.protein < b.protein : a.motif < b.motif;
  });
  auto compare = [](const MotifHit &a, const MotifHit &b) {
    if (a.motif == b.motif) return a.protein == b.protein ? 0 : (a.protein > b.protein ? 1 : -1);
    return a.motif > b.motif ? 1 : -1;
  };
  size_t i = 0, j = 0;
  double tp = 0.0, fn = 0.0, missed = 0;
  std::unordered_map<int, int> tp_bin, fn_bin;
  while (i < truth.size() && j < found.size()) {
    const int c = compare(truth[i], found[j]);
    if (c == 0) {
      tp += recall_weight(truth[i].dis, R);
      tp_bin[int(truth[i].dis * 100 / 10)]++;
      ++i;
      ++j;
    } else if (c == 1) {
      std::cout << "xnomo " << found[j].motif << " " << found[j].protein << " " << found[j].dis << " " << R << std::endl;
      ++j;
    } else {
      fn += recall_weight(truth[i].dis, R);
      fn_bin[int(truth[i].dis * 100 / 10)]++;
      std::cout << truth[i].dis << " " << recall_weight(truth[i].dis, R) << std::endl;
      ++missed;
      ++i;
    }
  }
  for (; i < truth.size(); ++i) {
    fn += recall_weight(truth[i].dis, R);
    fn_bin[int(truth[i].dis * 100 / 10)]++;
    ++missed;
  }
  std::cout << "ACCU: " << tp << " " << fn << " " << tp / (tp + fn) << "\t"
            << "size = " << missed << " " << truth.size() << " " << missed / (double)truth.size() << output_file
            << std::endl;
  std::ofstream fout((output_file + ".accuracy.txt").c_str());
  for (int b = 0; b < 500; ++b) {
    const bool has_fn = fn_bin.count(b), has_tp = tp_bin.count(b);
    if (has_fn && has_tp)
      fout << b << " " << tp_bin[b] / (fn_bin[b] + (double)tp_bin[b]) << " " << tp_bin[b] << " " << fn_bin[b] << std::endl;
    else if (has_fn)
      fout << b << " " << 0 << " fn " << fn_bin[b] << std::endl;
    else if (has_tp)
      fout << b << " " << 1 << " tp " << tp_bin[b] << std::endl;
  }
  return tp / (tp + fn);
}

// ---- FASTA, one sequence per line (ProteinDB, hclust/src/hclust/protein.hpp:41-71) --
struct ProteinStore {
  std::vector<std::string> names;
  std::vector<uint32_t> start;  // nprot + 1 offsets into residues
  std::vector<uint8_t> codes;   // residue codes after the AA20 round trip (E <-> Q swap)
  std::string letters;          // the stored letters (AA20[code'])
  uint32_t nprot() const { return start.empty() ? 0u : (uint32_t)start.size() - 1; }
};

inline ProteinStore read_protein_db(const std::string &path) {
  static const char AA20[] = "ARNDCEQGHILKMFPSTWYV";  // util.hpp:89
  ProteinStore db;
  std::ifstream fin(path.c_str());
  std::cout << "Read protein sequences from " << path << std::endl;
  std::string line;
  uint64_t total = 0;
  while (std::getline(fin, line)) {
    if (line.empty()) continue;
    if (line[0] == '>') {
      db.names.push_back(line.substr(1));
      continue;
    }
    db.start.push_back((uint32_t)db.codes.size());
    total += line.size();
    for (char ch : line) {
      int stored = hs_proteindb_code(ch);  // code of AA20[base[ch]]
      if (stored < 0) {                    // non-amino-acid letter -> random residue (:60-62)
        const int aa = rand() % 20;
        stored = hs_letter_to_code(AA20[aa]);
        db.letters.push_back(AA20[aa]);
      } else {
        const int aa = hs_letter_to_code(ch);
        db.letters.push_back(AA20[aa]);
      }
      db.codes.push_back((uint8_t)stored);
    }
  }
  db.start.push_back((uint32_t)db.codes.size());
  std::cout << "number of proteins " << db.nprot() << std::endl;
  std::cout << "total length " << total << std::endl;
  return db;
}

}  // namespace hscli
