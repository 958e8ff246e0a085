// oracle/_ref harness, TU 3: pcluster's UnionFind, KLSH, Kmer2Integer and FASTA
// reader compiled in place.  TEST INFRASTRUCTURE ONLY.
#include "pcluster/src/pcluster/util.hpp"
#include "pcluster/src/pcluster/lsh.hpp"
#include "pcluster/src/pcluster/read_proteins.hpp"
#include "pcluster/src/pcluster/union_find.hpp"
#include <cstring>

extern "C" {
// UnionFind (union_find.cpp:3-33) over ids[0..n), edges fed with the
// FindRoot-then-JoinUnion protocol; root_out[i] = FindRoot(ids[i]) at the end.
void ref_union_find(const uint32_t *ids, uint32_t n, const uint32_t *eu, const uint32_t *ev,
                    uint64_t ne, uint32_t *root_out) {
  std::vector<uint32_t> v(ids, ids + n);
  ProteinDB db("/dev/null");
  UnionFind uf(v, db);
  for (uint64_t e = 0; e < ne; ++e) {
    uint32_t x = eu[e], y = ev[e];
    uf.FindRoot(x);
    uf.FindRoot(y);
    uf.JoinUnion(x, y);
  }
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t x = ids[i];
    root_out[i] = (uint32_t)uf.FindRoot(x);
  }
}
// KLSH::GetHashValue (lsh.cpp:40-49) of a default-seeded KLSH(feat,bits,sigma)
uint64_t ref_klsh_hash(const double *p, uint32_t feat, uint32_t bits, double sigma) {
  KLSH k(feat, bits, sigma);
  std::vector<double> v(p, p + feat);
  return k.GetHashValue(v);
}
// Kmer2Integer (util.hpp:244-250)
uint32_t ref_kmer2integer(const char *kmer) { return Kmer2Integer(kmer); }
// ProteinDB::ReadFASTAFile (read_proteins.cpp:6-41): returns #proteins, writes
// "name\tSEQ\n" records to out (caller buffer).
uint32_t ref_read_fasta(const char *path, char *out, uint64_t cap) {
  ProteinDB db(path);  // the constructor already calls ReadFASTAFile()
  std::string s;
  for (uint32_t i = 0; i < db.num_of_proteins; ++i) {
    s += db.pro_names[i];
    s += '\t';
    s.append(db.pro_seqs[i].begin(), db.pro_seqs[i].end());
    s += '\n';
  }
  if (s.size() + 1 <= cap) memcpy(out, s.c_str(), s.size() + 1);
  return db.num_of_proteins;
}
}
