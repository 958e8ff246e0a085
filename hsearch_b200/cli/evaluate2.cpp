// evaluate2 -- drop-in for the reference's ground-truth sorter
// (hclust/src/hclust/evaluate2.cpp:75-96; everything after its `return 0` at :96 is dead
// code): reads `motif protein distance` triples from argv[1], echoes the file name, sorts
// by (motif, protein) and writes "<argv[1]>sort.txt" with tab-separated fields -- the file
// motif_both_points takes as -g.  Host-only text utility of the recall workflow (R1).
#include <algorithm>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

struct MotifRes {
  std::string motif, protein;
  double dis;
};

int main(int argc, const char **argv) {
  std::vector<MotifRes> rows;
  // like the reference, argv[1] is used unchecked: a missing file gives an empty list
  const std::string in = argc > 1 ? argv[1] : "";
  std::ifstream fin(in.c_str());
  std::cout << in << std::endl;
  MotifRes r;
  while (fin >> r.motif >> r.protein >> r.dis) rows.push_back(r);
  fin.close();
  // sortCMP (:34-39); (motif, protein) pairs are unique in a brute-force result, so the
  // unspecified order of equal elements under std::sort never shows
  std::sort(rows.begin(), rows.end(), [](const MotifRes &a, const MotifRes &b) {
    if (a.motif == b.motif) return a.protein < b.protein;
    return a.motif < b.motif;
  });
  std::ofstream fout((in + "sort.txt").c_str());
  for (const MotifRes &x : rows) fout << x.motif << "\t" << x.protein << "\t" << x.dis << std::endl;
  fout.close();
  return 0;
}
