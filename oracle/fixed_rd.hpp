// Determinism shim for compiling the reference's sources in place (oracle/_ref
// only; test infrastructure).  Force-included with `-include fixed_rd.hpp`:
// the reference seeds each LSH object from std::random_device
// (hclust/src/hclust/lsh.hpp:18-19); here the n-th random_device() call of the
// process returns seed_base + n, so table l of a run uses seed_base + l.
// The reference files themselves are not edited.
#pragma once
#include <random>
#include <cstdlib>
struct hs_fixed_rd {
  typedef unsigned int result_type;
  static unsigned long long &base() {
    static unsigned long long b = [] {
      const char *e = std::getenv("HS_REF_SEED");
      return e ? std::strtoull(e, nullptr, 10) : 12345ULL;
    }();
    return b;
  }
  static unsigned long long &counter() {
    static unsigned long long c = 0;
    return c;
  }
  static void reset(unsigned long long seed_base) {
    base() = seed_base;
    counter() = 0;
  }
  result_type operator()() { return (result_type)(base() + counter()++); }
  static constexpr result_type min() { return 0; }
  static constexpr result_type max() { return 0xffffffffu; }
};
#define random_device hs_fixed_rd
