// K1: p-stable LSH hash of residue-code fragments (H2-H4 of SURVEY.md 8a).
//
// Reference: LSH::DotProduct / HashBucketIndex / HashKey, hclust/src/hclust/
// lsh.hpp:33-59, over the build loop motif_both_points.cpp:212-216.
//
// Fast path (FP32): because a fragment is a string over 20 letters, the dot
// product a.v splits into `len` residue terms.  The host precomputes
// T[pos][code][proj] = sum_j table[code][j] * a[proj][8*pos+j] (FP64, rounded
// to FP32); a fragment's dot product is then `len` shared-memory lookups and
// FP32 adds per projection instead of 8*len FMAs.  Each projection carries a
// rigorous error bound eps (host, hash_host.cpp): if (dot+b)/W lies within eps
// of an integer the projection is recomputed in FP64 in the reference's exact
// operation order (sequential multiply, then add; division; floor).  Outside
// the guard band FP32 and FP64 buckets provably agree, so keys are bit-exact.
#pragma once
#include "common.cuh"

namespace hs {

// ---- packed digit-string key -------------------------------------------------
// HashKey (lsh.hpp:51-59) concatenates std::to_string(bucket) with no
// separator, so (1,23) and (12,3) share a bucket.  The packed key therefore
// encodes the *string*: one nibble per character ('0'..'9' -> 1..10, '-' ->
// 11), appended left to right into a KW*64-bit big integer (word 0 least
// significant).  All nibbles are non-zero, so two strings are equal iff their
// packed integers are equal, whatever their lengths.
template <int KW>
struct KeyBuilder {
  uint64_t w[KW];
  int nchars;
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int i = 0; i < KW; ++i) w[i] = 0;
    nchars = 0;
  }
  __device__ __forceinline__ void push(uint32_t nib) {
#pragma unroll
    for (int j = KW - 1; j > 0; --j) w[j] = (w[j] << 4) | (w[j - 1] >> 60);
    w[0] = (w[0] << 4) | (uint64_t)nib;
    ++nchars;
  }
  // std::to_string(int): the whole string (sign and digits, <= 11 characters) is assembled in one
  // word and appended with a single multi-word shift
  __device__ __forceinline__ void push_int(int v) {
    uint32_t u = v < 0 ? (uint32_t)(-(long long)v) : (uint32_t)v;
    uint64_t str = 0;
    int nc = 0;
    do {
      const uint32_t q = u / 10u;
      str |= (uint64_t)(u - q * 10u + 1u) << (4 * nc);
      u = q;
      ++nc;
    } while (u);
    if (v < 0) {
      str |= 11ull << (4 * nc);
      ++nc;
    }
    const int sh = 4 * nc;  // 4 .. 44
#pragma unroll
    for (int j = KW - 1; j > 0; --j) w[j] = (w[j] << sh) | (w[j - 1] >> (64 - sh));
    w[0] = (w[0] << sh) | str;
    nchars += nc;
  }
};

// FP64 bucket of one projection in the reference's operation order
// (lsh.hpp:33-49): dot += point[i] * a[i] with separate multiply and add,
// val = dot + b, floor(val / W).
__device__ __forceinline__ int exact_bucket_codes(const uint8_t *codes, int len,
                                                  const double *__restrict__ table64,
                                                  const double *__restrict__ a_row, double b, double W) {
  double dot = 0.0;
  for (int pos = 0; pos < len; ++pos) {
    const double *row = table64 + (int)codes[pos] * HS_CDIM;
#pragma unroll
    for (int j = 0; j < HS_CDIM; ++j) dot = __dadd_rn(dot, __dmul_rn(row[j], a_row[pos * HS_CDIM + j]));
  }
  double val = __dadd_rn(dot, b);
  return (int)floor(__ddiv_rn(val, W));
}

__device__ __forceinline__ int exact_bucket_point(const double *__restrict__ pt, int dim,
                                                  const double *__restrict__ a_row, double b, double W) {
  double dot = 0.0;
  for (int i = 0; i < dim; ++i) dot = __dadd_rn(dot, __dmul_rn(pt[i], a_row[i]));
  double val = __dadd_rn(dot, b);
  return (int)floor(__ddiv_rn(val, W));
}

struct HashChunkArgs {
  uint64_t *keys[8];  // per table of the chunk: [KW][N]   (packed-key path)
  int l0;             // first table of the chunk
  int ntab;           // tables in the chunk
  // dense bucket ranks (rank path): rank = lut[mixed-radix index of the bucket tuple]
  uint16_t *ranks[8];        // per table of the chunk: [N]
  const uint16_t *lut[8];    // per table of the chunk
  int lo[32], rng[32];       // per projection slot of the chunk: smallest bucket, bucket count
  float b32[32], eps32[32];  // per projection slot: offset b and guard band (FP32 fast path)
  uint8_t *rec;              // fragment records [N][rec_stride]
  uint32_t rec_stride, rec_rank_off;
  int full_rec;              // 1: this launch writes whole records (codes + ranks) through shared memory
  int k4_full;               // 1: K == 4 and all 4*NQ projection slots of the chunk are in use
};

// Mixed-radix index of one table's bucket tuple; returns false when a bucket lies
// outside the range setup_projection derived (cannot happen for residue strings).
__device__ __forceinline__ bool rank_tuple_push(uint32_t &tix, int bucket, int lo, int rng) {
  const int d = bucket - lo;
  tix = tix * (uint32_t)rng + (uint32_t)d;
  return d >= 0 && d < rng;
}

constexpr uint64_t kHashRangeAlign = 768 * 256;  // multiple of every hash block size (768, 256)
int launch_hash_fast(hs_ctx *ctx, bool want_buckets, uint64_t f0 = 0, uint64_t f1 = 0);
bool hash_single_launch_records(const hs_ctx *ctx);  // the hash launch writes whole fragment records
int launch_hash_exact(hs_ctx *ctx, bool want_buckets, bool audit);
int ensure_records(hs_ctx *ctx);  // fragment records hold at least the codes
int launch_hash_queries(hs_ctx *ctx, const double *d_q64, uint32_t Q, uint64_t *d_qkeys, uint8_t *d_qvalid);

}  // namespace hs
