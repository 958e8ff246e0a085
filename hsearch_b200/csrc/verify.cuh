// K3/K4/K5 interfaces (verify.cu): probe, candidate filter, exact verify + emit.
#pragma once
#include "common.cuh"

namespace hs {

// One unit of filter work: the members [m_begin, m_end) of one bucket (positions
// in the bucket-ordered store of `table`) against the queries
// qlist[q_begin .. q_end).  It owns blocks [block_begin, block_begin + ntiles).
struct WorkItem {
  uint32_t table;
  uint32_t m_begin, m_end;
  uint32_t q_begin, q_end;
  uint32_t block_begin;
};

// A (query, member) pair that passed the filter.
struct Survivor {
  uint32_t query;  // query index (search / brute force) or member position (self join)
  uint32_t table;
  uint32_t pos;    // member position in the table's bucket order
  uint32_t pad;    // bit 0: `query` is an index into the pipelined tensor filter's query list;
                   // bit 1: `pos` is already the fragment id
};

constexpr int kFilterThreads = 256;
constexpr int kFilterMembers = 4;                                  // members per thread
constexpr int kFilterTile = kFilterThreads * kFilterMembers;       // members per block
constexpr int kFilterQChunk = 8;                                   // query tables staged per step
constexpr uint32_t kQueriesPerItem = 256;

// Search: LSH query loop (dedup across tables).  AllPairs: all pairs i<j of the
// DB.  SelfJoin: pairs i<j inside one bucket (cluster).  Brute: explicit queries
// against the whole DB, no dedup.
enum FilterMode { kModeSearch = 0, kModeAllPairs = 1, kModeSelfJoin = 2, kModeBrute = 3 };

struct FilterArgs {
  const WorkItem *items;
  uint32_t nitems;
  const uint32_t *qlist;         // query indices (search / brute force)
  const float *tq;               // [Q][len][20] filter tables (search / brute force)
  uint32_t tq_base;              // tq row of query id x is x - tq_base
  const float *dsq32;            // [20][20] residue-pair table (self join)
  const uint8_t *const *stores;  // per table: position-major code*4 store
  uint64_t npad;
  int len;
  float thr;                     // pass iff filter distance <= thr
  Survivor *surv;
  unsigned long long surv_cap;
  unsigned long long *surv_count;
};

int launch_probe(hs_ctx *ctx, uint32_t table, const uint64_t *d_qkeys, const uint8_t *d_qvalid, uint32_t Q,
                 uint2 *d_qrange, uint32_t *d_qrank);
int launch_build_tq_points(hs_ctx *ctx, const double *d_q64, uint32_t Q, float *d_tq);
int launch_detect_query_codes(hs_ctx *ctx, const double *d_q64, uint32_t Q, uint8_t *d_qcodes, uint8_t *d_qrow);
int launch_build_tq_int(hs_ctx *ctx, const uint8_t *d_qcodes, uint32_t Q, float *d_tq);
int launch_filter(hs_ctx *ctx, const FilterArgs &args, uint32_t nblocks, int mode);

// Tensor-core filter (filter_tc.cu).  A TC work item holds <= kTcQueriesPerItem
// queries; its block_begin counts blocks of kTcTilesPerBlock 128-member tiles.
constexpr uint32_t kTcQueriesPerItem = 128;
constexpr uint32_t kTcTileMembers = 128;
constexpr uint32_t kTcTilesPerBlock = 16;
constexpr uint32_t kTcMinMembers = 64;
int tc_min_queries();           // buckets probed by fewer queries stay on the scalar filter
uint32_t tc_padded_k(uint32_t len);
int launch_tq_to_half(hs_ctx *ctx, const float *d_tq, uint64_t nq, void *d_tq16);
int launch_filter_tc(hs_ctx *ctx, const FilterArgs &args, const void *d_tq16, uint32_t tiles_per_block,
                     uint32_t nblocks, int mode);

// Pipelined tensor-core filter for the Euclidean metric (filter_mma.cu): one
// persistent CTA per SM walks a contiguous range of units; a unit is a chunk of
// one bucket's members against one item (<= qmax queries of that bucket).
struct MmaGeometry {
  int kp;        // K padded: 8*len coordinates + 2 constant columns, multiple of 16
  int nstages;   // A stages in shared memory
  int qmax;      // queries per item (B rows resident in shared memory)
  int cring;     // tiles of residue codes in flight
  size_t smem;   // dynamic shared memory per CTA
  double beta;   // relative error bound of the FP16/FP32 dot product
};
struct MmaItemHost {
  uint32_t table, q_begin, q_end, pad;
};
struct MmaUnitHost {
  uint32_t item, m_begin, m_end, pad;
};
// 1: the pipelined tensor filter stages threshold events (16 raw accumulators) and
// resolve_events_kernel finds the passing columns; 0 (default): the epilogue tests the columns
// itself.  Measured at bench C2: 41.6 ms (incl. 2-3 ms resolver) against 34.3 ms.
#ifndef HS_MMA_EVENTS
#define HS_MMA_EVENTS 0
#endif
#ifndef HS_MMA_UNIT_TILES
#define HS_MMA_UNIT_TILES 32
#endif
constexpr uint32_t kMmaUnitTiles = HS_MMA_UNIT_TILES;  // 128-member tiles per unit
constexpr uint32_t kMmaMinMembers = 64;
int mma_geometry(const hs_ctx *ctx, MmaGeometry *g);
bool mma_filter_usable(const hs_ctx *ctx);
int mma_upload_tables(hs_ctx *ctx);
int launch_build_qb_points(hs_ctx *ctx, const double *d_q64, uint32_t Q, void *d_qb16);
int launch_build_qb_codes(hs_ctx *ctx, uint64_t q0, uint32_t nq, void *d_qb16);
int launch_build_qb_qcodes(hs_ctx *ctx, const uint8_t *d_qcodes, uint32_t nq, void *d_qb16);
int launch_build_qb_store(hs_ctx *ctx, uint32_t table, uint32_t pos0, uint32_t nq, void *d_qb16);
int launch_filter_mma(hs_ctx *ctx, const FilterArgs &fa, const void *d_items, const void *d_units, uint32_t nunits,
                      uint32_t *d_unit_counter, uint32_t grid, const void *d_qb16, int mode);

// ---- lock-free union-find (device) ------------------------------------------------------
__device__ __forceinline__ uint32_t uf_find(uint32_t *parent, uint32_t x) {
  // path halving; parent[] only ever decreases, so racing updates stay valid
  volatile uint32_t *p = parent;
  while (true) {
    const uint32_t px = p[x];
    if (px == x) return x;
    const uint32_t ppx = p[px];
    if (ppx != px) p[x] = ppx;
    x = px;
  }
}
// JoinUnion (union_find.cpp:31-33) made order-independent: the larger root is
// hooked under the smaller, so the final root of a component is its min id.
__device__ __forceinline__ void uf_union(uint32_t *parent, uint32_t a, uint32_t b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    const uint32_t hi = a > b ? a : b, lo = a > b ? b : a;
    if (atomicCAS(parent + hi, hi, lo) == hi) return;
  }
}

struct ExactArgs {
  const Survivor *surv;
  unsigned long long nsurv;
  int mode;                      // FilterMode
  int metric, predicate;
  int len, dim, key_words, L;
  double R;
  const uint32_t *const *sorted_ids;  // per table (nullptr entry: identity)
  const uint8_t *codes;               // [N][len]
  const uint8_t *rec;                 // fragment records [N][rec_stride]: codes, then (rank path) u16 ranks
  uint32_t rec_stride, rec_rank_off;
  const uint32_t *qrank;              // rank path: [L][Q] bucket slot of every query (dedup); else nullptr
  uint64_t N, id_base;
  const double *table64;
  const int32_t *metric_tab;          // [20][20]
  const double *q64;                  // [Q][dim]   (Euclid, search / brute force)
  const uint8_t *qcodes;              // [Q][len]   (integer metric; Euclid: codes of embedded-string queries)
  const uint8_t *qrow;                // [Q] 1: the dense query equals the embedding of qcodes (Euclid)
  uint32_t Q;
  const uint64_t *const *keys;        // per table: [KW][N] original-order keys (dedup)
  const uint64_t *qkeys;              // [L][Q][KW]
  const uint8_t *qvalid;              // [L][Q]
  const uint32_t *qlist_mma;          // query list of the pipelined tensor filter (survivors with pad = 1)
  hs_hit *hits;
  unsigned long long hit_cap;
  unsigned long long *hit_count;
  uint32_t *parent;                   // self join: union-find forest
  unsigned long long *edge_count;
};
int launch_exact(hs_ctx *ctx, const ExactArgs &args);

}  // namespace hs
