// Internal definitions shared by the translation units of libhsearch_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/hsearch_b200.h"

namespace hs {

void set_error(const char *fmt, ...);

#define HS_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      hs::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return HS_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define HS_TRY(expr)            \
  do {                          \
    int _s = (expr);            \
    if (_s != HS_OK) return _s; \
  } while (0)

// Device buffer that grows but never shrinks (sized once for the 180 GB part;
// avoids cudaMalloc inside timed regions after the first call).
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return HS_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + (bytes >> 3) + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      e = cudaMalloc(&p, bytes);
      want = bytes;
    }
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
      p = nullptr;
      return HS_ERR_NOMEM;
    }
    cap = want;
    return HS_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T *as() const {
    return reinterpret_cast<T *>(p);
  }
};

// ---- geometry of the packed structures -------------------------------------
constexpr int kMaxKeyWords = HS_MAX_KEY_WORDS;
constexpr int kCodeScale = 4;  // bucket-ordered code store keeps code*4 (a byte offset into a float row)

// 64-bit hash of a packed key (hashed-key path of the index build, radix_sort.cu; probe, verify.cu).
// Equal keys hash equally; the build checks every bucket member's full key against its bucket's, so a
// collision of distinct keys is detected, never silently merged.
template <int NW>
__host__ __device__ __forceinline__ uint64_t key_hash(const uint64_t *k) {
  uint64_t h = 0x9E3779B97F4A7C15ull;
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    uint64_t x = h ^ k[w];
    x ^= x >> 30;
    x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27;
    x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    h = x + 0x9E3779B97F4A7C15ull * (uint64_t)(w + 1);
  }
  return h;
}

struct TableIndex {
  DevBuf sorted_ids;    // u32 [N]  fragment ids in bucket order
  DevBuf ukeys;         // u64 [key_words][nslots]  keys of the bucket slots, ascending
  bool hashed_keys = false;  // hashed-key path (radix_sort.cu): ukeys holds the ascending 64-bit hashes of the slots'
                             // key strings, one word per slot; a slot's key string is that of its first member
  DevBuf bstart;        // u32 [nslots+1]           bucket boundaries into sorted_ids
  DevBuf codes_sorted;  // u8  [len][npad]      code*4, position-major, bucket order
  uint64_t nb = 0;      // non-empty buckets (the reference's "table size")
  uint64_t nslots = 0;  // bucket slots: == nb on the packed-key path; on the rank path one slot
                        // per possible key string (empty slots have start == end)
};

struct SortScratch {
  DevBuf keys_alt[kMaxKeyWords];  // ping-pong key words
  DevBuf keys_cur[kMaxKeyWords];
  DevBuf vals_alt;
  DevBuf tile_hist;  // u32 [256][ntiles] (+1)
  DevBuf digit_hist; // u32 [ndigits][256]
  DevBuf flags;      // u32 [N] scan scratch
  DevBuf block_sums;
  DevBuf or_and;     // small reduction scratch
};

struct Stats : hs_stats {};

}  // namespace hs

// The opaque context (C name so the header's forward declaration matches).
struct hs_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  hs_params prm{};
  uint32_t dim = 0;

  // tables
  double table64[HS_AA * HS_CDIM];
  hs::DevBuf d_table64;  // f64 [20][8]
  // the table the pipelined tensor filter embeds residues with: the coordinates (Euclidean metric)
  // or a contracting embedding of the integer BLOSUM metric (tables.cpp: blosum_filter_embedding)
  double ftable64[HS_AA * HS_CDIM];
  hs::DevBuf d_ftable64;
  bool have_ftable = false;
  hs::DevBuf d_dsq32;    // f32 [20][20] squared residue distances (filter)
  hs::DevBuf d_metric;   // i32 [20][20] integer BLOSUM metric
  hs::DevBuf d_metric32; // f32 copy (filter tables of the integer metric)
  hs::DevBuf d_large;    // cluster: (begin, end) of large buckets

  // projection
  bool have_projection = false;
  std::vector<double> h_a, h_b;
  hs::DevBuf d_a64, d_b64;   // f64 [L][K][DIM], [L][K]
  hs::DevBuf d_a64t;         // f64 [L][DIM][qh_group]: the projection rows transposed for the query hash
  uint32_t qh_group = 1;     // K rounded up to a power of two
  hs::DevBuf d_T32;          // f32 [L][len][20][Kp]  residue-projection partial sums
  hs::DevBuf d_b32, d_eps32; // f32 [L][Kp]
  std::vector<float> h_b32, h_eps32;  // host copies (passed to the hash kernel as parameters)
  uint32_t Kp = 0;           // K rounded up to a multiple of 4
  uint32_t tpc = 1;          // tables per hash chunk
  uint32_t nchunks = 1;      // hash chunks
  uint32_t nq = 1;           // float4 quads of projections per chunk
  uint32_t key_words = 1;
  uint32_t max_chars = 0;

  // Dense bucket ranks (hash.cu, setup_projection): when every table has at most 65536
  // possible key strings, a fragment's bucket is the rank of its key string among them
  // (ascending packed key) -- a u16 instead of a KW*64-bit key.
  bool rank_mode = false;
  uint32_t rank_nr[HS_MAX_L];                 // possible key strings per table
  uint32_t rank_lut_off[HS_MAX_L];            // offset of the table's tuple->rank LUT in d_lut
  std::vector<int> rank_lo, rank_rng;         // [L][K] bucket range of every projection
  std::vector<uint64_t> h_rkeys[HS_MAX_L];    // [KW][nr] packed key of every rank (host copy)
  hs::DevBuf d_lut;                           // u16 tuple index -> rank, all tables
  hs::DevBuf d_rinfo;                         // i32 [L*K] lo, [L*K] rng, [L] lut offsets (audit kernel)
  hs::DevBuf d_ranks;                         // u16 [L][N] ranks, table-major (sort input)
  // Fragment records: codes, then (rank mode) the L u16 ranks; one aligned vector load
  // gives the exact stage / the bucket-order gather everything they need of a fragment.
  hs::DevBuf d_rec;                           // u8 [N][rec_stride]
  uint32_t rec_stride = 0, rec_rank_off = 0;
  bool have_rec = false;

  // database
  uint64_t N = 0, id_base = 0;
  uint64_t npad = 0;         // N rounded up to 16
  hs::DevBuf d_codes;        // u8 [N][len]
  bool hashed = false, indexed = false;
  hs::DevBuf d_keys[HS_MAX_L];     // u64 [key_words][N] per table, original order
  hs::DevBuf d_buckets;            // i32 [N][L][K] (only when requested)
  hs::TableIndex tables[HS_MAX_L];
  hs::DevBuf d_codes_pm;           // u8 [len][npad] code*4, original order (brute force / cluster)
  bool have_codes_pm = false;
  hs::SortScratch sort;

  // search scratch
  hs::DevBuf d_q64, d_qkeys, d_qvalid, d_qrange, d_tq, d_work, d_qlist, d_surv, d_hits, d_counters;
  hs::DevBuf d_hit_keys[3], d_hit_perm, d_hits_sorted, d_hits_sorted_alt;
  hs::DevBuf d_surv_blk;                // survivors regrouped by query block (pipelined verify / copy-out)
  hs::DevBuf d_cidt, d_cidt_alt, d_cdist, d_cdist_alt, d_coffsets;  // compact (CSR) hit output, double buffered
  hs::DevBuf d_seg_tab, d_seg_key, d_seg_dist, d_seg_ctl;           // segmented hit sort (hitsort.cu)
  cudaStream_t copy_stream = nullptr;   // D2H of sorted hit blocks, overlapped with the next block's search
  std::vector<cudaEvent_t> ev_chunk;
  hs::DevBuf d_misc, d_tabptrs, d_qcodes, d_hits_gathered, d_residues, d_starts;
  hs::DevBuf d_tq16, d_work_tc, d_qlist_tc;  // tensor-core filter: FP16 query tables, its work list
  hs::DevBuf d_tab16;                        // pipelined tensor filter: FP16 embedding rows + row norms
  hs::DevBuf d_qb16, d_mma_items, d_mma_units, d_mma_cta, d_qlist_mma;
  hs::DevBuf d_binctr;                       // survivor regrouping: per-bin counts and cursors
  hs::DevBuf d_events;                       // pipelined tensor filter: staged threshold events (filter_mma.cu)
  int num_sms = 0;
  uint64_t hit_qmax = 0;   // query ids of the current call are < hit_qmax (hit sort key width)
  uint64_t hit_idmax = 0;  // db ids of the hits being sorted are < hit_idmax (0: id_base + N)
  bool have_qcodes = false;
  hs::DevBuf d_qcodes_det, d_qrow;  // codes recovered from dense queries (Euclid exact stage)
  hs::DevBuf d_qrank;               // u32 [L][Q] bucket slot of every query (0xffffffff: none)
  void *h_pinned = nullptr;   // mapped pinned staging of read_back()
  void *d_pinned = nullptr;
  size_t h_pinned_cap = 0;
  void *h_up = nullptr;       // mapped pinned arena of upload(): host -> device control data
  void *d_up = nullptr;
  size_t up_used = 0;

  // cluster
  hs::DevBuf d_parent;

  // comm (comm.cu): NCCL for the small collectives, peer stores into rank 0's receive buffers
  // (mapped into every process with CUDA IPC) for the hit lists
  void *nccl_comm = nullptr;    // ctx stream: query broadcast, set-up exchanges
  void *nccl_comm2 = nullptr;   // gather stream: segment counts, completion barrier
  int rank = 0, nranks = 1;
  cudaStream_t gather_stream = nullptr;
  cudaEvent_t ev_gather[2] = {nullptr, nullptr}, ev_gather_in = nullptr;
  void *recv_local[2] = {nullptr, nullptr};   // rank 0: the receive buffers (hs_hit[recv_cap]), double buffered
  void *recv_mapped[2] = {nullptr, nullptr};  // every rank: rank 0's buffers as seen from this process
  uint64_t recv_cap = 0;
  uint32_t gather_seq = 0;      // gathers started so far (slot = seq & 1)
  bool gather_pending = false;
  unsigned long long h_gather_flag[2] = {0, 0};
  struct {                      // bulk transfer of the latest gather, started by the next filter launch
    bool pending = false, overflow = false;
    const hs_hit *hits = nullptr;
    uint64_t n = 0;
    int tbits = 0, slot = 0;
  } gather_def;
  hs::DevBuf d_segoff[2], d_segcnt, d_segcnt_all, d_segdst, d_gather_info;

  // debugging switches, read from the environment once in hs_create
  bool no_pipeline = false;      // HS_NO_PIPELINE: host-buffer searches in one pass (no query blocks)
  bool no_load_overlap = false;  // HS_NO_LOAD_OVERLAP: hs_load_fragments copies first, hashes later
  bool no_hash_sort = false;     // HS_NO_HASH_SORT: multi-word keys sorted on every word (no 64-bit key hash)
  int hash_sort_choice = -1;          // this index build: -1 undecided, 0 sort on all key words, 1 sort on the key hash
  bool force_hash_sort = false;       // HS_FORCE_HASH_SORT: multi-word keys always take the hashed sort (tests)
  bool force_hash_collision = false;  // HS_FORCE_HASH_COLLISION: test hook, the hashed sort reports a collision
  bool exact_rep = true;         // HS_EXACT_REP=0: exact stage without the 8-fold replicated residue-pair table
  bool no_lazy_stores = false;   // HS_NO_LAZY_STORES: hs_build_index always builds the bucket-ordered code stores
  uint64_t bypass_max = 0;       // candidates up to which a search on an index without code stores skips the filter
  bool stores_built = false;     // the tables' codes_sorted are valid (ensure_code_stores)
  bool plan_stats = false;       // HS_PLAN_STATS: print the filter work-list statistics
  uint32_t selfjoin_chunk = 1u << 16;  // HS_SELFJOIN_CHUNK: query members per tensor-filter pass of a large bucket (hs_cluster)
  bool surv_bins = false;        // HS_SURV_BINS: survivors regrouped by fragment-id block before the exact stage
  bool no_mma_int = false;       // HS_NO_MMA_INT: the integer metric stays on the one-hot tensor filter (filter_tc.cu)
  bool no_mma_filter = false;    // HS_NO_MMA_FILTER: keep the Euclidean metric off the pipelined tensor filter
  bool segsort = true;           // HS_SEGSORT=0: hit lists ordered by the radix passes only, never by the segmented sort (hitsort.cu)
  uint64_t segsort_min = 1u << 18;   // HS_SEGSORT_MIN: ... for lists of at least this many hits
  uint32_t segsort_buf = 1u << 30;   // HS_SEGSORT_BUF: keys per shared-memory buffer (test hook: forces the range path)
  uint32_t segsort_nblk = 0;         // HS_SEGSORT_NBLK: blocks of its partition pass (0: two per SM)
  bool segsort_radix = false;        // HS_SEGSORT_RADIX: its per-bin sort always takes the radix passes (A/B, tests)
  bool segsort_prof = false;         // HS_SEGSORT_PROF: per-kernel times of every segmented sort on stderr

  hs_stats stats{};
  hs_stats hash_stats{};   // counters / timing of the hash that produced the current keys
  cudaEvent_t ev[16];
  std::vector<cudaEvent_t> ev_pool;  // per-pass sort timing
};
