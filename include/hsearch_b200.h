/*
 * hsearch_b200.h -- C ABI of the B200-native HSEARCH hot path.
 *
 * The reference (acgtun/hsearch) has no plugin / FFI interface: it is a set of
 * single-threaded CLI programs.  The drop-in boundary is therefore (i) this C
 * ABI, which the flag-compatible CLI mains in hsearch_b200/cli/ call, and
 * (ii) the reference's argv grammar and text file formats, which those mains
 * keep.  Each entry point cites the reference code it replaces (paths relative
 * to the reference root).
 *
 * Conventions
 *   - plain pointers and sizes only; the caller allocates and frees every
 *     buffer, the library owns only the opaque hs_ctx;
 *   - every call returns 0 (HS_OK) or a negative hs_status; hs_last_error()
 *     gives the message of the calling thread's last failure; the library
 *     never throws across the boundary and never exits;
 *   - one hs_ctx per GPU; calls on one ctx are serialised by the caller; each
 *     ctx owns one CUDA stream;
 *   - there is NO CPU fallback: every compute entry point fails with
 *     HS_ERR_CUDA when no sm_100 device is usable;
 *   - residue codes are 0..19 in BLOSUM order A R N D C Q E G H I L K M F P S
 *     T W Y V (the row index of `coordinates`, hclust/src/hclust/util.hpp:21-42,
 *     i.e. base[letter-'A'], util.hpp:92).
 */
#ifndef HSEARCH_B200_H
#define HSEARCH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hs_ctx hs_ctx_t;

#define HS_AA 20
#define HS_CDIM 8            /* AACoordinateSize, util.hpp:94 */
#define HS_MAX_LEN 32        /* fragment length limit of this build */
#define HS_MAX_K 32
#define HS_MAX_L 64
#define HS_MAX_KEY_WORDS 4   /* packed digit-string key: <= 64 characters */

typedef enum {
  HS_OK = 0,
  HS_ERR_INVALID = -1,     /* bad argument / call order */
  HS_ERR_CUDA = -2,        /* CUDA runtime failure, or no usable sm_100 GPU */
  HS_ERR_CAPACITY = -3,    /* caller's hit buffer too small; *nhits = needed */
  HS_ERR_UNSUPPORTED = -4, /* parameter outside this build's limits */
  HS_ERR_NOMEM = -5,
  HS_ERR_COMM = -6         /* NCCL failure */
} hs_status;

/* Which 20x8 embedding table the DB fragments are embedded with (M1).
 * FULL   = util.hpp:21-42 as written (hclust2.cpp:49-62, kmer_search.cpp:52-62).
 * PRINT6 = the same table after the 6-significant-digit text round trip of
 *          protein2datapoints.cpp:23-29 -> motif_both_points.cpp:347-351, which
 *          is what motif_both_points really hashes. */
enum { HS_TABLE_FULL = 0, HS_TABLE_PRINT6 = 1 };
/* EUCLID_FP64 = PairwiseDistance(_square), motif_both_points.cpp:167-183.
 * BLOSUM_INT  = DistanceScore, BLOSUM-Metric/src/BLOSUM-metric/evaluate_correlation.cpp:34-41
 *               over distance_matrix.hpp:13-20. */
enum { HS_METRIC_EUCLID_FP64 = 0, HS_METRIC_BLOSUM_INT = 1 };
/* D2_LE_R2  : hit iff d2 <= R*R        (motif_both_points.cpp:204,239)
 * SQRT_LE_R : hit iff !(sqrt(d2) > R)  (motif_both_points_noLSH.cpp:46-47, hclust2.cpp:119-120)
 * ignored for BLOSUM_INT (hit iff d <= (int)R). */
enum { HS_PRED_D2_LE_R2 = 0, HS_PRED_SQRT_LE_R = 1 };
enum {
  HS_FLAG_SORT_HITS = 1u,   /* return hits in the reference's output order:
                               query, first-finding table, ascending db id */
  HS_FLAG_HASH_EXACT = 2u,  /* hash every projection in FP64 reference order
                               (no FP32 fast path); for validation */
  HS_FLAG_HASH_AUDIT = 4u,  /* after hashing, recompute every projection in
                               FP64 and count residual flips (must be 0) */
  HS_FLAG_SCALAR_FILTER = 8u /* keep every candidate on the scalar filter kernel
                               (no tcgen05 filter); for A/B validation */
};

typedef struct {
  uint32_t len;           /* fragment length in residues; DIM = 8*len (motif_both_points.cpp:337-338) */
  uint32_t K;             /* projections per table (hash_K) */
  uint32_t L;             /* tables (hash_L) */
  double W;               /* bucket width (hash_W) */
  double R;               /* distance threshold (hash_R) */
  uint32_t table_variant; /* HS_TABLE_* */
  uint32_t metric;        /* HS_METRIC_* */
  uint32_t predicate;     /* HS_PRED_* */
  uint32_t flags;         /* HS_FLAG_* */
} hs_params;

/* One verified pair.  dist2 = squared Euclidean distance exactly as the
 * reference's FP64 loop produces it (EUCLID_FP64), or the integer window
 * distance as a double (BLOSUM_INT).  db_id = id_base + local index. */
typedef struct {
  uint32_t query;
  uint32_t table_first; /* first table whose bucket held the pair (label[], motif_both_points.cpp:232-238) */
  uint64_t db_id;
  double dist2;
} hs_hit;

/* Counters and device timings (CUDA events on the ctx stream) of the most
 * recent hs_hash / hs_build_index / hs_search_* / hs_bruteforce_* / hs_cluster. */
typedef struct {
  uint64_t n_fragments;
  uint64_t guard_hits;      /* projections re-evaluated in FP64 (inside the FP32 guard band) */
  uint64_t guard_corrected; /* of those, FP32 bucket != FP64 bucket (flips the guard caught) */
  uint64_t residual_flips;  /* HS_FLAG_HASH_AUDIT: final bucket != FP64 bucket over ALL projections */
  uint64_t n_candidates;    /* (query, member) pairs examined by the filter kernel */
  uint64_t n_survivors;     /* pairs passed to the exact FP64 / dedup stage */
  uint64_t n_hits;
  uint64_t n_edges;         /* hs_cluster: near pairs found */
  uint64_t n_work_items;
  uint32_t key_words;       /* 64-bit words per packed key */
  uint32_t sort_passes;     /* radix passes actually executed (all tables) */
  uint32_t kernel_launches; /* kernels launched by the most recent call */
  uint32_t rank_path;       /* 1: buckets are dense u16 ranks of the key strings (<= 65536 possible
                               strings per table); 0: packed KW*64-bit key path */
  float ms_hash, ms_sort, ms_group, ms_permute;      /* index build stages */
  float ms_sort_upsweep, ms_sort_scan, ms_sort_downsweep; /* inside ms_sort, per kernel family */
  float ms_qhash, ms_probe, ms_filter, ms_exact, ms_hitsort; /* search stages */
  float ms_total;           /* whole call, first to last event */
  float ms_filter_tc;       /* inside ms_filter: the tensor-core (tcgen05) filter kernel */
  float ms_host;            /* search: host-side work-list construction between probe and filter */
  uint64_t n_candidates_tc; /* of n_candidates, pairs examined by the tensor-core filter */
  uint64_t hash_sort_fallbacks; /* index build, multi-word keys: tables whose sort on the 64-bit key hash met two
                               distinct keys with one hash and was redone on all key words (expected 0) */
  uint64_t segsort_lists;   /* search: hit lists put into the reference's order by the segmented sort (partition
                               by the top key bits, per-bin sort in shared memory; hitsort.cu) */
  uint64_t segsort_fallbacks; /* ... lists it handed back to the radix passes (a key field did not fit, or one
                               value of a bin's top 8 key bits alone exceeds the shared-memory buffer) */
} hs_stats;

/* ---- lifetime -------------------------------------------------------------- */
/* Replaces the per-process state of motif_both_points.cpp:252-383 (parameters
 * parsed at :302-335, K=L=4 forced at :380-381 -- here K and L are honoured). */
int hs_create(hs_ctx_t **out, int device, const hs_params *params);
void hs_destroy(hs_ctx_t *ctx);
const char *hs_last_error(void);
int hs_get_stats(hs_ctx_t *ctx, hs_stats *out);
/* 1 if a CUDA device of compute capability 10.x is visible, else 0. */
int hs_device_available(void);
/* The ctx's cudaStream_t (as void*), so that a caller can record its own CUDA
 * events around calls or order its own work after them. */
int hs_get_stream(hs_ctx_t *ctx, void **stream_out);

/* ---- embedding tables (M1..M3) ---------------------------------------------- */
/* out160 <- the 20x8 table of the given HS_TABLE_* variant. */
int hs_get_coordinates(uint32_t table_variant, double *out160);
/* out400 <- D[i][j] = B[i][i]+B[j][j]-2B[i][j] (distance_matrix.hpp:13-20). */
int hs_get_blosum_metric(int32_t *out400);
/* out160 <- the 20x8 table e the tensor filter embeds residues with under HS_METRIC_BLOSUM_INT:
 * |e(a) - e(b)|^2 <= D[a][b] for every residue pair, so the embedded squared distance of two fragments
 * never exceeds their integer window distance (evaluate_correlation.cpp:26-41) and a filter on it
 * loses no pair within R; the exact stage decides on the integers.  Host function. */
int hs_get_blosum_filter_embedding(double *out160);
/* Override the ctx's embedding table (default: params.table_variant). */
int hs_set_coordinates(hs_ctx_t *ctx, const double *table160);
/* letter -> code (base[letter-'A'], util.hpp:92); -1 for non-amino-acid letters. */
int hs_letter_to_code(char letter);
/* ProteinDB storage rule (protein.hpp:58-64): code after the AA20 round trip
 * (E and Q swap). */
int hs_proteindb_code(char letter);

/* ---- projection (H1) -------------------------------------------------------- */
/* LSH::LSH (lsh.hpp:10-31) for one table: default_random_engine(seed), per k
 * DIM normals then one uniform[0,W).  a[K][dim], b[K].  Host-side, libstdc++
 * <random>, exactly the reference's calls. */
int hs_generate_projection(uint64_t seed, uint32_t dim, uint32_t K, double W, double *a, double *b);
/* a[L][K][DIM], b[L][K] -> device; builds the FP32 residue-projection tables
 * and guard bands; fixes the packed-key width. */
int hs_set_projection(hs_ctx_t *ctx, const double *a, const double *b);

/* ---- database (E1, E2) ------------------------------------------------------ */
/* codes[N][len] host -> device.  Replaces the DB read loop
 * motif_both_points.cpp:343-354 (points are stored as 1-byte codes, not 8*len
 * doubles). */
int hs_load_fragments(hs_ctx_t *ctx, const uint8_t *codes, uint64_t N, uint64_t id_base);
/* Same, codes already in device memory (copied device-to-device). */
int hs_load_fragments_dev(hs_ctx_t *ctx, const void *codes_dev, uint64_t N, uint64_t id_base);
/* Sliding windows over a concatenated residue store (ProteinDB, protein.hpp:7-72;
 * window loop kmer_search.cpp:64-83 with the :73 bug fixed): every window start
 * j = 0, stride, 2*stride ... <= len_i - len of every protein becomes one
 * fragment, in protein order.  residues are codes; start_index has nprot+1
 * entries.  pos_out (optional, host, capacity pos_cap) receives the global
 * start position of each fragment; *nfrag the fragment count. */
int hs_extract_windows(hs_ctx_t *ctx, const uint8_t *residues, const uint32_t *start_index,
                       uint32_t nprot, uint32_t stride, uint64_t id_base, uint32_t *pos_out,
                       uint64_t pos_cap, uint64_t *nfrag);
uint64_t hs_num_fragments(hs_ctx_t *ctx);
/* ProteinDB::ProteinID (protein.hpp:28-39): protein_out[i] = the protein holding the global residue
 * position pos[i] -- the largest l with pos >= start_index[l], searched like the reference over all
 * nstart entries of start_index (nprot + 1: the end sentinel included, so a position at or past the
 * end gives nstart - 1).  With the positions hs_extract_windows returns this maps hits (fragment ids)
 * back to proteins.  Device binary search; host arrays. */
int hs_protein_id(hs_ctx_t *ctx, const uint32_t *start_index, uint32_t nstart, const uint32_t *pos, uint64_t n,
                  uint32_t *protein_out);
/* The data-point name protein2datapoints writes (protein2datapoints.cpp:61-65): the first
 * white-space-delimited token of the protein's header line, then #protein_index $offset @KMER *cnt.
 * Host only.  HS_ERR_CAPACITY when out_cap is too small. */
int hs_fragment_name(const char *protein_header, uint32_t protein_index, uint32_t offset, const char *kmer, uint32_t len,
                     uint64_t cnt, char *out, uint64_t out_cap);

/* ---- hash (H2..H4) ---------------------------------------------------------- */
/* Bucket ints floor((a.v+b)/W) for every fragment, table and projection, and
 * the packed digit-string keys.  buckets_out[N][L][K] (host) may be NULL.
 * Replaces LSH::HashKey over motif_both_points.cpp:212-216. */
int hs_hash(hs_ctx_t *ctx, int32_t *buckets_out);
/* FP64 audit of the keys hs_hash (or the overlapped hash of hs_load_fragments) produced:
 * every projection of every fragment is recomputed in FP64 in the reference's operation order
 * (lsh.hpp:33-49) and compared; *residual_flips = (fragment, table) keys that differ -- required
 * to be zero.  The same check HS_FLAG_HASH_AUDIT runs inside hs_hash, callable after the fact. */
int hs_hash_audit(hs_ctx_t *ctx, uint64_t *residual_flips);
/* keys_out[N][key_words] (host, word 0 = least significant) of table `table`. */
int hs_get_keys(hs_ctx_t *ctx, uint32_t table, uint64_t *keys_out);
/* Pack a reference key string (digits and '-') the way the device does:
 * right-aligned nibbles '0'..'9' -> 1..10, '-' -> 11.  words_out[key_words]. */
int hs_pack_key_string(const char *s, uint32_t key_words, uint64_t *words_out);

/* ---- index build (B1) ------------------------------------------------------- */
/* Radix sort of (key, id) per table + bucket grouping + bucket-ordered code
 * store.  Replaces the unordered_map insert of motif_both_points.cpp:212-216.
 * Calls hs_hash first if it has not run. */
int hs_build_index(hs_ctx_t *ctx);
/* sizes_out[L] = number of buckets per table ("table size", :217). */
int hs_table_sizes(hs_ctx_t *ctx, uint64_t *sizes_out);
/* ids_out[N] = fragment ids of table `table` in bucket order (ascending id
 * inside a bucket); starts_out[nb+1] bucket boundaries (either may be NULL). */
int hs_get_table(hs_ctx_t *ctx, uint32_t table, uint32_t *ids_out, uint32_t *starts_out);

/* ---- search (V1, V2, V3) ---------------------------------------------------- */
/* Query loop of motif_both_points.cpp:224-245 for Q queries given as dense
 * points [Q][DIM] (centres may be arbitrary real vectors) or as residue codes
 * [Q][len] (embedded with the ctx table).  hits (host, capacity cap) receives
 * *nhits hits; if *nhits > cap the call returns HS_ERR_CAPACITY and the first
 * cap hits are valid only as a set. */
int hs_search_points(hs_ctx_t *ctx, const double *qpoints, uint32_t Q, hs_hit *hits, uint64_t cap,
                     uint64_t *nhits);
int hs_search_codes(hs_ctx_t *ctx, const uint8_t *qcodes, uint32_t Q, hs_hit *hits, uint64_t cap,
                    uint64_t *nhits);
/* Device-resident variants: queries and the hit buffer are device pointers;
 * nothing crosses PCIe except the hit count. */
int hs_search_points_dev(hs_ctx_t *ctx, const void *qpoints_dev, uint32_t Q, void *hits_dev,
                         uint64_t cap, uint64_t *nhits);

/* Compact result layout (12 bytes per hit instead of the 24 of hs_hit): the hit list of
 * motif_both_points.cpp:224-245 as a per-query CSR.  The hits of query q are the entries
 * [offsets[q], offsets[q+1]) of idt / dist2, in the reference's output order (first table,
 * then ascending db id).  idt = local id | table_first << id_bits with local id = db_id -
 * id_base; the query of an entry is implied by its segment.  The caller owns the three
 * arrays (offsets: Q+1 entries; idt, dist2: cap entries); id_bits is set by the call. */
typedef struct {
  uint64_t *offsets; /* [Q+1] */
  uint32_t *idt;     /* [cap] */
  double *dist2;     /* [cap] */
  uint64_t cap;
  uint32_t id_bits;  /* out */
} hs_compact_hits;
/* hs_search_points with the compact result layout (host buffers; HS_FLAG_SORT_HITS is implied).
 * Half the device-to-host bytes of hs_search_points.  HS_ERR_UNSUPPORTED when local id and
 * table do not fit 32 bits together; HS_ERR_CAPACITY as for hs_search_points. */
int hs_search_points_compact(hs_ctx_t *ctx, const double *qpoints, uint32_t Q, hs_compact_hits *out, uint64_t *nhits);
/* Expands a compact result into the hs_hit records hs_search_points returns (host only):
 * hits_out[offsets[Q]] receives exactly the same bytes. */
int hs_expand_hits(const hs_compact_hits *in, uint32_t Q, uint64_t id_base, hs_hit *hits_out);

/* Order-independent 64-bit checksum of a hit list: the sum mod 2^64 over the hits of a mix of
 * (query, table_first, db_id, bit pattern of dist2).  Equal multisets of hits give equal sums
 * whatever their order or their split over ranks (partial sums add), so an N-GPU run is
 * compared with a 1-GPU run by one number.  _dev: the list is in device memory. */
int hs_hits_checksum(const hs_hit *hits, uint64_t n, uint64_t *sum_out);
int hs_hits_checksum_dev(hs_ctx_t *ctx, const void *hits_dev, uint64_t n, uint64_t *sum_out);

/* ---- brute force (G1) ------------------------------------------------------- */
/* All Q x N distances, hits only (motif_both_points_noLSH.cpp:36-56; the
 * non-hit dump :47-49 is not produced).  qcodes == NULL: all pairs i<j of the
 * DB (query = i, db_id = j).  Uses params.metric / predicate / R. */
int hs_bruteforce_codes(hs_ctx_t *ctx, const uint8_t *qcodes, uint32_t Q, hs_hit *hits, uint64_t cap,
                        uint64_t *nhits);
int hs_bruteforce_points(hs_ctx_t *ctx, const double *qpoints, uint32_t Q, hs_hit *hits, uint64_t cap,
                         uint64_t *nhits);
/* Device-resident variant (queries and hit buffer are device pointers); with cap == 0 only the
 * hit count is produced -- the denominator of the recall of an LSH search at full size. */
int hs_bruteforce_points_dev(hs_ctx_t *ctx, const void *qpoints_dev, uint32_t Q, void *hits_dev, uint64_t cap,
                             uint64_t *nhits);

/* ---- recall of a search against ground truth (R1) ---------------------------- */
/* evaulate() + weight() of motif_both_points.cpp:66-86,100-165 on binary hit lists instead
 * of the two text files: `truth` (any order; e.g. the output of hs_bruteforce_*) is joined
 * with `found`, which must be in the order hs_search_* returns with HS_FLAG_SORT_HITS
 * (query, first table, db id; pairs unique).  A truth pair that was found adds
 * weight(dis) to tp, a missed one to fn, dis = sqrt(dist2) (the integer distance for
 * BLOSUM_INT); recall = tp / (tp + fn).  tp_bin / fn_bin count the pairs by
 * int(dis*100/10), the rows of <out>.accuracy.txt (:151-163); n_extra = found pairs absent
 * from truth (the "xnomo" lines, :127-129).  A truth distance above R + 0.1 makes the
 * reference print "err" and exit (:67-70): here HS_ERR_INVALID.  Counts are exact; tp and
 * fn are summed in a fixed parallel order, within 1e-12 relative of the sequential sums. */
#define HS_RECALL_BINS 500
typedef struct {
  double tp, fn;
  uint64_t n_tp, n_fn, n_extra;
  uint64_t tp_bin[HS_RECALL_BINS], fn_bin[HS_RECALL_BINS];
} hs_recall;
int hs_evaluate_recall(hs_ctx_t *ctx, const hs_hit *truth, uint64_t n_truth, const hs_hit *found, uint64_t n_found,
                       uint32_t Q, hs_recall *out);
/* Device-resident variant: both lists are device pointers (recall at 10^8 .. 10^9 fragments
 * without a round trip through text or the host). */
int hs_evaluate_recall_dev(hs_ctx_t *ctx, const void *truth_dev, uint64_t n_truth, const void *found_dev,
                           uint64_t n_found, uint32_t Q, hs_recall *out);

/* ---- cluster (U1) ----------------------------------------------------------- */
/* Connected components of the near-pair graph: every pair sharing a bucket in
 * some table with distance within R is an edge; UnionFind
 * (pcluster/src/pcluster/union_find.cpp:3-33) over the edges.  label_out[N]
 * (host) = smallest local id of the fragment's component. */
int hs_cluster(hs_ctx_t *ctx, uint32_t *label_out);

/* UnionFind (pcluster/src/pcluster/union_find.cpp:3-33) over an explicit edge list: ids
 * 0..n-1, edges (eu[i], ev[i]) (host arrays).  label_out[n] = smallest id of each id's
 * component.  Used to merge the per-shard near-pair edges of a sharded cluster run. */
int hs_union_find(hs_ctx_t *ctx, uint32_t n, const uint32_t *eu, const uint32_t *ev, uint64_t ne, uint32_t *label_out);

/* Greedy centre clustering, Clustering() of hclust2.cpp:86-151 (= hclust3.cpp:87-152):
 * L rounds, round l over the buckets of table l; inside a bucket, in member (id) order,
 * every unprocessed fragment joins the first centre within R (sqrt predicate) or becomes
 * a candidate centre.  center_out[N]: the centre holding each fragment (itself when it
 * heads a cluster or stayed alone); round_out[N] (optional): the round in which it joined
 * (0xffffffff for cluster heads) -- a cluster's member list in the reference's order is
 * its head, then its members by (round, id); state_out[N] (optional): the reference's
 * merged[] flags (0 unprocessed, 1 centre, 2 joined).  Euclidean metric only. */
int hs_greedy_cluster(hs_ctx_t *ctx, uint32_t *center_out, uint32_t *round_out, uint8_t *state_out);

/* ---- sequence front ends (E4, E5, E6, KL1) ------------------------------------ */
/* ProteinDB::ReadFASTAFile (pcluster/src/pcluster/read_proteins.cpp:6-41) over a FASTA text
 * held in memory (host only).  residues[res_cap] receives the kept letters of all sequences
 * back to back, start[start_cap] the nseq+1 boundaries, name_begin/name_len[name_cap] the
 * byte range of each header's name inside `text` (header up to the first space).  The
 * reference's quirks are kept: a name per header line but a sequence only when non-empty
 * (*nnames may exceed *nseq); non-amino-acid letters become AA20[rand() % 20].  Any output
 * buffer may be NULL / too small: the counts are still returned, with HS_ERR_CAPACITY. */
int hs_parse_fasta(const char *text, uint64_t nbytes, char *residues, uint64_t res_cap, uint64_t *start,
                   uint64_t start_cap, uint64_t *name_begin, uint32_t *name_len, uint64_t name_cap, uint32_t *nseq,
                   uint32_t *nnames, uint64_t *nres);
/* The same parse on the device (fasta.cu): the text is copied to the GPU, two byte-parallel passes
 * classify every byte (header line or sequence line from the start of its line), count and place
 * the kept letters, the replaced letters and the header names; the replacement letters are drawn
 * with rand() on the host, one per replaced letter in file order as the reference does, and
 * patched in.  Outputs, counts, status and the state of rand() afterwards are those of
 * hs_parse_fasta; texts of 4 GB and more are refused (HS_ERR_UNSUPPORTED).  The context's loaded
 * database is not touched. */
int hs_parse_fasta_gpu(hs_ctx_t *ctx, const char *text, uint64_t nbytes, char *residues, uint64_t res_cap,
                       uint64_t *start, uint64_t start_cap, uint64_t *name_begin, uint32_t *name_len,
                       uint64_t name_cap, uint32_t *nseq, uint32_t *nnames, uint64_t *nres);
/* KLSH::KLSH (pcluster/src/pcluster/lsh.cpp:17-38): w[bits][feat] ~ N(0, sigma^2 as the
 * STANDARD DEVIATION), t[bits] ~ U[-1,1), b[bits] ~ U[0,2pi) from a default-constructed
 * std::default_random_engine (fixed seed: identical in every instance).  Host, libstdc++. */
int hs_klsh_generate(uint32_t feat, uint32_t bits, double sigma, double *w, double *t, double *b);
/* PreClustering (pcluster/src/pcluster/pcluster.cpp:11-35): per protein the 512-bin histogram
 * of reduced-alphabet 3-mers (Kmer2Integer, util.hpp:244-250) and its KLSH value
 * (GetHashValue, lsh.cpp:40-49; bit i = [cos(w_i.p + b_i) + t_i >= 0]).  residues: letters of
 * all proteins back to back, start[nprot+1].  feat_out[nprot][512] and valid_out[nprot]
 * (0: shorter than 3 residues, skipped by the reference) are optional.  Bits whose
 * cos(..)+t lies within 1e-9 of zero are recomputed on the host with libm's cos (the
 * function the reference calls); *n_host_fixed counts those proteins. */
int hs_kmer3_klsh(hs_ctx_t *ctx, const char *residues, const uint64_t *start, uint32_t nprot, const double *w,
                  const double *t, const double *b, uint32_t bits, uint32_t *feat_out, uint64_t *hash_out,
                  uint8_t *valid_out, uint64_t *n_host_fixed);
/* ORF::orf6 (orf/orf.cc:39-74, genetic code orf/orf.h:28-31): the six reading frames of
 * every DNA sequence (0-2 forward, 3-5 reverse complement), each translated codon by codon
 * up to its first stop.  Frame f of sequence s is written at
 * aa_out + 2*start[s] + 6*s + f*(len_s/3 + 1); aa_len[nseq][6] = residues written (the
 * frame is an ORF of the reference when >= 6, orf.cc:58).  aa_cap >= 2*start[nseq] + 6*nseq.
 * Letters other than A, C, G, T -> HS_ERR_INVALID (ERROR_INFO in the reference). */
int hs_orf6(hs_ctx_t *ctx, const char *dna, const uint64_t *start, uint32_t nseq, char *aa_out, uint64_t aa_cap,
            int32_t *aa_len);

/* ---- multi-GPU --------------------------------------------------
 * One process and one context per GPU; rank r loads the id block [id_base_r, id_base_r + N_r)
 * of the database with id_base ascending in rank order.  Nothing of this exists in the
 * reference (a single-threaded process); what it must reproduce is the output order of
 * motif_both_points.cpp:224-245 over the whole database. */
/* Join an NCCL communicator (nccl_unique_id: the 128-byte ncclUniqueId made by rank 0 with
 * hs_comm_unique_id and handed to the other processes by the launcher).  After this,
 * hs_search_* / hs_bruteforce_* on every rank take the queries of rank 0 (ncclBroadcast). */
int hs_comm_init(hs_ctx_t *ctx, const void *nccl_unique_id, int rank, int nranks);
int hs_comm_unique_id(void *out128);
/* Collective.  Rank 0 allocates two receive buffers of cap_hits hs_hit records (the largest
 * request of all ranks) and every rank maps them into its address space (CUDA IPC over NVLink).
 * From then on every search with a non-empty hit buffer, besides returning the rank's own hits
 * as usual, merges the lists of all ranks into rank 0's current receive buffer in the reference's
 * order (query, first table, ascending db id): the ranks exchange their hit counts per (query,
 * table) and each writes its hits straight to their final positions in rank 0's memory.  The
 * merge runs on its own stream and overlaps whatever the ranks do next (the next batch's hash and
 * index build); a search reuses the buffer of the search before the previous one. */
int hs_comm_reserve(hs_ctx_t *ctx, uint64_t cap_hits);
/* Completes the merge started by the rank's latest search.  *nhits_total = hits of all ranks;
 * on rank 0 *hits_dev points at the merged list in device memory (valid until the search after
 * the next); NULL on the other ranks.  HS_ERR_CAPACITY when the receive buffer or some rank's own
 * hit buffer was too small (then *nhits_total is the size hs_comm_reserve needs). */
int hs_comm_result(hs_ctx_t *ctx, const void **hits_dev, uint64_t *nhits_total);

#ifdef __cplusplus
}
#endif
#endif /* HSEARCH_B200_H */
