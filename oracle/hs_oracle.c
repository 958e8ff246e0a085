/*
 * hs_oracle.c -- CPU restatement of the HSEARCH hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  Nothing under hsearch_b200/ links, imports or executes it.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference).  The restatement is pinned against (i) the known-answer
 * vectors recorded in SURVEY.md section 8c and (ii) outputs of the reference's
 * own sources compiled in place by oracle/Makefile into oracle/_ref/ (see
 * oracle/ref_harness.cpp and tests/golden/make_golden.py).
 *
 * Third-party arithmetic restated here: GNU libstdc++ 13.3 <random>
 * (minstd_rand0, generate_canonical<double,53>, polar normal_distribution,
 * uniform_real_distribution), /usr/include/c++/13/bits/random.h and
 * random.tcc:1804-1846, 3349-3381.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_AA 20
#define ORC_CDIM 8 /* AACoordinateSize, hclust/src/hclust/util.hpp:94 */

/* ---- tables ---------------------------------------------------------------
 * M1: hclust/src/hclust/util.hpp:21-42 (coordinates[20][8], BLOSUM order
 * A R N D C Q E G H I L K M F P S T W Y V).  Frozen constants (MDS output of
 * IGC/distance2coordinate/BLOSUM.m); data, not code. */
static const double orc_coordinates[ORC_AA][ORC_CDIM] = {
    {-0.876280, 3.598596, 2.554616, -0.729216, 0.698828, 1.221507, -2.765205, -3.163091},
    {-4.111404, -1.936791, -2.682295, 0.942498, 6.924314, -1.195785, -1.639269, 0.615381},
    {-7.471612, -2.468058, 0.932738, -4.488355, 0.553080, -3.081577, 0.368010, 4.223792},
    {-8.317871, -0.848602, 1.752372, -1.407818, -4.874022, -1.493568, 5.256411, -2.561758},
    {5.421664, 11.791877, 2.675596, -5.622478, 4.322457, 3.946839, 2.229597, -1.901479},
    {-3.771796, -2.525005, -1.567736, 2.619391, 2.781873, 0.952486, 3.947072, -0.954304},
    {-6.585010, -2.752755, -1.649014, 1.605597, -1.833933, -0.730211, 2.313328, -3.239486},
    {-3.978253, -1.155062, 9.994796, -0.195264, -1.110059, -2.860194, -4.952672, -1.495210},
    {-2.630176, -8.283034, -4.773107, -6.479084, 0.070359, 4.318067, -1.847373, -0.086451},
    {4.548022, 5.189698, -3.999001, -0.186966, -3.275059, -1.882387, -0.627095, 0.049364},
    {5.341899, 4.436639, -3.552811, 1.250614, 0.266899, -2.609335, -0.694939, 0.812004},
    {-5.742562, -1.207887, -2.587323, 2.866228, 4.169821, -1.991698, -1.941954, -0.747156},
    {4.241223, 2.474317, -2.658336, 2.946054, 2.011534, -3.254331, 1.266004, -0.186966},
    {9.340442, -3.359172, -0.635377, -2.878570, -3.255191, -2.200202, -1.104637, -0.062654},
    {-6.150933, 3.182318, 0.122393, 7.788554, -3.094076, 6.831600, -1.992627, 1.807240},
    {-2.523437, 1.824168, 3.256463, -2.386830, 0.439791, 1.024198, 0.486894, 1.190316},
    {-0.823028, 3.115233, 2.075337, -0.585875, -1.471153, 0.518398, 1.846290, 6.269577},
    {13.592409, -8.961858, 6.548108, 4.623650, 2.128797, 0.808588, 2.631353, 0.521535},
    {7.173223, -6.765800, -2.811202, -1.654989, -1.878135, 3.104673, -1.272146, -0.635970},
    {3.323480, 4.651177, -2.996218, 1.972858, -3.576126, -1.427066, -1.507041, -0.454682}};

/* M3 source: pcluster/src/pcluster/util.hpp:109-130 (BLOSUM62, same order). */
static const int orc_blosum62[ORC_AA][ORC_AA] = {
    {4, -1, -2, -2, 0, -1, -1, 0, -2, -1, -1, -1, -1, -2, -1, 1, 0, -3, -2, 0},
    {-1, 5, 0, -2, -3, 1, 0, -2, 0, -3, -2, 2, -1, -3, -2, -1, -1, -3, -2, -3},
    {-2, 0, 6, 1, -3, 0, 0, 0, 1, -3, -3, 0, -2, -3, -2, 1, 0, -4, -2, -3},
    {-2, -2, 1, 6, -3, 0, 2, -1, -1, -3, -4, -1, -3, -3, -1, 0, -1, -4, -3, -3},
    {0, -3, -3, -3, 9, -3, -4, -3, -3, -1, -1, -3, -1, -2, -3, -1, -1, -2, -2, -1},
    {-1, 1, 0, 0, -3, 5, 2, -2, 0, -3, -2, 1, 0, -3, -1, 0, -1, -2, -1, -2},
    {-1, 0, 0, 2, -4, 2, 5, -2, 0, -3, -3, 1, -2, -3, -1, 0, -1, -3, -2, -2},
    {0, -2, 0, -1, -3, -2, -2, 6, -2, -4, -4, -2, -3, -3, -2, 0, -2, -2, -3, -3},
    {-2, 0, 1, -1, -3, 0, 0, -2, 8, -3, -3, -1, -2, -1, -2, -1, -2, -2, 2, -3},
    {-1, -3, -3, -3, -1, -3, -3, -4, -3, 4, 2, -3, 1, 0, -3, -2, -1, -3, -1, 3},
    {-1, -2, -3, -4, -1, -2, -3, -4, -3, 2, 4, -2, 2, 0, -3, -2, -1, -2, -1, 1},
    {-1, 2, 0, -1, -3, 1, 1, -2, -1, -3, -2, 5, -1, -3, -1, 0, -1, -3, -2, -2},
    {-1, -1, -2, -3, -1, 0, -2, -3, -2, 1, 2, -1, 5, 0, -2, -1, -1, -1, -1, 1},
    {-2, -3, -3, -3, -2, -3, -3, -3, -1, 0, 0, -3, 0, 6, -4, -2, -2, 1, 3, -1},
    {-1, -2, -2, -1, -3, -1, -1, -2, -2, -3, -3, -1, -2, -4, 7, -1, -1, -4, -3, -2},
    {1, -1, 1, 0, -1, 0, 0, 0, -1, -2, -2, 0, -1, -2, -1, 4, 1, -3, -2, -2},
    {0, -1, 0, -1, -1, -1, -1, -2, -2, -1, -1, -1, -1, -2, -1, 1, 5, -2, -2, 0},
    {-3, -3, -4, -4, -2, -2, -3, -2, -2, -3, -2, -3, -1, 1, -4, -3, -2, 11, 2, -3},
    {-2, -2, -2, -3, -2, -1, -2, -3, 2, -1, -1, -2, -1, 3, -3, -2, -2, 2, 7, -1},
    {0, -3, -3, -3, -1, -2, -2, -3, -3, 3, 1, -2, 1, -1, -2, -2, 0, -3, -1, 4}};

/* hclust/src/hclust/util.hpp:89,92 */
static const char orc_AA20[] = "ARNDCEQGHILKMFPSTWYV";
static const int orc_base[26] = {0, -1, 4, 3, 6, 13, 7, 8, 9, -1, 11, 10, 12,
                                 2, -1, 14, 5, 1, 15, 16, -1, 19, 17, -1, 18, -1};
/* pcluster/src/pcluster/util.hpp:103-104 */
static const int orc_reduced[26] = {0, -1, 3, 1, 1, 6, 4, 2, 5, -1, 1, 5, 5,
                                    2, -1, 7, 1, 1, 0, 0, -1, 5, 6, -1, 6, -1};

void orc_get_coordinates(double *out160) { memcpy(out160, orc_coordinates, sizeof(orc_coordinates)); }
void orc_get_blosum62(int *out400) { memcpy(out400, orc_blosum62, sizeof(orc_blosum62)); }
void orc_get_base(int *out26) { memcpy(out26, orc_base, sizeof(orc_base)); }
const char *orc_get_aa20(void) { return orc_AA20; }

/* protein2datapoints.cpp:23-29: Point::Output prints with ostream default
 * precision (== printf "%g", 6 significant digits); motif_both_points.cpp:
 * 347-351 reads the text back with operator>> (== strtod).  The table the
 * search really hashes is therefore the print-rounded one. */
void orc_get_coordinates_print6(double *out160) {
  char buf[64];
  for (int i = 0; i < ORC_AA; ++i)
    for (int j = 0; j < ORC_CDIM; ++j) {
      snprintf(buf, sizeof buf, "%g", orc_coordinates[i][j]);
      out160[i * ORC_CDIM + j] = strtod(buf, NULL);
    }
}

/* M3: BLOSUM-Metric/src/BLOSUM-metric/distance_matrix.hpp:13-20 */
void orc_blosum_metric(int *out400) {
  for (int i = 0; i < ORC_AA; ++i)
    for (int j = 0; j < ORC_AA; ++j)
      out400[i * ORC_AA + j] = orc_blosum62[i][i] + orc_blosum62[j][j] - 2 * orc_blosum62[i][j];
}

/* distance_matrix.hpp:36-50: number of triangle-inequality violations */
int orc_triangle_violations(const int *d400) {
  int cnt = 0;
  for (int i = 0; i < 20; i++)
    for (int j = 0; j < 20; j++)
      for (int k = 0; k < 20; k++)
        if (d400[i * 20 + j] + d400[j * 20 + k] < d400[i * 20 + k]) cnt++;
  return cnt;
}

/* ---- libstdc++ <random> restatement --------------------------------------- */
typedef struct {
  uint32_t x; /* minstd_rand0 state */
  int saved_available;
  double saved;
} orc_rng;

/* bits/random.h linear_congruential_engine<uint_fast32_t,16807,0,2147483647>::seed */
static void orc_rng_seed(orc_rng *r, uint64_t s) {
  uint64_t m = 2147483647ULL;
  uint64_t v = s % m;
  r->x = (uint32_t)(v == 0 ? 1 : v);
  r->saved_available = 0;
  r->saved = 0.0;
}
static uint32_t orc_rng_next(orc_rng *r) {
  r->x = (uint32_t)(((uint64_t)r->x * 16807ULL) % 2147483647ULL);
  return r->x;
}
/* random.tcc:3349-3381 with urng = minstd_rand0: r = 2147483646, log2r = 30,
 * m = (53 + 30 - 1) / 30 = 2 draws per double. */
static double orc_canonical(orc_rng *r) {
  const long double R = 2147483646.0L;
  double sum = 0.0, tmp = 1.0;
  for (int k = 2; k != 0; --k) {
    sum += (double)(orc_rng_next(r) - 1u) * tmp;
    tmp = (double)((long double)tmp * R);
  }
  double ret = sum / tmp;
  if (ret >= 1.0) ret = nextafter(1.0, 0.0);
  return ret;
}
/* random.tcc:1804-1846, Marsaglia polar with cached second variate */
static double orc_normal(orc_rng *r, double mean, double stddev) {
  double ret;
  if (r->saved_available) {
    r->saved_available = 0;
    ret = r->saved;
  } else {
    double x, y, r2;
    do {
      x = 2.0 * orc_canonical(r) - 1.0;
      y = 2.0 * orc_canonical(r) - 1.0;
      r2 = x * x + y * y;
    } while (r2 > 1.0 || r2 == 0.0);
    const double mult = sqrt(-2 * log(r2) / r2);
    r->saved = x * mult;
    r->saved_available = 1;
    ret = y * mult;
  }
  return ret * stddev + mean;
}
/* bits/random.h uniform_real_distribution::operator(): aurng()*(b-a)+a */
static double orc_uniform(orc_rng *r, double a, double b) { return orc_canonical(r) * (b - a) + a; }

/* H1: hclust/src/hclust/lsh.hpp:10-31.  One engine per LSH object; per k:
 * DIM normals into a[k][*], then one uniform[0,W) into b[k].  The normal
 * distribution object (with its cached variate) lives across the k loop. */
void orc_lsh_generate(uint64_t seed, uint32_t dim, uint32_t K, double W, double *a, double *b) {
  orc_rng r;
  orc_rng_seed(&r, seed);
  for (uint32_t k = 0; k < K; ++k) {
    for (uint32_t i = 0; i < dim; ++i) a[(size_t)k * dim + i] = orc_normal(&r, 0.0, 1.0);
    b[k] = orc_uniform(&r, 0.0, W);
  }
}

/* KL1 ctor: pcluster/src/pcluster/lsh.cpp:17-38.  `generator` is a
 * default-constructed member => minstd_rand0 seed 1.  stddev = sigma*sigma. */
void orc_klsh_generate(uint32_t feat, uint32_t bits, double sigma, double *w, double *t, double *b) {
  orc_rng r;
  orc_rng_seed(&r, 1);
  /* three distribution objects; only the normal one carries state */
  for (uint32_t i = 0; i < bits; ++i) {
    t[i] = orc_uniform(&r, -1.0, 1.0);
    b[i] = orc_uniform(&r, 0.0, 2.0 * M_PI);
    for (uint32_t j = 0; j < feat; ++j) w[(size_t)i * feat + j] = orc_normal(&r, 0.0, sigma * sigma);
  }
}
/* KL1 hash: pcluster/src/pcluster/lsh.cpp:8-15,40-49 */
uint64_t orc_klsh_hash(const double *p, uint32_t feat, uint32_t bits, const double *w, const double *t,
                       const double *b) {
  uint64_t h = 0;
  for (uint32_t i = 0; i < bits; ++i) {
    double sum = 0;
    for (uint32_t j = 0; j < feat; ++j) sum += p[j] * w[(size_t)i * feat + j];
    sum = sum + b[i];
    h |= (uint64_t)((cos(sum) + t[i]) >= 0 ? 1 : 0) << i;
  }
  return h;
}
/* E5: pcluster/src/pcluster/pcluster.cpp:26-32 + util.hpp:244-250 (HASHLEN 3,
 * BASEP = powers of 8): 512-bin histogram of reduced-alphabet 3-mers. */
void orc_kmer3_features(const char *seq, uint32_t n, double *feat512) {
  for (int i = 0; i < 512; ++i) feat512[i] = 0;
  if (n < 3) return;
  for (uint32_t i = 0; i + 3 <= n; ++i) {
    uint32_t h = 0, p = 1;
    for (int k = 0; k < 3; ++k) {
      h += orc_reduced[seq[i + k] - 'A'] * p;
      p *= 8;
    }
    feat512[h] += 1;
  }
}

/* ---- embedding + hash ------------------------------------------------------ */
/* M1: hclust2.cpp:49-62 / kmer2coordinates.cpp:49-71: concatenate the 8-vector
 * of each residue.  `table` is the 20x8 table to use (full or print6). */
void orc_embed(const uint8_t *codes, uint32_t len, const double *table160, double *point) {
  uint32_t k = 0;
  for (uint32_t i = 0; i < len; ++i)
    for (uint32_t j = 0; j < ORC_CDIM; ++j) point[k++] = table160[codes[i] * ORC_CDIM + j];
}

/* H2: lsh.hpp:33-42.  Sequential, separate multiply and add (no FMA: the
 * reference is built for baseline x86-64 without -mfma). */
#if defined(__GNUC__)
__attribute__((optimize("fp-contract=off")))
#endif
double orc_dot(const double *point, const double *a_row, uint32_t dim) {
  double dot = 0;
  for (uint32_t i = 0; i < dim; ++i) {
    dot += point[i] * a_row[i];
  }
  return dot;
}
/* H3: lsh.hpp:44-49 */
int orc_bucket(const double *point, const double *a_row, double b, double W, uint32_t dim) {
  double val = orc_dot(point, a_row, dim) + b;
  return (int)floor(val / W);
}
/* H4: lsh.hpp:51-59: decimal strings concatenated with no separator.
 * Returns the string length; `out` must hold 12*K+1 bytes. */
int orc_hash_key(const double *point, const double *a, const double *b, uint32_t K, double W,
                 uint32_t dim, char *out, int *buckets_out) {
  int n = 0;
  for (uint32_t k = 0; k < K; ++k) {
    int bk = orc_bucket(point, a + (size_t)k * dim, b[k], W, dim);
    if (buckets_out) buckets_out[k] = bk;
    n += sprintf(out + n, "%d", bk);
  }
  return n;
}

/* bucket ints for N code fragments, L tables: out[N][L][K].
 * a: [L][K][dim], b: [L][K]. */
void orc_hash_codes(const uint8_t *codes, uint64_t N, uint32_t len, const double *table160,
                    const double *a, const double *b, uint32_t K, uint32_t L, double W, int *out) {
  uint32_t dim = len * ORC_CDIM;
  double *pt = (double *)malloc(sizeof(double) * dim);
  for (uint64_t i = 0; i < N; ++i) {
    orc_embed(codes + i * len, len, table160, pt);
    for (uint32_t l = 0; l < L; ++l)
      for (uint32_t k = 0; k < K; ++k)
        out[(i * L + l) * K + k] =
            orc_bucket(pt, a + ((size_t)l * K + k) * dim, b[l * K + k], W, dim);
  }
  free(pt);
}
/* same for dense points [N][dim] */
void orc_hash_points(const double *pts, uint64_t N, uint32_t dim, const double *a, const double *b,
                     uint32_t K, uint32_t L, double W, int *out) {
  for (uint64_t i = 0; i < N; ++i)
    for (uint32_t l = 0; l < L; ++l)
      for (uint32_t k = 0; k < K; ++k)
        out[(i * L + l) * K + k] =
            orc_bucket(pts + i * dim, a + ((size_t)l * K + k) * dim, b[l * K + k], W, dim);
}

/* ---- distances ------------------------------------------------------------- */
/* V2: motif_both_points.cpp:176-183 (no sqrt) / :167-174 (sqrt) */
#if defined(__GNUC__)
__attribute__((optimize("fp-contract=off")))
#endif
double orc_dist2(const double *x, const double *y, uint32_t dim) {
  double dis = 0.0, r = 0.0;
  for (uint32_t i = 0; i < dim; ++i) {
    r = x[i] - y[i];
    dis += r * r;
  }
  return dis;
}
/* V3: BLOSUM-Metric/src/BLOSUM-metric/evaluate_correlation.cpp:34-41 with the
 * second index restated as s2[i]-'A' (the source writes 'B', an off-by-one
 * that reads base[-1] for 'A'; flagged in SURVEY.md 8a row V3).  Operates on
 * codes (= base[c-'A']) directly. */
int orc_distance_int(const uint8_t *x, const uint8_t *y, uint32_t len, const int *d400) {
  int d = 0;
  for (uint32_t i = 0; i < len; ++i) d += d400[x[i] * ORC_AA + y[i]];
  return d;
}
/* evaluate_correlation.cpp:26-32 */
int orc_similarity_int(const uint8_t *x, const uint8_t *y, uint32_t len) {
  int s = 0;
  for (uint32_t i = 0; i < len; ++i) s += orc_blosum62[x[i]][y[i]];
  return s;
}

/* ---- B1 + V1: Search() of motif_both_points.cpp:195-250 -------------------- */
typedef struct {
  uint32_t query;
  uint32_t table_first;
  uint64_t db_id;
  double dist2;
} orc_hit;

typedef struct {
  char s[200];
  uint32_t id;
} orc_keyrec;
static int orc_keyrec_cmp(const void *pa, const void *pb) {
  const orc_keyrec *a = (const orc_keyrec *)pa, *b = (const orc_keyrec *)pb;
  int c = strcmp(a->s, b->s);
  if (c) return c;
  return a->id < b->id ? -1 : (a->id > b->id);
}

/* pred: 0 => d2 <= R*R (motif_both_points.cpp:204,239)
 *       1 => !(sqrt(d2) > R) (motif_both_points_noLSH.cpp:46-47, hclust2.cpp:119-120) */
static int orc_is_hit(double d2, double R, int pred) {
  if (pred == 0) return d2 <= R * R;
  return !(sqrt(d2) > R);
}

/* The unordered_map<string, vector<uint32_t>> of :25,212-216 is restated as a
 * (string, id)-sorted array; bucket member lists keep ascending id order, which
 * is the reference's insertion order.  Hit order = query, table, ascending id,
 * first table wins (label[], :232-238).  Returns total hits (may exceed cap;
 * only the first cap are written).  table_sizes[L] = #buckets (:217). */
uint64_t orc_search(const double *db, uint64_t N, const double *queries, uint32_t Q, uint32_t dim,
                    const double *a, const double *b, uint32_t K, uint32_t L, double W, double R,
                    int pred, orc_hit *hits, uint64_t cap, uint64_t *table_sizes,
                    uint64_t *ncandidates) {
  if (12 * K + 1 > sizeof(((orc_keyrec *)0)->s)) return (uint64_t)-1;
  orc_keyrec **tabs = (orc_keyrec **)malloc(sizeof(orc_keyrec *) * L);
  for (uint32_t l = 0; l < L; ++l) {
    orc_keyrec *t = (orc_keyrec *)malloc(sizeof(orc_keyrec) * (N ? N : 1));
    for (uint64_t i = 0; i < N; ++i) {
      memset(t[i].s, 0, sizeof t[i].s);
      orc_hash_key(db + i * dim, a + (size_t)l * K * dim, b + l * K, K, W, dim, t[i].s, NULL);
      t[i].id = (uint32_t)i;
    }
    qsort(t, N, sizeof(orc_keyrec), orc_keyrec_cmp);
    uint64_t nb = 0;
    for (uint64_t i = 0; i < N; ++i)
      if (i == 0 || strcmp(t[i].s, t[i - 1].s) != 0) nb++;
    if (table_sizes) table_sizes[l] = nb;
    tabs[l] = t;
  }
  uint32_t *label = (uint32_t *)calloc(N ? N : 1, sizeof(uint32_t)); /* epoch-stamped label[] */
  uint64_t nh = 0, ncand = 0;
  char key[200];
  for (uint32_t q = 0; q < Q; ++q) {
    for (uint32_t l = 0; l < L; ++l) {
      memset(key, 0, sizeof key);
      orc_hash_key(queries + (size_t)q * dim, a + (size_t)l * K * dim, b + l * K, K, W, dim, key, NULL);
      /* lower bound */
      uint64_t lo = 0, hi = N;
      while (lo < hi) {
        uint64_t m = (lo + hi) / 2;
        if (strcmp(tabs[l][m].s, key) < 0) lo = m + 1; else hi = m;
      }
      for (uint64_t j = lo; j < N && strcmp(tabs[l][j].s, key) == 0; ++j) {
        uint32_t id = tabs[l][j].id;
        if (label[id] == q + 1) continue;
        double d2 = orc_dist2(db + (size_t)id * dim, queries + (size_t)q * dim, dim);
        label[id] = q + 1;
        ncand++;
        if (orc_is_hit(d2, R, pred)) {
          if (nh < cap) {
            hits[nh].query = q;
            hits[nh].table_first = l;
            hits[nh].db_id = id;
            hits[nh].dist2 = d2;
          }
          nh++;
        }
      }
    }
  }
  if (ncandidates) *ncandidates = ncand;
  for (uint32_t l = 0; l < L; ++l) free(tabs[l]);
  free(tabs);
  free(label);
  return nh;
}

/* G1: motif_both_points_noLSH.cpp:36-56 (hits only; the non-hit dump is I/O).
 * Hit order = query-major, ascending db id.  pred as above (reference: 1). */
uint64_t orc_bruteforce(const double *db, uint64_t N, const double *queries, uint32_t Q, uint32_t dim,
                        double R, int pred, orc_hit *hits, uint64_t cap) {
  uint64_t nh = 0;
  for (uint32_t q = 0; q < Q; ++q)
    for (uint64_t j = 0; j < N; ++j) {
      double d2 = orc_dist2(db + j * dim, queries + (size_t)q * dim, dim);
      if (orc_is_hit(d2, R, pred)) {
        if (nh < cap) {
          hits[nh].query = q;
          hits[nh].table_first = 0;
          hits[nh].db_id = j;
          hits[nh].dist2 = d2;
        }
        nh++;
      }
    }
  return nh;
}
/* G1 with the integer metric V3: hit iff DistanceScore <= R.  dist2 field
 * carries the integer distance.  qcodes == NULL => all pairs i<j of the DB
 * (query = i, db_id = j). */
uint64_t orc_bruteforce_int(const uint8_t *db, uint64_t N, const uint8_t *qcodes, uint32_t Q,
                            uint32_t len, int R, orc_hit *hits, uint64_t cap) {
  int d400[400];
  orc_blosum_metric(d400);
  uint64_t nh = 0;
  uint64_t nq = qcodes ? Q : N;
  for (uint64_t q = 0; q < nq; ++q) {
    const uint8_t *x = qcodes ? qcodes + q * len : db + q * len;
    for (uint64_t j = qcodes ? 0 : q + 1; j < N; ++j) {
      int d = orc_distance_int(x, db + j * len, len, d400);
      if (d <= R) {
        if (nh < cap) {
          hits[nh].query = (uint32_t)q;
          hits[nh].table_first = 0;
          hits[nh].db_id = j;
          hits[nh].dist2 = (double)d;
        }
        nh++;
      }
    }
  }
  return nh;
}

/* ---- U1: pcluster/src/pcluster/union_find.cpp:3-33 ------------------------- */
/* root map restated as a dense array over ids 0..n-1 (ids are dense here). */
static uint32_t orc_find_root(uint32_t *root, uint32_t *px) {
  uint32_t x = *px;
  uint32_t t = x;
  while (t != root[t]) t = root[t];
  while (x != root[x]) {
    uint32_t tmp = root[x];
    root[x] = t;
    x = tmp;
  }
  *px = x; /* FindRoot leaves its by-reference argument at the root (:22-26) */
  return t;
}
/* Feed edges with the FindRoot-then-JoinUnion protocol (JoinUnion: root[x] =
 * root[y], :31-33) and canonicalise the partition as min id per component. */
void orc_union_find_labels(uint32_t n, const uint32_t *eu, const uint32_t *ev, uint64_t ne,
                           uint32_t *label_out) {
  uint32_t *root = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
  for (uint32_t i = 0; i < n; ++i) root[i] = i;
  for (uint64_t e = 0; e < ne; ++e) {
    uint32_t x = eu[e], y = ev[e];
    orc_find_root(root, &x);
    orc_find_root(root, &y);
    root[x] = root[y];
  }
  uint32_t *mn = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
  for (uint32_t i = 0; i < n; ++i) mn[i] = 0xffffffffu;
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t x = i;
    uint32_t r = orc_find_root(root, &x);
    if (i < mn[r]) mn[r] = i;
  }
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t x = i;
    label_out[i] = mn[orc_find_root(root, &x)];
  }
  free(root);
  free(mn);
}

/* Cluster oracle (SURVEY.md 8c "composition"): reference LSH tables over the
 * fragments -> every in-bucket pair within R -> UnionFind -> min-id labels.
 * metric 0: Euclidean with sqrt(d2) <= R (hclust2.cpp:64-71,119-120) over the
 * embedded points; metric 1: integer DistanceScore <= R.  Returns #edges. */
uint64_t orc_cluster(const uint8_t *codes, uint64_t N, uint32_t len, const double *table160,
                     const double *a, const double *b, uint32_t K, uint32_t L, double W, double R,
                     int metric, uint32_t *label_out) {
  uint32_t dim = len * ORC_CDIM;
  double *pts = (double *)malloc(sizeof(double) * dim * (N ? N : 1));
  for (uint64_t i = 0; i < N; ++i) orc_embed(codes + i * len, len, table160, pts + i * dim);
  int d400[400];
  orc_blosum_metric(d400);
  uint64_t cap = 1024, ne = 0;
  uint32_t *eu = (uint32_t *)malloc(sizeof(uint32_t) * cap), *ev = (uint32_t *)malloc(sizeof(uint32_t) * cap);
  orc_keyrec *t = (orc_keyrec *)malloc(sizeof(orc_keyrec) * (N ? N : 1));
  for (uint32_t l = 0; l < L; ++l) {
    for (uint64_t i = 0; i < N; ++i) {
      memset(t[i].s, 0, sizeof t[i].s);
      orc_hash_key(pts + i * dim, a + (size_t)l * K * dim, b + l * K, K, W, dim, t[i].s, NULL);
      t[i].id = (uint32_t)i;
    }
    qsort(t, N, sizeof(orc_keyrec), orc_keyrec_cmp);
    uint64_t s = 0;
    while (s < N) {
      uint64_t e = s + 1;
      while (e < N && strcmp(t[e].s, t[s].s) == 0) e++;
      for (uint64_t i = s; i < e; ++i)
        for (uint64_t j = i + 1; j < e; ++j) {
          uint32_t u = t[i].id, v = t[j].id;
          int near;
          if (metric == 0) near = !(sqrt(orc_dist2(pts + (size_t)u * dim, pts + (size_t)v * dim, dim)) > R);
          else near = orc_distance_int(codes + (size_t)u * len, codes + (size_t)v * len, len, d400) <= (int)R;
          if (near) {
            if (ne == cap) {
              cap *= 2;
              eu = (uint32_t *)realloc(eu, sizeof(uint32_t) * cap);
              ev = (uint32_t *)realloc(ev, sizeof(uint32_t) * cap);
            }
            eu[ne] = u;
            ev[ne] = v;
            ne++;
          }
        }
      s = e;
    }
  }
  orc_union_find_labels((uint32_t)N, eu, ev, ne, label_out);
  free(eu); free(ev); free(t); free(pts);
  return ne;
}

/* ---- E1/E2: sequence store + stride-1 windows ------------------------------ */
/* E1: hclust/src/hclust/protein.hpp:58-64: letter -> base[] index, stored back
 * as AA20[idx] (E<->Q swap because AA20 is not in base[] order).  Returns the
 * code that re-reading the stored letter through base[] yields, i.e. the row
 * of `coordinates` the downstream embed uses (protein2datapoints.cpp:56). */
int orc_proteindb_code(char letter) {
  int AA = orc_base[letter - 'A'];
  if (AA < 0) return -1; /* reference: rand()%20, nondeterministic */
  char stored = orc_AA20[AA];
  return orc_base[stored - 'A'];
}
char orc_proteindb_stored_letter(char letter) {
  int AA = orc_base[letter - 'A'];
  return AA < 0 ? '?' : orc_AA20[AA];
}
/* E2: kmer_search.cpp:64-83 window enumeration (j in [0, len_i - L]) with the
 * :73 bug fixed (advance inside the window) and proteins shorter than the
 * window skipped (the source underflows an unsigned there, :70).  residues are
 * codes; out_codes [nfrag][L]; out_pos = global start position (:71,79).
 * Returns number of fragments. */
uint64_t orc_extract_windows(const uint8_t *residues, const uint32_t *start_index, uint32_t nprot,
                             uint32_t L, uint32_t stride, uint8_t *out_codes, uint32_t *out_pos) {
  uint64_t n = 0;
  for (uint32_t i = 0; i < nprot; ++i) {
    uint32_t plen = start_index[i + 1] - start_index[i];
    if (plen < L) continue;
    for (uint32_t j = 0; j + L <= plen; j += stride) {
      uint32_t pos = start_index[i] + j;
      if (out_codes) memcpy(out_codes + n * L, residues + pos, L);
      if (out_pos) out_pos[n] = pos;
      n++;
    }
  }
  return n;
}
/* protein.hpp:28-39 */
uint32_t orc_protein_id(const uint32_t *start_index, uint32_t nstart, uint32_t pos) {
  uint32_t l = 0, h = nstart - 1;
  while (l < h) {
    uint32_t m = (l + h + 1) / 2;
    if (pos >= start_index[m]) l = m; else h = m - 1;
  }
  return l;
}

/* ---- E6: orf/orf.cc:39-74, code table orf/orf.h:28-31 ----------------------- */
static const char orc_Base1[] = "TTTTTTTTTTTTTTTTCCCCCCCCCCCCCCCCAAAAAAAAAAAAAAAAGGGGGGGGGGGGGGGG";
static const char orc_Base2[] = "TTTTCCCCAAAAGGGGTTTTCCCCAAAAGGGGTTTTCCCCAAAAGGGGTTTTCCCCAAAAGGGG";
static const char orc_Base3[] = "TCAGTCAGTCAGTCAGTCAGTCAGTCAGTCAGTCAGTCAGTCAGTCAGTCAGTCAGTCAGTCAG";
static const char orc_AAs[] = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
static char orc_codon(const char *c) {
  for (int i = 0; i < 64; ++i)
    if (orc_Base1[i] == c[0] && orc_Base2[i] == c[1] && orc_Base3[i] == c[2]) return orc_AAs[i];
  return 0; /* std::map operator[] default for unknown codons (orf.cc:49) */
}
/* Writes up to 6 NUL-terminated strings of stride (n/3+2) into out; kept[f]=1
 * if frame f (0-2 forward, 3-5 reverse complement) produced >= 6 aa.  Returns
 * the number kept.  A codon not in the table maps to '\0', which std::string
 * appends as a byte; restated by stopping the C string there is NOT done:
 * inputs are restricted to ACGT (non-ACGT is an ERROR_INFO in orf.cc:26). */
int orc_orf6(const char *dna, int n, char *out, int *kept) {
  int stride = n / 3 + 2, nk = 0;
  char *rev = (char *)malloc(n + 1);
  for (int i = 0; i < n; ++i) {
    char c = dna[n - i - 1];
    rev[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
  }
  rev[n] = 0;
  int len = n - 3;
  for (int f = 0; f < 6; ++f) {
    const char *s = f < 3 ? dna : rev;
    int st = f % 3, m = 0;
    char *o = out + (size_t)f * stride;
    for (int i = st; i <= len; i += 3) {
      char aa = orc_codon(s + i);
      if (aa == '*') break;
      o[m++] = aa;
    }
    o[m] = 0;
    kept[f] = m >= 6;
    nk += kept[f];
  }
  free(rev);
  return nk;
}

/* ---- R1: motif_both_points.cpp:67-87 weight() ------------------------------ */
double orc_weight(double dis, double R) {
  (void)R; /* the dis > R + 0.1 branch exit(0)s in the reference (:68-71) */
  if (dis < 0.0000001) return 1;
  if (dis < 24) return 1;
  double w = 1 / (dis - 24);
  if (w > 1) return 1;
  if (w < 0) return 1;
  return w;
}

/* ---- R1: motif_both_points.cpp:100-165 evaulate() --------------------------- */
/* The merge-join of the ground truth with the search output, restated on binary lists that
 * are both sorted by (query, db id) -- the reference sorts by (motif name, protein name);
 * any common total order gives the same matched set.  dis[] are the distances as the
 * reference reads them from the third text column.  Sequential FP64 sums in list order
 * (:117-141); bins by int(dis*100/10) (:120,131,138); n_extra counts the "xnomo" lines
 * (:127-129).  Returns -1 where the reference would print "err" and exit (:67-70). */
int orc_evaluate(const orc_hit *truth, const double *tdis, uint64_t nt, const orc_hit *found, uint64_t nf,
                 double R, uint32_t nbins, double *tp_out, double *fn_out, uint64_t *n_tp, uint64_t *n_fn,
                 uint64_t *n_extra, uint64_t *tp_bin, uint64_t *fn_bin) {
  uint64_t i = 0, j = 0;
  double tp = 0.0, fn = 0.0;
  *n_tp = *n_fn = *n_extra = 0;
  for (uint32_t b = 0; b < nbins; ++b) tp_bin[b] = fn_bin[b] = 0;
  for (uint64_t t = 0; t < nt; ++t)
    if (tdis[t] > R + 0.1) return -1;
  while (i < nt && j < nf) {
    int cmp;
    if (truth[i].query != found[j].query) cmp = truth[i].query > found[j].query ? 1 : -1;
    else if (truth[i].db_id != found[j].db_id) cmp = truth[i].db_id > found[j].db_id ? 1 : -1;
    else cmp = 0;
    if (cmp == 0) {
      tp += orc_weight(tdis[i], R);
      int b = (int)(tdis[i] * 100 / 10);
      if (b >= 0 && (uint32_t)b < nbins) tp_bin[b]++;
      ++*n_tp;
      i++;
      j++;
    } else if (cmp == 1) {
      ++*n_extra;
      j++;
    } else {
      fn += orc_weight(tdis[i], R);
      int b = (int)(tdis[i] * 100 / 10);
      if (b >= 0 && (uint32_t)b < nbins) fn_bin[b]++;
      ++*n_fn;
      i++;
    }
  }
  while (i < nt) {
    fn += orc_weight(tdis[i], R);
    int b = (int)(tdis[i] * 100 / 10);
    if (b >= 0 && (uint32_t)b < nbins) fn_bin[b]++;
    ++*n_fn;
    i++;
  }
  *n_extra += nf - j;
  *tp_out = tp;
  *fn_out = fn;
  return 0;
}
