// Host-side constants and small host functions of the C ABI: embedding tables
// (M1-M3), letter maps (E1), projection generation (H1), key-string packing (H4).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <random>

#include "../../include/hsearch_b200.h"
#include "host_tables.h"

namespace hs {

// MDS embedding of the BLOSUM62-derived metric; frozen constants of
// hclust/src/hclust/util.hpp:21-42 (rows in BLOSUM order ARNDCQEGHILKMFPSTWYV).
const double kCoordinates[HS_AA][HS_CDIM] = {
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

// BLOSUM62, pcluster/src/pcluster/util.hpp:109-130 (same residue order).
const int kBlosum62[HS_AA][HS_AA] = {
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

// util.hpp:89 (storage alphabet, E and Q swapped relative to the table order)
const char kAA20[21] = "ARNDCEQGHILKMFPSTWYV";
// the table order itself (what code i means)
const char kCodeLetters[21] = "ARNDCQEGHILKMFPSTWYV";
// util.hpp:92
const int kBase[26] = {0, -1, 4, 3, 6, 13, 7, 8, 9, -1, 11, 10, 12,
                       2, -1, 14, 5, 1, 15, 16, -1, 19, 17, -1, 18, -1};
// pcluster/src/pcluster/util.hpp:103-104
const int kReduced[26] = {0, -1, 3, 1, 1, 6, 4, 2, 5, -1, 1, 5, 5,
                          2, -1, 7, 1, 1, 0, 0, -1, 5, 6, -1, 6, -1};

static thread_local char g_err[1024] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
const char *get_error() { return g_err; }

void coordinates_table(uint32_t variant, double *out160) {
  for (int i = 0; i < HS_AA; ++i)
    for (int j = 0; j < HS_CDIM; ++j) {
      double v = kCoordinates[i][j];
      if (variant == HS_TABLE_PRINT6) {
        // ostream default formatting == "%g" with precision 6, then operator>>
        char buf[64];
        snprintf(buf, sizeof buf, "%g", v);
        v = strtod(buf, nullptr);
      }
      out160[i * HS_CDIM + j] = v;
    }
}

void blosum_metric(int32_t *out400) {
  for (int i = 0; i < HS_AA; ++i)
    for (int j = 0; j < HS_AA; ++j)
      out400[i * HS_AA + j] = kBlosum62[i][i] + kBlosum62[j][j] - 2 * kBlosum62[i][j];
}

// A contracting 8-dimensional embedding of the integer BLOSUM metric D (BLOSUM-Metric/src/
// BLOSUM-metric/distance_matrix.hpp:13-20): points e(a) with |e(a) - e(b)|^2 <= D[a][b] for all
// residues, so that for two fragments sum_p |e(x_p) - e(y_p)|^2 <= sum_p D[x_p][y_p], the window
// distance of evaluate_correlation.cpp:26-41.  The pipelined tensor filter (filter_mma.cu) then
// serves the integer metric unchanged: it keeps every pair whose embedded squared distance is
// <= R, a superset of the pairs with integer distance <= R, and the exact stage decides on the
// integers.  Classical multidimensional scaling of sqrt(D) (Jacobi eigen-decomposition of the
// doubly centred matrix, eight largest eigenvalues), scaled down until every pair contracts; the
// result is checked pair by pair and the function reports failure rather than return a table
// that does not contract.  Tightness only affects the number of survivors (5e-6 of random pairs
// at length 10, R = 30), never the result.
bool blosum_filter_embedding(double *out160) {
  constexpr int n = HS_AA;
  int32_t D[n * n];
  blosum_metric(D);
  double B[n][n], V[n][n], rm[n], gm = 0.0;
  for (int i = 0; i < n; ++i) {
    rm[i] = 0.0;
    for (int j = 0; j < n; ++j) rm[i] += (double)D[i * n + j];
    rm[i] /= n;
    gm += rm[i];
  }
  gm /= n;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      // symmetrised: the table is symmetric, this guards the decomposition only
      const double d = 0.5 * ((double)D[i * n + j] + (double)D[j * n + i]);
      B[i][j] = -0.5 * (d - rm[i] - rm[j] + gm);
      V[i][j] = i == j ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) off += B[p][q] * B[p][q];
    if (off < 1e-22) break;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        if (fabs(B[p][q]) < 1e-300) continue;
        const double theta = (B[q][q] - B[p][p]) / (2.0 * B[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < n; ++k) {
          const double bkp = B[k][p], bkq = B[k][q];
          B[k][p] = c * bkp - sn * bkq;
          B[k][q] = sn * bkp + c * bkq;
        }
        for (int k = 0; k < n; ++k) {
          const double bpk = B[p][k], bqk = B[q][k];
          B[p][k] = c * bpk - sn * bqk;
          B[q][k] = sn * bpk + c * bqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - sn * vkq;
          V[k][q] = sn * vkp + c * vkq;
        }
      }
  }
  int order[n];
  for (int i = 0; i < n; ++i) order[i] = i;
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j)
      if (B[order[j]][order[j]] > B[order[i]][order[i]]) {
        const int t = order[i];
        order[i] = order[j];
        order[j] = t;
      }
  for (int a = 0; a < n; ++a)
    for (int j = 0; j < HS_CDIM; ++j) {
      const double lam = B[order[j]][order[j]];
      out160[a * HS_CDIM + j] = lam > 0.0 ? V[a][order[j]] * sqrt(lam) : 0.0;
    }
  auto d2 = [&](int a, int b) {
    double s = 0.0;
    for (int j = 0; j < HS_CDIM; ++j) {
      const double r = out160[a * HS_CDIM + j] - out160[b * HS_CDIM + j];
      s += r * r;
    }
    return s;
  };
  double s2 = 1e300;
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b) {
      if (a == b) continue;
      if (D[a * n + b] < 0) return false;
      const double e = d2(a, b);
      if (e > 0.0) s2 = fmin(s2, (double)D[a * n + b] / e);
    }
  if (!(s2 > 0.0) || !(s2 < 1e300)) return false;
  const double sc = sqrt(s2) * (1.0 - 1e-9);
  for (int i = 0; i < n * HS_CDIM; ++i) out160[i] *= sc;
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b)
      if (!(d2(a, b) <= (double)D[a * n + b] * (1.0 - 1e-10) + 0.0)) return false;
  return true;
}


}  // namespace hs

extern "C" {

const char *hs_last_error(void) { return hs::get_error(); }

int hs_get_coordinates(uint32_t table_variant, double *out160) {
  if (!out160 || table_variant > HS_TABLE_PRINT6) {
    hs::set_error("hs_get_coordinates: bad argument");
    return HS_ERR_INVALID;
  }
  hs::coordinates_table(table_variant, out160);
  return HS_OK;
}

int hs_get_blosum_metric(int32_t *out400) {
  if (!out400) return HS_ERR_INVALID;
  hs::blosum_metric(out400);
  return HS_OK;
}

int hs_get_blosum_filter_embedding(double *out160) {
  if (!out160) return HS_ERR_INVALID;
  if (!hs::blosum_filter_embedding(out160)) {
    hs::set_error("hs_get_blosum_filter_embedding: no contracting embedding of the metric was found");
    return HS_ERR_UNSUPPORTED;
  }
  return HS_OK;
}

int hs_letter_to_code(char letter) {
  if (letter >= 'a' && letter <= 'z') letter = (char)(letter - 'a' + 'A');
  if (letter < 'A' || letter > 'Z') return -1;
  return hs::kBase[letter - 'A'];
}

int hs_proteindb_code(char letter) {
  int aa = hs_letter_to_code(letter);
  if (aa < 0) return -1;
  return hs::kBase[hs::kAA20[aa] - 'A'];
}

int hs_generate_projection(uint64_t seed, uint32_t dim, uint32_t K, double W, double *a, double *b) {
  if (!a || !b || dim == 0 || K == 0) {
    hs::set_error("hs_generate_projection: bad argument");
    return HS_ERR_INVALID;
  }
  // Same objects, same call order as the reference constructor; the
  // distributions keep their state across the k loop.
  std::normal_distribution<double> normal(0.0, 1.0);
  std::uniform_real_distribution<double> uniform_width(0, W);
  std::default_random_engine generator((std::default_random_engine::result_type)seed);
  for (uint32_t k = 0; k < K; ++k) {
    for (uint32_t i = 0; i < dim; ++i) a[(size_t)k * dim + i] = normal(generator);
    b[k] = uniform_width(generator);
  }
  return HS_OK;
}

int hs_pack_key_string(const char *s, uint32_t key_words, uint64_t *w) {
  if (!s || !w || key_words == 0 || key_words > HS_MAX_KEY_WORDS) return HS_ERR_INVALID;
  size_t n = strlen(s);
  if (n > 16u * key_words) {
    hs::set_error("hs_pack_key_string: %zu characters do not fit %u words", n, key_words);
    return HS_ERR_UNSUPPORTED;
  }
  for (uint32_t i = 0; i < key_words; ++i) w[i] = 0;
  for (size_t i = 0; i < n; ++i) {
    uint64_t nib;
    if (s[i] >= '0' && s[i] <= '9') nib = (uint64_t)(s[i] - '0' + 1);
    else if (s[i] == '-') nib = 11;
    else return HS_ERR_INVALID;
    for (int j = (int)key_words - 1; j > 0; --j) w[j] = (w[j] << 4) | (w[j - 1] >> 60);
    w[0] = (w[0] << 4) | nib;
  }
  return HS_OK;
}
}
