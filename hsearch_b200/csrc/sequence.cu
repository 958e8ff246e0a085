// Sequence front ends either side of the hash path:
//
//   E4  ProteinDB::ReadFASTAFile, pcluster/src/pcluster/read_proteins.cpp:6-41   (host)
//   E5  3-mer feature histogram, pcluster/src/pcluster/pcluster.cpp:19-32 with
//       Kmer2Integer / REDUCEDAAINDEX / BASEP, pcluster/src/pcluster/util.hpp:103-106,244-250
//   KL1 KLSH::KLSH / GetHashValue, pcluster/src/pcluster/lsh.cpp:8-49           (ctor host)
//   E6  ORF::orf6 six-frame translation, orf/orf.cc:4-74, code table orf/orf.h:28-31
#include <ctype.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <random>
#include <string>
#include <vector>

#include "common.cuh"
#include "host_tables.h"

namespace hs {

constexpr int kFeat = 512;    // pow(8, HASHLEN), HASHLEN = 3 (pcluster.cpp:13)
constexpr int kHashLen = 3;
constexpr double kCosGuard = 1e-9;  // |cos(sum) + t| below this: bit recomputed on the host with libm

// REDUCEDAAINDEX (util.hpp:103-104); -1 for the six letters that are not amino acids
__constant__ int c_reduced[26] = {0, -1, 3, 1, 1, 6, 4, 2, 5, -1, 1, 5, 5, 2, -1, 7, 1, 1, 0, 0, -1, 5, 6, -1, 6, -1};
static const int h_reduced[26] = {0, -1, 3, 1, 1, 6, 4, 2, 5, -1, 1, 5, 5, 2, -1, 7, 1, 1, 0, 0, -1, 5, 6, -1, 6, -1};

// One block per protein: shared-memory histogram of the reduced-alphabet 3-mers, then one
// thread per hash bit: Dot() of lsh.cpp:8-15 in its order (separate multiply and add), + b,
// cos, + t, sign.  wT is the projection transposed ([feat][bits]) so that the bit threads
// read consecutive doubles.  flags[p]: 1 = hashed, 0 = shorter than HASHLEN (skipped,
// pcluster.cpp:22-24), |2 = a bit fell inside the cos guard band, |4 = letter outside A-Z /
// not an amino acid.
__global__ void __launch_bounds__(128)
kmer3_klsh_kernel(const char *__restrict__ residues, const uint64_t *__restrict__ start, uint32_t nprot,
                  const double *__restrict__ wT, const double *__restrict__ t, const double *__restrict__ b,
                  uint32_t bits, uint32_t *__restrict__ feat_out, uint64_t *__restrict__ hash_out,
                  uint8_t *__restrict__ flags) {
  __shared__ int s_feat[kFeat];
  __shared__ unsigned long long s_hash;
  __shared__ unsigned int s_flag;
  for (uint32_t p = blockIdx.x; p < nprot; p += gridDim.x) {
    const uint64_t s0 = start[p], n = start[p + 1] - s0;
    for (int i = threadIdx.x; i < kFeat; i += blockDim.x) s_feat[i] = 0;
    if (threadIdx.x == 0) {
      s_hash = 0ull;
      s_flag = n >= (uint64_t)kHashLen ? 1u : 0u;
    }
    __syncthreads();
    if (n >= (uint64_t)kHashLen) {
      for (uint64_t i = threadIdx.x; i + kHashLen <= n; i += blockDim.x) {
        int h = 0, mul = 1;
        bool bad = false;
#pragma unroll
        for (int k = 0; k < kHashLen; ++k) {
          const int c = (int)residues[s0 + i + k] - 'A';
          const int r = (c >= 0 && c < 26) ? c_reduced[c] : -1;
          bad = bad || r < 0;
          h += r * mul;
          mul *= 8;
        }
        if (bad) atomicOr(&s_flag, 4u);
        else atomicAdd(&s_feat[h], 1);
      }
      __syncthreads();
      if (threadIdx.x < bits) {
        const uint32_t bit = threadIdx.x;
        double sum = 0.0;
        for (int i = 0; i < kFeat; ++i) sum = __dadd_rn(sum, __dmul_rn((double)s_feat[i], wT[(size_t)i * bits + bit]));
        sum = __dadd_rn(sum, b[bit]);
        const double v = __dadd_rn(cos(sum), t[bit]);
        if (fabs(v) < kCosGuard) atomicOr(&s_flag, 2u);
        if (v >= 0) atomicOr(&s_hash, 1ull << bit);
      }
      if (feat_out)
        for (int i = threadIdx.x; i < kFeat; i += blockDim.x) feat_out[(size_t)p * kFeat + i] = (uint32_t)s_feat[i];
    } else if (feat_out) {
      for (int i = threadIdx.x; i < kFeat; i += blockDim.x) feat_out[(size_t)p * kFeat + i] = 0u;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      hash_out[p] = s_hash;
      flags[p] = (uint8_t)s_flag;
    }
    __syncthreads();
  }
}

// GetHashValue on the host (libm cos: the function the reference calls), for the rare
// proteins whose device result was inside the guard band.
static uint64_t klsh_hash_host(const char *seq, uint64_t n, const double *w, const double *t, const double *b,
                               uint32_t bits) {
  std::vector<double> p(kFeat, 0.0);
  for (uint64_t i = 0; i + kHashLen <= n; ++i) {
    int h = 0, mul = 1;
    for (int k = 0; k < kHashLen; ++k) {
      h += h_reduced[seq[i + k] - 'A'] * mul;
      mul *= 8;
    }
    p[h] += 1.0;
  }
  uint64_t hv = 0;
  for (uint32_t i = 0; i < bits; ++i) {
    volatile double sum = 0;
    for (int j = 0; j < kFeat; ++j) {
      volatile double prod = p[j] * w[(size_t)i * kFeat + j];  // no contraction into an FMA
      sum = sum + prod;
    }
    const double s = sum + b[i];
    hv |= (uint64_t)((std::cos(s) + t[i]) >= 0 ? 1 : 0) << i;
  }
  return hv;
}

// ---- E6 ----------------------------------------------------------------------------------
// Base1/2/3 of orf.h:28-30 enumerate the codons in T, C, A, G order; AAs (orf.h:31).
__constant__ char c_codon_aa[65] = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";

__device__ __forceinline__ int base_index(char c) {  // T C A G -> 0..3, else -1
  return c == 'T' ? 0 : c == 'C' ? 1 : c == 'A' ? 2 : c == 'G' ? 3 : -1;
}

// One thread per (sequence, frame): frames 0-2 read the strand, 3-5 its reverse complement
// (orf.cc:32-37,52-72); codon by codon until the first stop (orf.cc:44-50).  aa_len[s][f] =
// residues written; the frame is an ORF when it is >= 6 (orf.cc:58).  bad[s] = 1 when the
// sequence holds a letter other than A, C, G, T (ERROR_INFO in the reference, orf.cc:26).
__global__ void orf6_kernel(const char *__restrict__ dna, const uint64_t *__restrict__ start, uint32_t nseq,
                            char *__restrict__ aa_out, int32_t *__restrict__ aa_len, uint8_t *__restrict__ bad) {
  const uint64_t tix = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tix >= (uint64_t)nseq * 6) return;
  const uint32_t s = (uint32_t)(tix / 6);
  const int f = (int)(tix % 6);
  const uint64_t s0 = start[s];
  const int64_t n = (int64_t)(start[s + 1] - s0);
  const char *seq = dna + s0;
  char *out = aa_out + 2 * s0 + 6ull * s + (uint64_t)f * (uint64_t)(n / 3 + 1);
  const int st = f % 3;
  const bool rev = f >= 3;
  int m = 0;
  bool isbad = false;
  for (int64_t i = st; i + 3 <= n; i += 3) {
    int idx = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      // reverse strand: position j of the reverse complement is the complement of seq[n-1-j]
      const char c = rev ? seq[n - 1 - (i + k)] : seq[i + k];
      int bi = base_index(c);
      if (bi < 0) isbad = true;
      if (rev && bi >= 0) bi = bi ^ 2;  // T<->A (0<->2), C<->G (1<->3)
      idx = idx * 4 + (bi < 0 ? 0 : bi);
    }
    if (isbad) break;
    const char aa = c_codon_aa[idx];
    if (aa == '*') break;
    out[m++] = aa;
  }
  aa_len[(size_t)s * 6 + f] = m;
  if (isbad) bad[s] = 1;
}

}  // namespace hs

using namespace hs;

extern "C" {

int hs_klsh_generate(uint32_t feat, uint32_t bits, double sigma, double *w, double *t, double *b) {
  if (!w || !t || !b || feat == 0 || bits == 0 || bits > 64) {
    set_error("hs_klsh_generate: bad argument");
    return HS_ERR_INVALID;
  }
  // KLSH::KLSH (lsh.cpp:17-38): `generator` is a default-constructed member, i.e.
  // minstd_rand0 with seed 1 in every instance; stddev of the normals = sigma * sigma
  std::default_random_engine generator;
  std::normal_distribution<double> normal(0.0, sigma * sigma);
  std::uniform_real_distribution<double> uniform_1(-1.0, 1.0);
  std::uniform_real_distribution<double> uniform_pi(0.0, 2.0 * M_PI);
  for (uint32_t i = 0; i < bits; ++i) {
    t[i] = uniform_1(generator);
    b[i] = uniform_pi(generator);
    for (uint32_t j = 0; j < feat; ++j) w[(size_t)i * feat + j] = normal(generator);
  }
  return HS_OK;
}

int hs_kmer3_klsh(hs_ctx_t *ctx, const char *residues, const uint64_t *start, uint32_t nprot, const double *w,
                  const double *t, const double *b, uint32_t bits, uint32_t *feat_out, uint64_t *hash_out,
                  uint8_t *valid_out, uint64_t *n_host_fixed) {
  if (!ctx || !start || !w || !t || !b || !hash_out || bits == 0 || bits > 64 || (!residues && nprot && start[nprot])) {
    set_error("hs_kmer3_klsh: bad argument");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  if (n_host_fixed) *n_host_fixed = 0;
  if (nprot == 0) return HS_OK;
  const uint64_t total = start[nprot];
  std::vector<double> wT((size_t)kFeat * bits);
  for (uint32_t i = 0; i < bits; ++i)
    for (int j = 0; j < kFeat; ++j) wT[(size_t)j * bits + i] = w[(size_t)i * kFeat + j];
  DevBuf d_res, d_start, d_w, d_tb, d_feat, d_hash, d_flags;
  auto cleanup = [&]() {
    for (DevBuf *p : {&d_res, &d_start, &d_w, &d_tb, &d_feat, &d_hash, &d_flags}) p->release();
  };
  int rc = d_res.reserve(total + 16);
  if (rc == HS_OK) rc = d_start.reserve(sizeof(uint64_t) * ((size_t)nprot + 1));
  if (rc == HS_OK) rc = d_w.reserve(sizeof(double) * wT.size());
  if (rc == HS_OK) rc = d_tb.reserve(sizeof(double) * 2 * bits);
  if (rc == HS_OK && feat_out) rc = d_feat.reserve(sizeof(uint32_t) * (size_t)nprot * kFeat);
  if (rc == HS_OK) rc = d_hash.reserve(sizeof(uint64_t) * nprot);
  if (rc == HS_OK) rc = d_flags.reserve(nprot);
  if (rc != HS_OK) {
    cleanup();
    return rc;
  }
  cudaStream_t st = ctx->stream;
  cudaError_t e = cudaSuccess;
  if (total) e = cudaMemcpyAsync(d_res.p, residues, total, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_start.p, start, sizeof(uint64_t) * ((size_t)nprot + 1), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_w.p, wT.data(), sizeof(double) * wT.size(), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_tb.p, t, sizeof(double) * bits, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_tb.as<double>() + bits, b, sizeof(double) * bits, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    const unsigned grid = (unsigned)std::min<uint64_t>(nprot, (uint64_t)ctx->num_sms * 16);
    kmer3_klsh_kernel<<<grid, 128, 0, st>>>(d_res.as<char>(), d_start.as<uint64_t>(), nprot, d_w.as<double>(),
                                            d_tb.as<double>(), d_tb.as<double>() + bits, bits,
                                            feat_out ? d_feat.as<uint32_t>() : nullptr, d_hash.as<uint64_t>(),
                                            d_flags.as<uint8_t>());
    e = cudaGetLastError();
    ctx->stats.kernel_launches++;
  }
  std::vector<uint8_t> flags(nprot);
  if (e == cudaSuccess) e = cudaMemcpyAsync(hash_out, d_hash.p, sizeof(uint64_t) * nprot, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(flags.data(), d_flags.p, nprot, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && feat_out)
    e = cudaMemcpyAsync(feat_out, d_feat.p, sizeof(uint32_t) * (size_t)nprot * kFeat, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cleanup();
  if (e != cudaSuccess) {
    set_error("hs_kmer3_klsh: %s", cudaGetErrorString(e));
    return HS_ERR_CUDA;
  }
  uint64_t fixed = 0;
  for (uint32_t p = 0; p < nprot; ++p) {
    if (flags[p] & 4u) {
      set_error("hs_kmer3_klsh: protein %u holds a letter that is not one of the 20 amino acids", p);
      return HS_ERR_INVALID;
    }
    if (flags[p] & 2u) {  // inside the cos guard band: libm decides, as in the reference
      hash_out[p] = klsh_hash_host(residues + start[p], start[p + 1] - start[p], w, t, b, bits);
      ++fixed;
    }
    if (valid_out) valid_out[p] = flags[p] & 1u;
  }
  if (n_host_fixed) *n_host_fixed = fixed;
  return HS_OK;
}

int hs_orf6(hs_ctx_t *ctx, const char *dna, const uint64_t *start, uint32_t nseq, char *aa_out, uint64_t aa_cap,
            int32_t *aa_len) {
  if (!ctx || !start || !aa_len || (!aa_out && aa_cap) || (!dna && nseq && start[nseq])) {
    set_error("hs_orf6: bad argument");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  if (nseq == 0) return HS_OK;
  const uint64_t total = start[nseq];
  const uint64_t need = 2 * total + 6ull * nseq;
  if (aa_cap < need) {
    set_error("hs_orf6: aa_out needs %llu bytes (2 * total length + 6 * nseq)", (unsigned long long)need);
    return HS_ERR_CAPACITY;
  }
  DevBuf d_dna, d_start, d_out, d_len, d_bad;
  auto cleanup = [&]() {
    for (DevBuf *p : {&d_dna, &d_start, &d_out, &d_len, &d_bad}) p->release();
  };
  int rc = d_dna.reserve(total + 16);
  if (rc == HS_OK) rc = d_start.reserve(sizeof(uint64_t) * ((size_t)nseq + 1));
  if (rc == HS_OK) rc = d_out.reserve(need + 16);
  if (rc == HS_OK) rc = d_len.reserve(sizeof(int32_t) * 6 * (size_t)nseq);
  if (rc == HS_OK) rc = d_bad.reserve(nseq);
  if (rc != HS_OK) {
    cleanup();
    return rc;
  }
  cudaStream_t st = ctx->stream;
  cudaError_t e = cudaSuccess;
  if (total) e = cudaMemcpyAsync(d_dna.p, dna, total, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_start.p, start, sizeof(uint64_t) * ((size_t)nseq + 1), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_out.p, 0, need, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_bad.p, 0, nseq, st);
  if (e == cudaSuccess) {
    const uint64_t nthreads = (uint64_t)nseq * 6;
    orf6_kernel<<<(unsigned)((nthreads + 127) / 128), 128, 0, st>>>(d_dna.as<char>(), d_start.as<uint64_t>(), nseq,
                                                                    d_out.as<char>(), d_len.as<int32_t>(),
                                                                    d_bad.as<uint8_t>());
    e = cudaGetLastError();
    ctx->stats.kernel_launches++;
  }
  std::vector<uint8_t> bad(nseq);
  if (e == cudaSuccess) e = cudaMemcpyAsync(aa_out, d_out.p, need, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(aa_len, d_len.p, sizeof(int32_t) * 6 * (size_t)nseq, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(bad.data(), d_bad.p, nseq, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cleanup();
  if (e != cudaSuccess) {
    set_error("hs_orf6: %s", cudaGetErrorString(e));
    return HS_ERR_CUDA;
  }
  for (uint32_t s = 0; s < nseq; ++s)
    if (bad[s]) {
      set_error("hs_orf6: sequence %u holds a letter other than A, C, G, T", s);
      return HS_ERR_INVALID;
    }
  return HS_OK;
}

// ProteinDB::ReadFASTAFile (read_proteins.cpp:6-41) over an in-memory FASTA text.  Quirks
// kept: a name is recorded for every header line but a sequence only when it is non-empty,
// so *nnames may exceed *nseq; letters of AA20 are kept, every other alphabetic character
// becomes AA20[rand() % 20] (lower case included: AA20 is upper case), the rest is dropped.
int hs_parse_fasta(const char *text, uint64_t nbytes, char *residues, uint64_t res_cap, uint64_t *start,
                   uint64_t start_cap, uint64_t *name_begin, uint32_t *name_len, uint64_t name_cap, uint32_t *nseq,
                   uint32_t *nnames, uint64_t *nres) {
  if (!text || !nseq || !nnames || !nres) {
    set_error("hs_parse_fasta: null argument");
    return HS_ERR_INVALID;
  }
  static const char AA20[] = "ARNDCEQGHILKMFPSTWYV";  // pcluster/src/pcluster/util.hpp:97
  uint64_t r = 0, ns = 0, nn = 0, seq_begin = 0;
  bool overflow = false;
  auto close_seq = [&]() {
    if (r != seq_begin) {
      if (start && ns + 1 < start_cap) {
        start[ns] = seq_begin;
        start[ns + 1] = r;
      } else {
        overflow = true;
      }
      ++ns;
      seq_begin = r;
    }
  };
  uint64_t pos = 0;
  while (pos < nbytes) {
    uint64_t eol = pos;
    while (eol < nbytes && text[eol] != '\n') ++eol;
    // getline yields the line without '\n'; an empty line has line[0] == '\0' (not '>')
    if (eol > pos && text[pos] == '>') {
      close_seq();
      uint64_t sp = pos;
      while (sp < eol && text[sp] != ' ') ++sp;  // header up to the first space
      if (name_begin && name_len && nn < name_cap) {
        name_begin[nn] = pos + 1;
        name_len[nn] = (uint32_t)(sp - pos - 1);
      } else {
        overflow = true;
      }
      ++nn;
    } else {
      for (uint64_t i = pos; i < eol; ++i) {
        const char c = text[i];
        char out = 0;
        if (c != '\0' && strchr(AA20, c)) out = (char)toupper((unsigned char)c);
        else if (isalpha((unsigned char)c)) out = AA20[rand() % 20];
        if (out) {
          if (residues && r < res_cap) residues[r] = out;
          else overflow = true;
          ++r;
        }
      }
    }
    pos = eol + 1;
  }
  close_seq();
  *nseq = (uint32_t)ns;
  *nnames = (uint32_t)nn;
  *nres = r;
  if (overflow) {
    set_error("hs_parse_fasta: buffers too small (%llu sequences, %llu names, %llu residues)",
              (unsigned long long)ns, (unsigned long long)nn, (unsigned long long)r);
    return HS_ERR_CAPACITY;
  }
  return HS_OK;
}
}
