// E4 on the device: ProteinDB::ReadFASTAFile (pcluster/src/pcluster/read_proteins.cpp:6-41) over a
// FASTA text of up to 4 GB, byte-parallel.  Same outputs and quirks as the host parser
// hs_parse_fasta (sequence.cu): a name per header line (up to the first space), a sequence only
// when non-empty, letters of AA20 kept, every other letter replaced by AA20[rand() % 20] -- the
// k-th replaced letter in file order takes the k-th rand() value, drawn on the host after the
// device has counted and ranked them, so that the output is the reference's under the same srand.
//
// Two passes over the text in chunks of 4096 bytes (256 threads x 16 bytes):
//   count: per chunk the kept letters, the letters to replace and the header lines;
//   emit : with the exclusive scans of those counts every thread knows where its letters, its
//          replacement slots and its headers go.
// Whether a byte belongs to a header line is decided by the start of its line: inside a chunk by a
// block-wide scan over "last line start so far", at the chunk's first bytes by looking back to
// the previous newline.
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "internal.cuh"
#include "sort.cuh"

namespace hs {

constexpr int kFaThreads = 256;
constexpr int kFaBytes = 16;                       // per thread
constexpr int kFaChunk = kFaThreads * kFaBytes;    // per block

// 0: dropped, 1: kept (a letter of AA20 = "ARNDCEQGHILKMFPSTWYV", util.hpp:97), 2: replaced (isalpha)
__device__ __forceinline__ int fa_class(unsigned char c) {
  // bit i set: letter 'A' + i is in AA20 (all but B J O U X Z)
  constexpr uint32_t kAA = 0x03FFFFFFu & ~((1u << ('B' - 'A')) | (1u << ('J' - 'A')) | (1u << ('O' - 'A')) |
                                           (1u << ('U' - 'A')) | (1u << ('X' - 'A')) | (1u << ('Z' - 'A')));
  if (c >= 'A' && c <= 'Z') return ((kAA >> (c - 'A')) & 1u) ? 1 : 2;
  if (c >= 'a' && c <= 'z') return 2;
  return 0;
}

struct FaOut {
  char *residues;        // [nres]
  uint32_t *rand_pos;    // [nrnd] residue index of the k-th letter to replace
  uint64_t *name_begin;  // [nhdr]
  uint32_t *name_len;    // [nhdr]
  uint32_t *hdr_res;     // [nhdr] kept + replaced letters before the header
};

template <bool EMIT>
__global__ void __launch_bounds__(kFaThreads)
fasta_pass_kernel(const unsigned char *__restrict__ text, uint64_t n, uint32_t *__restrict__ cnt /* [3][nchunks] */,
                  uint32_t nchunks, FaOut out) {
  __shared__ uint32_t s_state[kFaThreads / 32];
  __shared__ uint32_t s_cnt[3][kFaThreads / 32];
  __shared__ unsigned long long s_back;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint64_t base = (uint64_t)blockIdx.x * kFaChunk;
  const uint64_t i0 = base + (uint64_t)tid * kFaBytes;

  unsigned char b[kFaBytes];
  unsigned char prev = '\n';   // the byte before the segment ('\n' before the text: byte 0 starts a line)
  if (i0 < n) {
    if (i0 + kFaBytes <= n) {
      const uint4 v = *reinterpret_cast<const uint4 *>(text + i0);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < kFaBytes; ++j) b[j] = (unsigned char)(w[j >> 2] >> (8 * (j & 3)));
    } else {
#pragma unroll
      for (int j = 0; j < kFaBytes; ++j) b[j] = i0 + j < n ? text[i0 + j] : (unsigned char)'\n';
    }
    if (i0 > 0) prev = text[i0 - 1];
  } else {
#pragma unroll
    for (int j = 0; j < kFaBytes; ++j) b[j] = '\n';
  }

  // last line start inside the segment: 0 none, 2 sequence line, 3 header line
  uint32_t st = 0;
  {
    unsigned char p = prev;
#pragma unroll
    for (int j = 0; j < kFaBytes; ++j) {
      if (p == '\n' && i0 + j < n) st = b[j] == '>' ? 3u : 2u;
      p = b[j];
    }
  }
  // exclusive scan of "the later line start wins" over the block
  uint32_t inc = st;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d && inc == 0) inc = o;
  }
  if (lane == 31) s_state[wid] = inc;
  // state of the line that contains the chunk's first byte when that byte does not start a line
  if (tid == 0) s_back = 0ull;
  __syncthreads();
  const bool need_back = base > 0 && base < n && text[base - 1] != '\n';
  if (need_back) {
    // look back for the previous line start, 256 bytes at a time
    for (uint64_t w = 0;; ++w) {
      const uint64_t off = w * kFaThreads + (uint64_t)tid + 1;   // candidate start index = base - off
      bool is_start = false;
      if (off <= base) {
        const uint64_t k = base - off;
        is_start = k == 0 || text[k - 1] == '\n';
        if (is_start) atomicMax(&s_back, k + 1);
      }
      const bool last = (w + 1) * kFaThreads >= base;
      if (__syncthreads_or(is_start ? 1 : 0) || last) break;
    }
  }
  __syncthreads();
  uint32_t chunk_state = 2u;
  if (need_back) chunk_state = text[s_back - 1] == '>' ? 3u : 2u;
  uint32_t before = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) before = 0;
  for (int w2 = wid - 1; w2 >= 0 && before == 0; --w2) before = s_state[w2];
  if (before == 0) before = chunk_state;
  bool header = before == 3u;

  // walk the segment
  uint32_t nres = 0, nrnd = 0, nhdr = 0;
  uint32_t cls = 0;      // 2 bits per byte: class of the bytes that are sequence letters
  uint32_t hmask = 0;    // header line starts
  {
    unsigned char p = prev;
#pragma unroll
    for (int j = 0; j < kFaBytes; ++j) {
      const unsigned char c = b[j];
      if (i0 + j < n) {
        if (p == '\n') {
          header = c == '>';
          if (header) {
            ++nhdr;
            hmask |= 1u << j;
          }
        }
        if (!header) {
          const int k = fa_class(c);
          cls |= (uint32_t)k << (2 * j);
          nres += k != 0;
          nrnd += k == 2;
        }
      }
      p = c;
    }
  }
  // block-wide exclusive sums of the three counters
  uint32_t v[3] = {nres, nrnd, nhdr}, ex[3], tot[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    uint32_t x = v[q];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += o;
    }
    if (lane == 31) s_cnt[q][wid] = x;
    ex[q] = x - v[q];
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    uint32_t add = 0, t = 0;
    for (int w2 = 0; w2 < kFaThreads / 32; ++w2) {
      if (w2 < wid) add += s_cnt[q][w2];
      t += s_cnt[q][w2];
    }
    ex[q] += add;
    tot[q] = t;
  }
  if (!EMIT) {
    if (tid == 0) {
      cnt[blockIdx.x] = tot[0];
      cnt[(size_t)nchunks + blockIdx.x] = tot[1];
      cnt[2 * (size_t)nchunks + blockIdx.x] = tot[2];
    }
    return;
  }
  uint32_t o_res = cnt[blockIdx.x] + ex[0];
  uint32_t o_rnd = cnt[(size_t)nchunks + blockIdx.x] + ex[1];
  uint32_t o_hdr = cnt[2 * (size_t)nchunks + blockIdx.x] + ex[2];
#pragma unroll
  for (int j = 0; j < kFaBytes; ++j) {
    if ((hmask >> j) & 1u) {
      const uint64_t i = i0 + j;
      uint64_t e = i + 1;
      while (e < n && text[e] != ' ' && text[e] != '\n') ++e;   // header up to the first space
      out.name_begin[o_hdr] = i + 1;
      out.name_len[o_hdr] = (uint32_t)(e - i - 1);
      out.hdr_res[o_hdr] = o_res;
      ++o_hdr;
    }
    const uint32_t k = (cls >> (2 * j)) & 3u;
    if (k == 1u) {
      out.residues[o_res++] = (char)b[j];
    } else if (k == 2u) {
      out.rand_pos[o_rnd++] = o_res;
      out.residues[o_res++] = '?';
    }
  }
}

__global__ void fasta_patch_kernel(char *__restrict__ residues, const uint32_t *__restrict__ rand_pos,
                                   const char *__restrict__ letters, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) residues[rand_pos[i]] = letters[i];
}

// Sequence boundaries: the distinct values of {0} u {hdr_res[h]} u {nres}.  flag[j], j < nhdr: header
// j closes a non-empty sequence; flag[nhdr]: the text ends inside a non-empty sequence.
__global__ void fasta_bound_flags_kernel(const uint32_t *__restrict__ hdr_res, uint32_t nhdr, uint32_t nres,
                                         uint32_t *__restrict__ flag) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > nhdr) return;
  const uint32_t v = j < nhdr ? hdr_res[j] : nres;
  const uint32_t p = j ? hdr_res[j - 1] : 0u;
  flag[j] = v > p ? 1u : 0u;
}
__global__ void fasta_bound_scatter_kernel(const uint32_t *__restrict__ hdr_res, uint32_t nhdr, uint32_t nres,
                                           const uint32_t *__restrict__ flag, const uint32_t *__restrict__ scanned,
                                           uint64_t *__restrict__ start) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0) start[0] = 0;
  if (j > nhdr || !flag[j]) return;
  start[1 + scanned[j]] = j < nhdr ? hdr_res[j] : nres;
}

int parse_fasta_gpu_impl(hs_ctx *ctx, const char *text, uint64_t nbytes, char *residues, uint64_t res_cap, uint64_t *start,
                         uint64_t start_cap, uint64_t *name_begin, uint32_t *name_len, uint64_t name_cap, uint32_t *nseq,
                         uint32_t *nnames, uint64_t *nres) {
  *nseq = *nnames = 0;
  *nres = 0;
  if (nbytes == 0) return HS_OK;
  if (nbytes >= (1ull << 32) - kFaChunk) {
    set_error("hs_parse_fasta_gpu: texts of 4 GB and more are parsed in pieces (cut at a header line)");
    return HS_ERR_UNSUPPORTED;
  }
  stats_begin(ctx);
  const uint32_t nchunks = (uint32_t)((nbytes + kFaChunk - 1) / kFaChunk);
  HS_TRY(ctx->d_residues.reserve(nbytes + 64));             // the text
  HS_TRY(ctx->d_starts.reserve(sizeof(uint32_t) * (3 * (size_t)nchunks + 8)));
  unsigned char *d_text = ctx->d_residues.as<unsigned char>();
  uint32_t *d_cnt = ctx->d_starts.as<uint32_t>();
  uint32_t *d_tot = d_cnt + 3 * (size_t)nchunks;
  HS_CUDA(cudaMemcpyAsync(d_text, text, nbytes, cudaMemcpyHostToDevice, ctx->stream));
  FaOut none{};
  fasta_pass_kernel<false><<<nchunks, kFaThreads, 0, ctx->stream>>>(d_text, nbytes, d_cnt, nchunks, none);
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches++;
  for (int q = 0; q < 3; ++q)
    HS_TRY(exclusive_scan_u32(ctx, d_cnt + (size_t)q * nchunks, d_cnt + (size_t)q * nchunks, nchunks, d_tot + q));
  uint32_t tot[3];
  HS_TRY(read_back(ctx, d_tot, tot, sizeof tot));
  const uint32_t n_res = tot[0], n_rnd = tot[1], n_hdr = tot[2];
  // the reference draws one rand() per replaced letter, in file order, whatever happens later
  std::vector<char> letters(n_rnd);
  static const char AA20[] = "ARNDCEQGHILKMFPSTWYV";  // pcluster/src/pcluster/util.hpp:97
  for (uint32_t i = 0; i < n_rnd; ++i) letters[i] = AA20[rand() % 20];

  // device outputs, all in the context's scratch buffer (the loaded database is not touched)
  const size_t o_rand = 0, o_nlen = o_rand + sizeof(uint32_t) * (size_t)n_rnd, o_hres = o_nlen + sizeof(uint32_t) * (size_t)n_hdr;
  const size_t o_flag = o_hres + sizeof(uint32_t) * (size_t)n_hdr, o_scan = o_flag + sizeof(uint32_t) * ((size_t)n_hdr + 1);
  size_t o_nbeg = o_scan + sizeof(uint32_t) * ((size_t)n_hdr + 1);
  o_nbeg = (o_nbeg + 15) & ~(size_t)15;
  const size_t o_start = o_nbeg + sizeof(uint64_t) * (size_t)n_hdr;
  const size_t o_resid = o_start + sizeof(uint64_t) * ((size_t)n_hdr + 2);   // [residues | replacement letters]
  HS_TRY(ctx->d_misc.reserve(o_resid + (size_t)n_res + n_rnd + 64));
  char *d_misc = ctx->d_misc.as<char>();
  FaOut fo;
  fo.residues = d_misc + o_resid;
  fo.rand_pos = reinterpret_cast<uint32_t *>(d_misc + o_rand);
  fo.name_len = reinterpret_cast<uint32_t *>(d_misc + o_nlen);
  fo.hdr_res = reinterpret_cast<uint32_t *>(d_misc + o_hres);
  fo.name_begin = reinterpret_cast<uint64_t *>(d_misc + o_nbeg);
  uint32_t *d_flag = reinterpret_cast<uint32_t *>(d_misc + o_flag), *d_scan = reinterpret_cast<uint32_t *>(d_misc + o_scan);
  uint64_t *d_start = reinterpret_cast<uint64_t *>(d_misc + o_start);
  fasta_pass_kernel<true><<<nchunks, kFaThreads, 0, ctx->stream>>>(d_text, nbytes, d_cnt, nchunks, fo);
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches++;
  if (n_rnd) {
    char *d_letters = fo.residues + n_res;
    HS_CUDA(cudaMemcpyAsync(d_letters, letters.data(), n_rnd, cudaMemcpyHostToDevice, ctx->stream));
    fasta_patch_kernel<<<(n_rnd + 255) / 256, 256, 0, ctx->stream>>>(fo.residues, fo.rand_pos, d_letters, n_rnd);
    HS_CUDA(cudaGetLastError());
    ctx->stats.kernel_launches++;
  }
  const unsigned gb = (n_hdr + 1 + 255) / 256;
  fasta_bound_flags_kernel<<<gb, 256, 0, ctx->stream>>>(fo.hdr_res, n_hdr, n_res, d_flag);
  HS_TRY(exclusive_scan_u32(ctx, d_flag, d_scan, (uint64_t)n_hdr + 1, d_tot + 3));
  fasta_bound_scatter_kernel<<<gb, 256, 0, ctx->stream>>>(fo.hdr_res, n_hdr, n_res, d_flag, d_scan, d_start);
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches += 2;
  uint32_t n_seq = 0;
  HS_TRY(read_back(ctx, d_tot + 3, &n_seq, sizeof n_seq));
  *nseq = n_seq;
  *nnames = n_hdr;
  *nres = n_res;
  const bool overflow = (n_res && (!residues || res_cap < n_res)) || (n_seq && (!start || start_cap < (uint64_t)n_seq + 1)) ||
                        (n_hdr && (!name_begin || !name_len || name_cap < n_hdr));
  if (overflow) {
    set_error("hs_parse_fasta_gpu: buffers too small (%u sequences, %u names, %u residues)", n_seq, n_hdr, n_res);
    return HS_ERR_CAPACITY;
  }
  if (n_res) HS_CUDA(cudaMemcpyAsync(residues, fo.residues, n_res, cudaMemcpyDeviceToHost, ctx->stream));
  if (n_seq) HS_CUDA(cudaMemcpyAsync(start, d_start, sizeof(uint64_t) * ((size_t)n_seq + 1), cudaMemcpyDeviceToHost, ctx->stream));
  if (n_hdr) {
    HS_CUDA(cudaMemcpyAsync(name_begin, fo.name_begin, sizeof(uint64_t) * (size_t)n_hdr, cudaMemcpyDeviceToHost, ctx->stream));
    HS_CUDA(cudaMemcpyAsync(name_len, fo.name_len, sizeof(uint32_t) * (size_t)n_hdr, cudaMemcpyDeviceToHost, ctx->stream));
  }
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  return HS_OK;
}

}  // namespace hs

extern "C" int hs_parse_fasta_gpu(hs_ctx_t *ctx, const char *text, uint64_t nbytes, char *residues, uint64_t res_cap,
                                  uint64_t *start, uint64_t start_cap, uint64_t *name_begin, uint32_t *name_len,
                                  uint64_t name_cap, uint32_t *nseq, uint32_t *nnames, uint64_t *nres) {
  if (!ctx || (!text && nbytes) || !nseq || !nnames || !nres) {
    hs::set_error("hs_parse_fasta_gpu: null argument");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  return hs::parse_fasta_gpu_impl(ctx, text, nbytes, residues, res_cap, start, start_cap, name_begin, name_len, name_cap,
                                  nseq, nnames, nres);
}
