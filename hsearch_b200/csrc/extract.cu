// E2: sliding-window fragment extractor over a concatenated residue store.
// Reference: ProteinDB (hclust/src/hclust/protein.hpp:7-72) + the window loop of
// BuildLSHTalbe (kmer_search.cpp:64-83) with the :73 bug fixed (the window
// advances through its residues) and proteins shorter than the window skipped
// (the source underflows an unsigned there, :70).
#include <vector>

#include "common.cuh"
#include "internal.cuh"

namespace hs {

// one thread per fragment; frag_start[p] = first fragment of protein p
__global__ void extract_windows_kernel(const uint8_t *__restrict__ residues, const uint32_t *__restrict__ start_index,
                                       const uint32_t *__restrict__ frag_start, uint32_t nprot, uint32_t stride,
                                       int len, uint64_t nfrag, uint8_t *__restrict__ codes,
                                       uint32_t *__restrict__ pos_out) {
  const uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nfrag) return;
  uint32_t lo = 0, hi = nprot;  // last protein with frag_start <= f
  while (hi - lo > 1) {
    const uint32_t m = (lo + hi) >> 1;
    if ((uint64_t)frag_start[m] <= f) lo = m; else hi = m;
  }
  const uint32_t pos = start_index[lo] + (uint32_t)(f - frag_start[lo]) * stride;
  for (int i = 0; i < len; ++i) codes[f * len + i] = residues[pos + i];
  if (pos_out) pos_out[f] = pos;
}

int extract_windows_impl(hs_ctx *ctx, const uint8_t *residues, const uint32_t *start_index, uint32_t nprot,
                         uint32_t stride, uint64_t id_base, uint32_t *pos_out, uint64_t pos_cap, uint64_t *nfrag) {
  const uint32_t len = ctx->prm.len;
  std::vector<uint32_t> frag_start((size_t)nprot + 1);
  uint64_t n = 0;
  for (uint32_t p = 0; p < nprot; ++p) {
    frag_start[p] = (uint32_t)n;
    if (start_index[p + 1] < start_index[p]) {
      set_error("hs_extract_windows: start_index not ascending at protein %u", p);
      return HS_ERR_INVALID;
    }
    const uint32_t plen = start_index[p + 1] - start_index[p];
    if (plen >= len) n += (uint64_t)(plen - len) / stride + 1;
    if (n >= (1ull << 32) - 4096) {
      set_error("hs_extract_windows: more than 2^32-4096 fragments");
      return HS_ERR_UNSUPPORTED;
    }
  }
  frag_start[nprot] = (uint32_t)n;
  *nfrag = n;
  if (pos_out && pos_cap < n) {
    set_error("hs_extract_windows: pos_out holds %llu entries, %llu needed", (unsigned long long)pos_cap,
              (unsigned long long)n);
    return HS_ERR_CAPACITY;
  }
  const uint64_t total = nprot ? start_index[nprot] : 0;
  ctx->N = n;
  ctx->id_base = id_base;
  ctx->npad = (n + 15) & ~15ull;
  ctx->hashed = ctx->indexed = ctx->have_codes_pm = ctx->have_rec = false;
  HS_TRY(ctx->d_codes.reserve((size_t)n * len + 64));
  if (n == 0) return HS_OK;
  HS_TRY(ctx->d_residues.reserve(total + 16));
  HS_TRY(ctx->d_starts.reserve(sizeof(uint32_t) * 2 * ((size_t)nprot + 1)));
  uint32_t *d_start = ctx->d_starts.as<uint32_t>();
  uint32_t *d_fstart = d_start + nprot + 1;
  HS_CUDA(cudaMemcpyAsync(ctx->d_residues.p, residues, total, cudaMemcpyHostToDevice, ctx->stream));
  HS_CUDA(cudaMemcpyAsync(d_start, start_index, sizeof(uint32_t) * (nprot + 1), cudaMemcpyHostToDevice, ctx->stream));
  HS_CUDA(cudaMemcpyAsync(d_fstart, frag_start.data(), sizeof(uint32_t) * (nprot + 1), cudaMemcpyHostToDevice,
                          ctx->stream));
  uint32_t *d_pos = nullptr;
  if (pos_out) {
    HS_TRY(ctx->d_misc.reserve(sizeof(uint32_t) * n));
    d_pos = ctx->d_misc.as<uint32_t>();
  }
  extract_windows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
      ctx->d_residues.as<uint8_t>(), d_start, d_fstart, nprot, stride, (int)len, n, ctx->d_codes.as<uint8_t>(), d_pos);
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches++;
  if (pos_out) HS_CUDA(cudaMemcpyAsync(pos_out, d_pos, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
  return validate_codes(ctx, "hs_extract_windows");
}

// ProteinDB::ProteinID (hclust/src/hclust/protein.hpp:28-39): the largest l with pos >= start_index[l],
// searched over ALL nstart entries of start_index (the final sentinel included, as the reference
// does: a position at or past the end maps to nstart - 1).
__global__ void protein_id_kernel(const uint32_t *__restrict__ start_index, uint32_t nstart, const uint32_t *__restrict__ pos,
                                  uint64_t n, uint32_t *__restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t p = pos[i];
  uint32_t l = 0, h = nstart - 1;
  while (l < h) {
    const uint32_t m = l + (h - l + 1) / 2;
    if (p >= __ldg(start_index + m)) l = m; else h = m - 1;
  }
  out[i] = l;
}

int protein_id_impl(hs_ctx *ctx, const uint32_t *start_index, uint32_t nstart, const uint32_t *pos, uint64_t n,
                    uint32_t *out) {
  if (n == 0) return HS_OK;
  HS_TRY(ctx->d_starts.reserve(sizeof(uint32_t) * (size_t)nstart));
  HS_TRY(ctx->d_misc.reserve(sizeof(uint32_t) * 2 * n));
  uint32_t *d_pos = ctx->d_misc.as<uint32_t>(), *d_out = d_pos + n;
  HS_CUDA(cudaMemcpyAsync(ctx->d_starts.p, start_index, sizeof(uint32_t) * nstart, cudaMemcpyHostToDevice, ctx->stream));
  HS_CUDA(cudaMemcpyAsync(d_pos, pos, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
  protein_id_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_starts.as<uint32_t>(), nstart, d_pos, n, d_out);
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches++;
  HS_CUDA(cudaMemcpyAsync(out, d_out, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  return HS_OK;
}

}  // namespace hs
