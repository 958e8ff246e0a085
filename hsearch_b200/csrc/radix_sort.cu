// K2: hand-written LSD radix sort of (packed key, fragment id) pairs, bucket
// grouping, and the bucket-ordered code store.  Replaces the
// unordered_map<string, vector<uint32_t>> insert of motif_both_points.cpp:212-216
// (HashTable, :25): after a stable sort by key, one bucket is one contiguous
// run of ids in ascending id order -- the reference's insertion order.
//
// Sort structure (per 8-bit digit pass): upsweep (per-tile digit histogram) ->
// device-wide exclusive scan of the digit-major histogram -> downsweep (stable
// in-tile ranking with warp match, shared-memory reorder, coalesced scatter).
// Keys are KW 64-bit words stored word-major (SoA); a pass ranks on one word
// and moves all words.  Digit windows are placed only over bits that actually
// vary across the input (OR/AND reduction), so constant nibbles cost nothing.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "hash.cuh"
#include "sort.cuh"
#include "internal.cuh"

namespace hs {

constexpr int kSortThreads = 512;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096 keys per tile
constexpr int kSortWarps = kSortThreads / 32;

// Lanes of the warp holding the same digit d (0..256; 256 = out of range), from 9 ballots:
// match.any walks the distinct values of the warp one by one, which for random 8-bit digits is
// ~28 rounds per call; the ballot form is a fixed 9 votes + 9 logic ops.
__device__ __forceinline__ uint32_t digit_peers(uint32_t d) {
  uint32_t peers = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < 9; ++b) {
    const bool bit = (d >> b) & 1u;
    const uint32_t v = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? v : ~v;
  }
  return peers;
}

// ---- OR / AND reduction of the key words -------------------------------------
__global__ void key_bits_kernel(const uint64_t *__restrict__ keys, uint64_t n, int nw,
                                unsigned long long *__restrict__ or_and /* [nw][2] */) {
  for (int w = 0; w < nw; ++w) {
    unsigned long long o = 0ull, a = ~0ull;
    const uint64_t *kw = keys + (uint64_t)w * n;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
      const unsigned long long v = kw[i];
      o |= v;
      a &= v;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      o |= __shfl_xor_sync(0xffffffffu, o, s);
      a &= __shfl_xor_sync(0xffffffffu, a, s);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicOr(or_and + 2 * w, o);
      atomicAnd(or_and + 2 * w + 1, a);
    }
  }
}

// ---- upsweep: per-tile digit histogram, digit-major output --------------------
__global__ void __launch_bounds__(kSortThreads)
radix_upsweep_kernel(const uint64_t *__restrict__ kword, uint64_t n, int shift, uint32_t mask,
                     uint32_t *__restrict__ tile_hist, uint32_t ntiles) {
  __shared__ uint32_t s_hist[256];
  const int tid = threadIdx.x;
  if (tid < 256) s_hist[tid] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * kSortTile;
#pragma unroll 4
  for (int j = 0; j < kSortItems; ++j) {
    const uint64_t idx = base + (uint64_t)j * kSortThreads + tid;
    if (idx < n) {
      const uint32_t d = (uint32_t)(kword[idx] >> shift) & mask;
      atomicAdd(&s_hist[d], 1u);
    }
  }
  __syncthreads();
  if (tid < 256) tile_hist[(uint64_t)tid * ntiles + blockIdx.x] = s_hist[tid];
}

// ---- device-wide exclusive scan (u32), three phases ---------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

template <int NWARPS = 8>
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t *s_warp /*[NWARPS]*/, uint32_t &total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, s);
    if (lane >= s) x += y;
  }
  if (lane == 31) s_warp[wid] = x;
  __syncthreads();
  uint32_t woff = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < NWARPS; ++i) {
    const uint32_t t = s_warp[i];
    if (i < wid) woff += t;
    tot += t;
  }
  total = tot;
  __syncthreads();
  return woff + x - v;
}

__global__ void __launch_bounds__(kScanThreads)
scan_reduce_kernel(const uint32_t *__restrict__ in, uint64_t n, uint32_t *__restrict__ block_sums) {
  __shared__ uint32_t s_warp[8];
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile;
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    const uint64_t idx = base + (uint64_t)j * kScanThreads + threadIdx.x;
    if (idx < n) s += in[idx];
  }
  uint32_t total;
  block_exclusive_scan_256(s, s_warp, total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of block_sums in place; total -> *total_out
__global__ void __launch_bounds__(kScanThreads)
scan_sums_kernel(uint32_t *__restrict__ block_sums, uint32_t nblocks, uint32_t *__restrict__ total_out) {
  __shared__ uint32_t s_warp[8];
  uint32_t carry = 0;
  for (uint32_t base = 0; base < nblocks; base += kScanThreads) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < nblocks ? block_sums[i] : 0u;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan_256(v, s_warp, total);
    if (i < nblocks) block_sums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(kScanThreads)
scan_downsweep_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint64_t n,
                      const uint32_t *__restrict__ block_sums) {
  __shared__ uint32_t s_warp[8];
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    v[j] = (base + j < n) ? in[base + j] : 0u;
    s += v[j];
  }
  uint32_t total;
  uint32_t off = block_exclusive_scan_256(s, s_warp, total) + block_sums[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    if (base + j < n) out[base + j] = off;
    off += v[j];
  }
}

int exclusive_scan_u32(hs_ctx *ctx, const uint32_t *in, uint32_t *out, uint64_t n, uint32_t *d_total) {
  if (n == 0) {
    if (d_total) HS_CUDA(cudaMemsetAsync(d_total, 0, sizeof(uint32_t), ctx->stream));
    return HS_OK;
  }
  const uint32_t nblocks = (uint32_t)((n + kScanTile - 1) / kScanTile);
  HS_TRY(ctx->sort.block_sums.reserve(sizeof(uint32_t) * (nblocks + 1)));
  uint32_t *sums = ctx->sort.block_sums.as<uint32_t>();
  scan_reduce_kernel<<<nblocks, kScanThreads, 0, ctx->stream>>>(in, n, sums);
  scan_sums_kernel<<<1, kScanThreads, 0, ctx->stream>>>(sums, nblocks, d_total);
  scan_downsweep_kernel<<<nblocks, kScanThreads, 0, ctx->stream>>>(in, out, n, sums);
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches += 3;
  return HS_OK;
}

// ---- downsweep: stable rank + reorder + scatter -------------------------------
constexpr size_t kDownsweepSmem = sizeof(uint64_t) * kSortTile + sizeof(uint32_t) * kSortTile;

template <int NW>
__global__ void __launch_bounds__(kSortThreads, 2)
radix_downsweep_kernel(KeyPtrs in, const uint32_t *__restrict__ val_in, KeyPtrs out,
                       uint32_t *__restrict__ val_out, uint64_t n, int word, int shift, uint32_t mask,
                       const uint32_t *__restrict__ tile_off, uint32_t ntiles) {
  extern __shared__ __align__(16) unsigned char sort_smem[];
  uint64_t *s_key = reinterpret_cast<uint64_t *>(sort_smem);                               // 32 KB
  uint32_t *s_val = reinterpret_cast<uint32_t *>(sort_smem + sizeof(uint64_t) * kSortTile);  // 16 KB
  __shared__ uint32_t s_whist[kSortWarps][257];         // per-warp digit counters (+ tail bin)
  __shared__ uint32_t s_dstart[257];
  __shared__ uint32_t s_goff[256];
  __shared__ uint32_t s_warp[kSortWarps];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int i = tid; i < kSortWarps * 257; i += kSortThreads) (&s_whist[0][0])[i] = 0;
  __syncthreads();

  const uint64_t tile_base = (uint64_t)blockIdx.x * kSortTile;
  const uint64_t warp_base = tile_base + (uint64_t)wid * (32 * kSortItems);
  const uint64_t *kw = in.w[word];

  uint64_t key[kSortItems];
  uint32_t lp[kSortItems];  // rank within (warp, digit), then position in the tile
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const uint64_t idx = warp_base + (uint64_t)j * 32 + lane;
    key[j] = idx < n ? kw[idx] : ~0ull;
  }
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const uint64_t idx = warp_base + (uint64_t)j * 32 + lane;
    const uint32_t d = idx < n ? ((uint32_t)(key[j] >> shift) & mask) : 256u;
    const uint32_t peers = digit_peers(d);
    const uint32_t pre = s_whist[wid][d];
    __syncwarp();
    lp[j] = pre + __popc(peers & lt_mask);
    if ((peers & lt_mask) == 0) s_whist[wid][d] = pre + __popc(peers);
    __syncwarp();
  }
  __syncthreads();

  // per digit: exclusive prefix over warps, tile count (threads 0..255 <-> digits)
  uint32_t cnt = 0;
  if (tid < 256) {
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t c = s_whist[w][tid];
      s_whist[w][tid] = cnt;
      cnt += c;
    }
  }
  uint32_t total_valid;
  const uint32_t dstart = block_exclusive_scan_256<kSortWarps>(cnt, s_warp, total_valid);
  if (tid < 256) {
    s_dstart[tid] = dstart;
    s_goff[tid] = tile_off[(uint64_t)tid * ntiles + blockIdx.x] - dstart;  // dst = goff[d] + p (mod 2^32)
  }
  if (tid == 0) {
    s_dstart[256] = total_valid;
    uint32_t run = 0;
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t c = s_whist[w][256];
      s_whist[w][256] = run;
      run += c;
    }
  }
  __syncthreads();

#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const uint64_t idx = warp_base + (uint64_t)j * 32 + lane;
    const uint32_t d = idx < n ? ((uint32_t)(key[j] >> shift) & mask) : 256u;
    lp[j] += s_dstart[d] + s_whist[wid][d];
    s_key[lp[j]] = key[j];
    s_val[lp[j]] = val_in ? (idx < n ? val_in[idx] : 0u) : (uint32_t)idx;
  }
  __syncthreads();

#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const uint32_t p = (uint32_t)j * kSortThreads + tid;
    if (p < total_valid) {
      const uint64_t k = s_key[p];
      const uint32_t dst = s_goff[(uint32_t)(k >> shift) & mask] + p;
      out.w[word][dst] = k;
      val_out[dst] = s_val[p];
    }
  }
  if (NW > 1) {
    // remaining key words ride along through the same staging buffer; the
    // destination of slot p is recomputed from the ranking word kept in key[]
    uint32_t dst[kSortItems];
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
      const uint32_t p = (uint32_t)j * kSortThreads + tid;
      dst[j] = p < total_valid ? s_goff[(uint32_t)(s_key[p] >> shift) & mask] + p : 0u;
    }
#pragma unroll
    for (int w2 = 0; w2 < NW; ++w2) {
      if (w2 == word) continue;
      __syncthreads();
      const uint64_t *src = in.w[w2];
#pragma unroll
      for (int j = 0; j < kSortItems; ++j) {
        const uint64_t idx = warp_base + (uint64_t)j * 32 + lane;
        s_key[lp[j]] = idx < n ? src[idx] : 0ull;
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < kSortItems; ++j) {
        const uint32_t p = (uint32_t)j * kSortThreads + tid;
        if (p < total_valid) out.w[w2][dst[j]] = s_key[p];
      }
    }
  }
}

// ---- bucket grouping ------------------------------------------------------------
template <int NW>
__global__ void head_flags_kernel(KeyPtrs keys, uint64_t n, uint32_t *__restrict__ flags) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool head = (i == 0);
  if (!head) {
#pragma unroll
    for (int w = 0; w < NW; ++w) head = head || (keys.w[w][i] != keys.w[w][i - 1]);
  }
  flags[i] = head ? 1u : 0u;
}

template <int NW>
__global__ void bucket_scatter_kernel(KeyPtrs keys, uint64_t n, const uint32_t *__restrict__ flags,
                                      const uint32_t *__restrict__ scanned, uint64_t nb,
                                      uint64_t *__restrict__ ukeys /* [NW][nb] */,
                                      uint32_t *__restrict__ bstart /* [nb+1] */) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) bstart[nb] = (uint32_t)n;
  if (i >= n) return;
  if (flags[i]) {
    const uint32_t b = scanned[i];
    bstart[b] = (uint32_t)i;
#pragma unroll
    for (int w = 0; w < NW; ++w) ukeys[(uint64_t)w * nb + b] = keys.w[w][i];
  }
}

// ---- bucket-ordered, position-major code store ---------------------------------
// out[pos][i] = 4 * code[ids[i]][pos]; four consecutive i per thread so that stores are
// 32-bit and coalesced.  The codes come from the fragment records (one aligned 16-byte
// load per 16 residues, a single 32-byte sector per fragment at len <= 16) instead of
// `len` byte gathers.  ids == nullptr: identity order.
template <int NV>
__global__ void __launch_bounds__(256)
permute_rec_kernel(const uint8_t *__restrict__ rec, uint32_t RS, const uint32_t *__restrict__ ids, uint64_t n,
                   uint64_t npad, int len, uint8_t *__restrict__ out) {
  const uint64_t i4 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  uint32_t w[4][4 * NV];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const uint64_t i = i4 + m;
    if (i < n) {
      const uint64_t id = ids ? (uint64_t)__ldg(ids + i) : i;
      const uint4 *src = reinterpret_cast<const uint4 *>(rec + id * RS);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const uint4 r = __ldg(src + v);
        w[m][4 * v + 0] = r.x;
        w[m][4 * v + 1] = r.y;
        w[m][4 * v + 2] = r.z;
        w[m][4 * v + 3] = r.w;
      }
    } else {
#pragma unroll
      for (int v = 0; v < 4 * NV; ++v) w[m][v] = 0u;
    }
  }
#pragma unroll
  for (int pos = 0; pos < 16 * NV; ++pos) {
    if (pos < len) {
      uint32_t packed = 0;
#pragma unroll
      for (int m = 0; m < 4; ++m) packed |= ((w[m][pos >> 2] >> (8 * (pos & 3))) & 0xffu) << (8 * m);
      // codes are < 20: scaling all four bytes at once cannot carry across bytes
      *reinterpret_cast<uint32_t *>(out + (uint64_t)pos * npad + i4) = packed * (uint32_t)kCodeScale;
    }
  }
}

__global__ void iota_kernel(uint32_t *out, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint32_t)i;
}

// ---- host orchestration ---------------------------------------------------------
struct Pass {
  int word, shift;
  uint32_t mask;
};

static void plan_passes(const unsigned long long *or_and, int nw, std::vector<Pass> &passes) {
  for (int w = 0; w < nw; ++w) {
    uint64_t vary = or_and[2 * w] & ~or_and[2 * w + 1];
    int bit = 0;
    while (vary >> bit) {
      if (!((vary >> bit) & 1ull)) {
        ++bit;
        continue;
      }
      const int width = std::min(8, 64 - bit);
      passes.push_back({w, bit, (uint32_t)((1u << width) - 1u)});
      bit += width;
      if (bit >= 64) break;
    }
  }
}

template <int NW>
static void launch_downsweep(hs_ctx *ctx, const KeyPtrs &in, const uint32_t *vin, const KeyPtrs &out,
                             uint32_t *vout, uint64_t n, const Pass &p, const uint32_t *tile_off,
                             uint32_t ntiles) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(radix_downsweep_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)kDownsweepSmem);
    attr_set = true;
  }
  radix_downsweep_kernel<NW><<<ntiles, kSortThreads, kDownsweepSmem, ctx->stream>>>(
      in, vin, out, vout, n, p.word, p.shift, p.mask, tile_off, ntiles);
}

// Stable LSD sort of n (key, value) pairs.  keys_in: nw word arrays (not
// modified).  Values are the implicit index 0..n-1 when vals_in == nullptr.
// The sorted values land in v_final (v_tmp is the ping-pong partner); the
// sorted key words are left in the scratch (or in keys_in when no bit varies)
// and returned through *sorted_keys, valid until the next sort.
int radix_sort_pairs(hs_ctx *ctx, const KeyPtrs &keys_in, const uint32_t *vals_in, uint64_t n, int nw,
                     uint32_t *v_final, uint32_t *v_tmp, KeyPtrs *sorted_keys, uint64_t vary_hint) {
  SortScratch &S = ctx->sort;
  *sorted_keys = keys_in;
  if (n == 0) return HS_OK;
  for (int w = 0; w < nw; ++w) {
    HS_TRY(S.keys_cur[w].reserve(sizeof(uint64_t) * n));
    HS_TRY(S.keys_alt[w].reserve(sizeof(uint64_t) * n));
  }
  HS_TRY(S.or_and.reserve(sizeof(unsigned long long) * 2 * kMaxKeyWords));
  unsigned long long h_init[2 * kMaxKeyWords], h_oa[2 * kMaxKeyWords];
  for (int w = 0; w < kMaxKeyWords; ++w) {
    h_init[2 * w] = 0ull;
    h_init[2 * w + 1] = ~0ull;
  }
  unsigned long long *d_oa = S.or_and.as<unsigned long long>();
  if (vary_hint && nw == 1) {
    // the caller knows which bits of the (single) key word can vary: no reduction pass, no sync
    memcpy(h_oa, h_init, sizeof h_oa);
    h_oa[0] = vary_hint;
    h_oa[1] = 0ull;
  } else {
    HS_CUDA(cudaMemcpyAsync(d_oa, h_init, sizeof h_init, cudaMemcpyHostToDevice, ctx->stream));
    for (int w = 0; w < nw; ++w) {
      key_bits_kernel<<<148 * 4, 256, 0, ctx->stream>>>(keys_in.w[w], n, 1, d_oa + 2 * w);
      ctx->stats.kernel_launches++;
    }
    HS_TRY(read_back(ctx, d_oa, h_oa, sizeof h_oa));
  }
  std::vector<Pass> passes;
  plan_passes(h_oa, nw, passes);

  KeyPtrs A, B;
  for (int w = 0; w < kMaxKeyWords; ++w) {
    A.w[w] = w < nw ? S.keys_cur[w].as<uint64_t>() : nullptr;
    B.w[w] = w < nw ? S.keys_alt[w].as<uint64_t>() : nullptr;
  }
  if (passes.empty()) {
    if (vals_in) {
      HS_CUDA(cudaMemcpyAsync(v_final, vals_in, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
      iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(v_final, n);
      ctx->stats.kernel_launches++;
    }
    HS_CUDA(cudaGetLastError());
    return HS_OK;
  }
  const uint32_t ntiles = (uint32_t)((n + kSortTile - 1) / kSortTile);
  HS_TRY(S.tile_hist.reserve(sizeof(uint32_t) * ((uint64_t)ntiles * 256 + 1)));
  uint32_t *tile_hist = S.tile_hist.as<uint32_t>();

  KeyPtrs src = keys_in, dst = A;
  const uint32_t *vsrc = vals_in;
  // values ping-pong so that the last pass lands in v_final
  uint32_t *vdst = (passes.size() & 1) ? v_final : v_tmp;
  while (ctx->ev_pool.size() < 4 * passes.size()) {
    cudaEvent_t e;
    HS_CUDA(cudaEventCreate(&e));
    ctx->ev_pool.push_back(e);
  }
  for (size_t pi = 0; pi < passes.size(); ++pi) {
    const Pass &p = passes[pi];
    cudaEvent_t *pe = &ctx->ev_pool[4 * pi];
    HS_CUDA(cudaEventRecord(pe[0], ctx->stream));
    radix_upsweep_kernel<<<ntiles, kSortThreads, 0, ctx->stream>>>(src.w[p.word], n, p.shift, p.mask, tile_hist,
                                                                  ntiles);
    ctx->stats.kernel_launches++;
    HS_CUDA(cudaEventRecord(pe[1], ctx->stream));
    HS_TRY(exclusive_scan_u32(ctx, tile_hist, tile_hist, (uint64_t)ntiles * 256, nullptr));
    HS_CUDA(cudaEventRecord(pe[2], ctx->stream));
    switch (nw) {
      case 1: launch_downsweep<1>(ctx, src, vsrc, dst, vdst, n, p, tile_hist, ntiles); break;
      case 2: launch_downsweep<2>(ctx, src, vsrc, dst, vdst, n, p, tile_hist, ntiles); break;
      case 3: launch_downsweep<3>(ctx, src, vsrc, dst, vdst, n, p, tile_hist, ntiles); break;
      default: launch_downsweep<4>(ctx, src, vsrc, dst, vdst, n, p, tile_hist, ntiles); break;
    }
    ctx->stats.kernel_launches++;
    ctx->stats.sort_passes++;
    HS_CUDA(cudaGetLastError());
    HS_CUDA(cudaEventRecord(pe[3], ctx->stream));
    vsrc = vdst;
    vdst = (vdst == v_final) ? v_tmp : v_final;
    src = dst;
    dst = (dst.w[0] == A.w[0]) ? B : A;
  }
  *sorted_keys = src;
  HS_CUDA(cudaEventSynchronize(ctx->ev_pool[4 * passes.size() - 1]));
  for (size_t pi = 0; pi < passes.size(); ++pi) {
    cudaEvent_t *pe = &ctx->ev_pool[4 * pi];
    float a = 0.f, b = 0.f, c = 0.f;
    cudaEventElapsedTime(&a, pe[0], pe[1]);
    cudaEventElapsedTime(&b, pe[1], pe[2]);
    cudaEventElapsedTime(&c, pe[2], pe[3]);
    ctx->stats.ms_sort_upsweep += a;
    ctx->stats.ms_sort_scan += b;
    ctx->stats.ms_sort_downsweep += c;
  }
  return HS_OK;
}

// Number of radix passes a sort of these keys on all their words would take (one reduction pass
// over the keys: which bits vary).
static int count_full_passes(hs_ctx *ctx, const KeyPtrs &keys_in, uint64_t n, int nw, size_t *npasses) {
  SortScratch &S = ctx->sort;
  HS_TRY(S.or_and.reserve(sizeof(unsigned long long) * 2 * kMaxKeyWords));
  unsigned long long h_init[2 * kMaxKeyWords], h_oa[2 * kMaxKeyWords];
  for (int w = 0; w < kMaxKeyWords; ++w) {
    h_init[2 * w] = 0ull;
    h_init[2 * w + 1] = ~0ull;
  }
  unsigned long long *d_oa = S.or_and.as<unsigned long long>();
  HS_TRY(upload(ctx, d_oa, h_init, sizeof h_init));
  for (int w = 0; w < nw; ++w) {
    key_bits_kernel<<<148 * 4, 256, 0, ctx->stream>>>(keys_in.w[w], n, 1, d_oa + 2 * w);
    ctx->stats.kernel_launches++;
  }
  HS_TRY(read_back(ctx, d_oa, h_oa, sizeof h_oa));
  std::vector<Pass> passes;
  plan_passes(h_oa, nw, passes);
  *npasses = passes.size();
  return HS_OK;
}

static int sort_table(hs_ctx *ctx, uint32_t table, KeyPtrs *sorted_keys) {
  const uint64_t n = ctx->N;
  const int nw = (int)ctx->key_words;
  TableIndex &T = ctx->tables[table];
  HS_TRY(T.sorted_ids.reserve(sizeof(uint32_t) * n));
  HS_TRY(ctx->sort.vals_alt.reserve(sizeof(uint32_t) * n));
  KeyPtrs in;
  for (int w = 0; w < kMaxKeyWords; ++w)
    in.w[w] = w < nw ? ctx->d_keys[table].as<uint64_t>() + (uint64_t)w * n : nullptr;
  return radix_sort_pairs(ctx, in, nullptr, n, nw, T.sorted_ids.as<uint32_t>(), ctx->sort.vals_alt.as<uint32_t>(),
                          sorted_keys);
}

template <int NW>
static int group_inst(hs_ctx *ctx, uint32_t table, const KeyPtrs &keys) {
  const uint64_t n = ctx->N;
  SortScratch &S = ctx->sort;
  TableIndex &T = ctx->tables[table];
  HS_TRY(S.flags.reserve(sizeof(uint32_t) * 2 * n + 16));
  uint32_t *flags = S.flags.as<uint32_t>();
  uint32_t *scanned = flags + n;
  HS_TRY(S.or_and.reserve(sizeof(unsigned long long) * 2 * kMaxKeyWords));
  uint32_t *d_total = S.or_and.as<uint32_t>();
  const unsigned grid = (unsigned)((n + 255) / 256);
  head_flags_kernel<NW><<<grid, 256, 0, ctx->stream>>>(keys, n, flags);
  ctx->stats.kernel_launches++;
  HS_TRY(exclusive_scan_u32(ctx, flags, scanned, n, d_total));
  uint32_t nb = 0;
  HS_TRY(read_back(ctx, d_total, &nb, sizeof nb));
  T.nb = nb;
  T.nslots = nb;
  HS_TRY(T.ukeys.reserve(sizeof(uint64_t) * NW * (uint64_t)std::max<uint32_t>(nb, 1)));
  HS_TRY(T.bstart.reserve(sizeof(uint32_t) * ((uint64_t)nb + 1)));
  bucket_scatter_kernel<NW><<<grid, 256, 0, ctx->stream>>>(keys, n, flags, scanned, nb, T.ukeys.as<uint64_t>(),
                                                         T.bstart.as<uint32_t>());
  ctx->stats.kernel_launches++;
  HS_CUDA(cudaGetLastError());
  return HS_OK;
}

int build_code_store(hs_ctx *ctx, const uint32_t *ids, DevBuf &out) {
  const uint64_t n = ctx->N;
  const size_t bytes = (size_t)ctx->prm.len * ctx->npad + 256;  // slack: 128-byte tile reads may overrun the last row
  if (out.cap < bytes) {
    HS_TRY(out.reserve(bytes));
    HS_CUDA(cudaMemsetAsync(out.p, 0, out.cap, ctx->stream));
  }
  HS_TRY(ensure_records(ctx));
  const uint64_t nthreads = (n + 3) / 4;
  const unsigned grid = (unsigned)((nthreads + 255) / 256);
  if (ctx->prm.len <= 16)
    permute_rec_kernel<1><<<grid, 256, 0, ctx->stream>>>(ctx->d_rec.as<uint8_t>(), ctx->rec_stride, ids, n, ctx->npad,
                                                         (int)ctx->prm.len, out.as<uint8_t>());
  else
    permute_rec_kernel<2><<<grid, 256, 0, ctx->stream>>>(ctx->d_rec.as<uint8_t>(), ctx->rec_stride, ids, n, ctx->npad,
                                                         (int)ctx->prm.len, out.as<uint8_t>());
  ctx->stats.kernel_launches++;
  HS_CUDA(cudaGetLastError());
  return HS_OK;
}

// ---- bucket-order code stores of all tables in one L2-blocked gather -------------------------
// A random 32-byte record gather costs 128 bytes of DRAM traffic on B200 (ncu), and every
// table gathers the same records.  Members of a bucket are in ascending id order, so the part
// of a bucket whose ids fall into one block of the record array is one contiguous run of the
// bucket.  The gather therefore walks the record array block by block (16 MB: L2-resident) and,
// inside a block, all tables and all buckets, so that a record line fetched from DRAM is
// served from L2 to several of the ~16 gathers that want it (4 records per line x L tables;
// ncu: DRAM reads 52.9 GB -> 18.8 GB at L = 4).
constexpr uint32_t kGatherBlockBytes = 16u << 20;  // measured best on B200 (8-16 MB; 4 and 32 MB are slower)
constexpr int kGatherSlots = 256;     // slots per thread block
#ifndef HS_GATHER_PART
#define HS_GATHER_PART 2048  // measured at 100 M fragments: 512 -> 10.9, 1024 -> 10.1, 2048 -> 5.4, 4096 -> 6.0, 8192 -> 7.8 ms
#endif
constexpr uint32_t kGatherPart = HS_GATHER_PART;  // members per slot at most
constexpr int kGatherThreads = 256;

struct GatherTab {
  const uint32_t *ids;     // [N] bucket order
  const uint32_t *bstart;  // [nslots + 1]
  uint8_t *out;            // [len][npad]
  uint32_t nslots;         // slots: non-empty buckets, cut into parts of <= kGatherPart members
  uint32_t ngroups;        // thread blocks of this table per record block
  uint32_t groups_before;  // slot groups of the tables before this one
  uint64_t run_off;        // offset of this table's run boundaries in the runs array
};

// runs[c * nslots + b] = first position of slot b whose id is >= c * chunk (c = 0..nchunks).
// One thread per slot: its ids are ascending, so only the record-block boundaries between its
// first and its last id need a search.
__global__ void gather_runs_kernel(const GatherTab *__restrict__ tabs, uint32_t ntab, uint32_t nchunks, uint32_t chunk,
                                   uint32_t *__restrict__ runs) {
  const uint32_t gg = blockIdx.x;
  uint32_t l = 0;
  while (l + 1 < ntab && tabs[l + 1].groups_before <= gg) ++l;
  const GatherTab T = tabs[l];
  const uint32_t b = (gg - T.groups_before) * kGatherSlots + threadIdx.x;
  if (b >= T.nslots) return;
  const uint32_t lo0 = T.bstart[b], hi0 = T.bstart[b + 1];
  const uint32_t c_first = hi0 > lo0 ? T.ids[lo0] / chunk : 0u;
  const uint32_t c_last = hi0 > lo0 ? T.ids[hi0 - 1] / chunk : 0u;
  uint32_t *out = runs + T.run_off + b;
  uint32_t lo = lo0;
  for (uint32_t c = 0; c <= nchunks; ++c) {
    uint32_t r;
    if (c <= c_first) r = lo0;
    else if (c > c_last) r = hi0;
    else {
      const uint32_t key = c * chunk;  // (c <= c_last: c * chunk <= last id < 2^32)
      uint32_t hi = hi0;
      while (lo < hi) {  // lower bound, continuing from the previous boundary
        const uint32_t m = lo + ((hi - lo) >> 1);
        if (T.ids[m] < key) lo = m + 1; else hi = m;
      }
      r = lo;
    }
    out[(uint64_t)c * T.nslots] = r;
  }
}

template <int NV>
__global__ void __launch_bounds__(kGatherThreads)
gather_blocked_kernel(const GatherTab *__restrict__ tabs, uint32_t ntab, uint32_t total_groups,
                      const uint32_t *__restrict__ runs, const uint8_t *__restrict__ rec, uint32_t RS, uint64_t npad,
                      int len) {
  __shared__ uint32_t s_start[kGatherSlots];
  __shared__ uint32_t s_pref[kGatherSlots + 1];
  __shared__ uint32_t s_warp[kGatherThreads / 32];
  // blockIdx.x = c * total_groups + (table, slot group): all work of record block c is adjacent
  const uint32_t c = blockIdx.x / total_groups, gg = blockIdx.x - c * total_groups;
  uint32_t l = 0;
  while (l + 1 < ntab && tabs[l + 1].groups_before <= gg) ++l;
  const GatherTab T = tabs[l];
  const uint32_t b = (gg - T.groups_before) + threadIdx.x * T.ngroups;
  uint32_t st = 0, n = 0;
  if (b < T.nslots) {
    st = runs[T.run_off + (uint64_t)c * T.nslots + b];
    n = runs[T.run_off + (uint64_t)(c + 1) * T.nslots + b] - st;
  }
  uint32_t total;
  const uint32_t ex = block_exclusive_scan_256<kGatherThreads / 32>(n, s_warp, total);
  s_start[threadIdx.x] = st;
  s_pref[threadIdx.x] = ex;
  if (threadIdx.x == 0) s_pref[kGatherSlots] = total;
  __syncthreads();
  for (uint32_t e = threadIdx.x; e < total; e += kGatherThreads) {
    // slot of element e: last slot with prefix <= e
    uint32_t lo = 0, hi = kGatherSlots;
    while (hi - lo > 1) {
      const uint32_t m = (lo + hi) >> 1;
      if (s_pref[m] <= e) lo = m; else hi = m;
    }
    const uint32_t j = s_start[lo] + (e - s_pref[lo]);
    const uint32_t id = __ldg(T.ids + j);
    const uint4 *src = reinterpret_cast<const uint4 *>(rec + (uint64_t)id * RS);
    uint32_t w[4 * NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const uint4 r = __ldg(src + v);
      w[4 * v + 0] = r.x;
      w[4 * v + 1] = r.y;
      w[4 * v + 2] = r.z;
      w[4 * v + 3] = r.w;
    }
#pragma unroll
    for (int pos = 0; pos < 16 * NV; ++pos)
      if (pos < len) T.out[(uint64_t)pos * npad + j] = (uint8_t)(((w[pos >> 2] >> (8 * (pos & 3))) & 0xffu) * kCodeScale);
  }
}

// Builds codes_sorted of every table.  Returns HS_OK with *done = false when the blocked
// gather does not apply (small DB, or too many bucket slots for the run table).
int build_code_stores_blocked(hs_ctx *ctx, bool *done) {
  *done = false;
  const uint64_t n = ctx->N;
  const uint32_t L = ctx->prm.L, RS = ctx->rec_stride;
  const char *e = getenv("HS_NO_BLOCKED_GATHER");
  if ((e && atoi(e)) || n * RS < 4ull * kGatherBlockBytes) return HS_OK;
  uint32_t block_bytes = kGatherBlockBytes;
  if (const char *m = getenv("HS_GATHER_MB")) block_bytes = (uint32_t)std::max(1, atoi(m)) << 20;
  const uint32_t chunk = block_bytes / RS;
  const uint32_t nchunks = (uint32_t)((n + chunk - 1) / chunk);
  std::vector<GatherTab> tabs(L);
  uint64_t run_total = 0;
  uint32_t groups = 0;
  const size_t bytes = (size_t)ctx->prm.len * ctx->npad + 256;
  // slot boundaries per table: the non-empty buckets, cut into parts of <= kGatherPart members
  std::vector<uint32_t> vstart, bs;
  std::vector<uint64_t> voff(L);
  for (uint32_t l = 0; l < L; ++l) {
    TableIndex &T = ctx->tables[l];
    // many small buckets (small W / large K): the slot table would be built on the host from
    // millions of boundaries, and small buckets gain nothing from the blocking
    if (T.nslots > (1u << 18) || (T.nslots + n / kGatherPart) * (uint64_t)(nchunks + 1) > (64ull << 20)) return HS_OK;
    bs.resize(T.nslots + 1);
    HS_TRY(read_back(ctx, T.bstart.p, bs.data(), sizeof(uint32_t) * (T.nslots + 1)));
    voff[l] = vstart.size();
    for (uint64_t b = 0; b < T.nslots; ++b)
      for (uint32_t p = bs[b]; p < bs[b + 1]; p += kGatherPart) vstart.push_back(p);
    vstart.push_back((uint32_t)n);
  }
  HS_TRY(ctx->sort.block_sums.reserve(sizeof(uint32_t) * vstart.size() + 16));
  HS_TRY(upload(ctx, ctx->sort.block_sums.p, vstart.data(), sizeof(uint32_t) * vstart.size()));
  for (uint32_t l = 0; l < L; ++l) {
    TableIndex &T = ctx->tables[l];
    if (T.codes_sorted.cap < bytes) {
      HS_TRY(T.codes_sorted.reserve(bytes));
      HS_CUDA(cudaMemsetAsync(T.codes_sorted.p, 0, T.codes_sorted.cap, ctx->stream));
    }
    const uint64_t nv = (l + 1 < L ? voff[l + 1] : vstart.size()) - voff[l] - 1;
    tabs[l].ids = T.sorted_ids.as<uint32_t>();
    tabs[l].bstart = ctx->sort.block_sums.as<uint32_t>() + voff[l];
    tabs[l].out = T.codes_sorted.as<uint8_t>();
    tabs[l].nslots = (uint32_t)nv;
    tabs[l].ngroups = (uint32_t)((nv + kGatherSlots - 1) / kGatherSlots);
    tabs[l].groups_before = groups;
    tabs[l].run_off = run_total;
    groups += tabs[l].ngroups;
    run_total += (uint64_t)(nchunks + 1) * nv;
  }
  if (groups == 0 || run_total > (64ull << 20) || (uint64_t)groups * nchunks > 0x7fffffffull) return HS_OK;
  HS_TRY(ensure_records(ctx));
  HS_TRY(ctx->sort.flags.reserve(sizeof(uint32_t) * run_total + 16));
  HS_TRY(ctx->d_misc.reserve(sizeof(GatherTab) * L));
  HS_TRY(upload(ctx, ctx->d_misc.p, tabs.data(), sizeof(GatherTab) * L));
  const GatherTab *d_tabs = ctx->d_misc.as<GatherTab>();
  uint32_t *runs = ctx->sort.flags.as<uint32_t>();
  gather_runs_kernel<<<groups, kGatherSlots, 0, ctx->stream>>>(d_tabs, L, nchunks, chunk, runs);
  if (ctx->prm.len <= 16)
    gather_blocked_kernel<1><<<groups * nchunks, kGatherThreads, 0, ctx->stream>>>(d_tabs, L, groups, runs, ctx->d_rec.as<uint8_t>(), RS, ctx->npad, (int)ctx->prm.len);
  else
    gather_blocked_kernel<2><<<groups * nchunks, kGatherThreads, 0, ctx->stream>>>(d_tabs, L, groups, runs, ctx->d_rec.as<uint8_t>(), RS, ctx->npad, (int)ctx->prm.len);
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches += 2;
  *done = true;
  return HS_OK;
}

// ---- rank path: sort of u16 bucket ranks ----------------------------------------
// Same upsweep / scan / downsweep structure as above, specialised for 16-bit keys:
// at most two 8-bit passes, the first takes the implicit index as value, the last
// writes only the ids.
constexpr int kRkThreads = 256;
constexpr int kRkItems = 16;
constexpr int kRkTile = kRkThreads * kRkItems;  // 4096 ranks per tile
constexpr int kRkWarps = kRkThreads / 32;

__global__ void __launch_bounds__(kRkThreads)
rank_upsweep_kernel(const uint16_t *__restrict__ keys, uint64_t n, int shift, uint32_t mask,
                    uint32_t *__restrict__ tile_hist, uint32_t ntiles) {
  __shared__ uint32_t s_hist[256];
  const int tid = threadIdx.x;
  s_hist[tid] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * kRkTile;
  if (base + kRkTile <= n) {
    // whole tile: two 16-byte loads (8 ranks each) per thread
    const uint4 *src = reinterpret_cast<const uint4 *>(keys + base);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint4 v = __ldg(src + j * kRkThreads + tid);
      const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        atomicAdd(&s_hist[(ww[q] >> shift) & mask], 1u);
        atomicAdd(&s_hist[(ww[q] >> (16 + shift)) & mask], 1u);
      }
    }
  } else {
    for (int j = 0; j < kRkItems; ++j) {
      const uint64_t idx = base + (uint64_t)j * kRkThreads + tid;
      if (idx < n) atomicAdd(&s_hist[((uint32_t)keys[idx] >> shift) & mask], 1u);
    }
  }
  __syncthreads();
  tile_hist[(uint64_t)tid * ntiles + blockIdx.x] = s_hist[tid];
}

template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(kRkThreads, 4)
rank_downsweep_kernel(const uint16_t *__restrict__ key_in, const uint32_t *__restrict__ val_in,
                      uint16_t *__restrict__ key_out, uint32_t *__restrict__ val_out, uint64_t n, int shift,
                      uint32_t mask, const uint32_t *__restrict__ tile_off, uint32_t ntiles) {
  __shared__ uint32_t s_val[kRkTile];            // 16 KB
  __shared__ uint16_t s_key[kRkTile];            // 8 KB
  __shared__ uint32_t s_whist[kRkWarps][257];    // per-warp digit counters (+ tail bin)
  __shared__ uint32_t s_dstart[257];
  __shared__ uint32_t s_goff[256];
  __shared__ uint32_t s_warp[kRkWarps];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int i = tid; i < kRkWarps * 257; i += kRkThreads) (&s_whist[0][0])[i] = 0;
  __syncthreads();

  const uint64_t tile_base = (uint64_t)blockIdx.x * kRkTile;
  const uint64_t warp_base = tile_base + (uint64_t)wid * (32 * kRkItems);

  uint32_t key[kRkItems];
  uint32_t lp[kRkItems];
#pragma unroll
  for (int j = 0; j < kRkItems; ++j) {
    const uint64_t idx = warp_base + (uint64_t)j * 32 + lane;
    key[j] = idx < n ? (uint32_t)key_in[idx] : 0xffffffffu;
  }
#pragma unroll
  for (int j = 0; j < kRkItems; ++j) {
    const uint32_t d = key[j] != 0xffffffffu ? ((key[j] >> shift) & mask) : 256u;
    const uint32_t peers = digit_peers(d);
    const uint32_t pre = s_whist[wid][d];
    __syncwarp();
    lp[j] = pre + __popc(peers & lt_mask);
    if ((peers & lt_mask) == 0) s_whist[wid][d] = pre + __popc(peers);
    __syncwarp();
  }
  __syncthreads();

  uint32_t cnt = 0;
#pragma unroll
  for (int w = 0; w < kRkWarps; ++w) {
    const uint32_t c = s_whist[w][tid];
    s_whist[w][tid] = cnt;
    cnt += c;
  }
  uint32_t total_valid;
  const uint32_t dstart = block_exclusive_scan_256<kRkWarps>(cnt, s_warp, total_valid);
  s_dstart[tid] = dstart;
  s_goff[tid] = tile_off[(uint64_t)tid * ntiles + blockIdx.x] - dstart;  // dst = goff[d] + p (mod 2^32)
  if (tid == 0) {
    s_dstart[256] = total_valid;
    uint32_t run = 0;
    for (int w = 0; w < kRkWarps; ++w) {
      const uint32_t c = s_whist[w][256];
      s_whist[w][256] = run;
      run += c;
    }
  }
  __syncthreads();

#pragma unroll
  for (int j = 0; j < kRkItems; ++j) {
    const uint64_t idx = warp_base + (uint64_t)j * 32 + lane;
    const uint32_t d = key[j] != 0xffffffffu ? ((key[j] >> shift) & mask) : 256u;
    const uint32_t p = lp[j] + s_dstart[d] + s_whist[wid][d];
    s_key[p] = (uint16_t)key[j];
    s_val[p] = FIRST ? (uint32_t)idx : (idx < n ? val_in[idx] : 0u);
  }
  __syncthreads();

#pragma unroll
  for (int j = 0; j < kRkItems; ++j) {
    const uint32_t p = (uint32_t)j * kRkThreads + tid;
    if (p < total_valid) {
      const uint32_t k = s_key[p];
      const uint32_t dst = s_goff[(k >> shift) & mask] + p;
      key_out[dst] = (uint16_t)k;  // (the last pass's sorted ranks give the bucket boundaries)
      val_out[dst] = s_val[p];
    }
  }
}

// Bucket boundaries from the sorted ranks: slot r starts at the first position whose rank
// is >= r (empty slots get start == end); *nb counts the non-empty slots.  Eight ranks per
// thread (one 16-byte load).
__global__ void __launch_bounds__(256)
rank_bounds_kernel(const uint16_t *__restrict__ sorted, uint64_t n, uint32_t nr, uint32_t *__restrict__ bstart,
                   unsigned int *__restrict__ nb) {
  const uint64_t i8 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  unsigned int heads = 0;
  if (i8 < n) {
    uint32_t k[8];
    if (i8 + 8 <= n) {
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(sorted + i8));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        k[2 * j] = w[j] & 0xffffu;
        k[2 * j + 1] = w[j] >> 16;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) k[j] = i8 + j < n ? (uint32_t)sorted[i8 + j] : 0u;
    }
    int prev = i8 ? (int)sorted[i8 - 1] : -1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (i8 + j < n) {
        const int cur = (int)k[j];
        if (cur != prev) {
          ++heads;
          for (int r = prev + 1; r <= cur; ++r) bstart[r] = (uint32_t)(i8 + j);
          prev = cur;
        }
      }
    }
    if (i8 + 8 >= n)
      for (uint32_t r = (uint32_t)prev + 1; r <= nr; ++r) bstart[r] = (uint32_t)n;
  }
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) heads += __shfl_xor_sync(0xffffffffu, heads, sft);
  if (heads && (threadIdx.x & 31) == 0) atomicAdd(nb, heads);
}

// Rank path of build_table_index: two-pass (or one-pass) sort of the table's u16 ranks,
// bucket boundaries from the rank histogram (one slot per possible key string).
static int build_table_index_ranks(hs_ctx *ctx, uint32_t table, cudaEvent_t ev_sort_end, cudaEvent_t ev_group_end,
                                   bool with_store) {
  const uint64_t n = ctx->N;
  SortScratch &S = ctx->sort;
  TableIndex &T = ctx->tables[table];
  const uint32_t nr = ctx->rank_nr[table];
  const uint32_t KW = ctx->key_words;
  const uint16_t *ranks = ctx->d_ranks.as<uint16_t>() + (size_t)table * ctx->npad;
  HS_TRY(T.sorted_ids.reserve(sizeof(uint32_t) * n));
  HS_TRY(S.vals_alt.reserve(sizeof(uint32_t) * n));
  HS_TRY(S.keys_alt[0].reserve(sizeof(uint16_t) * n + 16));
  HS_TRY(S.keys_cur[0].reserve(sizeof(uint16_t) * n + 16));
  const uint32_t ntiles = (uint32_t)((n + kRkTile - 1) / kRkTile);
  HS_TRY(S.tile_hist.reserve(sizeof(uint32_t) * ((uint64_t)ntiles * 256 + 1)));
  uint32_t *tile_hist = S.tile_hist.as<uint32_t>();
  uint32_t *ids = T.sorted_ids.as<uint32_t>();
  uint32_t *vtmp = S.vals_alt.as<uint32_t>();
  uint16_t *ktmp = S.keys_alt[0].as<uint16_t>();
  uint16_t *ksorted = S.keys_cur[0].as<uint16_t>();
  while (ctx->ev_pool.size() < 8) {
    cudaEvent_t e;
    HS_CUDA(cudaEventCreate(&e));
    ctx->ev_pool.push_back(e);
  }
  int hibits = 0;
  while (((uint64_t)256 << hibits) < nr) ++hibits;
  const int npass = nr > 256 ? 2 : 1;
  for (int pi = 0; pi < npass; ++pi) {
    cudaEvent_t *pe = &ctx->ev_pool[4 * pi];
    const int shift = 8 * pi;
    const uint32_t mask = pi == 0 ? 0xffu : ((1u << hibits) - 1u);
    const uint16_t *kin = pi == 0 ? ranks : ktmp;
    HS_CUDA(cudaEventRecord(pe[0], ctx->stream));
    rank_upsweep_kernel<<<ntiles, kRkThreads, 0, ctx->stream>>>(kin, n, shift, mask, tile_hist, ntiles);
    ctx->stats.kernel_launches++;
    HS_CUDA(cudaEventRecord(pe[1], ctx->stream));
    HS_TRY(exclusive_scan_u32(ctx, tile_hist, tile_hist, (uint64_t)ntiles * 256, nullptr));
    HS_CUDA(cudaEventRecord(pe[2], ctx->stream));
    if (npass == 1)
      rank_downsweep_kernel<true, true><<<ntiles, kRkThreads, 0, ctx->stream>>>(kin, nullptr, ksorted, ids, n, shift, mask, tile_hist, ntiles);
    else if (pi == 0)
      rank_downsweep_kernel<true, false><<<ntiles, kRkThreads, 0, ctx->stream>>>(kin, nullptr, ktmp, vtmp, n, shift, mask, tile_hist, ntiles);
    else
      rank_downsweep_kernel<false, true><<<ntiles, kRkThreads, 0, ctx->stream>>>(kin, vtmp, ksorted, ids, n, shift, mask, tile_hist, ntiles);
    ctx->stats.kernel_launches++;
    ctx->stats.sort_passes++;
    HS_CUDA(cudaGetLastError());
    HS_CUDA(cudaEventRecord(pe[3], ctx->stream));
  }
  HS_CUDA(cudaEventRecord(ev_sort_end, ctx->stream));

  // bucket boundaries from the sorted ranks
  HS_TRY(T.bstart.reserve(sizeof(uint32_t) * ((uint64_t)nr + 1)));
  HS_TRY(S.or_and.reserve(sizeof(unsigned long long) * 2 * kMaxKeyWords));
  unsigned int *d_nb = S.or_and.as<unsigned int>();
  HS_CUDA(cudaMemsetAsync(d_nb, 0, sizeof(unsigned int), ctx->stream));
  rank_bounds_kernel<<<(unsigned)(((n + 7) / 8 + 255) / 256), 256, 0, ctx->stream>>>(ksorted, n, nr, T.bstart.as<uint32_t>(), d_nb);
  ctx->stats.kernel_launches++;
  HS_TRY(T.ukeys.reserve(sizeof(uint64_t) * KW * (uint64_t)nr));
  HS_TRY(upload(ctx, T.ukeys.p, ctx->h_rkeys[table].data(), sizeof(uint64_t) * KW * nr));
  unsigned int h_nb = 0;
  HS_CUDA(cudaEventRecord(ev_group_end, ctx->stream));
  HS_TRY(read_back(ctx, d_nb, &h_nb, sizeof h_nb));
  T.nb = h_nb;
  T.nslots = nr;
  for (int pi = 0; pi < npass; ++pi) {
    cudaEvent_t *pe = &ctx->ev_pool[4 * pi];
    float a = 0.f, b = 0.f, c = 0.f;
    cudaEventElapsedTime(&a, pe[0], pe[1]);
    cudaEventElapsedTime(&b, pe[1], pe[2]);
    cudaEventElapsedTime(&c, pe[2], pe[3]);
    ctx->stats.ms_sort_upsweep += a;
    ctx->stats.ms_sort_scan += b;
    ctx->stats.ms_sort_downsweep += c;
  }
  if (with_store) HS_TRY(build_code_store(ctx, ids, T.codes_sorted));
  return HS_OK;
}

// ---- hashed-key path: multi-word keys (K >= 8 or a small W: 2-4 words, up to 16 radix passes per table
// over all words) are sorted on a 64-bit hash of the key instead: 8 passes over one word.  The reference's
// HashTable is an unordered_map (motif_both_points.cpp:212-218, lsh.hpp:51-59): the order of the buckets
// is immaterial, a bucket is the ascending id list of the fragments sharing a key string.  Equal keys
// hash equally, so after the stable sort every hash group is in ascending id order; a group is a bucket
// iff all its members carry the same full key, which hashed_buckets_kernel checks member by member
// against the group's head (any mismatch -> the caller falls back to the sort on all words).
template <int NW>
__global__ void key_hash_kernel(KeyPtrs keys, uint64_t n, uint64_t *__restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t k[NW];
#pragma unroll
  for (int w = 0; w < NW; ++w) k[w] = keys.w[w][i];
  out[i] = key_hash<NW>(k);
}

// Every member that is not the head of its hash group compares its full key with the head's (the
// heads' keys are re-read by the whole group: cached); singletons -- nearly all groups at large K --
// cost nothing.  The probe fetches a slot's full key through the head's id (verify.cu).
template <int NW>
__global__ void hashed_buckets_kernel(KeyPtrs keys /* by fragment id */, uint64_t n, const uint32_t *__restrict__ ids,
                                      const uint32_t *__restrict__ flags, const uint32_t *__restrict__ scanned,
                                      const uint32_t *__restrict__ bstart, unsigned int *__restrict__ mismatches) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || flags[i]) return;
  const uint32_t id = ids[i];
  const uint32_t head = ids[bstart[scanned[i] - 1]];
  bool same = true;
#pragma unroll
  for (int w = 0; w < NW; ++w) same = same && (keys.w[w][id] == keys.w[w][head]);
  if (!same) atomicAdd(mismatches, 1u);
}

template <int NW>
static int build_table_index_hashed(hs_ctx *ctx, uint32_t table, cudaEvent_t ev_sort_end, bool *ok, bool *declined) {
  const uint64_t n = ctx->N;
  SortScratch &S = ctx->sort;
  TableIndex &T = ctx->tables[table];
  *ok = false;
  *declined = false;
  HS_TRY(T.sorted_ids.reserve(sizeof(uint32_t) * n));
  HS_TRY(S.vals_alt.reserve(sizeof(uint32_t) * n));
  HS_TRY(S.keys_cur[1].reserve(sizeof(uint64_t) * n));
  KeyPtrs in;
  for (int w = 0; w < kMaxKeyWords; ++w)
    in.w[w] = w < NW ? ctx->d_keys[table].as<uint64_t>() + (uint64_t)w * n : nullptr;
  // worth it?  bytes moved per key: P passes over NW words + id against 8 passes over one word + id,
  // plus the hash pass and the full-key check (two gathers of NW words)
  // (decided on the first table of a build: the tables' keys are statistically alike, and the
  // reduction pass reads every key word)
  if (ctx->hash_sort_choice < 0) {
    size_t full_passes = 0;
    HS_TRY(count_full_passes(ctx, in, n, NW, &full_passes));
    ctx->hash_sort_choice = (ctx->force_hash_sort || full_passes * (8 * NW + 4) > 8 * 12 + 8 * NW + 64 * NW) ? 1 : 0;
  }
  if (!ctx->hash_sort_choice) {
    *declined = true;
    return HS_OK;
  }
  const unsigned grid = (unsigned)((n + 255) / 256);
  key_hash_kernel<NW><<<grid, 256, 0, ctx->stream>>>(in, n, S.keys_cur[1].as<uint64_t>());
  ctx->stats.kernel_launches++;
  KeyPtrs hin, hsorted;
  for (int w = 0; w < kMaxKeyWords; ++w) hin.w[w] = nullptr;
  hin.w[0] = S.keys_cur[1].as<uint64_t>();
  HS_TRY(radix_sort_pairs(ctx, hin, nullptr, n, 1, T.sorted_ids.as<uint32_t>(), S.vals_alt.as<uint32_t>(), &hsorted, ~0ull));
  HS_CUDA(cudaEventRecord(ev_sort_end, ctx->stream));
  HS_TRY(group_inst<1>(ctx, table, hsorted));   // ukeys = the ascending hashes, bstart
  unsigned int *mism = S.or_and.as<unsigned int>() + 2;
  HS_CUDA(cudaMemsetAsync(mism, 0, sizeof(unsigned int), ctx->stream));
  uint32_t *flags = S.flags.as<uint32_t>();
  hashed_buckets_kernel<NW><<<grid, 256, 0, ctx->stream>>>(in, n, T.sorted_ids.as<uint32_t>(), flags, flags + n,
                                                          T.bstart.as<uint32_t>(), mism);
  ctx->stats.kernel_launches++;
  HS_CUDA(cudaGetLastError());
  unsigned int h_mism = 0;
  HS_TRY(read_back(ctx, mism, &h_mism, sizeof h_mism));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  *ok = h_mism == 0 && !ctx->force_hash_collision;
  T.hashed_keys = *ok;
  return HS_OK;
}

// with_store == false: the caller builds the code stores of all tables afterwards
// (build_code_stores_blocked, or build_table_store per table).
int build_table_index(hs_ctx *ctx, uint32_t table, cudaEvent_t ev_sort_end, cudaEvent_t ev_group_end, bool with_store) {
  if (ctx->rank_mode) return build_table_index_ranks(ctx, table, ev_sort_end, ev_group_end, with_store);
  ctx->tables[table].hashed_keys = false;
  if (ctx->key_words >= 2 && !ctx->no_hash_sort && ctx->N) {
    bool ok = false, declined = false;
    switch (ctx->key_words) {
      case 2: HS_TRY(build_table_index_hashed<2>(ctx, table, ev_sort_end, &ok, &declined)); break;
      case 3: HS_TRY(build_table_index_hashed<3>(ctx, table, ev_sort_end, &ok, &declined)); break;
      default: HS_TRY(build_table_index_hashed<4>(ctx, table, ev_sort_end, &ok, &declined)); break;
    }
    if (ok) {
      HS_CUDA(cudaEventRecord(ev_group_end, ctx->stream));
      if (with_store) HS_TRY(build_table_store(ctx, table));
      return HS_OK;
    }
    if (!declined) ctx->stats.hash_sort_fallbacks++;   // two distinct keys with one hash: sort on all words
  }
  KeyPtrs sorted;
  HS_TRY(sort_table(ctx, table, &sorted));
  HS_CUDA(cudaEventRecord(ev_sort_end, ctx->stream));
  switch (ctx->key_words) {
    case 1: HS_TRY(group_inst<1>(ctx, table, sorted)); break;
    case 2: HS_TRY(group_inst<2>(ctx, table, sorted)); break;
    case 3: HS_TRY(group_inst<3>(ctx, table, sorted)); break;
    default: HS_TRY(group_inst<4>(ctx, table, sorted)); break;
  }
  HS_CUDA(cudaEventRecord(ev_group_end, ctx->stream));
  if (with_store) HS_TRY(build_table_store(ctx, table));
  return HS_OK;
}

int build_table_store(hs_ctx *ctx, uint32_t table) {
  return build_code_store(ctx, ctx->tables[table].sorted_ids.as<uint32_t>(), ctx->tables[table].codes_sorted);
}

}  // namespace hs
