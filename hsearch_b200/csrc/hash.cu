// K1 kernels: fragment hash (FP32 fast path with FP64 guard), all-FP64 hash /
// audit, and the exact query hash.  See hash.cuh for the method.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "hash.cuh"
#include "internal.cuh"

namespace hs {

constexpr int kHashThreads = 256;

__device__ __noinline__ int exact_bucket_codes_cold(const uint8_t *codes, int len, const double *__restrict__ table64,
                                                    const double *__restrict__ a_row, double b, double W) {
  return exact_bucket_codes(codes, len, table64, a_row, b, W);
}

// counters[0] guard_hits, [1] guard_corrected, [2] key overflow / bucket out of range, [3] residual flips
// RANK: write dense u16 bucket ranks (and the fragment records) instead of packed keys.
// REP: copies of every 16-byte table cell (8: copy j lives in bank group j and lane i reads
// copy i & 7, so the random-row LDS.128 gathers are bank-conflict free; 1: plain layout for
// tables too large to replicate).  Persistent blocks: the table is staged once per block.
template <int NQ, int KW, bool RANK, int REP, int NT, bool K4, int G>
__global__ void __launch_bounds__(NT)
hash_fast_kernel(const uint8_t *__restrict__ codes, uint64_t N, int len,
                 const float *__restrict__ T32,    // [len][20][4*NQ] of this chunk
                 const float *__restrict__ b32,    // [4*NQ]
                 const float *__restrict__ eps32,  // [4*NQ]
                 float invW, const double *__restrict__ table64, const double *__restrict__ a64,
                 const double *__restrict__ b64, double W, int K, int Kp, int L, int dim,
                 HashChunkArgs args, int32_t *__restrict__ buckets_out,
                 unsigned long long *__restrict__ counters, uint64_t f0 /* multiple of NT */, uint64_t f1,
                 unsigned int *__restrict__ tile_counter /* preset to 2 * gridDim.x * G */) {
  constexpr int P = 4 * NQ;
  // 16-byte words of the next tile's codes held in registers: NT * 16 * NPF bytes >= NT * len
  // (the replicated-table path is taken for len <= 16 only)
  constexpr int NPF = REP > 1 ? 1 : 2;
  // G groups of GT = NT / G threads share the table but walk their own tiles (GT fragments each)
  // with their own buffers and named barriers, so that one group computes while the other
  // waits for its codes or copies its records out
  constexpr int GT = NT / G;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4 *sT = reinterpret_cast<float4 *>(smem_raw);                             // len*20*NQ*REP cells
  const int tid = threadIdx.x;
  const int gid = tid / GT, gt = tid - gid * GT;
  const size_t codes_bytes = ((size_t)GT * len + 15) & ~(size_t)15;
  uint8_t *sC = smem_raw + (size_t)len * HS_AA * NQ * REP * sizeof(float4) + (size_t)gid * codes_bytes;  // GT*len bytes
  uint8_t *sRec = smem_raw + (size_t)len * HS_AA * NQ * REP * sizeof(float4) + (size_t)G * codes_bytes +
                  (size_t)gid * GT * args.rec_stride;                            // GT*rec_stride (full_rec)
  auto group_sync = [&]() {
    if (G == 1) __syncthreads();
    else if (GT == 32) __syncwarp();  // one warp per tile: no block-level barrier at all
    else asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "r"(GT) : "memory");
  };

  const bool full_rec = RANK && args.full_rec;
  const uint32_t RS = args.rec_stride;
  {
    const int n4 = len * HS_AA * NQ;
    const float4 *src = reinterpret_cast<const float4 *>(T32);
    if (REP > 1) {
      // replicated layout: undo the row rotation of the global table, so that the hot loop
      // addresses quad j with a compile-time offset
      for (int i = tid; i < n4 * REP; i += NT) {
        const int cell = i / REP, j = cell % NQ, rc = cell / NQ, c = rc % HS_AA;
        sT[i] = src[rc * NQ + ((j + (c >> 1)) & (NQ - 1))];
      }
    } else {
      for (int i = tid; i < n4; i += NT) sT[i] = src[i];
    }
  }
  if (G > 1) __syncthreads();  // the table is staged by all threads; the loop below only syncs groups
  const float4 *myT = sT + (REP > 1 ? (tid & (REP - 1)) : 0);
  unsigned int my_guard = 0, my_corr = 0, my_over = 0;
  // this launch hashes the fragments [f0, f1)
  const uint64_t tile0 = f0 / GT;
  const uint64_t ntiles = (f1 + GT - 1) / GT;
  const uint64_t total_bytes = f1 * (uint64_t)len;
  // the tile's code bytes are contiguous in global memory (16-byte aligned start): whole 16-byte
  // words are prefetched into registers one tile ahead, the ragged end of the DB byte by byte
  uint4 pf[NPF];
  auto prefetch = [&](uint64_t tile) {
    const uint64_t byte0 = tile * GT * (uint64_t)len;
#pragma unroll
    for (int v = 0; v < NPF; ++v) {
      const uint64_t off = byte0 + ((uint64_t)v * GT + gt) * 16;
      pf[v] = make_uint4(0u, 0u, 0u, 0u);
      if (tile < ntiles && ((uint64_t)v * GT + gt) * 16 < (uint64_t)GT * len && off + 16 <= total_bytes)
        pf[v] = __ldg(reinterpret_cast<const uint4 *>(codes + off));
    }
  };
  // Tiles are handed out dynamically: a block that starts late (its SM was held by another
  // kernel -- the hit merge of the previous batch, another context's work) takes fewer tiles
  // instead of making the launch wait for it.  Every group's first two tiles are static; from
  // then on thread 0 of the group takes the tile after next from a global counter, one
  // iteration ahead of its use, so that the atomic's latency is never waited for.
  __shared__ unsigned int s_next[G];
  const uint64_t tstride = (uint64_t)gridDim.x * G;
  uint64_t tile = tile0 + (uint64_t)blockIdx.x * G + gid, tile_nxt = tile + tstride;
  unsigned int pending = 0;
  if (gt == 0) pending = atomicAdd(tile_counter, 1u);
  prefetch(tile);
  for (; tile < ntiles;) {
    const uint64_t frag0 = tile * GT;
    const uint64_t nfrag = min((uint64_t)GT, f1 - frag0);
    group_sync();  // the previous tile's sC / sRec are no longer read
    if (gt == 0) {
      s_next[gid] = pending;                     // the tile after next (taken one iteration ago)
      pending = atomicAdd(tile_counter, 1u);
    }
    {
      const uint64_t byte0 = frag0 * (uint64_t)len;
      const uint32_t nbytes = (uint32_t)(nfrag * (uint64_t)len);
#pragma unroll
      for (int v = 0; v < NPF; ++v) {
        const uint32_t o = ((uint32_t)v * GT + gt) * 16u;
        if (o + 16u <= nbytes) *reinterpret_cast<uint4 *>(sC + o) = pf[v];
      }
      for (uint32_t i = (nbytes & ~15u) + gt; i < nbytes; i += GT) sC[i] = codes[byte0 + i];
    }
    if (full_rec) {
      uint4 *z = reinterpret_cast<uint4 *>(sRec);
      for (uint32_t i = gt; i < (uint32_t)GT * RS / 16; i += GT) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    prefetch(tile_nxt);
    group_sync();
    const uint64_t tile_nn = tile0 + (uint64_t)s_next[gid];
    if ((uint64_t)gt < nfrag) {
      const uint64_t frag = frag0 + gt;
      const uint8_t *myc = sC + gt * len;
      uint8_t *myrec = sRec + (size_t)gt * RS;

      float acc[P];
#pragma unroll
      for (int s = 0; s < P; ++s) acc[s] = 0.f;
      for (int pos = 0; pos < len; ++pos) {
        const int c = myc[pos];
        if (full_rec) myrec[pos] = (uint8_t)c;
        const float4 *row = myT + (size_t)(pos * HS_AA + c) * (NQ * REP);
        // the quads of row c are stored rotated by c/2 (setup_projection): spreads the lanes of
        // the unreplicated layout over the banks when they all want logical quad j
        const int rot = REP > 1 ? 0 : (c >> 1);
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
          const float4 v = row[((j + rot) & (NQ - 1)) * REP];
          acc[4 * j + 0] += v.x;
          acc[4 * j + 1] += v.y;
          acc[4 * j + 2] += v.z;
          acc[4 * j + 3] += v.w;
        }
      }

      KeyBuilder<KW> kb;
      kb.reset();
      uint32_t tix = 0;
      bool in_range = true;
      // projection slot s = projection k of table t of the chunk
      auto project = [&](int s, int t, int k, bool last_k) {
        const float val = acc[s] + args.b32[s];
        const float tt = val * invW;
        const float f = floorf(tt);
        int bucket = (int)f;
        const float e = args.eps32[s];
        if ((tt - f) < e || ((f + 1.0f) - tt) < e) {
          const int l = args.l0 + t;
          // out of line: the FP64 evaluation runs for ~3e-5 of the projections; inlined 16 times it
          // made the kernel 10 k instructions long and the hot path jump across it
          const int ex = exact_bucket_codes_cold(myc, len, table64, a64 + ((size_t)l * K + k) * dim, b64[l * K + k], W);
          ++my_guard;
          if (ex != bucket) ++my_corr;
          bucket = ex;
        }
        if (RANK) in_range = rank_tuple_push(tix, bucket, args.lo[s], args.rng[s]) && in_range;
        else kb.push_int(bucket);
        if (buckets_out) buckets_out[(frag * L + (args.l0 + t)) * K + k] = bucket;
        if (last_k) {
          if (RANK) {
            uint16_t rank = 0;
            if (in_range) rank = __ldg(args.lut[t] + tix);
            else ++my_over;
            args.ranks[t][frag] = rank;
            if (full_rec)
              *reinterpret_cast<uint16_t *>(myrec + args.rec_rank_off + 2 * (args.l0 + t)) = rank;
            else
              *reinterpret_cast<uint16_t *>(args.rec + frag * RS + args.rec_rank_off + 2 * (args.l0 + t)) = rank;
            tix = 0;
            in_range = true;
          } else {
            if (kb.nchars > 16 * KW) ++my_over;
            uint64_t *dst = args.keys[t];
#pragma unroll
            for (int w = 0; w < KW; ++w) dst[(uint64_t)w * N + frag] = kb.w[w];
            kb.reset();
          }
        }
      };
      if constexpr (K4) {
        // K = 4 and every slot of the chunk in use (the reference's configuration): the
        // (table, projection) walk is fully static
#pragma unroll
        for (int s = 0; s < P; ++s) project(s, s >> 2, s & 3, (s & 3) == 3);
      } else {
        int t = 0, k = 0;  // uniform across the block
#pragma unroll
        for (int s = 0; s < P; ++s) {
          if (t < args.ntab && k < K) project(s, t, k, k == K - 1);
          if (++k == Kp) {
            k = 0;
            ++t;
          }
        }
      }
    }
    if (full_rec) {
      // the tile's records are contiguous in global memory: coalesced 16-byte copies
      group_sync();
      const uint32_t nvec = (uint32_t)(nfrag * RS / 16);
      const uint4 *src = reinterpret_cast<const uint4 *>(sRec);
      uint4 *dst = reinterpret_cast<uint4 *>(args.rec + frag0 * RS);
      for (uint32_t i = gt; i < nvec; i += GT) dst[i] = src[i];
    }
    tile = tile_nxt;
    tile_nxt = tile_nn;
  }
  if (my_guard) atomicAdd(counters + 0, (unsigned long long)my_guard);
  if (my_corr) atomicAdd(counters + 1, (unsigned long long)my_corr);
  if (my_over) atomicAdd(counters + 2, (unsigned long long)my_over);
}

// Fragment records without ranks: codes copied into rows of rec_stride bytes (zero padded).
__global__ void __launch_bounds__(kHashThreads)
build_records_kernel(const uint8_t *__restrict__ codes, uint64_t N, int len, uint32_t RS, uint8_t *__restrict__ rec) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint8_t *sRec = smem_raw;
  const int tid = threadIdx.x;
  const uint64_t frag0 = (uint64_t)blockIdx.x * kHashThreads;
  const uint64_t nfrag = min((uint64_t)kHashThreads, N - frag0);
  uint4 *z = reinterpret_cast<uint4 *>(sRec);
  for (uint32_t i = tid; i < (uint32_t)kHashThreads * RS / 16; i += kHashThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  const uint64_t byte0 = frag0 * (uint64_t)len;
  const uint32_t nbytes = (uint32_t)(nfrag * (uint64_t)len);
  for (uint32_t i = tid; i < nbytes; i += kHashThreads) {
    const uint32_t f = i / (uint32_t)len, p = i - f * (uint32_t)len;
    sRec[f * RS + p] = codes[byte0 + i];
  }
  __syncthreads();
  const uint32_t nvec = (uint32_t)(nfrag * RS / 16);
  uint4 *dst = reinterpret_cast<uint4 *>(rec + frag0 * RS);
  for (uint32_t i = tid; i < nvec; i += kHashThreads) dst[i] = z[i];
}

// All-FP64 hash in reference order.  audit == 0: writes keys (and buckets).
// audit == 1: compares the recomputed key (rank path: the recomputed rank) with the
// stored one and counts mismatching (fragment, table) keys in counters[3].
template <int KW>
__global__ void __launch_bounds__(kHashThreads)
hash_exact_kernel(const uint8_t *__restrict__ codes, uint64_t N, int len,
                  const double *__restrict__ table64, const double *__restrict__ a64,
                  const double *__restrict__ b64, double W, int K, int L, int dim, int l0, int l1,
                  uint64_t *const *__restrict__ keys, int32_t *__restrict__ buckets_out, int audit,
                  const int *__restrict__ rinfo /* rank path, audit only */, const uint16_t *__restrict__ lut,
                  const uint16_t *__restrict__ ranks, uint64_t rstride, unsigned long long *__restrict__ counters) {
  const uint64_t frag = (uint64_t)blockIdx.x * kHashThreads + threadIdx.x;
  if (frag >= N) return;
  uint8_t c[HS_MAX_LEN];
  for (int i = 0; i < len; ++i) c[i] = codes[frag * len + i];
  unsigned int over = 0, flips = 0;
  for (int l = l0; l < l1; ++l) {
    KeyBuilder<KW> kb;
    kb.reset();
    uint32_t tix = 0;
    bool in_range = true;
    for (int k = 0; k < K; ++k) {
      const int bucket = exact_bucket_codes(c, len, table64, a64 + ((size_t)l * K + k) * dim, b64[l * K + k], W);
      kb.push_int(bucket);
      if (rinfo) in_range = rank_tuple_push(tix, bucket, rinfo[l * K + k], rinfo[L * K + l * K + k]) && in_range;
      if (!audit && buckets_out) buckets_out[(frag * L + l) * K + k] = bucket;
    }
    if (kb.nchars > 16 * KW) ++over;
    if (rinfo) {  // rank path (audit only)
      const uint32_t off = (uint32_t)rinfo[2 * L * K + l];
      if (!in_range || lut[off + tix] != ranks[(uint64_t)l * rstride + frag]) ++flips;
      continue;
    }
    uint64_t *dst = keys[l];
    if (audit) {
      bool same = true;
#pragma unroll
      for (int w = 0; w < KW; ++w) same = same && (dst[(uint64_t)w * N + frag] == kb.w[w]);
      if (!same) ++flips;
    } else {
#pragma unroll
      for (int w = 0; w < KW; ++w) dst[(uint64_t)w * N + frag] = kb.w[w];
    }
  }
  if (over) atomicAdd(counters + 2, (unsigned long long)over);
  if (flips) atomicAdd(counters + 3, (unsigned long long)flips);
}

// Query hash (motif_both_points.cpp:227): one thread per (query, table), FP64
// in reference order.  qkeys [L][Q][KW]; qvalid[L][Q] = 0 when the key string
// is longer than any DB key can be (then it matches no bucket).
// GP = K rounded up to a power of two (<= 32) lanes evaluate the K projections of one (query,
// table) pair, lane k the k-th: the query point is read once per group (broadcast), the projection
// rows from a transposed copy a64t[l][i][k] (consecutive lanes, consecutive doubles).  Each lane runs
// the reference's sequential multiply-add over the dim coordinates (exact_bucket_point's order);
// the group's first lane then strings the K buckets together.  (One thread per pair, K projections
// in sequence, read the points with a 640-byte stride between lanes: 5 ms at K = 16, L = 32.)
template <int KW>
__global__ void hash_queries_kernel(const double *__restrict__ q64, uint32_t Q, int dim,
                                    const double *__restrict__ a64t, const double *__restrict__ b64,
                                    double W, int K, int GP, int L, uint64_t *__restrict__ qkeys,
                                    uint8_t *__restrict__ qvalid) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t pair = t / (uint32_t)GP;
  const int k = (int)(t % (uint32_t)GP);
  const bool live = pair < (uint64_t)Q * (uint32_t)L;
  const uint32_t l = live ? (uint32_t)(pair / Q) : 0u, q = live ? (uint32_t)(pair - (uint64_t)l * Q) : 0u;
  int bucket = 0;
  if (live && k < K) {
    const double *pt = q64 + (size_t)q * dim;
    const double *arow = a64t + (size_t)l * dim * GP + k;
    double dot = 0.0;
    for (int i = 0; i < dim; ++i) dot = __dadd_rn(dot, __dmul_rn(pt[i], arow[(size_t)i * GP]));
    const double val = __dadd_rn(dot, b64[l * K + k]);
    bucket = (int)floor(__ddiv_rn(val, W));
  }
  KeyBuilder<KW> kb;
  kb.reset();
  const int gbase = (threadIdx.x & 31) & ~(GP - 1);
  for (int kk = 0; kk < K; ++kk) kb.push_int(__shfl_sync(0xffffffffu, bucket, gbase + kk));
  if (live && k == 0) {
#pragma unroll
    for (int w = 0; w < KW; ++w) qkeys[((size_t)l * Q + q) * KW + w] = kb.w[w];
    qvalid[(size_t)l * Q + q] = kb.nchars <= 16 * KW ? 1 : 0;
  }
}

// ---- host side ---------------------------------------------------------------
constexpr int kHashRepThreads = 1024;           // one persistent block per SM on the replicated-table path
constexpr size_t kHashRepBudget = 200 * 1024;   // replicated table + tile buffers

static size_t hash_smem_bytes(const hs_ctx *ctx, int NQ, int rep, int nt, bool full_rec, int groups = 1) {
  const size_t len = ctx->prm.len;
  return len * HS_AA * NQ * rep * sizeof(float4) + groups * ((((size_t)nt / groups) * len + 15) & ~(size_t)15) +
         (full_rec ? (size_t)nt * ctx->rec_stride : 0);
}

template <int NQ, int KW, bool RANK, int REP, int NT, bool K4 = false, int G = 1>
static int launch_fast_rep(hs_ctx *ctx, int chunk, const HashChunkArgs &args, int32_t *buckets,
                           unsigned long long *counters, uint64_t f0, uint64_t f1) {
  const int P = 4 * NQ;
  const int len = (int)ctx->prm.len;
  constexpr int GT = NT / G;
  const size_t smem = hash_smem_bytes(ctx, NQ, REP, NT, args.full_rec != 0, G);
  auto kern = hash_fast_kernel<NQ, KW, RANK, REP, NT, K4, G>;
  // (the kernel also has a little static shared memory: opt in before the sum reaches 48 KB)
  if (smem + 2048 > 48 * 1024) HS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  HS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
  if (per_sm < 1) per_sm = 1;
  const uint64_t ntiles = (f1 + GT - 1) / GT - f0 / GT;
  const unsigned grid = (unsigned)std::min<uint64_t>((ntiles + G - 1) / G, (uint64_t)ctx->num_sms * per_sm);
  const float *T = ctx->d_T32.as<float>() + (size_t)chunk * len * HS_AA * P;
  // dynamic tile counter (counters slot 28): the first two tiles of every group are static
  unsigned int *tile_counter = reinterpret_cast<unsigned int *>(counters + 28);
  const unsigned int tc0 = 2u * grid * G;
  HS_TRY(upload(ctx, tile_counter, &tc0, sizeof tc0));
  kern<<<grid, NT, smem, ctx->stream>>>(
      ctx->d_codes.as<uint8_t>(), ctx->N, len, T, ctx->d_b32.as<float>() + (size_t)chunk * P,
      ctx->d_eps32.as<float>() + (size_t)chunk * P, (float)(1.0 / ctx->prm.W), ctx->d_table64.as<double>(),
      ctx->d_a64.as<double>(), ctx->d_b64.as<double>(), ctx->prm.W, (int)ctx->prm.K, (int)ctx->Kp,
      (int)ctx->prm.L, (int)ctx->dim, args, buckets, counters, f0, f1, tile_counter);
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches++;
  return HS_OK;
}

template <int NQ, int KW, bool RANK>
static int launch_fast_inst(hs_ctx *ctx, int chunk, const HashChunkArgs &args, int32_t *buckets,
                            unsigned long long *counters, uint64_t f0, uint64_t f1) {
  // 768 threads leave 85 registers per thread: 16 accumulators without spills
  // the replicated-table kernel exists for the rank path with up to 16 projections per launch
  // (the reference's configurations); everything else takes the plain layout
  if constexpr (RANK && NQ <= 4) {
    if (ctx->prm.len <= 16 && hash_smem_bytes(ctx, NQ, 8, kHashRepThreads, args.full_rec != 0) <= kHashRepBudget) {
      if (args.k4_full)  // K = 4, all slots in use: statically unrolled (table, projection) walk
        return launch_fast_rep<NQ, KW, RANK, 8, kHashRepThreads, true, 8>(ctx, chunk, args, buckets, counters, f0, f1);
      return launch_fast_rep<NQ, KW, RANK, 8, kHashRepThreads, false, 8>(ctx, chunk, args, buckets, counters, f0, f1);
    }
  }
  return launch_fast_rep<NQ, KW, RANK, 1, kHashThreads>(ctx, chunk, args, buckets, counters, f0, f1);
}

template <int NQ>
static int launch_fast_kw(hs_ctx *ctx, int chunk, const HashChunkArgs &a, int32_t *b, unsigned long long *c,
                          uint64_t f0, uint64_t f1) {
  if (ctx->rank_mode) return launch_fast_inst<NQ, 1, true>(ctx, chunk, a, b, c, f0, f1);
  switch (ctx->key_words) {
    case 1: return launch_fast_inst<NQ, 1, false>(ctx, chunk, a, b, c, f0, f1);
    case 2: return launch_fast_inst<NQ, 2, false>(ctx, chunk, a, b, c, f0, f1);
    case 3: return launch_fast_inst<NQ, 3, false>(ctx, chunk, a, b, c, f0, f1);
    default: return launch_fast_inst<NQ, 4, false>(ctx, chunk, a, b, c, f0, f1);
  }
}

int ensure_records(hs_ctx *ctx) {
  if (ctx->have_rec || ctx->N == 0) return HS_OK;
  HS_TRY(ctx->d_rec.reserve((size_t)ctx->N * ctx->rec_stride + 64));
  const unsigned grid = (unsigned)((ctx->N + kHashThreads - 1) / kHashThreads);
  const size_t smem = (size_t)kHashThreads * ctx->rec_stride;
  if (smem > 48 * 1024)
    HS_CUDA(cudaFuncSetAttribute(build_records_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  build_records_kernel<<<grid, kHashThreads, smem, ctx->stream>>>(ctx->d_codes.as<uint8_t>(), ctx->N,
                                                                 (int)ctx->prm.len, ctx->rec_stride,
                                                                 ctx->d_rec.as<uint8_t>());
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches++;
  ctx->have_rec = true;
  return HS_OK;
}

bool hash_single_launch_records(const hs_ctx *ctx) {
  return ctx->rank_mode && ctx->nchunks == 1 && hash_smem_bytes(ctx, (int)ctx->nq, 1, kHashThreads, true) <= 200 * 1024;
}

// Hashes the fragments [f0, f1) (f0 a multiple of kHashRangeAlign); the whole DB when f1 == 0.
int launch_hash_fast(hs_ctx *ctx, bool want_buckets, uint64_t f0, uint64_t f1) {
  if (f1 == 0) f1 = ctx->N;
  int32_t *buckets = want_buckets ? ctx->d_buckets.as<int32_t>() : nullptr;
  unsigned long long *counters = ctx->d_counters.as<unsigned long long>();
  const uint32_t K = ctx->prm.K;
  // the single-chunk rank launch writes whole records through shared memory; otherwise the
  // records must exist before the per-table ranks are stored into them
  const bool full_rec = hash_single_launch_records(ctx);
  if (ctx->rank_mode) {
    HS_TRY(ctx->d_ranks.reserve(sizeof(uint16_t) * (size_t)ctx->prm.L * ctx->npad + 64));  // table stride npad: 16-byte aligned rows
    if (full_rec) HS_TRY(ctx->d_rec.reserve((size_t)ctx->N * ctx->rec_stride + 64));
    else if (f0 == 0 && f1 == ctx->N) HS_TRY(ensure_records(ctx));
    else if (!ctx->have_rec) {
      set_error("launch_hash_fast: ranged hashing needs the single-launch record path");
      return HS_ERR_UNSUPPORTED;
    }
  }
  for (uint32_t chunk = 0; chunk < ctx->nchunks; ++chunk) {
    HashChunkArgs args;
    memset(&args, 0, sizeof args);
    args.l0 = (int)(chunk * ctx->tpc);
    args.ntab = (int)std::min<uint32_t>(ctx->tpc, ctx->prm.L - args.l0);
    for (int t = 0; t < args.ntab; ++t) {
      const uint32_t l = (uint32_t)args.l0 + t;
      if (ctx->rank_mode) {
        args.ranks[t] = ctx->d_ranks.as<uint16_t>() + (size_t)l * ctx->npad;
        args.lut[t] = ctx->d_lut.as<uint16_t>() + ctx->rank_lut_off[l];
        for (uint32_t k = 0; k < K; ++k) {
          args.lo[t * ctx->Kp + k] = ctx->rank_lo[(size_t)l * K + k];
          args.rng[t * ctx->Kp + k] = ctx->rank_rng[(size_t)l * K + k];
        }
      } else {
        args.keys[t] = ctx->d_keys[l].as<uint64_t>();
      }
    }
    for (uint32_t sl = 0; sl < 4 * ctx->nq; ++sl) {
      args.b32[sl] = ctx->h_b32[(size_t)chunk * 4 * ctx->nq + sl];
      args.eps32[sl] = ctx->h_eps32[(size_t)chunk * 4 * ctx->nq + sl];
    }
    args.rec = ctx->d_rec.as<uint8_t>();
    args.rec_stride = ctx->rec_stride;
    args.rec_rank_off = ctx->rec_rank_off;
    args.full_rec = full_rec ? 1 : 0;
    args.k4_full = (K == 4 && (uint32_t)args.ntab * 4 == 4 * ctx->nq) ? 1 : 0;
    switch (ctx->nq) {
      case 1: HS_TRY((launch_fast_kw<1>(ctx, chunk, args, buckets, counters, f0, f1))); break;
      case 2: HS_TRY((launch_fast_kw<2>(ctx, chunk, args, buckets, counters, f0, f1))); break;
      case 4: HS_TRY((launch_fast_kw<4>(ctx, chunk, args, buckets, counters, f0, f1))); break;
      default: HS_TRY((launch_fast_kw<8>(ctx, chunk, args, buckets, counters, f0, f1))); break;
    }
  }
  if (full_rec && f1 == ctx->N) ctx->have_rec = true;
  return HS_OK;
}

template <int KW>
static int launch_exact_inst(hs_ctx *ctx, int32_t *buckets, int audit, uint64_t *const *d_keyptrs,
                             unsigned long long *counters) {
  const unsigned grid = (unsigned)((ctx->N + kHashThreads - 1) / kHashThreads);
  const bool rank = ctx->rank_mode;  // (the all-FP64 hash itself never runs on the rank path)
  hash_exact_kernel<KW><<<grid, kHashThreads, 0, ctx->stream>>>(
      ctx->d_codes.as<uint8_t>(), ctx->N, (int)ctx->prm.len, ctx->d_table64.as<double>(),
      ctx->d_a64.as<double>(), ctx->d_b64.as<double>(), ctx->prm.W, (int)ctx->prm.K, (int)ctx->prm.L,
      (int)ctx->dim, 0, (int)ctx->prm.L, d_keyptrs, buckets, audit, rank ? ctx->d_rinfo.as<int>() : nullptr,
      ctx->d_lut.as<uint16_t>(), ctx->d_ranks.as<uint16_t>(), ctx->npad, counters);
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches++;
  return HS_OK;
}

int launch_hash_exact(hs_ctx *ctx, bool want_buckets, bool audit) {
  int32_t *buckets = want_buckets ? ctx->d_buckets.as<int32_t>() : nullptr;
  unsigned long long *counters = ctx->d_counters.as<unsigned long long>();
  // table of per-table key pointers in device memory
  uint64_t *h_ptrs[HS_MAX_L];
  for (uint32_t l = 0; l < ctx->prm.L; ++l) h_ptrs[l] = ctx->d_keys[l].as<uint64_t>();
  HS_TRY(ctx->d_misc.reserve(sizeof h_ptrs));
  HS_CUDA(cudaMemcpyAsync(ctx->d_misc.p, h_ptrs, sizeof(uint64_t *) * ctx->prm.L, cudaMemcpyHostToDevice,
                          ctx->stream));
  uint64_t *const *d_ptrs = ctx->d_misc.as<uint64_t *>();
  switch (ctx->key_words) {
    case 1: return launch_exact_inst<1>(ctx, buckets, audit, d_ptrs, counters);
    case 2: return launch_exact_inst<2>(ctx, buckets, audit, d_ptrs, counters);
    case 3: return launch_exact_inst<3>(ctx, buckets, audit, d_ptrs, counters);
    default: return launch_exact_inst<4>(ctx, buckets, audit, d_ptrs, counters);
  }
}

int launch_hash_queries(hs_ctx *ctx, const double *d_q64, uint32_t Q, uint64_t *d_qkeys, uint8_t *d_qvalid) {
  const uint64_t n = (uint64_t)Q * ctx->prm.L * ctx->qh_group;
  if (n == 0) return HS_OK;
  const unsigned grid = (unsigned)((n + 127) / 128);
#define HS_QH(KWV)                                                                                        \
  hash_queries_kernel<KWV><<<grid, 128, 0, ctx->stream>>>(d_q64, Q, (int)ctx->dim, ctx->d_a64t.as<double>(), \
                                                          ctx->d_b64.as<double>(), ctx->prm.W, (int)ctx->prm.K, \
                                                          (int)ctx->qh_group, (int)ctx->prm.L, d_qkeys, d_qvalid)
  switch (ctx->key_words) {
    case 1: HS_QH(1); break;
    case 2: HS_QH(2); break;
    case 3: HS_QH(3); break;
    default: HS_QH(4); break;
  }
#undef HS_QH
  HS_CUDA(cudaGetLastError());
  ctx->stats.kernel_launches++;
  return HS_OK;
}

// std::to_string(int) strings of a bucket tuple, packed like KeyBuilder does (hash.cuh).
static void host_pack_tuple(const int *vals, uint32_t K, uint32_t KW, uint64_t *w) {
  for (uint32_t i = 0; i < KW; ++i) w[i] = 0;
  char buf[16];
  for (uint32_t k = 0; k < K; ++k) {
    const int n = snprintf(buf, sizeof buf, "%d", vals[k]);
    for (int i = 0; i < n; ++i) {
      const uint64_t nib = buf[i] == '-' ? 11u : (uint64_t)(buf[i] - '0') + 1u;
      for (uint32_t j = KW - 1; j > 0; --j) w[j] = (w[j] << 4) | (w[j - 1] >> 60);
      w[0] = (w[0] << 4) | nib;
    }
  }
}

constexpr uint32_t kMaxRanks = 65536;  // ranks are u16

// Dense bucket ranks.  Every projection's bucket lies in [lo, lo + rng) for every possible
// residue string (setup_projection), so a table has at most prod(rng) bucket tuples.  When
// that is <= 65536 for every table, all tuples are enumerated on the host, their key strings
// (HashKey, lsh.hpp:51-59: concatenated without separator, so distinct tuples may collide)
// are packed and sorted, and lut[tuple] = rank of the tuple's string.  The device then
// carries a u16 rank per (fragment, table) instead of a KW*64-bit packed key; rank order is
// packed-key order, so bucket order, bucket contents and the probe are unchanged.
static int setup_ranks(hs_ctx *ctx) {
  const uint32_t K = ctx->prm.K, L = ctx->prm.L, KW = ctx->key_words, len = ctx->prm.len;
  ctx->rank_mode = false;
  ctx->have_rec = false;
  ctx->rec_rank_off = (len + 1u) & ~1u;
  ctx->rec_stride = (len + 15u) & ~15u;
  const char *e = getenv("HS_NO_RANK");
  if ((ctx->prm.flags & HS_FLAG_HASH_EXACT) || (e && atoi(e))) return HS_OK;
  std::vector<uint32_t> prod(L);
  uint64_t total = 0;
  for (uint32_t l = 0; l < L; ++l) {
    uint64_t p = 1;
    for (uint32_t k = 0; k < K; ++k) {
      p *= (uint64_t)ctx->rank_rng[(size_t)l * K + k];
      if (p > kMaxRanks) return HS_OK;  // too many possible buckets: packed-key path
    }
    prod[l] = (uint32_t)p;
    total += p;
  }
  std::vector<uint16_t> lut(total);
  std::vector<int> rinfo(2 * (size_t)L * K + L);
  uint64_t off = 0;
  std::vector<int> vals(K);
  for (uint32_t l = 0; l < L; ++l) {
    const uint32_t np = prod[l];
    std::vector<uint64_t> keys((size_t)np * KW);
    for (uint32_t t = 0; t < np; ++t) {
      uint32_t rem = t;
      for (int k = (int)K - 1; k >= 0; --k) {  // first projection is the most significant digit
        const uint32_t r = (uint32_t)ctx->rank_rng[(size_t)l * K + k];
        vals[k] = ctx->rank_lo[(size_t)l * K + k] + (int)(rem % r);
        rem /= r;
      }
      host_pack_tuple(vals.data(), K, KW, &keys[(size_t)t * KW]);
    }
    std::vector<uint32_t> order(np);
    for (uint32_t t = 0; t < np; ++t) order[t] = t;
    auto less = [&](uint32_t x, uint32_t y) {
      for (int w = (int)KW - 1; w >= 0; --w) {
        const uint64_t a = keys[(size_t)x * KW + w], b = keys[(size_t)y * KW + w];
        if (a != b) return a < b;
      }
      return false;
    };
    std::sort(order.begin(), order.end(), less);
    std::vector<uint64_t> &rk = ctx->h_rkeys[l];
    rk.clear();
    uint32_t nr = 0;
    std::vector<uint64_t> uniq;  // [nr][KW]
    for (uint32_t i = 0; i < np; ++i) {
      if (i == 0 || less(order[i - 1], order[i])) {
        for (uint32_t w = 0; w < KW; ++w) uniq.push_back(keys[(size_t)order[i] * KW + w]);
        ++nr;
      }
      lut[off + order[i]] = (uint16_t)(nr - 1);
    }
    rk.resize((size_t)KW * nr);  // word-major like the device tables
    for (uint32_t r = 0; r < nr; ++r)
      for (uint32_t w = 0; w < KW; ++w) rk[(size_t)w * nr + r] = uniq[(size_t)r * KW + w];
    ctx->rank_nr[l] = nr;
    ctx->rank_lut_off[l] = (uint32_t)off;
    rinfo[2 * (size_t)L * K + l] = (int)off;
    off += np;
  }
  for (size_t i = 0; i < (size_t)L * K; ++i) {
    rinfo[i] = ctx->rank_lo[i];
    rinfo[(size_t)L * K + i] = ctx->rank_rng[i];
  }
  HS_TRY(ctx->d_lut.reserve(sizeof(uint16_t) * lut.size() + 16));
  HS_TRY(ctx->d_rinfo.reserve(sizeof(int) * rinfo.size()));
  HS_CUDA(cudaMemcpyAsync(ctx->d_lut.p, lut.data(), sizeof(uint16_t) * lut.size(), cudaMemcpyHostToDevice, ctx->stream));
  HS_CUDA(cudaMemcpyAsync(ctx->d_rinfo.p, rinfo.data(), sizeof(int) * rinfo.size(), cudaMemcpyHostToDevice, ctx->stream));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->rank_mode = true;
  ctx->rec_stride = (ctx->rec_rank_off + 2u * L + 15u) & ~15u;
  return HS_OK;
}

// Build the device-side projection data: FP64 matrix, FP32 residue-projection
// tables per chunk, guard bands, key width.
int setup_projection(hs_ctx *ctx, const double *a, const double *b) {
  const uint32_t K = ctx->prm.K, L = ctx->prm.L, len = ctx->prm.len, dim = ctx->dim;
  const double W = ctx->prm.W;
  ctx->h_a.assign(a, a + (size_t)L * K * dim);
  ctx->h_b.assign(b, b + (size_t)L * K);
  HS_TRY(ctx->d_a64.reserve(sizeof(double) * L * K * dim));
  HS_TRY(ctx->d_b64.reserve(sizeof(double) * L * K));
  HS_CUDA(cudaMemcpyAsync(ctx->d_a64.p, a, sizeof(double) * L * K * dim, cudaMemcpyHostToDevice, ctx->stream));
  HS_CUDA(cudaMemcpyAsync(ctx->d_b64.p, b, sizeof(double) * L * K, cudaMemcpyHostToDevice, ctx->stream));
  {
    // transposed copy for the query hash: a64t[l][i][k], k padded to a power of two (hash_queries_kernel)
    uint32_t gp = 1;
    while (gp < K) gp <<= 1;
    ctx->qh_group = gp;
    std::vector<double> at((size_t)L * dim * gp, 0.0);
    for (uint32_t l = 0; l < L; ++l)
      for (uint32_t k = 0; k < K; ++k)
        for (uint32_t i = 0; i < dim; ++i) at[((size_t)l * dim + i) * gp + k] = a[((size_t)l * K + k) * dim + i];
    HS_TRY(ctx->d_a64t.reserve(sizeof(double) * at.size()));
    HS_CUDA(cudaMemcpyAsync(ctx->d_a64t.p, at.data(), sizeof(double) * at.size(), cudaMemcpyHostToDevice, ctx->stream));
    HS_CUDA(cudaStreamSynchronize(ctx->stream));   // `at` is a temporary
  }

  ctx->Kp = (K + 3u) & ~3u;
  ctx->tpc = std::max<uint32_t>(1u, std::min<uint32_t>(L, 32u / ctx->Kp));
  if (ctx->tpc > 8) ctx->tpc = 8;
  ctx->nchunks = (L + ctx->tpc - 1) / ctx->tpc;
  uint32_t quads = ctx->tpc * ctx->Kp / 4;
  ctx->nq = quads <= 1 ? 1 : quads <= 2 ? 2 : quads <= 4 ? 4 : 8;
  const uint32_t P = 4 * ctx->nq;

  std::vector<float> T32((size_t)ctx->nchunks * len * HS_AA * P, 0.f), b32((size_t)ctx->nchunks * P, 0.f),
      eps32((size_t)ctx->nchunks * P, 0.f);
  uint32_t max_chars = 0;
  ctx->rank_lo.assign((size_t)L * K, 0);
  ctx->rank_rng.assign((size_t)L * K, 1);
  for (uint32_t l = 0; l < L; ++l) {
    const uint32_t chunk = l / ctx->tpc, t = l % ctx->tpc;
    uint32_t chars = 0;
    for (uint32_t k = 0; k < K; ++k) {
      const double *row = a + ((size_t)l * K + k) * dim;
      const uint32_t slot = t * ctx->Kp + k;
      double A = 0.0, Bmax = 0.0;
      for (uint32_t pos = 0; pos < len; ++pos) {
        double amax = 0.0, tmax = 0.0;
        for (int c = 0; c < HS_AA; ++c) {
          double s = 0.0, sa = 0.0;
          for (int j = 0; j < HS_CDIM; ++j) {
            const double term = ctx->table64[c * HS_CDIM + j] * row[pos * HS_CDIM + j];
            s += term;
            sa += fabs(term);
          }
          // quads of a row are rotated by c/2 (bank-conflict avoidance, hash_fast_kernel)
          const uint32_t pq = ((slot >> 2) + ((uint32_t)c >> 1)) & (ctx->nq - 1);
          T32[(((size_t)chunk * len + pos) * HS_AA + c) * P + pq * 4 + (slot & 3)] = (float)s;
          amax = std::max(amax, sa);
          tmax = std::max(tmax, fabs(s));
        }
        A += amax;
        Bmax += tmax;
      }
      const double bb = b[(size_t)l * K + k];
      b32[(size_t)chunk * P + slot] = (float)bb;
      // |t_fp32 - t_real| <= (len + 6) * 2^-24 * (A + |b|) / W  (see DESIGN.md);
      // doubled, plus an absolute floor.
      const double Et = (double)(len + 6) * ldexp(1.0, -24) * (A + fabs(bb)) / W;
      eps32[(size_t)chunk * P + slot] = (float)(2.0 * Et + 1e-6);
      // bucket range over all possible fragments -> characters of to_string
      // (the FP64 reference sum differs from the real value by < 1e-13 relative; a bucket
      // outside [lo, hi] is still detected on the device and reported, never mis-keyed)
      const double slack = 1e-9 * (Bmax + fabs(bb)) / W + 1e-9;
      const long long lo = (long long)floor((-Bmax + bb) / W - slack);
      const long long hi = (long long)floor((Bmax + bb) / W + slack);
      char buf[32];
      uint32_t c1 = (uint32_t)snprintf(buf, sizeof buf, "%lld", lo);
      uint32_t c2 = (uint32_t)snprintf(buf, sizeof buf, "%lld", hi);
      chars += std::max(c1, c2);
      const long long lo_c = std::max<long long>(lo, -2000000000ll), hi_c = std::min<long long>(hi, 2000000000ll);
      ctx->rank_lo[(size_t)l * K + k] = (int)lo_c;
      ctx->rank_rng[(size_t)l * K + k] = (int)std::min<long long>(hi_c - lo_c + 1, 1ll << 30);
    }
    max_chars = std::max(max_chars, chars);
  }
  ctx->max_chars = max_chars;
  ctx->key_words = (max_chars + 15) / 16;
  if (ctx->key_words == 0) ctx->key_words = 1;
  if (ctx->key_words > HS_MAX_KEY_WORDS) {
    set_error("hs_set_projection: key strings need up to %u characters (> %d); raise W or lower K", max_chars,
              16 * HS_MAX_KEY_WORDS);
    return HS_ERR_UNSUPPORTED;
  }
  HS_TRY(setup_ranks(ctx));
  HS_TRY(ctx->d_T32.reserve(T32.size() * sizeof(float)));
  HS_TRY(ctx->d_b32.reserve(b32.size() * sizeof(float)));
  HS_TRY(ctx->d_eps32.reserve(eps32.size() * sizeof(float)));
  HS_CUDA(cudaMemcpyAsync(ctx->d_T32.p, T32.data(), T32.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  HS_CUDA(cudaMemcpyAsync(ctx->d_b32.p, b32.data(), b32.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  HS_CUDA(cudaMemcpyAsync(ctx->d_eps32.p, eps32.data(), eps32.size() * sizeof(float), cudaMemcpyHostToDevice,
                          ctx->stream));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));  // host vectors go out of scope
  ctx->h_b32 = b32;
  ctx->h_eps32 = eps32;
  ctx->have_projection = true;
  ctx->hashed = false;
  ctx->indexed = false;
  return HS_OK;
}

}  // namespace hs
