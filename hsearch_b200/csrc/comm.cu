// Multi-GPU exchange step of the search path, one process per GPU.
//
// Queries are broadcast from rank 0 (NCCL).  The hit lists are MERGED INTO RANK 0's MEMORY BY
// THE RANKS THAT PRODUCED THEM: rank r owns the ids [id_base_r, id_base_r + N_r) with id_base
// ascending in rank order, so the reference's output order (query, first table, ascending db
// id; motif_both_points.cpp:224-245) of the union is, for every (query, table) segment, the
// ranks' segments one after the other.  After its local sort every rank
//   1. counts its hits per (query, table) segment,
//   2. all-gathers the counts (NCCL, 4 bytes per segment and rank),
//   3. turns them into the final position of each of its segments (one scan), and
//   4. writes its hits straight to those positions of rank 0's receive buffer, which is mapped
//      into every process with CUDA IPC: peer stores over NVLink, no staging copy, no second
//      pass on rank 0, no NCCL channel moving the bulk data.
// Steps 2-4 run on a second stream, so the transfer of batch i overlaps the hash and index build
// of batch i+1; hs_comm_result() completes the pending gather.  NCCL is loaded lazily with
// dlopen so that single-GPU use needs no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "host_tables.h"
#include "internal.cuh"

namespace hs {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.lib) return HS_OK;
  void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!lib) {
    set_error("hs_comm: cannot load libnccl.so.2: %s", dlerror());
    return HS_ERR_COMM;
  }
#define HS_SYM(field, name)                                          \
  *(void **)(&g_nccl.field) = dlsym(lib, name);                      \
  if (!g_nccl.field) {                                               \
    set_error("hs_comm: libnccl has no symbol %s", name);            \
    return HS_ERR_COMM;                                              \
  }
  HS_SYM(GetUniqueId, "ncclGetUniqueId")
  HS_SYM(CommInitRank, "ncclCommInitRank")
  HS_SYM(CommDestroy, "ncclCommDestroy")
  HS_SYM(Broadcast, "ncclBroadcast")
  HS_SYM(AllGather, "ncclAllGather")
  HS_SYM(AllReduce, "ncclAllReduce")
  HS_SYM(GetErrorString, "ncclGetErrorString")
#undef HS_SYM
  g_nccl.lib = lib;
  return HS_OK;
}

#define HS_NCCL(expr)                                                                          \
  do {                                                                                         \
    ncclResult_t _r = (expr);                                                                  \
    if (_r != ncclSuccess) {                                                                   \
      set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r));      \
      return HS_ERR_COMM;                                                                      \
    }                                                                                          \
  } while (0)

int comm_broadcast(hs_ctx *ctx, void *d_buf, size_t bytes) {
  if (ctx->nranks <= 1 || bytes == 0) return HS_OK;
  HS_NCCL(g_nccl.Broadcast(d_buf, d_buf, bytes, ncclUint8, 0, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  return HS_OK;
}

// n 32-bit words of every rank, in rank order, to every rank (cluster labels: cluster.cu)
int comm_allgather_u32(hs_ctx *ctx, const uint32_t *d_send, uint32_t *d_recv, uint64_t n) {
  if (ctx->nranks <= 1) {
    set_error("hs_comm: all-gather on a context without a communicator");
    return HS_ERR_COMM;
  }
  HS_NCCL(g_nccl.AllGather(d_send, d_recv, n, ncclUint32, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  return HS_OK;
}

// ---- kernels ---------------------------------------------------------------------------------
// off[s] = first index of the sorted key list whose segment (key >> shift) is >= s, s in [0, S];
// cnt[s] = off[s+1] - off[s] is written by the thread that closes segment s.
__global__ void seg_offsets_kernel(const uint64_t *__restrict__ keys, uint64_t n, int shift, uint32_t S,
                                   uint64_t *__restrict__ off) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  const uint64_t s_prev = i == 0 ? 0ull : (keys[i - 1] >> shift) + 1ull;
  const uint64_t s_here = i == n ? (uint64_t)S : (keys[i] >> shift);
  for (uint64_t s = s_prev; s <= s_here && s <= S; ++s) off[s] = i;
}
__global__ void seg_counts_kernel(const uint64_t *__restrict__ off, uint32_t S, uint32_t *__restrict__ cnt) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < S) cnt[s] = (uint32_t)(off[s + 1] - off[s]);
}

// One block: for every segment the total over ranks and the hits of lower ranks, scanned into
// the final position of this rank's segment.  info[0] = grand total, info[1] = 1 when it
// exceeds the receive buffer (then nothing is written by the scatter).
constexpr int kDestThreads = 1024;
__global__ void __launch_bounds__(kDestThreads)
seg_dest_kernel(const uint32_t *__restrict__ cnt_all /* [G][S] */, uint32_t S, int G, int rank, uint64_t cap,
                uint64_t *__restrict__ dst /* [S] */, unsigned long long *__restrict__ info) {
  __shared__ unsigned long long s_warp[kDestThreads / 32];
  __shared__ unsigned long long s_carry;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_carry = 0ull;
  __syncthreads();
  for (uint32_t s0 = 0; s0 < S; s0 += kDestThreads) {
    const uint32_t s = s0 + tid;
    unsigned long long tot = 0ull, before = 0ull;
    if (s < S)
      for (int r = 0; r < G; ++r) {
        const unsigned long long c = cnt_all[(size_t)r * S + s];
        if (r < rank) before += c;
        tot += c;
      }
    unsigned long long inc = tot;  // inclusive scan over the block
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      unsigned long long w = lane < kDestThreads / 32 ? s_warp[lane] : 0ull, wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long v = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += v;
      }
      if (lane < kDestThreads / 32) s_warp[lane] = wi - w;  // exclusive prefix of the warp
    }
    __syncthreads();
    const unsigned long long carry = s_carry;
    if (s < S) dst[s] = carry + s_warp[wid] + (inc - tot) + before;
    __syncthreads();
    if (tid == kDestThreads - 1) s_carry = carry + s_warp[wid] + inc;
    __syncthreads();
  }
  if (tid == 0) {
    info[0] = s_carry;
    info[1] = s_carry > cap ? 1ull : 0ull;
  }
}

// Every hit of the local sorted list goes to its final position of the merged list (rank 0's
// memory, mapped here).  A warp takes 32 consecutive hits; inside a segment they land on 32
// consecutive 24-byte records, i.e. 96 consecutive 8-byte words: the warp transposes them through
// shared memory and writes three fully contiguous 256-byte rows (well-formed NVLink packets; a hit
// per lane would be 8-byte pieces at a 24-byte stride).  Warps that straddle a segment boundary
// write hit by hit.  Streaming (evict-first) loads: the list is read once.  Few thread blocks: the
// transfer is bound by the link, not by the SMs.
constexpr int kScatterWarps = 8;
__global__ void __launch_bounds__(kScatterWarps * 32, 4)
scatter_merged_kernel(const hs_hit *__restrict__ hits, uint64_t n, int tbits, const uint64_t *__restrict__ off,
                      const uint64_t *__restrict__ dst, const unsigned long long *__restrict__ info,
                      hs_hit *__restrict__ out) {
  if (info[1]) return;  // receive buffer too small: reported by hs_comm_result
  __shared__ unsigned long long s_w[kScatterWarps][96];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint64_t nwarps = (uint64_t)gridDim.x * kScatterWarps;
  constexpr int U = 4;  // groups of 32 hits per warp and iteration: their loads are all in flight before the first store
  for (uint64_t i0 = ((uint64_t)blockIdx.x * kScatterWarps + wid) * (32 * U); i0 < n; i0 += nwarps * (32 * U)) {
    unsigned long long w0[U], w1[U], w2[U];
    uint64_t pos[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t i = i0 + (uint64_t)u * 32 + lane;
      w0[u] = w1[u] = w2[u] = 0ull;
      if (i < n) {
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(hits + i);
        w0[u] = __ldcs(src);      // query | table_first << 32
        w1[u] = __ldcs(src + 1);  // db id
        w2[u] = __ldcs(src + 2);  // dist2
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t i = i0 + (uint64_t)u * 32 + lane;
      pos[u] = 0;
      if (i < n) {
        const uint64_t s = ((w0[u] & 0xffffffffull) << tbits) | (w0[u] >> 32);
        pos[u] = __ldg(dst + s) + (i - __ldg(off + s));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t i = i0 + (uint64_t)u * 32 + lane;
      const uint64_t pos0 = __shfl_sync(0xffffffffu, pos[u], 0);
      const bool run = __all_sync(0xffffffffu, i < n && pos[u] == pos0 + (uint64_t)lane);
      if (run) {
        s_w[wid][3 * lane + 0] = w0[u];
        s_w[wid][3 * lane + 1] = w1[u];
        s_w[wid][3 * lane + 2] = w2[u];
        __syncwarp();
        unsigned long long *o = reinterpret_cast<unsigned long long *>(out + pos0);
        o[lane] = s_w[wid][lane];
        o[32 + lane] = s_w[wid][32 + lane];
        o[64 + lane] = s_w[wid][64 + lane];
        __syncwarp();
      } else if (i < n) {
        unsigned long long *o = reinterpret_cast<unsigned long long *>(out + pos[u]);
        o[0] = w0[u];
        o[1] = w1[u];
        o[2] = w2[u];
      }
    }
  }
}

static int bits_for(uint64_t nvalues) {
  int b = 1;
  while (b < 64 && (nvalues - 1) >> b) ++b;
  return b;
}

// ---- gather ----------------------------------------------------------------------------------
// Called by the search after the local sort, on the ctx stream: d_hits = the rank's n hits in
// reference order, keys = their sorted one-word keys (query | table | id; the table field starts
// at bit tshift).  Segment offsets are taken from the keys now (the sort scratch is reused by
// whatever runs next); the rest runs on the gather stream.
int comm_gather_flush(hs_ctx *ctx);
int comm_gather_start(hs_ctx *ctx, const hs_hit *d_hits, uint64_t n, const uint64_t *d_keys, int tshift, uint32_t Q,
                      bool local_overflow) {
  const int G = ctx->nranks;
  HS_TRY(comm_gather_flush(ctx));   // (an older transfer no filter launch has picked up)
  if (!ctx->recv_cap) {
    set_error("hs_search on a context that joined a communicator: call hs_comm_reserve first");
    return HS_ERR_INVALID;
  }
  const int tbits = bits_for((uint64_t)ctx->prm.L + 1);
  const uint64_t S64 = (uint64_t)Q << tbits;
  if (S64 >= (1ull << 31)) {
    set_error("hs_comm: %llu (query, table) segments exceed this build's limit", (unsigned long long)S64);
    return HS_ERR_UNSUPPORTED;
  }
  const uint32_t S = (uint32_t)S64;
  const int slot = (int)(ctx->gather_seq & 1u);
  HS_TRY(ctx->d_segoff[slot].reserve(sizeof(uint64_t) * ((size_t)S + 1)));
  HS_TRY(ctx->d_segcnt.reserve(sizeof(uint32_t) * std::max<size_t>(S, 1)));
  HS_TRY(ctx->d_segcnt_all.reserve(sizeof(uint32_t) * std::max<size_t>((size_t)S * G, 1)));
  HS_TRY(ctx->d_segdst.reserve(sizeof(uint64_t) * std::max<size_t>(S, 1)));
  uint64_t *off = ctx->d_segoff[slot].as<uint64_t>();
  seg_offsets_kernel<<<(unsigned)((n + 256) / 256), 256, 0, ctx->stream>>>(d_keys, n, tshift, S, off);
  HS_CUDA(cudaGetLastError());
  HS_CUDA(cudaEventRecord(ctx->ev_gather_in, ctx->stream));
  ctx->stats.kernel_launches += 1;

  cudaStream_t gs = ctx->gather_stream;
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm2;
  unsigned long long *info = ctx->d_gather_info.as<unsigned long long>() + 2 * slot;
  HS_CUDA(cudaStreamWaitEvent(gs, ctx->ev_gather_in, 0));
  if (S) {
    seg_counts_kernel<<<(S + 255) / 256, 256, 0, gs>>>(off, S, ctx->d_segcnt.as<uint32_t>());
    HS_NCCL(g_nccl.AllGather(ctx->d_segcnt.p, ctx->d_segcnt_all.p, S, ncclUint32, comm, gs));
  }
  seg_dest_kernel<<<1, kDestThreads, 0, gs>>>(ctx->d_segcnt_all.as<uint32_t>(), S, G, ctx->rank, ctx->recv_cap,
                                              ctx->d_segdst.as<uint64_t>(), info);
  HS_CUDA(cudaGetLastError());
  // The bulk transfer is deferred to the next search's filter launch (comm_gather_flush): while a
  // sender pushes into rank 0's congested link its own memory-bound kernels crawl (measured at 8
  // ranks: the next batch's rank sort 6.8 -> 15.6 ms, bucket grouping 0.4 -> 3.3 ms), whereas the
  // tensor filter barely touches HBM and hides the transfer.  hs_comm_result flushes it too.
  ctx->gather_def.pending = true;
  ctx->gather_def.hits = d_hits;
  ctx->gather_def.n = n;
  ctx->gather_def.tbits = tbits;
  ctx->gather_def.slot = slot;
  ctx->gather_def.overflow = local_overflow;
  ctx->stats.kernel_launches += 3;
  ctx->gather_seq++;
  ctx->gather_pending = true;
  return HS_OK;
}

// Starts the deferred bulk transfer of the latest gather (no-op when there is none).
int comm_gather_flush(hs_ctx *ctx) {
  if (!ctx->gather_def.pending) return HS_OK;
  ctx->gather_def.pending = false;
  const int G = ctx->nranks, slot = ctx->gather_def.slot;
  const uint64_t n = ctx->gather_def.n;
  cudaStream_t gs = ctx->gather_stream;
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm2;
  unsigned long long *info = ctx->d_gather_info.as<unsigned long long>() + 2 * slot;
  if (n) {
    // Peer stores are latency-bound per thread block (4 K hits in flight, ~40 GB/s), and the senders
    // share rank 0's ingress (900 GB/s): as many blocks as fill this sender's share, no more -- the
    // blocks start before the filter's and keep their SMs from it for the length of the transfer.
    const unsigned want = (unsigned)std::min(32, std::max(4, 900 / (39 * std::max(1, G - 1)) + 1));
    const unsigned grid = (unsigned)std::min<uint64_t>((n + 1023) / 1024, want);
    scatter_merged_kernel<<<grid, kScatterWarps * 32, 0, gs>>>(ctx->gather_def.hits, n, ctx->gather_def.tbits, ctx->d_segoff[slot].as<uint64_t>(),
                                                ctx->d_segdst.as<uint64_t>(), info,
                                                reinterpret_cast<hs_hit *>(ctx->recv_mapped[slot]));
  }
  HS_CUDA(cudaGetLastError());
  // every rank's stores have left its GPU when its kernel has ended; the all-reduce that follows in
  // stream order is the barrier after which rank 0 may read the merged list
  // (its value counts the ranks whose own hit buffer overflowed: their lists are missing)
  ctx->h_gather_flag[slot] = ctx->gather_def.overflow ? 1ull : 0ull;
  HS_CUDA(cudaMemcpyAsync(ctx->d_gather_info.as<unsigned long long>() + 4, &ctx->h_gather_flag[slot], sizeof(unsigned long long),
                          cudaMemcpyHostToDevice, gs));
  HS_NCCL(g_nccl.AllReduce(ctx->d_gather_info.as<unsigned long long>() + 4, ctx->d_gather_info.as<unsigned long long>() + 5,
                           1, ncclUint64, ncclSum, comm, gs));
  HS_CUDA(cudaEventRecord(ctx->ev_gather[slot], gs));
  return HS_OK;
}

void comm_destroy(hs_ctx *ctx) {
  if (ctx->gather_stream) {
    comm_gather_flush(ctx);
    cudaStreamSynchronize(ctx->gather_stream);
  }
  for (int i = 0; i < 2; ++i) {
    if (ctx->recv_mapped[i] && ctx->recv_mapped[i] != ctx->recv_local[i]) cudaIpcCloseMemHandle(ctx->recv_mapped[i]);
    if (ctx->recv_local[i]) cudaFree(ctx->recv_local[i]);
    ctx->recv_mapped[i] = ctx->recv_local[i] = nullptr;
    if (ctx->ev_gather[i]) cudaEventDestroy(ctx->ev_gather[i]);
    ctx->ev_gather[i] = nullptr;
    ctx->d_segoff[i].release();
  }
  if (ctx->ev_gather_in) cudaEventDestroy(ctx->ev_gather_in);
  ctx->ev_gather_in = nullptr;
  ctx->d_segcnt.release();
  ctx->d_segcnt_all.release();
  ctx->d_segdst.release();
  ctx->d_gather_info.release();
  ctx->recv_cap = 0;
  if (ctx->nccl_comm2 && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm2);
  if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
  ctx->nccl_comm = ctx->nccl_comm2 = nullptr;
  if (ctx->gather_stream) cudaStreamDestroy(ctx->gather_stream);
  ctx->gather_stream = nullptr;
  ctx->nranks = 1;
  ctx->rank = 0;
}

}  // namespace hs

using namespace hs;

extern "C" {

int hs_comm_unique_id(void *out128) {
  if (!out128) return HS_ERR_INVALID;
  HS_TRY(load_nccl());
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) {
    set_error("ncclGetUniqueId failed");
    return HS_ERR_COMM;
  }
  memcpy(out128, &id, sizeof id);
  return HS_OK;
}

int hs_comm_init(hs_ctx_t *ctx, const void *nccl_unique_id, int rank, int nranks) {
  if (!ctx || !nccl_unique_id || nranks < 1 || rank < 0 || rank >= nranks) {
    set_error("hs_comm_init: bad argument");
    return HS_ERR_INVALID;
  }
  if (ctx->nccl_comm) {
    set_error("hs_comm_init: the context already joined a communicator");
    return HS_ERR_INVALID;
  }
  HS_TRY(load_nccl());
  HS_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, nccl_unique_id, sizeof id);
  ncclComm_t comm;
  ncclResult_t r = g_nccl.CommInitRank(&comm, nranks, id, rank);
  if (r != ncclSuccess) {
    set_error("ncclCommInitRank: %s", g_nccl.GetErrorString(r));
    return HS_ERR_COMM;
  }
  ctx->nccl_comm = comm;
  ctx->rank = rank;
  ctx->nranks = nranks;
  if (nranks == 1) return HS_OK;
  // a second communicator for the gather stream, so that the query broadcast of the next search
  // (ctx stream) and a gather still in flight never share one: its id travels over the first
  HS_TRY(ctx->d_gather_info.reserve(256));
  HS_CUDA(cudaMemsetAsync(ctx->d_gather_info.p, 0, 256, ctx->stream));
  ncclUniqueId id2;
  memset(&id2, 0, sizeof id2);
  if (rank == 0 && g_nccl.GetUniqueId(&id2) != ncclSuccess) {
    set_error("ncclGetUniqueId failed");
    return HS_ERR_COMM;
  }
  HS_CUDA(cudaMemcpyAsync((char *)ctx->d_gather_info.p + 64, &id2, sizeof id2, cudaMemcpyHostToDevice, ctx->stream));
  HS_NCCL(g_nccl.Broadcast((char *)ctx->d_gather_info.p + 64, (char *)ctx->d_gather_info.p + 64, sizeof id2, ncclUint8, 0, comm,
                           ctx->stream));
  HS_CUDA(cudaMemcpyAsync(&id2, (char *)ctx->d_gather_info.p + 64, sizeof id2, cudaMemcpyDeviceToHost, ctx->stream));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  ncclComm_t comm2;
  r = g_nccl.CommInitRank(&comm2, nranks, id2, rank);
  if (r != ncclSuccess) {
    set_error("ncclCommInitRank (gather communicator): %s", g_nccl.GetErrorString(r));
    return HS_ERR_COMM;
  }
  ctx->nccl_comm2 = comm2;
  // lowest priority: where the merge's blocks and the next batch's kernels both wait for an SM, the batch goes first
  int prio_lo = 0, prio_hi = 0;
  HS_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  HS_CUDA(cudaStreamCreateWithPriority(&ctx->gather_stream, cudaStreamNonBlocking, prio_lo));
  for (int i = 0; i < 2; ++i) HS_CUDA(cudaEventCreateWithFlags(&ctx->ev_gather[i], cudaEventDisableTiming));
  HS_CUDA(cudaEventCreateWithFlags(&ctx->ev_gather_in, cudaEventDisableTiming));
  return HS_OK;
}

int hs_comm_reserve(hs_ctx_t *ctx, uint64_t cap_hits) {
  if (!ctx || ctx->nranks <= 1 || !ctx->nccl_comm2) {
    set_error("hs_comm_reserve: the context has not joined a communicator of more than one rank");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  HS_CUDA(cudaStreamSynchronize(ctx->gather_stream));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  // every rank must ask for the same capacity: the maximum wins
  unsigned long long *scr = ctx->d_gather_info.as<unsigned long long>() + 6;
  HS_CUDA(cudaMemcpyAsync(scr, &cap_hits, sizeof cap_hits, cudaMemcpyHostToDevice, ctx->stream));
  HS_NCCL(g_nccl.AllReduce(scr, scr, 1, ncclUint64, ncclMax, comm, ctx->stream));
  HS_CUDA(cudaMemcpyAsync(&cap_hits, scr, sizeof cap_hits, cudaMemcpyDeviceToHost, ctx->stream));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  if (cap_hits == 0) cap_hits = 1;
  if (cap_hits <= ctx->recv_cap) return HS_OK;
  for (int i = 0; i < 2; ++i) {
    if (ctx->recv_mapped[i] && ctx->recv_mapped[i] != ctx->recv_local[i]) cudaIpcCloseMemHandle(ctx->recv_mapped[i]);
    if (ctx->recv_local[i]) cudaFree(ctx->recv_local[i]);
    ctx->recv_mapped[i] = ctx->recv_local[i] = nullptr;
  }
  cudaIpcMemHandle_t hnd[2];
  memset(hnd, 0, sizeof hnd);
  int ok = 1;
  if (ctx->rank == 0) {
    for (int i = 0; i < 2 && ok; ++i) {
      if (cudaMalloc(&ctx->recv_local[i], sizeof(hs_hit) * cap_hits) != cudaSuccess ||
          cudaIpcGetMemHandle(&hnd[i], ctx->recv_local[i]) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
      }
    }
  }
  // handles (and rank 0's verdict) to every rank over the first communicator
  char *xfer = (char *)ctx->d_gather_info.p + 64;
  static_assert(2 * sizeof(cudaIpcMemHandle_t) + sizeof(int) <= 192, "exchange area");
  HS_CUDA(cudaMemcpyAsync(xfer, hnd, sizeof hnd, cudaMemcpyHostToDevice, ctx->stream));
  HS_CUDA(cudaMemcpyAsync(xfer + sizeof hnd, &ok, sizeof ok, cudaMemcpyHostToDevice, ctx->stream));
  HS_NCCL(g_nccl.Broadcast(xfer, xfer, sizeof hnd + sizeof ok, ncclUint8, 0, comm, ctx->stream));
  HS_CUDA(cudaMemcpyAsync(hnd, xfer, sizeof hnd, cudaMemcpyDeviceToHost, ctx->stream));
  HS_CUDA(cudaMemcpyAsync(&ok, xfer + sizeof hnd, sizeof ok, cudaMemcpyDeviceToHost, ctx->stream));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  if (!ok) {
    set_error("hs_comm_reserve: rank 0 could not allocate 2 x %llu hits", (unsigned long long)cap_hits);
    return HS_ERR_NOMEM;
  }
  for (int i = 0; i < 2; ++i) {
    if (ctx->rank == 0) ctx->recv_mapped[i] = ctx->recv_local[i];
    else HS_CUDA(cudaIpcOpenMemHandle(&ctx->recv_mapped[i], hnd[i], cudaIpcMemLazyEnablePeerAccess));
  }
  ctx->recv_cap = cap_hits;
  return HS_OK;
}

int hs_comm_result(hs_ctx_t *ctx, const void **hits_dev, uint64_t *nhits_total) {
  if (!ctx || !nhits_total) {
    set_error("hs_comm_result: null argument");
    return HS_ERR_INVALID;
  }
  if (hits_dev) *hits_dev = nullptr;
  *nhits_total = 0;
  if (ctx->nranks <= 1 || ctx->gather_seq == 0) {
    set_error("hs_comm_result: no gathered search on this context");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  HS_TRY(comm_gather_flush(ctx));
  const int slot = (int)((ctx->gather_seq - 1) & 1u);
  unsigned long long info[2] = {0, 0}, lost = 0;
  HS_CUDA(cudaMemcpyAsync(info, ctx->d_gather_info.as<unsigned long long>() + 2 * slot, sizeof info, cudaMemcpyDeviceToHost,
                          ctx->gather_stream));
  HS_CUDA(cudaMemcpyAsync(&lost, ctx->d_gather_info.as<unsigned long long>() + 5, sizeof lost, cudaMemcpyDeviceToHost,
                          ctx->gather_stream));
  HS_CUDA(cudaStreamSynchronize(ctx->gather_stream));
  ctx->gather_pending = false;
  *nhits_total = info[0];
  if (lost) {
    set_error("%llu rank(s) overflowed their own hit buffer: their hits are missing from the merged list", lost);
    return HS_ERR_CAPACITY;
  }
  if (info[1]) {
    set_error("receive buffer too small: %llu hits over all ranks, capacity %llu (hs_comm_reserve)", info[0],
              (unsigned long long)ctx->recv_cap);
    return HS_ERR_CAPACITY;
  }
  if (hits_dev && ctx->rank == 0) *hits_dev = ctx->recv_local[slot];
  return HS_OK;
}
}
