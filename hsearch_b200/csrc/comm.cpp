// Multi-GPU exchange step of the search path: queries are
// broadcast from rank 0 and every rank's hit list is gathered to rank 0.  NCCL
// is loaded lazily with dlopen so that single-GPU use needs no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "host_tables.h"

namespace hs {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
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
  HS_SYM(Send, "ncclSend")
  HS_SYM(Recv, "ncclRecv")
  HS_SYM(GroupStart, "ncclGroupStart")
  HS_SYM(GroupEnd, "ncclGroupEnd")
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

// Gather every rank's first min(*nhits, cap) hits to rank 0 (ctx->d_hits_gathered).
// On return *nhits = total over ranks (all ranks).
int comm_gather_hits(hs_ctx *ctx, hs_hit *d_hits, uint64_t *nhits, uint64_t cap) {
  const int G = ctx->nranks;
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  HS_TRY(ctx->d_misc.reserve(sizeof(uint64_t) * (G + 1)));
  uint64_t *d_counts = ctx->d_misc.as<uint64_t>();
  const uint64_t mine = *nhits < cap ? *nhits : cap;
  HS_CUDA(cudaMemcpyAsync(d_counts + G, &mine, sizeof mine, cudaMemcpyHostToDevice, ctx->stream));
  HS_NCCL(g_nccl.AllGather(d_counts + G, d_counts, 1, ncclUint64, comm, ctx->stream));
  std::vector<uint64_t> counts(G);
  HS_CUDA(cudaMemcpyAsync(counts.data(), d_counts, sizeof(uint64_t) * G, cudaMemcpyDeviceToHost, ctx->stream));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  uint64_t total = 0;
  for (int r = 0; r < G; ++r) total += counts[r];
  if (ctx->rank == 0) HS_TRY(ctx->d_hits_gathered.reserve(sizeof(hs_hit) * (total ? total : 1)));
  HS_NCCL(g_nccl.GroupStart());
  if (ctx->rank == 0) {
    uint64_t off = counts[0];
    for (int r = 1; r < G; ++r) {
      if (counts[r])
        HS_NCCL(g_nccl.Recv(ctx->d_hits_gathered.as<hs_hit>() + off, counts[r] * sizeof(hs_hit), ncclUint8, r, comm,
                            ctx->stream));
      off += counts[r];
    }
  } else if (mine) {
    HS_NCCL(g_nccl.Send(d_hits, mine * sizeof(hs_hit), ncclUint8, 0, comm, ctx->stream));
  }
  HS_NCCL(g_nccl.GroupEnd());
  if (ctx->rank == 0 && mine)
    HS_CUDA(cudaMemcpyAsync(ctx->d_hits_gathered.p, d_hits, mine * sizeof(hs_hit), cudaMemcpyDeviceToDevice,
                            ctx->stream));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  *nhits = total;
  return HS_OK;
}

void comm_destroy(hs_ctx *ctx) {
  if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
  ctx->nranks = 1;
  ctx->rank = 0;
}

}  // namespace hs

extern "C" {

int hs_comm_unique_id(void *out128) {
  if (!out128) return HS_ERR_INVALID;
  HS_TRY(hs::load_nccl());
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (hs::g_nccl.GetUniqueId(&id) != ncclSuccess) {
    hs::set_error("ncclGetUniqueId failed");
    return HS_ERR_COMM;
  }
  memcpy(out128, &id, sizeof id);
  return HS_OK;
}

int hs_comm_init(hs_ctx_t *ctx, const void *nccl_unique_id, int rank, int nranks) {
  if (!ctx || !nccl_unique_id || nranks < 1 || rank < 0 || rank >= nranks) {
    hs::set_error("hs_comm_init: bad argument");
    return HS_ERR_INVALID;
  }
  HS_TRY(hs::load_nccl());
  HS_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, nccl_unique_id, sizeof id);
  ncclComm_t comm;
  ncclResult_t r = hs::g_nccl.CommInitRank(&comm, nranks, id, rank);
  if (r != ncclSuccess) {
    hs::set_error("ncclCommInitRank: %s", hs::g_nccl.GetErrorString(r));
    return HS_ERR_COMM;
  }
  ctx->nccl_comm = comm;
  ctx->rank = rank;
  ctx->nranks = nranks;
  return HS_OK;
}
}
