// Hit-list utilities: the order-independent checksum that compares an N-GPU search with the
// 1-GPU search of the same database (SURVEY.md 7 "1-GPU vs N-GPU checksums"), on the device and
// on the host with the same arithmetic.
#include <algorithm>

#include "common.cuh"
#include "internal.cuh"

namespace hs {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
  x ^= x >> 30;
  x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27;
  x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}
// One hit -> 64 bits; every field (query, first table, db id, the bit pattern of dist2) matters.
__host__ __device__ __forceinline__ uint64_t hit_digest(uint32_t query, uint32_t table_first, uint64_t db_id, uint64_t d2bits) {
  const uint64_t qt = ((uint64_t)query << 32) | (uint64_t)table_first;
  return mix64(qt ^ mix64(db_id + 0x9e3779b97f4a7c15ull + mix64(d2bits ^ 0xd1b54a32d192ed03ull)));
}

__global__ void __launch_bounds__(256) hits_checksum_kernel(const hs_hit *__restrict__ hits, uint64_t n,
                                                            unsigned long long *__restrict__ sum) {
  unsigned long long mine = 0ull;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const hs_hit h = hits[i];
    mine += hit_digest(h.query, h.table_first, h.db_id, (uint64_t)__double_as_longlong(h.dist2));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
  __shared__ unsigned long long s_part[8];
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0ull;
    for (int w = 0; w < 8; ++w) t += s_part[w];
    atomicAdd(sum, t);  // sums mod 2^64 commute: the result does not depend on the order
  }
}

constexpr int kChecksumSlot = 31;  // d_counters slot

}  // namespace hs

using namespace hs;

extern "C" {

int hs_hits_checksum(const hs_hit *hits, uint64_t n, uint64_t *sum_out) {
  if ((!hits && n) || !sum_out) {
    set_error("hs_hits_checksum: null argument");
    return HS_ERR_INVALID;
  }
  uint64_t s = 0;
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t bits;
    memcpy(&bits, &hits[i].dist2, sizeof bits);
    s += hit_digest(hits[i].query, hits[i].table_first, hits[i].db_id, bits);
  }
  *sum_out = s;
  return HS_OK;
}

int hs_hits_checksum_dev(hs_ctx_t *ctx, const void *hits_dev, uint64_t n, uint64_t *sum_out) {
  if (!ctx || (!hits_dev && n) || !sum_out) {
    set_error("hs_hits_checksum_dev: null argument");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  unsigned long long *d_sum = ctx->d_counters.as<unsigned long long>() + kChecksumSlot;
  HS_CUDA(cudaMemsetAsync(d_sum, 0, sizeof(unsigned long long), ctx->stream));
  if (n) {
    const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->num_sms * 8);
    hits_checksum_kernel<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const hs_hit *>(hits_dev), n, d_sum);
    HS_CUDA(cudaGetLastError());
  }
  unsigned long long h = 0;
  HS_TRY(read_back(ctx, d_sum, &h, sizeof h));
  *sum_out = h;
  return HS_OK;
}
}
