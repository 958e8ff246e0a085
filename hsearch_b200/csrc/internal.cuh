// Helpers shared between api.cu and cluster.cu.
#pragma once
#include "common.cuh"
#include "verify.cuh"
namespace hs {
void stats_begin(hs_ctx *ctx);
float ev_ms(cudaEvent_t a, cudaEvent_t b);
float filter_threshold(const hs_ctx *ctx);
const uint8_t *const *dev_stores(hs_ctx *ctx);
const uint32_t *const *dev_sorted_ids(hs_ctx *ctx);
const uint64_t *const *dev_keys(hs_ctx *ctx);
void fill_exact_common(hs_ctx *ctx, ExactArgs &ea, uint32_t Q);
struct MmaLaunch {
  uint32_t grid = 0;               // CTAs of the pipelined tensor filter (0: not used)
  uint32_t nunits = 0;
  const uint32_t *qlist = nullptr; // its query list (device)
};
int run_filter(hs_ctx *ctx, FilterArgs &fa, uint32_t nblocks, int mode, uint64_t *nsurv_out,
               FilterArgs *fa_tc = nullptr, uint32_t nblocks_tc = 0, const MmaLaunch *ml = nullptr);
int ensure_identity_store(hs_ctx *ctx);
int ensure_code_stores(hs_ctx *ctx);
bool selfjoin_uses_mma(const hs_ctx *ctx, uint32_t n);
int selfjoin_bucket_mma(hs_ctx *ctx, uint32_t table, uint32_t mb, uint32_t me, uint32_t q_lo, uint32_t q_hi,
                        uint64_t *nsurv, uint64_t *npairs, bool *used);
// control traffic moved by kernels through mapped pinned memory instead of the copy engines (api.cu)
int read_back(hs_ctx *ctx, const void *d_src, void *h_dst, size_t bytes);  // device -> host, synchronising
int upload(hs_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);     // host -> device, asynchronous, h_src may be a temporary
int validate_codes(hs_ctx *ctx, const char *who);  // residue codes of the loaded DB are 0..19 (synchronising)
}  // namespace hs
