// K2 interfaces (radix_sort.cu).
#pragma once
#include "common.cuh"
namespace hs {
struct KeyPtrs {
  uint64_t *w[kMaxKeyWords];
};
int exclusive_scan_u32(hs_ctx *ctx, const uint32_t *in, uint32_t *out, uint64_t n, uint32_t *d_total);
int radix_sort_pairs(hs_ctx *ctx, const KeyPtrs &keys_in, const uint32_t *vals_in, uint64_t n, int nw,
                     uint32_t *v_final, uint32_t *v_tmp, KeyPtrs *sorted_keys, uint64_t vary_hint = 0);
int build_table_index(hs_ctx *ctx, uint32_t table, cudaEvent_t ev_sort_end, cudaEvent_t ev_group_end, bool with_store);
int build_table_store(hs_ctx *ctx, uint32_t table);
int build_code_stores_blocked(hs_ctx *ctx, bool *done);
int build_code_store(hs_ctx *ctx, const uint32_t *ids, DevBuf &out);
}  // namespace hs
