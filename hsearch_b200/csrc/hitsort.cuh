// Segmented hit sort (hitsort.cu): partition by the top key bits, per-bin sort in shared memory.
#pragma once
#include "common.cuh"
namespace hs {
struct SegSortRequest {
  bool compact = false;          // compact (CSR) output instead of hs_hit records
  hs_hit *hits_out = nullptr;    // plain output
  uint32_t *idt = nullptr;       // compact output: local id | table << id_bits, dist2, offsets[qa .. qb]
  double *dist2 = nullptr;
  uint64_t *offsets = nullptr;
  uint32_t qa = 0, qb = 0;
  uint64_t base = 0;
  int id_bits = 0;
};
// Sorts d_hits[0, n) by (query, first table, db id) into the requested format.  *used = false: nothing
// was produced (switched off, the key fields do not fit, or a bin could not be processed) and the
// caller must take the radix path; d_hits is never modified.
int sort_hits_segmented(hs_ctx *ctx, const hs_hit *d_hits, uint64_t n, const SegSortRequest &rq, bool *used);
}  // namespace hs
