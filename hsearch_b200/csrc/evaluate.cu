// R1: recall of a hit list against a ground-truth list, on the device.
//
// The reference's `evaulate` (hclust/src/hclust/motif_both_points.cpp:100-165) reads both lists
// back from text, sorts them by (motif, protein) and merge-joins them: a ground-truth pair that
// the search also reported adds weight(dis) (motif_both_points.cpp:66-86) to tp, one that it
// missed adds it to fn, and each is counted in the bin int(dis*100/10) that becomes a row of
// <out>.accuracy.txt (:151-163); reported pairs that are not in the ground truth are only
// printed ("xnomo", :127-129).  Pairs are unique in both lists, so the merge-join is a set
// intersection and needs no common order: here every ground-truth hit looks itself up in the
// found list, which is in the order hs_search_* returns it (query, first table, db id --
// HS_FLAG_SORT_HITS), by binary search inside its query's segment, once per table.  Counts and
// bins are integers (exact); the weighted sums add the weights that equal 1 as a count and the
// fractional ones per thread -> warp -> block -> grid in a fixed order, so they are
// reproducible but not in the reference's sequential order: tolerance 1e-12 relative.
#include <algorithm>

#include "common.cuh"

namespace hs {
namespace {

constexpr int kRecallThreads = 256;

struct RecallCounters {
  unsigned long long n_tp, n_fn, ones_tp, ones_fn, n_over, n_badq, unsorted, pad;
  unsigned long long tp_bin[HS_RECALL_BINS], fn_bin[HS_RECALL_BINS];
};

__device__ __forceinline__ bool hit_less(uint32_t qa, uint32_t ta, uint64_t ia, uint32_t qb, uint32_t tb, uint64_t ib) {
  if (qa != qb) return qa < qb;
  if (ta != tb) return ta < tb;
  return ia < ib;
}

// found must be strictly ascending by (query, table_first, db id)
__global__ void recall_order_kernel(const hs_hit *__restrict__ found, uint64_t nf, RecallCounters *c) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (i >= nf) return;
  const hs_hit a = found[i - 1], b = found[i];
  if (!hit_less(a.query, a.table_first, a.db_id, b.query, b.table_first, b.db_id)) atomicAdd(&c->unsorted, 1ull);
}

// seg[q] = first index of found whose query >= q, q in [0, Q]
__global__ void recall_segments_kernel(const hs_hit *__restrict__ found, uint64_t nf, uint32_t Q, uint64_t *__restrict__ seg) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q > Q) return;
  uint64_t lo = 0, hi = nf;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (found[mid].query < q) lo = mid + 1;
    else hi = mid;
  }
  seg[q] = lo;
}

// weight(), motif_both_points.cpp:66-86, for dis <= R + 0.1 (the caller rejects the rest)
__device__ __forceinline__ double recall_weight(double dis) {
  if (dis < 0.0000001) return 1.0;
  if (dis < 24) return 1.0;
  const double w = 1 / (dis - 24);
  if (w > 1) return 1.0;
  if (w < 0) return 1.0;
  return w;
}

__global__ void __launch_bounds__(kRecallThreads)
recall_join_kernel(const hs_hit *__restrict__ truth, uint64_t nt, const hs_hit *__restrict__ found,
                   const uint64_t *__restrict__ seg, uint32_t Q, uint32_t ntab, int euclid, double R,
                   RecallCounters *c, double *__restrict__ part_tp, double *__restrict__ part_fn) {
  __shared__ unsigned int s_tp[HS_RECALL_BINS], s_fn[HS_RECALL_BINS];
  __shared__ double s_wtp[kRecallThreads / 32], s_wfn[kRecallThreads / 32];
  for (int i = threadIdx.x; i < HS_RECALL_BINS; i += blockDim.x) s_tp[i] = s_fn[i] = 0;
  __syncthreads();
  double wtp = 0.0, wfn = 0.0;
  unsigned int ntp = 0, nfn = 0, otp = 0, ofn = 0, nover = 0, nbadq = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nt; i += (uint64_t)gridDim.x * blockDim.x) {
    const hs_hit g = truth[i];
    if (g.query >= Q) {
      ++nbadq;
      continue;
    }
    const double dis = euclid ? sqrt(g.dist2) : g.dist2;
    if (dis > R + 0.1) {  // the reference prints "err" and exits (:67-70)
      ++nover;
      continue;
    }
    const uint64_t s0 = seg[g.query], s1 = seg[g.query + 1];
    bool hit = false;
    for (uint32_t t = 0; t < ntab && !hit; ++t) {
      uint64_t lo = s0, hi = s1;
      while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        const uint32_t tf = found[mid].table_first;
        const uint64_t id = found[mid].db_id;
        if (tf < t || (tf == t && id < g.db_id)) lo = mid + 1;
        else hi = mid;
      }
      hit = lo < s1 && found[lo].table_first == t && found[lo].db_id == g.db_id;
    }
    const double w = recall_weight(dis);
    const int bin = int(dis * 100 / 10);
    if (hit) {
      ++ntp;
      if (w == 1.0) ++otp;
      else wtp += w;
      if (bin >= 0 && bin < HS_RECALL_BINS) atomicAdd(&s_tp[bin], 1u);
    } else {
      ++nfn;
      if (w == 1.0) ++ofn;
      else wfn += w;
      if (bin >= 0 && bin < HS_RECALL_BINS) atomicAdd(&s_fn[bin], 1u);
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    wtp += __shfl_down_sync(0xffffffffu, wtp, o);
    wfn += __shfl_down_sync(0xffffffffu, wfn, o);
    ntp += __shfl_down_sync(0xffffffffu, ntp, o);
    nfn += __shfl_down_sync(0xffffffffu, nfn, o);
    otp += __shfl_down_sync(0xffffffffu, otp, o);
    ofn += __shfl_down_sync(0xffffffffu, ofn, o);
    nover += __shfl_down_sync(0xffffffffu, nover, o);
    nbadq += __shfl_down_sync(0xffffffffu, nbadq, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_wtp[threadIdx.x >> 5] = wtp;
    s_wfn[threadIdx.x >> 5] = wfn;
    if (ntp) atomicAdd(&c->n_tp, (unsigned long long)ntp);
    if (nfn) atomicAdd(&c->n_fn, (unsigned long long)nfn);
    if (otp) atomicAdd(&c->ones_tp, (unsigned long long)otp);
    if (ofn) atomicAdd(&c->ones_fn, (unsigned long long)ofn);
    if (nover) atomicAdd(&c->n_over, (unsigned long long)nover);
    if (nbadq) atomicAdd(&c->n_badq, (unsigned long long)nbadq);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int wdx = 0; wdx < kRecallThreads / 32; ++wdx) {
      a += s_wtp[wdx];
      b += s_wfn[wdx];
    }
    part_tp[blockIdx.x] = a;
    part_fn[blockIdx.x] = b;
  }
  for (int i = threadIdx.x; i < HS_RECALL_BINS; i += blockDim.x) {
    if (s_tp[i]) atomicAdd(&c->tp_bin[i], (unsigned long long)s_tp[i]);
    if (s_fn[i]) atomicAdd(&c->fn_bin[i], (unsigned long long)s_fn[i]);
  }
}

int evaluate_recall_dev(hs_ctx *ctx, const hs_hit *d_truth, uint64_t nt, const hs_hit *d_found, uint64_t nf, uint32_t Q,
                        hs_recall *out) {
  HS_CUDA(cudaSetDevice(ctx->device));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)sms * 8, (nt + kRecallThreads - 1) / kRecallThreads));
  DevBuf scratch;
  const size_t seg_off = (sizeof(RecallCounters) + 255) & ~size_t(255);
  const size_t part_off = seg_off + ((sizeof(uint64_t) * ((size_t)Q + 2) + 255) & ~size_t(255));
  const size_t total = part_off + sizeof(double) * 2 * grid;
  HS_TRY(scratch.reserve(total));
  auto *c = scratch.as<RecallCounters>();
  auto *seg = reinterpret_cast<uint64_t *>(scratch.as<char>() + seg_off);
  auto *part = reinterpret_cast<double *>(scratch.as<char>() + part_off);
  cudaStream_t st = ctx->stream;
  int rc = HS_OK;
  RecallCounters h{};
  std::vector<double> hp(2 * (size_t)grid);
  do {
    if (cudaMemsetAsync(c, 0, sizeof(RecallCounters), st) != cudaSuccess) { rc = HS_ERR_CUDA; break; }
    if (nf > 1) recall_order_kernel<<<(unsigned)((nf - 1 + 255) / 256), 256, 0, st>>>(d_found, nf, c);
    recall_segments_kernel<<<(Q + 1 + 255) / 256, 256, 0, st>>>(d_found, nf, Q, seg);
    if (nt)
      recall_join_kernel<<<grid, kRecallThreads, 0, st>>>(d_truth, nt, d_found, seg, Q, ctx->prm.L,
                                                          ctx->prm.metric == HS_METRIC_EUCLID_FP64, ctx->prm.R, c, part,
                                                          part + grid);
    if (cudaGetLastError() != cudaSuccess) { rc = HS_ERR_CUDA; break; }
    if (cudaMemcpyAsync(&h, c, sizeof h, cudaMemcpyDeviceToHost, st) != cudaSuccess) { rc = HS_ERR_CUDA; break; }
    if (nt && cudaMemcpyAsync(hp.data(), part, sizeof(double) * 2 * grid, cudaMemcpyDeviceToHost, st) != cudaSuccess) {
      rc = HS_ERR_CUDA;
      break;
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) { rc = HS_ERR_CUDA; break; }
  } while (0);
  scratch.release();
  if (rc != HS_OK) {
    set_error("hs_evaluate_recall: CUDA failure: %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
  }
  ctx->stats.kernel_launches += 2 + (nt ? 1 : 0);
  if (h.unsorted) {
    set_error("hs_evaluate_recall: the found list is not in (query, first table, db id) order (%llu inversions)", h.unsorted);
    return HS_ERR_INVALID;
  }
  if (h.n_badq) {
    set_error("hs_evaluate_recall: %llu ground-truth hits name a query >= Q", h.n_badq);
    return HS_ERR_INVALID;
  }
  if (h.n_over) {  // weight() prints "err <dis>" and exits (motif_both_points.cpp:67-70)
    set_error("hs_evaluate_recall: %llu ground-truth distances exceed R + 0.1", h.n_over);
    return HS_ERR_INVALID;
  }
  double ftp = 0.0, ffn = 0.0;
  if (nt)
    for (unsigned b = 0; b < grid; ++b) {
      ftp += hp[b];
      ffn += hp[grid + b];
    }
  out->tp = (double)h.ones_tp + ftp;
  out->fn = (double)h.ones_fn + ffn;
  out->n_tp = h.n_tp;
  out->n_fn = h.n_fn;
  out->n_extra = nf - h.n_tp;
  for (int i = 0; i < HS_RECALL_BINS; ++i) {
    out->tp_bin[i] = h.tp_bin[i];
    out->fn_bin[i] = h.fn_bin[i];
  }
  return HS_OK;
}

}  // namespace
}  // namespace hs

using namespace hs;

extern "C" {

int hs_evaluate_recall_dev(hs_ctx_t *ctx, const void *truth_dev, uint64_t n_truth, const void *found_dev, uint64_t n_found,
                           uint32_t Q, hs_recall *out) {
  if (!ctx || !out || (n_truth && !truth_dev) || (n_found && !found_dev)) {
    set_error("hs_evaluate_recall_dev: null argument");
    return HS_ERR_INVALID;
  }
  return evaluate_recall_dev(ctx, static_cast<const hs_hit *>(truth_dev), n_truth, static_cast<const hs_hit *>(found_dev),
                             n_found, Q, out);
}

int hs_evaluate_recall(hs_ctx_t *ctx, const hs_hit *truth, uint64_t n_truth, const hs_hit *found, uint64_t n_found,
                       uint32_t Q, hs_recall *out) {
  if (!ctx || !out || (n_truth && !truth) || (n_found && !found)) {
    set_error("hs_evaluate_recall: null argument");
    return HS_ERR_INVALID;
  }
  HS_CUDA(cudaSetDevice(ctx->device));
  DevBuf lists;
  HS_TRY(lists.reserve(sizeof(hs_hit) * (n_truth + n_found) + 64));
  hs_hit *d_truth = lists.as<hs_hit>(), *d_found = d_truth + n_truth;
  cudaError_t e = cudaSuccess;
  if (n_truth) e = cudaMemcpyAsync(d_truth, truth, sizeof(hs_hit) * n_truth, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && n_found)
    e = cudaMemcpyAsync(d_found, found, sizeof(hs_hit) * n_found, cudaMemcpyHostToDevice, ctx->stream);
  int rc;
  if (e != cudaSuccess) {
    set_error("hs_evaluate_recall: copy failed: %s", cudaGetErrorString(e));
    rc = HS_ERR_CUDA;
  } else {
    rc = evaluate_recall_dev(ctx, d_truth, n_truth, d_found, n_found, Q, out);
  }
  cudaStreamSynchronize(ctx->stream);
  lists.release();
  return rc;
}
}
