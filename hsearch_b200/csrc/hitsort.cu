// Hits in the reference's output order (query, first table, ascending db id:
// motif_both_points.cpp:224-245) without a full multi-pass radix sort.
//
// The one-word hit key  query | table | db id  is split at `shift`: the top bits (the query,
// the table field's top bits, ...) form at most 2^15 BINS, the rb <= 32 bits below are the
// key inside a bin.  One partition pass puts every hit into its bin, then each bin is sorted
// where it sits:
//
//   seg_hist_kernel     per-block histograms of the bins (shared-memory atomics), bin-major table
//   exclusive_scan_u32  over the (bin, block) table: the start of every block's run in every bin
//   seg_scatter_kernel  each block re-reads its chunk of hits and writes (low key u32, dist2 f64)
//                       to the next free position of the hit's bin (shared-memory cursors)
//   seg_sort_kernel     one 1024-thread block per SM, bins handed out by a counter.  A bin's keys fall
//                       into 4096 buckets by their leading bits: histogram, block scan and one scatter
//                       of (key, position) pairs put every bucket in place in shared memory, each bucket
//                       (a few keys when ids are spread evenly) is insertion-sorted by one thread, and
//                       the records are written in their final format.  A bin larger than the buffer is
//                       taken in ranges of its top 8 key bits, each range <= the buffer.  A range whose
//                       keys pile up in few buckets is sorted by stable LSD radix passes (8-bit digits,
//                       ballot ranking) instead, its distances placed by binary search (the keys of a
//                       bin are unique -- a (query, fragment) pair is reported once).
//
// Traffic: 24 + 24 + 12 + 16 bytes read, 12 + 12..24 written per hit, against 6 passes x 24 bytes
// plus keys and a random 24-byte gather for the radix sort (api.cu, sort_hits: the fallback
// whenever a field does not fit, one value of a bin's top 8 key bits alone exceeds the buffer, or
// the caller needs the sorted one-word keys).  Measured at bench C2 (87 M hits): 6.6 ms against
// 7.9 ms (profiles/r02c_hitsort.md).  The kernels run unchanged under the CPU emulation of
// tests/emu/ (tests/test_emu_segsort.py).
#include <stdlib.h>

#include <algorithm>

#include "internal.cuh"
#include "sort.cuh"
#include "hitsort.cuh"

namespace hs {

constexpr int kSegThreads = 1024;
constexpr int kSegWarps = kSegThreads / 32;
constexpr int kSegItems = 4;
constexpr int kSegTile = kSegThreads * kSegItems;     // keys ranked per step of a radix pass
constexpr uint32_t kSegBufMax = 22528;                // (key, position) pairs of the per-bin sort's buffer = keys per radix buffer (176 KB)
constexpr int kSegBucketBits = 12;
constexpr int kSegBuckets = 1 << kSegBucketBits;      // buckets of the per-bin bucket sort
constexpr uint32_t kSegBucketMax = 48;                // keys in the fullest bucket up to which a range takes the bucket path
#ifndef HS_SEG_BIN_BITS
#define HS_SEG_BIN_BITS 15
#endif
constexpr int kSegMaxBinBits = HS_SEG_BIN_BITS;       // <= 32768 bins (cursors of the partition: 128 KB)

struct SegFields {
  int tshift, qshift;     // key = query << qshift | table << tshift | db id
  int qbits;
  int shift;              // bin = key >> shift
  int rb;                 // bits of the key inside a bin (<= 32)
  uint32_t nbins;
  uint32_t nblk;          // blocks of the partition kernels
  uint64_t chunk;         // hits per partition block
};

// bin and low key of a hit; false when a field does not fit its bits (the caller falls back)
__device__ __forceinline__ bool seg_key(const hs_hit &h, const SegFields &f, uint32_t &bin, uint32_t &low) {
  if ((h.db_id >> f.tshift) || ((uint64_t)h.table_first >> (f.qshift - f.tshift)) || ((uint64_t)h.query >> f.qbits))
    return false;
  const uint64_t key = ((uint64_t)h.query << f.qshift) | ((uint64_t)h.table_first << f.tshift) | h.db_id;
  const uint64_t b = key >> f.shift;
  if (b >= (uint64_t)f.nbins) return false;
  bin = (uint32_t)b;
  low = f.rb >= 32 ? (uint32_t)key : (uint32_t)(key & ((1ull << f.rb) - 1ull));
  return true;
}

// (query, table, db id) of a hit: two 8-byte loads (hs_hit is 24 bytes: 8-byte aligned only)
__device__ __forceinline__ hs_hit seg_load_key(const hs_hit *p) {
  const uint2 qt = __ldg(reinterpret_cast<const uint2 *>(p));
  hs_hit h;
  h.query = qt.x;
  h.table_first = qt.y;
  h.db_id = __ldg(reinterpret_cast<const unsigned long long *>(p) + 1);
  h.dist2 = 0.0;
  return h;
}

__global__ void __launch_bounds__(kSegThreads)
seg_hist_kernel(const hs_hit *__restrict__ hits, uint64_t n, SegFields f, uint32_t *__restrict__ hist /*[nbins][nblk]*/,
                unsigned int *__restrict__ flags) {
  extern __shared__ uint32_t seg_cnt[];
  for (uint32_t b = threadIdx.x; b < f.nbins; b += kSegThreads) seg_cnt[b] = 0u;
  __syncthreads();
  const uint64_t beg = (uint64_t)blockIdx.x * f.chunk;
  const uint64_t end = beg + f.chunk < n ? beg + f.chunk : n;
  bool bad = false;
  for (uint64_t i = beg + threadIdx.x; i < end; i += kSegThreads) {
    const hs_hit h = seg_load_key(hits + i);
    uint32_t bin, low;
    if (seg_key(h, f, bin, low)) atomicAdd(&seg_cnt[bin], 1u);
    else bad = true;
  }
  if (bad) flags[0] = 1u;
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < f.nbins; b += kSegThreads) hist[(uint64_t)b * f.nblk + blockIdx.x] = seg_cnt[b];
}

__global__ void __launch_bounds__(kSegThreads)
seg_scatter_kernel(const hs_hit *__restrict__ hits, uint64_t n, SegFields f, const uint32_t *__restrict__ start /*[nbins][nblk], scanned*/,
                   uint32_t *__restrict__ pkey, double *__restrict__ pdist) {
  extern __shared__ uint32_t seg_cur[];
  for (uint32_t b = threadIdx.x; b < f.nbins; b += kSegThreads) seg_cur[b] = start[(uint64_t)b * f.nblk + blockIdx.x];
  __syncthreads();
  const uint64_t beg = (uint64_t)blockIdx.x * f.chunk;
  const uint64_t end = beg + f.chunk < n ? beg + f.chunk : n;
  // four hits in flight per thread: with few blocks (long runs per bin, see sort_hits_segmented) the
  // kernel lives on memory-level parallelism
  for (uint64_t i0 = beg + threadIdx.x; i0 < end; i0 += 4 * kSegThreads) {
    hs_hit h[4];
    double d[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t i = i0 + (uint64_t)u * kSegThreads;
      if (i < end) {
        h[u] = seg_load_key(hits + i);
        d[u] = __ldg(&hits[i].dist2);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t i = i0 + (uint64_t)u * kSegThreads;
      if (i >= end) continue;
      uint32_t bin, low;
      if (!seg_key(h[u], f, bin, low)) continue;  // (flagged by seg_hist_kernel: the result is discarded)
      const uint32_t pos = atomicAdd(&seg_cur[bin], 1u);
      pkey[pos] = low;
      pdist[pos] = d[u];
    }
  }
}

// Lanes of the warp holding the same digit d (0..256; 256 = no key), from 9 ballots (radix_sort.cu).
__device__ __forceinline__ uint32_t seg_digit_peers(uint32_t d) {
  uint32_t peers = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < 9; ++b) {
    const bool bit = (d >> b) & 1u;
    const uint32_t v = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? v : ~v;
  }
  return peers;
}

template <int NT>
struct SegSortShared {
  uint32_t whist[NT / 32][257];  // per-warp digit counters of a step (+ tail bin), then run starts
  uint32_t cnt[256];               // digit histogram of a pass / top-8-bit histogram of a large bin
  uint32_t gbase[256];             // next free output position of every digit
  uint32_t wsum[8];
  uint32_t wsum32[32];             // bucket path: per-warp sums of the bucket scan
  uint32_t r_lo[256], r_hi[256];   // ranges of a large bin
  uint32_t nranges, count, bin, maxc;
};

// One stable LSD pass over n keys in shared memory: in -> out by the digit (key >> shift) & mask.
template <int NT>
__device__ __forceinline__ void seg_radix_pass(SegSortShared<NT> &sh, const uint32_t *in, uint32_t *out, uint32_t n, int shift,
                                               uint32_t mask) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  if (tid < 256) sh.cnt[tid] = 0u;
  __syncthreads();
  for (uint32_t i = tid; i < n; i += NT) atomicAdd(&sh.cnt[(in[i] >> shift) & mask], 1u);
  __syncthreads();
  // exclusive scan of the 256 digit counts (threads 0..255)
  uint32_t c = 0, incl = 0;
  if (tid < 256) {
    c = sh.cnt[tid];
    incl = c;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, s);
      if (lane >= s) incl += y;
    }
    if (lane == 31) sh.wsum[wid] = incl;
  }
  __syncthreads();
  if (tid < 256) {
    uint32_t off = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w)
      if (w < wid) off += sh.wsum[w];
    sh.gbase[tid] = off + incl - c;
  }
  // steps of (NT * kSegItems) keys: warp w ranks the keys [w * 128, w * 128 + 128) of the step, item-major
  for (uint32_t tile = 0; tile < n; tile += (NT * kSegItems)) {
    for (int i = tid; i < (NT / 32) * 257; i += NT) (&sh.whist[0][0])[i] = 0u;
    __syncthreads();  // (also orders the gbase writes above / of the previous step)
    uint32_t key[kSegItems], lp[kSegItems], dg[kSegItems];
#pragma unroll
    for (int j = 0; j < kSegItems; ++j) {
      const uint32_t idx = tile + (uint32_t)wid * (32 * kSegItems) + (uint32_t)j * 32 + lane;
      key[j] = idx < n ? in[idx] : 0u;
      dg[j] = idx < n ? ((key[j] >> shift) & mask) : 256u;
    }
#pragma unroll
    for (int j = 0; j < kSegItems; ++j) {
      const uint32_t peers = seg_digit_peers(dg[j]);
      const uint32_t pre = sh.whist[wid][dg[j]];
      __syncwarp();
      lp[j] = pre + __popc(peers & lt_mask);
      if ((peers & lt_mask) == 0u) sh.whist[wid][dg[j]] = pre + __popc(peers);
      __syncwarp();
    }
    __syncthreads();
    if (tid < 256) {
      uint32_t run = sh.gbase[tid];
#pragma unroll
      for (int w = 0; w < (NT / 32); ++w) {
        const uint32_t cw = sh.whist[w][tid];
        sh.whist[w][tid] = run;
        run += cw;
      }
      sh.gbase[tid] = run;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kSegItems; ++j)
      if (dg[j] < 256u) out[sh.whist[wid][dg[j]] + lp[j]] = key[j];
    __syncthreads();
  }
  __syncthreads();
}

struct SegOut {
  // exactly one of the two formats
  hs_hit *hits;            // plain: hs_hit records
  uint32_t *idt;           // compact: local id | table << id_bits
  double *dist2;           //          and the distance
  uint64_t id_base;
  int id_bits;
};

template <int NT>
__global__ void __launch_bounds__(NT, 1)
seg_sort_kernel(const uint32_t *__restrict__ pkey, const double *__restrict__ pdist, uint64_t n, SegFields f,
                const uint32_t *__restrict__ start /*[nbins][nblk], scanned*/, SegOut o, uint32_t bufcap, uint32_t bufmax,
                uint32_t force_radix, unsigned int *__restrict__ bin_counter, unsigned int *__restrict__ flags) {
  extern __shared__ __align__(16) uint32_t seg_buf[];
  __shared__ SegSortShared<NT> sh;
  uint32_t *buf0 = seg_buf, *buf1 = seg_buf + bufmax;   // radix path: bufmax keys each (bufcap <= bufmax)
  uint2 *pair = reinterpret_cast<uint2 *>(seg_buf);     // bucket path: bufmax (key, position in the bin) pairs, the same bytes
  static_assert((NT / 32) * 257 >= 2 * kSegBuckets && kSegBuckets % NT == 0, "bucket tables live in the per-warp counters");
  uint32_t *bstart = &sh.whist[0][0], *bcur = bstart + kSegBuckets;
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const int hb = f.rb > 8 ? f.rb - 8 : 0;   // a large bin is cut by the top 8 bits of its low keys
  const int npass = (f.rb + 7) / 8;
  const uint64_t idmask = (1ull << f.tshift) - 1ull;
  const uint32_t tmask = (1u << (f.qshift - f.tshift)) - 1u;
  for (;;) {
    __syncthreads();
    if (tid == 0) sh.bin = atomicAdd(bin_counter, 1u);
    __syncthreads();
    const uint32_t bin = sh.bin;
    if (bin >= f.nbins) break;
    const uint64_t beg = start[(uint64_t)bin * f.nblk];
    const uint64_t end = bin + 1 < f.nbins ? (uint64_t)start[(uint64_t)(bin + 1) * f.nblk] : n;
    if (end <= beg) continue;
    const uint32_t s = (uint32_t)(end - beg);
    const uint32_t *seg = pkey + beg;
    const bool single = s <= bufcap;
    if (single) {
      if (tid == 0) {
        sh.nranges = 1u;
        sh.r_lo[0] = 0u;
        sh.r_hi[0] = 256u;
      }
    } else {
      if (tid < 256) sh.cnt[tid] = 0u;
      __syncthreads();
      for (uint32_t j = tid; j < s; j += NT) atomicAdd(&sh.cnt[seg[j] >> hb], 1u);
      __syncthreads();
      if (tid == 0) {
        uint32_t nr = 0, lo = 0, acc = 0;
        bool fail = false;
        for (uint32_t b = 0; b < 256u; ++b) {
          const uint32_t cb = sh.cnt[b];
          if (cb > bufcap) {
            fail = true;
            break;
          }
          if (acc + cb > bufcap) {
            sh.r_lo[nr] = lo;
            sh.r_hi[nr] = b;
            ++nr;
            lo = b;
            acc = 0;
          }
          acc += cb;
        }
        sh.r_lo[nr] = lo;
        sh.r_hi[nr] = 256u;
        ++nr;
        if (fail) {
          flags[1] = 1u;  // one value of the top 8 bits alone exceeds the buffer: the caller falls back
          nr = 0;
        }
        sh.nranges = nr;
      }
    }
    __syncthreads();
    const uint32_t nranges = sh.nranges;
    uint64_t outpos = beg;
    for (uint32_t r = 0; r < nranges; ++r) {
      const uint32_t lo = sh.r_lo[r], hi = sh.r_hi[r];
      // ---- bucket path: the keys of the range fall into kSegBuckets buckets by their leading bits; a
      // histogram, a scan and one scatter of (key, position) pairs put every bucket in place, each
      // bucket (a handful of keys when the keys are spread evenly) is insertion-sorted by one thread.
      // No ranking, no passes; the radix passes below remain for ranges whose keys pile up.
      const uint32_t kbase = single ? 0u : (lo << hb);
      const uint64_t span = single ? (1ull << f.rb) : ((uint64_t)(hi - lo) << hb);   // keys of the range: [kbase, kbase + span)
      int sbits = 0;
      while (sbits < 33 && ((span - 1ull) >> sbits)) ++sbits;
      const int bshift = sbits > kSegBucketBits ? sbits - kSegBucketBits : 0;
      for (int i = tid; i < 2 * kSegBuckets; i += NT) bstart[i] = 0u;
      if (tid == 0) sh.maxc = 0u;
      __syncthreads();
#pragma unroll 4
      for (uint32_t j = tid; j < s; j += NT) {
        const uint32_t k = seg[j];
        if (!single) {
          const uint32_t b8 = k >> hb;
          if (b8 < lo || b8 >= hi) continue;
        }
        atomicAdd(&bstart[(k - kbase) >> bshift], 1u);
      }
      __syncthreads();
      uint32_t nn;
      {
        // exclusive scan of the bucket counts (kSegBuckets / NT consecutive buckets per thread) and their maximum
        constexpr int PER = kSegBuckets / NT;
        uint32_t c[PER], sum = 0, mx = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
          c[i] = bstart[tid * PER + i];
          sum += c[i];
          mx = c[i] > mx ? c[i] : mx;
        }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += y;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          const uint32_t y = __shfl_xor_sync(0xffffffffu, mx, d);
          mx = y > mx ? y : mx;
        }
        if (lane == 31) sh.wsum32[tid >> 5] = incl;
        if (lane == 0 && mx) atomicMax(&sh.maxc, mx);
        __syncthreads();
        uint32_t woff = 0, total = 0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) {
          const uint32_t t = sh.wsum32[w];
          if (w < (tid >> 5)) woff += t;
          total += t;
        }
        uint32_t run = woff + incl - sum;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
          bstart[tid * PER + i] = run;
          bcur[tid * PER + i] = run;
          run += c[i];
        }
        nn = total;
        __syncthreads();
      }
      if (nn == 0u) continue;  // (block-uniform)
      if (sh.maxc <= kSegBucketMax && !force_radix) {
#pragma unroll 4
        for (uint32_t j = tid; j < s; j += NT) {
          const uint32_t k = seg[j];
          if (!single) {
            const uint32_t b8 = k >> hb;
            if (b8 < lo || b8 >= hi) continue;
          }
          const uint32_t pos = atomicAdd(&bcur[(k - kbase) >> bshift], 1u);
          pair[pos] = make_uint2(k, j);
        }
        __syncthreads();
        for (int bk = tid; bk < kSegBuckets; bk += NT) {
          const uint32_t b0 = bstart[bk], b1 = bcur[bk];
          for (uint32_t i = b0 + 1; i < b1; ++i) {
            const uint2 x = pair[i];
            uint32_t q = i;
            while (q > b0 && pair[q - 1].x > x.x) {
              pair[q] = pair[q - 1];
              --q;
            }
            pair[q] = x;
          }
        }
        __syncthreads();
#pragma unroll 4
        for (uint32_t i = tid; i < nn; i += NT) {
          const uint2 e = pair[i];
          const uint64_t full = ((uint64_t)bin << f.shift) | (uint64_t)e.x;
          const uint32_t query = (uint32_t)(full >> f.qshift);
          const uint32_t table = (uint32_t)(full >> f.tshift) & tmask;
          const uint64_t dbid = full & idmask;
          const double d = pdist[beg + e.y];
          if (o.hits) {
            hs_hit *dst_hit = o.hits + outpos + i;   // (24-byte records: 8-byte stores)
            *reinterpret_cast<uint2 *>(dst_hit) = make_uint2(query, table);
            dst_hit->db_id = dbid;
            dst_hit->dist2 = d;
          } else {
            o.idt[outpos + i] = (uint32_t)(dbid - o.id_base) | (table << o.id_bits);
            o.dist2[outpos + i] = d;
          }
        }
        outpos += nn;
        __syncthreads();
        continue;
      }
      // ---- radix path (keys piled up in a few buckets): the keys of the range into buf0, LSD passes, binary search
      if (single) {
        for (uint32_t j = tid; j < s; j += NT) buf0[j] = seg[j];
        __syncthreads();
      } else {
        if (tid == 0) sh.count = 0u;
        __syncthreads();
        // the keys of the range, in any order (warp-aggregated append)
        for (uint32_t j0 = (uint32_t)(tid & ~31); j0 < s; j0 += NT) {
          const uint32_t j = j0 + lane;
          uint32_t k = 0;
          bool in = false;
          if (j < s) {
            k = seg[j];
            const uint32_t b = k >> hb;
            in = b >= lo && b < hi;
          }
          const uint32_t m = __ballot_sync(0xffffffffu, in);
          if (m) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&sh.count, (uint32_t)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (in) buf0[base + __popc(m & lt_mask)] = k;
          }
        }
        __syncthreads();
      }
      uint32_t *src = buf0, *dst = buf1;
      for (int p = 0; p < npass; ++p) {
        const int bits = f.rb - 8 * p < 8 ? f.rb - 8 * p : 8;
        seg_radix_pass<NT>(sh, src, dst, nn, 8 * p, (1u << bits) - 1u);
        uint32_t *t = src;
        src = dst;
        dst = t;
      }
      const uint32_t *sorted = src;
      // the keys in their final format, in order
      for (uint32_t i = tid; i < nn; i += NT) {
        const uint64_t full = ((uint64_t)bin << f.shift) | (uint64_t)sorted[i];
        const uint32_t query = (uint32_t)(full >> f.qshift);
        const uint32_t table = (uint32_t)(full >> f.tshift) & tmask;
        const uint64_t dbid = full & idmask;
        if (o.hits) {
          hs_hit *dst_hit = o.hits + outpos + i;   // (24-byte records: 8-byte stores)
          *reinterpret_cast<uint2 *>(dst_hit) = make_uint2(query, table);
          dst_hit->db_id = dbid;
        } else {
          o.idt[outpos + i] = (uint32_t)(dbid - o.id_base) | (table << o.id_bits);
        }
      }
      // every hit's distance to the rank of its key
      for (uint32_t j = tid; j < s; j += NT) {
        const uint32_t k = seg[j];
        if (!single) {
          const uint32_t b = k >> hb;
          if (b < lo || b >= hi) continue;
        }
        uint32_t a = 0, z = nn;   // lower bound of k in sorted[0, nn)
        while (a < z) {
          const uint32_t mid = (a + z) >> 1;
          if (sorted[mid] < k) a = mid + 1;
          else z = mid;
        }
        const double d = pdist[beg + j];
        if (o.hits) o.hits[outpos + a].dist2 = d;
        else o.dist2[outpos + a] = d;
      }
      outpos += nn;
      __syncthreads();
    }
  }
}

// offsets[q] = base + first entry (in the sorted order) whose query is >= q, q in [qa, qb]
__global__ void seg_query_offsets_kernel(const uint32_t *__restrict__ start, uint64_t n, SegFields f, uint32_t qa, uint32_t qb,
                                   uint64_t base, uint64_t *__restrict__ offsets) {
  const uint64_t q = (uint64_t)qa + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q > qb) return;
  const uint64_t bin = q << (f.qshift - f.shift);
  offsets[q] = base + (bin < (uint64_t)f.nbins ? (uint64_t)start[bin * f.nblk] : n);
}

static int seg_bits_for(uint64_t nvalues) {  // bits that hold 0 .. nvalues-1
  int b = 1;
  while (b < 64 && (nvalues - 1) >> b) ++b;
  return b;
}

int sort_hits_segmented(hs_ctx *ctx, const hs_hit *d_hits, uint64_t n, const SegSortRequest &rq, bool *used) {
  *used = false;
  if (!ctx->segsort || n == 0 || n < ctx->segsort_min || n >= (1ull << 32)) return HS_OK;
  SegFields f;
  const int tbits = seg_bits_for((uint64_t)ctx->prm.L + 1);
  const int ibits = seg_bits_for(std::max<uint64_t>(ctx->hit_idmax ? ctx->hit_idmax : ctx->id_base + ctx->N, 2));
  f.qbits = seg_bits_for(std::max<uint64_t>(ctx->hit_qmax, 1));
  f.tshift = ibits;
  f.qshift = ibits + tbits;
  const int kbits = f.qshift + f.qbits;
  if (kbits > 64 || f.qbits > kSegMaxBinBits) return HS_OK;
  const int pb = std::min(kbits, kSegMaxBinBits);
  f.shift = kbits - pb;
  f.rb = f.shift;
  if (f.rb > 32) return HS_OK;
  const uint64_t qmax = std::max<uint64_t>(ctx->hit_qmax, 1);
  f.nbins = (uint32_t)((((qmax << f.qshift) - 1ull) >> f.shift) + 1ull);
  if (f.nbins > (1u << kSegMaxBinBits)) return HS_OK;
  const uint64_t want_blk = (n + 8191) / 8192;
  // Blocks of the partition: every (bin, block) pair is one run of the partitioned arrays, written a hit at
  // a time -- the fewer blocks, the longer the runs and the fewer cache lines are open at once
  // (HS_SEGSORT_NBLK, default two per SM)
  const uint64_t max_blk = ctx->segsort_nblk ? ctx->segsort_nblk : 2ull * (uint64_t)ctx->num_sms;
  f.nblk = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(want_blk, max_blk));
  f.chunk = (n + f.nblk - 1) / f.nblk;
  const uint64_t ntab = (uint64_t)f.nbins * f.nblk;

  const size_t part_smem = sizeof(uint32_t) * f.nbins;
  // the per-bin sort: one 1024-thread block per SM, shared memory for kSegBufMax (key, position) pairs
  // (two blocks of 512 threads with half the buffer each were measured: twice as slow)
  const uint32_t bufmax = kSegBufMax;
  const size_t sort_smem = sizeof(uint32_t) * 2 * bufmax;
  // (per device: a process may drive several GPUs, each with its own context)
  HS_CUDA(cudaFuncSetAttribute(seg_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(uint32_t) << kSegMaxBinBits)));
  HS_CUDA(cudaFuncSetAttribute(seg_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(uint32_t) << kSegMaxBinBits)));
  HS_CUDA(cudaFuncSetAttribute(seg_sort_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem));
  HS_TRY(ctx->d_seg_tab.reserve(sizeof(uint32_t) * (ntab + 1)));
  HS_TRY(ctx->d_seg_key.reserve(sizeof(uint32_t) * n));
  HS_TRY(ctx->d_seg_dist.reserve(sizeof(double) * n));
  HS_TRY(ctx->d_seg_ctl.reserve(sizeof(unsigned int) * 4));
  uint32_t *tab = ctx->d_seg_tab.as<uint32_t>();
  uint32_t *pkey = ctx->d_seg_key.as<uint32_t>();
  double *pdist = ctx->d_seg_dist.as<double>();
  unsigned int *ctl = ctx->d_seg_ctl.as<unsigned int>();  // [0] field overflow, [1] range overflow, [2] bin counter
  HS_CUDA(cudaMemsetAsync(ctl, 0, sizeof(unsigned int) * 4, ctx->stream));
  cudaEvent_t pe[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // HS_SEGSORT_PROF: per-kernel times to stderr
  if (ctx->segsort_prof)
    for (cudaEvent_t &e : pe) HS_CUDA(cudaEventCreate(&e));
  if (ctx->segsort_prof) HS_CUDA(cudaEventRecord(pe[0], ctx->stream));
  seg_hist_kernel<<<f.nblk, kSegThreads, part_smem, ctx->stream>>>(d_hits, n, f, tab, ctl);
  HS_CUDA(cudaGetLastError());
  if (ctx->segsort_prof) HS_CUDA(cudaEventRecord(pe[1], ctx->stream));
  HS_TRY(exclusive_scan_u32(ctx, tab, tab, ntab, nullptr));
  if (ctx->segsort_prof) HS_CUDA(cudaEventRecord(pe[2], ctx->stream));
  seg_scatter_kernel<<<f.nblk, kSegThreads, part_smem, ctx->stream>>>(d_hits, n, f, tab, pkey, pdist);
  if (ctx->segsort_prof) HS_CUDA(cudaEventRecord(pe[3], ctx->stream));
  SegOut o;
  o.hits = rq.compact ? nullptr : rq.hits_out;
  o.idt = rq.idt;
  o.dist2 = rq.dist2;
  o.id_base = ctx->id_base;
  o.id_bits = rq.id_bits;
  const uint32_t bufcap = std::min<uint32_t>(bufmax, ctx->segsort_buf);  // (0, a test hook: no bin fits, every list is handed back)
  const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)f.nbins, (uint64_t)ctx->num_sms);
  seg_sort_kernel<1024><<<grid, 1024, sort_smem, ctx->stream>>>(pkey, pdist, n, f, tab, o, bufcap, bufmax, ctx->segsort_radix ? 1u : 0u,
                                                                  ctl + 2, ctl);
  ctx->stats.kernel_launches += 3;
  if (rq.compact) {
    const uint32_t nq = rq.qb - rq.qa + 1;
    seg_query_offsets_kernel<<<(nq + 255) / 256, 256, 0, ctx->stream>>>(tab, n, f, rq.qa, rq.qb, rq.base, rq.offsets);
    ctx->stats.kernel_launches++;
  }
  HS_CUDA(cudaGetLastError());
  if (ctx->segsort_prof) HS_CUDA(cudaEventRecord(pe[4], ctx->stream));
  unsigned int h_flags[2] = {0, 0};
  HS_TRY(read_back(ctx, ctl, h_flags, sizeof h_flags));
  if (ctx->segsort_prof) {
    float ms[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&ms[i], pe[i], pe[i + 1]);
    fprintf(stderr, "segsort: n=%llu bins=%u blocks=%u radix=%d rb=%d  hist %.3f  scan %.3f  scatter %.3f  sort %.3f ms  flags %u %u\n",
            (unsigned long long)n, f.nbins, f.nblk, ctx->segsort_radix ? 1 : 0, f.rb, ms[0], ms[1], ms[2], ms[3], h_flags[0], h_flags[1]);
    for (cudaEvent_t e : pe) cudaEventDestroy(e);
  }
  if (h_flags[0] || h_flags[1]) {
    ctx->stats.segsort_fallbacks++;
    return HS_OK;  // the caller sorts with the radix path (d_hits is untouched)
  }
  ctx->stats.segsort_lists++;
  *used = true;
  return HS_OK;
}

}  // namespace hs
