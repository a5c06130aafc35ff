// probe_batch.cu -- the throughput path for ONE join: a persistent kernel probes a
// whole key column and writes dense (compacted) result columns.
//
// It is the fused GPU form of the micro-bench loop (simd_micro_bench.cpp:83-116:
// Probe + while(HasNext) Next over every chunk) followed by a Compactor on every sparse
// result chunk (compactor.cpp:5-41):
//   * a CTA iteration owns one chunk (tile) of kPbTile probe rows, kPbKeysPerThread per
//     thread, loaded with coalesced strided loads so that all table accesses of a thread
//     are independent and in flight together (memory-level parallelism for the gather);
//     the NEXT tile's keys are requested into a second register set before this tile's
//     gathers are consumed;
//   * tiles are handed out by a global atomic counter, two iterations ahead.  This keeps
//     all CTAs inside one narrow, moving window of the key column -- essential behind the
//     partition pass: with static round-robin tiles the persistent CTAs drift apart by
//     hundreds of tiles, the window spreads over many table slices and L2 thrashes
//     (measured: 52 GB instead of 13 GB of DRAM reads for 2^29 keys, profiles/);
//   * tables without duplicate keys: every thread walks its own probe sequences / chains
//     to the first match (the walk past a match only ever finds duplicates,
//     linear_probing_ht.cpp:101-109), then the CTA compacts once;
//   * tables with duplicates: a "round" is one Next() -- every active lane compares its
//     current slot / chain entry (ScanInnerJoin), matches are compacted, lanes advance
//     (AdvancePointers) and finished lanes retire;
//   * compaction step: matches are ranked with warp ballots + a 32-entry scan of the
//     (key-slice, warp) counts, the CTA reserves a contiguous range of the global output
//     with ONE atomicAdd, and lanes store key / payload at consecutive positions.
// Output row order is unspecified; the multiset equals the reference's.
//
// Measured dead ends (B200, kept out of the code, evidence under profiles/):
//   - streaming the keys through a TMA (cp.async.bulk + mbarrier) shared-memory ring: any
//     sizeable shared-memory carve-out shrinks L1 and HALVES the gather rate of this chip
//     (L2-resident 16 MiB gather: 290 G/s with <= 16 KiB smem per CTA, 138 G/s with 40 KiB);
//   - L2::128B prefetch on the table loads: every miss then costs 4 sectors, no gain;
//   - cudaLimitMaxL2FetchGranularity = 32: accepted, no effect on the 37 G/s big-table ceiling.
#include <cstddef>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "partition.cuh"

namespace ccb {

constexpr int kPbThreads = 256;
constexpr int kPbKeysPerThread = 4;
constexpr int kPbWarps = kPbThreads / 32;
constexpr int kPbTile = kPbThreads * kPbKeysPerThread;
static_assert(kPbKeysPerThread == 4, "the interleaved walk tests act[0..3]");
static_assert(kPbWarps * kPbKeysPerThread == 32, "rank scan assumes 32 (slice, warp) counters");

struct ProbeArgs {
  const uint64_t *slots;
  const uint2 *dir;
  const int64_t *ckeys;
  uint64_t mask;
  const int64_t *keys;
  size_t n;
  int64_t *out_key;
  int64_t *out_payload;
  uint64_t *out_rowid;
  size_t cap;
  cc_probe_result *res;
  unsigned long long *tile_counter;  // zero-initialised per launch
  // segmented key column (behind the single-pass partition): partition p owns rows [p * seg_cap, p * seg_cap + seg_cursors[p]);
  // seg_prefix[0 .. seg_parts] = exclusive prefix of the per-partition tile counts.  seg_parts == 0: plain dense column.
  const uint32_t *seg_prefix;
  const unsigned long long *seg_cursors;
  unsigned long long seg_cap;
  int seg_parts;
  int seg_inner = 0, seg_outer_stride = 0;  // walk order != memory order (SegIn::region, partition.cuh); 0 = identity
  const int *gate;  // optional device-side switch: run only if (*gate != 0) == gate_want
  int gate_want;
  // payload columns (PAY kernels only): pay[c][i] belongs to the table entry at index i (LP slot / chain position)
  int n_pay = 0;
  const int64_t *pay[CC_MAX_PAYLOAD_COLS] = {nullptr, nullptr, nullptr, nullptr};
  int64_t *out_pay[CC_MAX_PAYLOAD_COLS] = {nullptr, nullptr, nullptr, nullptr};  // result columns, any may be NULL
  unsigned long long *col_sum = nullptr;  // device array [CC_MAX_PAYLOAD_COLS]: wrapping sums of the payloads of all result rows
};
static_assert(CC_MAX_PAYLOAD_COLS == 4, "ProbeArgs initialisers and the unrolled column loops assume 4 payload columns");

// Cache mode HINT (MODE bit 1): L2 eviction priorities (createpolicy + L2::cache_hint) -- the
// probe keys and the result columns are touched exactly once (evict_first, no L1 allocation),
// the table is what must stay resident (evict_last).
struct CachePolicy {
  uint64_t first, last;
};
__device__ __forceinline__ CachePolicy make_policies() {
  CachePolicy p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p.first));
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p.last));
  return p;
}
template <int MODE>
__device__ __forceinline__ uint64_t ld_table_u64(const uint64_t *p, const CachePolicy &pol) {
  uint64_t v;
  if (MODE & 2)
    asm volatile("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol.last));
  else
    asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
template <int MODE>
__device__ __forceinline__ uint2 ld_table_u32x2(const uint2 *p, const CachePolicy &pol) {
  uint2 v;
  if (MODE & 2)
    asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol.last));
  else
    asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
template <int MODE>
__device__ __forceinline__ uint64_t ld_stream_u64(const int64_t *p, const CachePolicy &pol) {
  uint64_t v;
  if (MODE & 2)
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol.first));
  else
    asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
template <int MODE>
__device__ __forceinline__ void st_stream_u64(void *p, uint64_t v, const CachePolicy &pol) {
  if (MODE & 2)
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol.first) : "memory");
  else
    asm volatile("st.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// memory region of walk segment p (SegIn::region)
__device__ __forceinline__ uint32_t seg_region(const ProbeArgs &a, uint32_t p) {
  return a.seg_inner ? (p % (uint32_t) a.seg_inner) * (uint32_t) a.seg_outer_stride + p / (uint32_t) a.seg_inner : p;
}

struct ProbeShared {
  uint32_t cnt[32];
  unsigned long long base;
  unsigned long long off_a, off_b;  // first row of the tiles handed out by the global counter
  uint32_t rows_a, rows_b;          // their row counts (0 = no such tile)
};

// thread 0: translate a global tile index into (first row, row count); false = past the last tile.
// Segmented mode: every partition region holds a whole number T = seg_cap / kPbTile of tiles, so the mapping is one
// division and one shared-memory read; tiles beyond a region's fill have rows == 0 and are skipped by the caller.
__device__ __forceinline__ bool tile_lookup(const ProbeArgs &a, const ProbeShared &sh, unsigned long long g, unsigned long long &off,
                                            uint32_t &rows) {
  off = 0;
  rows = 0;
  if (a.seg_parts == 0) {
    const unsigned long long ntiles = (a.n + kPbTile - 1) / kPbTile;
    if (g >= ntiles) return false;
    off = g * (unsigned long long) kPbTile;
    rows = (uint32_t) (a.n - off < (unsigned long long) kPbTile ? a.n - off : (unsigned long long) kPbTile);
    return true;
  }
  const uint32_t T = (uint32_t) (a.seg_cap / kPbTile);
  if (g >= (unsigned long long) T * (unsigned long long) a.seg_parts) return false;
  const uint32_t p = (uint32_t) g / T, local = (uint32_t) g - p * T;
  const unsigned long long first = (unsigned long long) local * kPbTile;
  // region fill count through the read-only L1 path (the same few regions are asked for again and again); a shared-memory
  // copy would grow the carve-out, which costs this chip gather throughput
  const uint32_t reg = seg_region(a, p);
  unsigned long long cnt = __ldg(a.seg_cursors + reg);
  if (cnt > a.seg_cap) cnt = a.seg_cap;
  off = (unsigned long long) reg * a.seg_cap + first;
  rows = cnt > first ? (uint32_t) (cnt - first < (unsigned long long) kPbTile ? cnt - first : (unsigned long long) kPbTile) : 0u;
  return true;
}

// thread 0: next non-empty tile from the global counter (rows == 0: none left).  `g` is an index already requested.
__device__ __forceinline__ void tile_fetch(const ProbeArgs &a, const ProbeShared &sh, unsigned long long g, unsigned long long &off,
                                           uint32_t &rows) {
  while (tile_lookup(a, sh, g, off, rows) && rows == 0) g = atomicAdd(a.tile_counter, 1ull);
}

// Compaction step: rank the matches of the CTA, reserve a contiguous range of the global output
// with ONE atomicAdd and store key / payload / row id at consecutive positions.
template <int MODE>
__device__ __forceinline__ void emit_matches(const ProbeArgs &a, ProbeShared &sh, const CachePolicy &pol, const bool (&m)[kPbKeysPerThread],
                                             const uint64_t (&k)[kPbKeysPerThread], const uint64_t (&v)[kPbKeysPerThread],
                                             size_t tbase, uint64_t &ksum, uint64_t &psum) {
  const unsigned w = threadIdx.x >> 5;
  const unsigned lt = lanemask_lt();
  unsigned bal[kPbKeysPerThread];
#pragma unroll
  for (int j = 0; j < kPbKeysPerThread; ++j) {
    bal[j] = __ballot_sync(0xffffffffu, m[j]);
    if (lane_id() == 0) sh.cnt[j * kPbWarps + w] = __popc(bal[j]);
  }
  __syncthreads();
  if (w == 0) {
    uint32_t c = sh.cnt[lane_id()];
    uint32_t incl = warp_incl_scan_u32(c);
    sh.cnt[lane_id()] = incl - c;
    if (lane_id() == 31) sh.base = incl ? atomicAdd((unsigned long long *) &a.res->n_matches, (unsigned long long) incl) : 0ull;
  }
  __syncthreads();
  const uint64_t base = sh.base;
  const bool fits = base + kPbTile <= a.cap;  // CTA-uniform: the whole tile fits, no per-row capacity check
#pragma unroll
  for (int j = 0; j < kPbKeysPerThread; ++j) {
    if (m[j]) {
      uint64_t dst = base + sh.cnt[j * kPbWarps + w] + __popc(bal[j] & lt);
      ksum += k[j];
      psum += v[j];
      if (fits || dst < a.cap) {
        if (a.out_key) st_stream_u64<MODE>(a.out_key + dst, k[j], pol);
        if (a.out_payload) st_stream_u64<MODE>(a.out_payload + dst, v[j], pol);
        if (a.out_rowid) st_stream_u64<MODE>(a.out_rowid + dst, tbase + (size_t) j * kPbThreads + threadIdx.x, pol);
      }
    }
  }
}

// The same compaction step for a table with payload columns: a match at table index idx[j] also emits pay[c][idx[j]].
// Column by column, the (up to four) gathers of a thread are issued together before their stores.
template <int MODE>
__device__ __forceinline__ void emit_matches_pay(const ProbeArgs &a, ProbeShared &sh, const CachePolicy &pol, const bool (&m)[kPbKeysPerThread],
                                                 const uint64_t (&k)[kPbKeysPerThread], const uint64_t (&v)[kPbKeysPerThread],
                                                 const uint64_t (&idx)[kPbKeysPerThread], size_t tbase, uint64_t &ksum, uint64_t &psum,
                                                 uint64_t (&csum)[CC_MAX_PAYLOAD_COLS]) {
  const unsigned w = threadIdx.x >> 5;
  const unsigned lt = lanemask_lt();
  unsigned bal[kPbKeysPerThread];
#pragma unroll
  for (int j = 0; j < kPbKeysPerThread; ++j) {
    bal[j] = __ballot_sync(0xffffffffu, m[j]);
    if (lane_id() == 0) sh.cnt[j * kPbWarps + w] = __popc(bal[j]);
  }
  __syncthreads();
  if (w == 0) {
    uint32_t c = sh.cnt[lane_id()];
    uint32_t incl = warp_incl_scan_u32(c);
    sh.cnt[lane_id()] = incl - c;
    if (lane_id() == 31) sh.base = incl ? atomicAdd((unsigned long long *) &a.res->n_matches, (unsigned long long) incl) : 0ull;
  }
  __syncthreads();
  const uint64_t base = sh.base;
  uint64_t dst[kPbKeysPerThread];
  bool wr[kPbKeysPerThread];
#pragma unroll
  for (int j = 0; j < kPbKeysPerThread; ++j) {
    dst[j] = base + sh.cnt[j * kPbWarps + w] + __popc(bal[j] & lt);
    wr[j] = m[j] && dst[j] < a.cap;
    if (m[j]) {
      ksum += k[j];
      psum += v[j];
    }
    if (wr[j]) {
      if (a.out_key) st_stream_u64<MODE>(a.out_key + dst[j], k[j], pol);
      if (a.out_payload) st_stream_u64<MODE>(a.out_payload + dst[j], v[j], pol);
      if (a.out_rowid) st_stream_u64<MODE>(a.out_rowid + dst[j], tbase + (size_t) j * kPbThreads + threadIdx.x, pol);
    }
  }
#pragma unroll
  for (int c = 0; c < CC_MAX_PAYLOAD_COLS; ++c) {
    if (c < a.n_pay) {
      uint64_t t[kPbKeysPerThread];
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j) t[j] = m[j] ? ld_table_u64<MODE>((const uint64_t *) a.pay[c] + idx[j], pol) : 0ull;
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j) {
        csum[c] += t[j];
        if (wr[j] && a.out_pay[c]) st_stream_u64<MODE>(a.out_pay[c] + dst[j], t[j], pol);
      }
    }
  }
}

template <int MODE>
__device__ __forceinline__ void load_tile_keys(const ProbeArgs &a, const CachePolicy &pol, unsigned long long off, uint32_t rows,
                                               uint64_t (&kn)[kPbKeysPerThread]) {
  const int64_t *p = a.keys + off + threadIdx.x;
#pragma unroll
  for (int j = 0; j < kPbKeysPerThread; ++j)
    kn[j] = (uint32_t) (j * kPbThreads) + threadIdx.x < rows ? ld_stream_u64<MODE>(p + j * kPbThreads, pol) : 0;
}

// W32: the table has at most 2^32 slots / buckets, so slot arithmetic runs in 32 bits
// PAY: the table carries payload columns -- a match also gathers pay[c][table index of the match]
template <int KIND, bool UNIQUE, int MODE, bool W32, bool PAY = false>
__global__ void __launch_bounds__(kPbThreads, PAY ? 3 : 4) probe_batch_kernel(ProbeArgs a) {
  __shared__ ProbeShared sh;
  const CachePolicy pol = make_policies();
  uint64_t ksum = 0, psum = 0;
  uint64_t csum[CC_MAX_PAYLOAD_COLS] = {0, 0, 0, 0};
  if (a.gate && ((*a.gate != 0) != (a.gate_want != 0))) return;  // device-side strategy switch (CTA-uniform)
  // dynamic tile scheduling: one tile is being processed, the keys of the next are in flight, and thread 0
  // fetches the one after that while the CTA works (published through the emit barriers)
  if (threadIdx.x == 0) {
    tile_fetch(a, sh, atomicAdd(a.tile_counter, 1ull), sh.off_a, sh.rows_a);
    tile_fetch(a, sh, atomicAdd(a.tile_counter, 1ull), sh.off_b, sh.rows_b);
  }
  __syncthreads();
  unsigned long long off = sh.off_a, noff = sh.off_b;
  uint32_t rows = sh.rows_a, nrows = sh.rows_b;
  __syncthreads();
  uint64_t kn[kPbKeysPerThread];
  load_tile_keys<MODE>(a, pol, off, rows, kn);
  while (rows > 0) {
    const size_t tbase = (size_t) off;
    // thread 0 requests the tile after next NOW and translates / publishes it just before the emit barriers, so the
    // atomic's L2 round trip is off warp 0's path to the barrier
    unsigned long long g_after = 0;
    if (threadIdx.x == 0) g_after = atomicAdd(a.tile_counter, 1ull);
    uint64_t k[kPbKeysPerThread], v[kPbKeysPerThread], pos[kPbKeysPerThread];
    uint32_t end[kPbKeysPerThread];
    bool act[kPbKeysPerThread];
    // ---- Probe (chaining_ht.cpp:44-55 / linear_probing_ht.cpp:45-57): all table loads of a thread in flight together
#pragma unroll
    for (int j = 0; j < kPbKeysPerThread; ++j) {
      act[j] = (uint32_t) (j * kPbThreads) + threadIdx.x < rows;
      k[j] = kn[j];
      pos[j] = W32 ? (uint64_t) ((uint32_t) murmurhash64(k[j]) & (uint32_t) a.mask) : (murmurhash64(k[j]) & a.mask);
    }
    if (KIND == CC_HT_LP) {
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j) v[j] = act[j] ? ld_table_u64<MODE>(a.slots + (W32 ? (uint32_t) pos[j] : pos[j]), pol) : kEmptyU;
      load_tile_keys<MODE>(a, pol, noff, nrows, kn);  // next tile's keys, behind this tile's gathers
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j) act[j] = v[j] != kEmptyU;
    } else {
      uint2 d[kPbKeysPerThread];
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j) d[j] = act[j] ? ld_table_u32x2<MODE>(a.dir + pos[j], pol) : make_uint2(0, 0);
      load_tile_keys<MODE>(a, pol, noff, nrows, kn);
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j) {
        pos[j] = d[j].x;
        end[j] = d[j].x + d[j].y;
        act[j] = d[j].y != 0;
        v[j] = act[j] ? ld_table_u64<MODE>((const uint64_t *) a.ckeys + pos[j], pol) : 0;
      }
    }
    if (UNIQUE) {
      // walk all of the thread's probe sequences / chains together: one dependent load per lap for
      // every key that has neither matched nor run out, instead of finishing key 0 before key 1
      bool m[kPbKeysPerThread];
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j) {
        m[j] = act[j] && (v[j] == k[j]);
        act[j] = act[j] && !m[j];
      }
      while (act[0] | act[1] | act[2] | act[3]) {
#pragma unroll
        for (int j = 0; j < kPbKeysPerThread; ++j) {
          if (act[j]) {
            if (KIND == CC_HT_LP) {
              if (W32) {
                uint32_t p32 = ((uint32_t) pos[j] + 1u) & (uint32_t) a.mask;
                pos[j] = p32;
                v[j] = ld_table_u64<MODE>(a.slots + p32, pol);
              } else {
                pos[j] = (pos[j] + 1) & a.mask;
                v[j] = ld_table_u64<MODE>(a.slots + pos[j], pol);
              }
            } else {
              pos[j] += 1;
              act[j] = pos[j] != end[j];
              if (act[j]) v[j] = ld_table_u64<MODE>((const uint64_t *) a.ckeys + pos[j], pol);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < kPbKeysPerThread; ++j) {
          if (act[j]) {
            if (KIND == CC_HT_LP) act[j] = v[j] != kEmptyU;
            m[j] = act[j] && (v[j] == k[j]);
            act[j] = act[j] && !m[j];
          }
        }
      }
      if (threadIdx.x == 0) tile_fetch(a, sh, g_after, sh.off_a, sh.rows_a);
      if (PAY)
        emit_matches_pay<MODE>(a, sh, pol, m, k, v, pos, tbase, ksum, psum, csum);  // pos[j] = table index of key j's match
      else
        emit_matches<MODE>(a, sh, pol, m, k, v, tbase, ksum, psum);
    } else {
      // ---- tables with duplicate keys: a lane keeps walking past its matches (linear_probing_ht.cpp:101-109).
      // One round = up to kRoundEntries Next() steps of every live key at once: the thread loads the next entries of
      // a key together (one or two sectors), records WHICH of them match (the payload of a match is the key itself,
      // chaining_ht.cpp:34 stores keys only), the CTA ranks the per-thread match counts with one exclusive scan,
      // reserves the output range with one atomicAdd and every thread stores its rows at consecutive positions.
      constexpr int kRoundEntries = KIND == CC_HT_CHAIN ? 8 : 4;
      if (threadIdx.x == 0) tile_fetch(a, sh, g_after, sh.off_a, sh.rows_a);  // published by the barriers of the first round
      bool any;
      do {
        uint32_t mm[kPbKeysPerThread];
        uint64_t pos0[kPbKeysPerThread];  // PAY: table index the round started at (match bit q <-> entry pos0 + q)
        uint32_t cnt = 0;
        bool mine = false;
#pragma unroll
        for (int j = 0; j < kPbKeysPerThread; ++j) {
          mm[j] = 0;
          if (PAY) pos0[j] = pos[j];
          if (act[j]) {
            uint64_t e[kRoundEntries];
            if (KIND == CC_HT_CHAIN) {
#pragma unroll
              for (int q = 0; q < kRoundEntries; ++q)
                e[q] = (uint32_t) pos[j] + q < end[j] ? ld_table_u64<MODE>((const uint64_t *) a.ckeys + pos[j] + q, pol) : ~k[j];
#pragma unroll
              for (int q = 0; q < kRoundEntries; ++q) mm[j] |= ((uint32_t) pos[j] + q < end[j] && e[q] == k[j]) ? (1u << q) : 0u;
              pos[j] = end[j] - (uint32_t) pos[j] > (uint32_t) kRoundEntries ? pos[j] + kRoundEntries : (uint64_t) end[j];
              act[j] = pos[j] != end[j];
            } else {
              e[0] = v[j];  // the current slot was already fetched
#pragma unroll
              for (int q = 1; q < kRoundEntries; ++q) e[q] = ld_table_u64<MODE>(a.slots + ((pos[j] + q) & a.mask), pol);
              bool open = true;
#pragma unroll
              for (int q = 0; q < kRoundEntries; ++q) {
                open = open && e[q] != kEmptyU;  // the walk ends at the first empty slot
                mm[j] |= (open && e[q] == k[j]) ? (1u << q) : 0u;
              }
              pos[j] = (pos[j] + kRoundEntries) & a.mask;
              act[j] = open;
              if (open) {
                v[j] = ld_table_u64<MODE>(a.slots + pos[j], pol);
                act[j] = v[j] != kEmptyU;
              }
            }
            cnt += __popc(mm[j]);
          }
          mine |= act[j];
        }
        // rank: exclusive scan of the per-thread match counts over the CTA
        uint32_t incl = warp_incl_scan_u32(cnt);
        const unsigned w = threadIdx.x >> 5;
        if (lane_id() == 31) sh.cnt[w] = incl;
        __syncthreads();
        if (w == 0) {
          uint32_t c = lane_id() < kPbWarps ? sh.cnt[lane_id()] : 0;
          uint32_t ci = warp_incl_scan_u32(c);
          sh.cnt[lane_id()] = ci - c;
          if (lane_id() == 31) sh.base = ci ? atomicAdd((unsigned long long *) &a.res->n_matches, (unsigned long long) ci) : 0ull;
        }
        __syncthreads();
        uint64_t dst = sh.base + sh.cnt[w] + (incl - cnt);
#pragma unroll
        for (int j = 0; j < kPbKeysPerThread; ++j) {
          uint32_t bits = mm[j];
          const uint32_t c = __popc(bits);
          ksum += k[j] * c;
          psum += k[j] * c;  // payload of a match == the matched build key == the probe key
          while (bits) {
            const uint32_t q = (uint32_t) __ffs((int) bits) - 1u;
            bits &= bits - 1;
            if (dst < a.cap) {
              if (a.out_key) st_stream_u64<MODE>(a.out_key + dst, k[j], pol);
              if (a.out_payload) st_stream_u64<MODE>(a.out_payload + dst, k[j], pol);
              if (a.out_rowid) st_stream_u64<MODE>(a.out_rowid + dst, tbase + (size_t) j * kPbThreads + threadIdx.x, pol);
            }
            if (PAY) {
              const uint64_t at = KIND == CC_HT_LP ? ((pos0[j] + q) & a.mask) : pos0[j] + q;
#pragma unroll
              for (int c = 0; c < CC_MAX_PAYLOAD_COLS; ++c) {
                if (c < a.n_pay) {
                  const uint64_t t = ld_table_u64<MODE>((const uint64_t *) a.pay[c] + at, pol);
                  csum[c] += t;
                  if (dst < a.cap && a.out_pay[c]) st_stream_u64<MODE>(a.out_pay[c] + dst, t, pol);
                }
              }
            }
            ++dst;
          }
        }
        any = __syncthreads_or(mine);
      } while (any);
    }
    // sh.off_a / rows_a were written before the barriers inside the emit: visible to everyone now
    off = noff;
    rows = nrows;
    noff = sh.off_a;
    nrows = sh.rows_a;
    __syncthreads();  // protects sh.off_a / sh.cnt / sh.base against the next iteration
  }
  ksum = warp_sum_u64(ksum);
  psum = warp_sum_u64(psum);
  if (lane_id() == 0 && (ksum | psum)) {
    atomicAdd((unsigned long long *) &a.res->key_sum, (unsigned long long) ksum);
    atomicAdd((unsigned long long *) &a.res->payload_sum, (unsigned long long) psum);
  }
  if (PAY) {
#pragma unroll
    for (int c = 0; c < CC_MAX_PAYLOAD_COLS; ++c) {
      const uint64_t t = warp_sum_u64(csum[c]);
      if (lane_id() == 0 && t && c < a.n_pay) atomicAdd(a.col_sum + c, (unsigned long long) t);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Lean kernel for tables WITHOUT duplicate keys and at most 2^32 slots / buckets (the C1 / C4 / C5 shape).  Same tiles, same
// dynamic tile scheduling and the same compaction as probe_batch_kernel, restructured for instruction count -- the generic
// kernel spends ~180 SASS instructions per key (ncu: 12.2 G warp instructions per 2^31 keys, issue slots the limiter):
//   * 32-bit slot arithmetic throughout, one IMAD.WIDE per table address (forced with inline PTX: the optimiser otherwise
//     rewrites zext(trunc(hash)) * 8 into 64-bit shift/mask sequences);
//   * lane state in two bit masks (um = still walking, mm = matched), recomputed from the loaded slots after every lap --
//     a resolved key keeps its last slot value, so the recomputation is idempotent and needs no per-key branches;
//   * the walk past the home slot only loads for the keys that need it (~13 % at load factor 0.25);
//   * the payload of a match is the matched build key == the probe key, so nothing but the key is kept after the compare
//     and one checksum serves both result columns;
//   * ranks come from one counter per WARP (8 instead of 32 shared-memory words), output pointers are formed once per
//     tile, and the capacity check is one 32-bit compare per row;
//   * all shared state is double-buffered by tile parity, which removes the third barrier per tile;
//   * segmented key columns are walked with a forward-only cursor over the per-partition tile prefix (tile indices handed
//     to a CTA only grow): no division, and no empty slack tiles that would cost an extra round trip to the tile counter.
struct __align__(16) LeanShared {
  uint32_t cnt[2][kPbWarps];
  unsigned long long base[2];
  unsigned long long off[2];
  uint32_t rows[2];
  // thread 0's cursor over the segmented key column
  uint32_t seg_p, seg_lo, seg_hi, seg_total;
  unsigned long long seg_cnt;
  // LP tables: per-warp queue of the keys whose home slot holds another key (~13 % at load factor 0.25); after the
  // tail walk its first entries are the tail's matches
  uint64_t queue[kPbWarps][kPbKeysPerThread * 32];
  // per-warp parking area of the matches of one tile (at most 128 keys), flushed to the output one iteration later
  uint64_t stage[kPbWarps][kPbKeysPerThread * 32];
};

__device__ __forceinline__ const uint64_t *elem_ptr_u64(const void *base, uint32_t idx) {  // base + idx * 8 in ONE IMAD.WIDE.U32
  uint64_t r;
  asm("mad.wide.u32 %0, %1, 8, %2;" : "=l"(r) : "r"(idx), "l"(base));
  return reinterpret_cast<const uint64_t *>(r);
}
__device__ __forceinline__ uint32_t home_slot32(uint64_t key, uint32_t mask) {  // (uint32_t) murmurhash64(key) & mask
  uint64_t x = key;
  x ^= x >> 32;
  x *= 0xd6e8feb86659fd93ULL;
  x ^= x >> 32;
  x *= 0xd6e8feb86659fd93ULL;
  uint32_t lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x));
  return (lo ^ hi) & mask;
}

// one aligned 32-byte sector (4 table slots, slot index s % 4 == 0) in ONE request: the load/store unit hands one request per
// lane and cycle to the memory system whatever its width, and that request rate is what bounds this kernel
template <int MODE>
__device__ __forceinline__ void ld_table_sector(const uint64_t *slots, uint32_t s, const CachePolicy &pol, uint64_t (&x)[4]) {
  const uint64_t *pa = elem_ptr_u64(slots, s);
  if (MODE & 2)
    asm volatile("ld.global.nc.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(x[0]), "=l"(x[1]), "=l"(x[2]), "=l"(x[3]) : "l"(pa), "l"(pol.last));
  else
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(x[0]), "=l"(x[1]), "=l"(x[2]), "=l"(x[3]) : "l"(pa));
}

// thread 0: translate tile index g (dense numbering over the non-empty tiles) into (first row, row count); rows == 0: no tile left
__device__ __forceinline__ void lean_tile(const ProbeArgs &a, LeanShared &sh, unsigned long long g, unsigned long long &off, uint32_t &rows) {
  off = 0;
  rows = 0;
  if (a.seg_parts == 0) {
    const unsigned long long ntiles = (a.n + kPbTile - 1) / kPbTile;
    if (g >= ntiles) return;
    off = g * (unsigned long long) kPbTile;
    rows = (uint32_t) (a.n - off < (unsigned long long) kPbTile ? a.n - off : (unsigned long long) kPbTile);
    return;
  }
  if (g >= sh.seg_total) return;
  uint32_t p = sh.seg_p, lo = sh.seg_lo, hi = sh.seg_hi;
  if ((uint32_t) g >= hi) {
    do {
      ++p;
      lo = hi;
      hi = __ldg(a.seg_prefix + p + 1);
    } while ((uint32_t) g >= hi);
    unsigned long long c = __ldg(a.seg_cursors + seg_region(a, p));
    sh.seg_cnt = c < a.seg_cap ? c : a.seg_cap;
    sh.seg_p = p;
    sh.seg_lo = lo;
    sh.seg_hi = hi;
  }
  const unsigned long long first = (unsigned long long) ((uint32_t) g - lo) * kPbTile;
  const unsigned long long left = sh.seg_cnt - first;
  off = (unsigned long long) seg_region(a, p) * a.seg_cap + first;
  rows = left < (unsigned long long) kPbTile ? (uint32_t) left : (uint32_t) kPbTile;
}

// write the `cnt` rows a warp parked in shared memory to output rows [first, first + cnt): lane l stores rows l, l + 32, ...
template <int MODE, int OUT>
__device__ __forceinline__ void lean_flush(const ProbeArgs &a, const CachePolicy &pol, const uint64_t *stage, unsigned long long first,
                                           uint32_t cnt, unsigned outsel, unsigned lane) {
  if (OUT == 0) return;
  // rows of this warp that still fit: everything in the common case, else what is left below the capacity
  const uint32_t room = first + kPbTile <= a.cap ? 0xFFFFFFFFu : (a.cap > first ? (uint32_t) (a.cap - first) : 0u);
  const uint32_t n = cnt < room ? cnt : room;
  const int64_t *ok = a.out_key + first, *op = a.out_payload + first;
  asm("" : "+l"(ok));
  asm("" : "+l"(op));
  for (uint32_t r = lane; r < n; r += 32) {
    const uint64_t key = stage[r];
    if (OUT == 3 || (outsel & 1u)) st_stream_u64<MODE>((void *) elem_ptr_u64(ok, r), key, pol);
    if (OUT == 3 || (outsel & 2u)) st_stream_u64<MODE>((void *) elem_ptr_u64(op, r), key, pol);
  }
}

template <int MODE>
__device__ __forceinline__ void lean_load_keys(const ProbeArgs &a, const CachePolicy &pol, unsigned long long off, uint32_t rows,
                                               uint64_t (&kn)[kPbKeysPerThread]) {
  const int64_t *p = a.keys + off + threadIdx.x;
  asm("" : "+l"(p));  // keep ONE base pointer; the four loads use immediate offsets
#pragma unroll
  for (int j = 0; j < kPbKeysPerThread; ++j)
    kn[j] = (uint32_t) (j * kPbThreads) + threadIdx.x < rows ? ld_stream_u64<MODE>(p + j * kPbThreads, pol) : 0;
}

// OUT: which result columns exist, as a compile-time constant for the two hot shapes -- 0 = none (count + checksums only),
// 3 = key + payload; -1 = decided at run time (row ids, single columns)
#ifndef CCB_LEAN_MIN_BLOCKS
#define CCB_LEAN_MIN_BLOCKS 4
#endif
template <int KIND, int MODE, int OUT>
__global__ void __launch_bounds__(kPbThreads, CCB_LEAN_MIN_BLOCKS) probe_unique_kernel(ProbeArgs a) {
  __shared__ LeanShared sh;
  const CachePolicy pol = make_policies();
  if (a.gate && ((*a.gate != 0) != (a.gate_want != 0))) return;  // device-side strategy switch (CTA-uniform)
  const uint32_t mask = (uint32_t) a.mask;
  const unsigned lane = lane_id(), w = threadIdx.x >> 5, lt = lanemask_lt();
  const unsigned outsel = OUT >= 0 ? (unsigned) OUT : ((a.out_key ? 1u : 0u) | (a.out_payload ? 2u : 0u) | (a.out_rowid ? 4u : 0u));
  uint64_t ksum = 0;
  if (threadIdx.x == 0) {
    if (a.seg_parts) {
      unsigned long long c = a.seg_cursors[seg_region(a, 0)];
      sh.seg_cnt = c < a.seg_cap ? c : a.seg_cap;
      sh.seg_p = 0;
      sh.seg_lo = 0;
      sh.seg_hi = a.seg_prefix[1];
      sh.seg_total = a.seg_prefix[a.seg_parts];
    }
    lean_tile(a, sh, atomicAdd(a.tile_counter, 1ull), sh.off[0], sh.rows[0]);
    lean_tile(a, sh, atomicAdd(a.tile_counter, 1ull), sh.off[1], sh.rows[1]);
  }
  __syncthreads();
  unsigned long long off = sh.off[0], noff = sh.off[1];
  uint32_t rows = sh.rows[0], nrows = sh.rows[1];
  __syncthreads();
  uint64_t kn[kPbKeysPerThread];
  lean_load_keys<MODE>(a, pol, off, rows, kn);
  unsigned par = 0;
  uint32_t prev_woff = 0, prev_cnt = 0;   // this warp's share of the previous tile's output range
  unsigned long long pending = 0;         // thread 0: output base of the previous tile (result of its atomicAdd)
  while (rows > 0) {
    // thread 0 asks for the tile after next now and publishes it before the emit barrier
    unsigned long long g_after = 0;
    if (threadIdx.x == 0) g_after = atomicAdd(a.tile_counter, 1ull);
    uint64_t k[kPbKeysPerThread], v[kPbKeysPerThread];
    uint32_t p[kPbKeysPerThread], e[kPbKeysPerThread];
    // ---- Probe (linear_probing_ht.cpp:45-57 / chaining_ht.cpp:44-55): all home-slot loads of a thread in flight together
#pragma unroll
    for (int j = 0; j < kPbKeysPerThread; ++j) {
      k[j] = kn[j];
      p[j] = home_slot32(k[j], mask);
    }
    if (KIND == CC_HT_LP) {
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j)
        v[j] = (uint32_t) (j * kPbThreads) + threadIdx.x < rows ? ld_table_u64<MODE>(elem_ptr_u64(a.slots, p[j]), pol) : kEmptyU;
      lean_load_keys<MODE>(a, pol, noff, nrows, kn);  // next tile's keys, behind this tile's gathers
    } else {
      uint2 d[kPbKeysPerThread];
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j)
        d[j] = (uint32_t) (j * kPbThreads) + threadIdx.x < rows ? ld_table_u32x2<MODE>(reinterpret_cast<const uint2 *>(elem_ptr_u64(a.dir, p[j])), pol)
                                                                : make_uint2(0u, 0u);
      lean_load_keys<MODE>(a, pol, noff, nrows, kn);
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j) {
        p[j] = d[j].x;
        e[j] = d[j].y ? d[j].x + d[j].y : d[j].x + 1u;  // empty bucket: looks like a one-entry chain whose entry never matches
        v[j] = d[j].y ? ld_table_u64<MODE>(elem_ptr_u64(a.ckeys, p[j]), pol) : ~k[j];
      }
    }
    // ---- walk (LPScanStructure::Next, linear_probing_ht.cpp:62-115 / AdvancePointers, chaining_ht.cpp:109-124) up to
    // the first match: beyond it only duplicates could follow, and this table has none.
    unsigned um, mm;
    uint32_t ntail = 0;  // LP: matches found by the tail walk, parked in sh.queue[w][0 .. ntail)
    if (KIND == CC_HT_LP) {
      // The CTA waits at the emit barrier for its SLOWEST key, and among 1024 keys the longest probe sequence is ~7 slots:
      // walking them slot by slot costs ~7 dependent L2 round trips per tile.  Instead the unresolved keys of the warp are
      // queued densely, one lane takes one key and examines a whole sector (4 slots, one request) per round trip.
      um = 0;
      mm = 0;
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j) {
        const bool hit = v[j] == k[j];
        mm |= hit ? (1u << j) : 0u;
        um |= (!hit && v[j] != kEmptyU) ? (1u << j) : 0u;
      }
      uint64_t *const q = sh.queue[w];
      uint32_t nq = 0;
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j) {
        const unsigned bu = __ballot_sync(0xffffffffu, (um & (1u << j)) != 0u);
        if (um & (1u << j)) q[nq + __popc(bu & lt)] = k[j];
        nq += __popc(bu);
      }
      __syncwarp();
      for (uint32_t b = 0; b < nq; b += 32) {
        const bool act = b + lane < nq;
        const uint64_t key = act ? q[b + lane] : 0;
        uint32_t at = (home_slot32(key, mask) + 1u) & mask;  // the home slot itself holds another key
        bool matched = false, open = act;
        while (open) {
          const uint32_t s0 = at & ~3u;
          uint64_t x[4];
          ld_table_sector<MODE>(a.slots, s0, pol, x);
          unsigned eq = 0, stop = 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            eq |= x[i] == key ? (1u << i) : 0u;
            stop |= (x[i] == key || x[i] == kEmptyU) ? (1u << i) : 0u;
          }
          stop &= 0xFu << (at & 3u);  // slots before `at` were examined earlier
          if (stop) {
            matched = (eq & stop & (0u - stop)) != 0u;  // the first slot that ends the walk: the key, or an empty slot
            open = false;
          } else {
            at = (s0 + 4u) & mask;
          }
        }
        __syncwarp();  // every lane has read its queue entry before the match list overwrites the front of the queue
        const unsigned bm = __ballot_sync(0xffffffffu, matched);
        if (matched) q[ntail + __popc(bm & lt)] = key;
        ntail += __popc(bm);
      }
      __syncwarp();
    } else {
      // chains of a duplicate-free table are short: all unresolved keys of a thread advance together, one entry per lap
      for (;;) {
        um = 0;
        mm = 0;
#pragma unroll
        for (int j = 0; j < kPbKeysPerThread; ++j) {
          const bool hit = v[j] == k[j];
          mm |= hit ? (1u << j) : 0u;
          um |= (!hit && p[j] + 1u != e[j]) ? (1u << j) : 0u;
        }
        if (!um) break;
#pragma unroll
        for (int j = 0; j < kPbKeysPerThread; ++j) {
          if (um & (1u << j)) {
            p[j] += 1u;
            v[j] = ld_table_u64<MODE>(elem_ptr_u64(a.ckeys, p[j]), pol);
          }
        }
      }
    }
    // ---- compaction, decoupled from the global round trip.  The warp PARKS its matches densely in its own region of
    // shared memory and thread 0 issues the ONE atomicAdd that reserves the tile's output range; nobody waits for it.
    // The parked rows are flushed one iteration later, when the reservation has long returned, with fully coalesced
    // stores (a warp store covers 32 consecutive output rows).  One barrier per tile: every warp derives its own offset
    // from the 8 published warp counts instead of waiting for a scan by warp 0.
    unsigned bal[kPbKeysPerThread];
    uint32_t wtotal = 0;
#pragma unroll
    for (int j = 0; j < kPbKeysPerThread; ++j) {
      bal[j] = __ballot_sync(0xffffffffu, (mm & (1u << j)) != 0u);
      wtotal += __popc(bal[j]);
    }
    wtotal += ntail;
    if (lane == 0) sh.cnt[par][w] = wtotal;
    if (threadIdx.x == 0) {
      lean_tile(a, sh, g_after, sh.off[par], sh.rows[par]);
      sh.base[par] = pending;  // output base of the PREVIOUS tile (reserved one iteration ago)
    }
    __syncthreads();
    uint32_t woff = 0, total = 0;
    {
      const uint4 c0 = *reinterpret_cast<const uint4 *>(&sh.cnt[par][0]), c1 = *reinterpret_cast<const uint4 *>(&sh.cnt[par][4]);
      const uint32_t c[kPbWarps] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
      for (int i = 0; i < kPbWarps; ++i) {
        woff += (unsigned) i < w ? c[i] : 0u;
        total += c[i];
      }
    }
    lean_flush<MODE, OUT>(a, pol, sh.stage[w], sh.base[par] + prev_woff, prev_cnt, outsel, lane);
    __syncwarp();
    if (OUT != 0) {
      uint32_t run = 0;
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j) {
        if (mm & (1u << j)) {
          ksum += k[j];
          sh.stage[w][run + __popc(bal[j] & lt)] = k[j];
        }
        run += __popc(bal[j]);
      }
      if (KIND == CC_HT_LP) {
        for (uint32_t i = lane; i < ntail; i += 32) {
          const uint64_t key = sh.queue[w][i];
          ksum += key;
          sh.stage[w][run + i] = key;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j) ksum += (mm & (1u << j)) ? k[j] : 0ull;
      if (KIND == CC_HT_LP)
        for (uint32_t i = lane; i < ntail; i += 32) ksum += sh.queue[w][i];
    }
    __syncwarp();  // parked rows are visible to the flushing lanes; the queue may be refilled
    if (threadIdx.x == 0) pending = total ? atomicAdd((unsigned long long *) &a.res->n_matches, (unsigned long long) total) : 0ull;
    prev_woff = woff;
    prev_cnt = wtotal;
    // the tile published before the barrier of this iteration becomes the prefetched one
    off = noff;
    rows = nrows;
    noff = sh.off[par];
    nrows = sh.rows[par];
    par ^= 1u;
  }
  // flush the rows parked by the last iteration
  if (threadIdx.x == 0) sh.base[par] = pending;
  __syncthreads();
  lean_flush<MODE, OUT>(a, pol, sh.stage[w], sh.base[par] + prev_woff, prev_cnt, outsel, lane);
  ksum = warp_sum_u64(ksum);
  if (lane == 0 && ksum) {
    atomicAdd((unsigned long long *) &a.res->key_sum, (unsigned long long) ksum);
    atomicAdd((unsigned long long *) &a.res->payload_sum, (unsigned long long) ksum);  // payload == matched build key == probe key
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// probe_unique_lp_kernel: the lean kernel for LP tables with the tail walk DEFERRED through a persistent per-warp queue.
// probe_unique_kernel resolves the ~13 % of keys whose home slot holds another key inside the same iteration: a second
// dependent L2 round trip per tile for all of them, a third for the ~3 % that need another sector (nearly every warp has one),
// and the CTA waits at the tile barrier for its slowest warp (ncu, round 1: 17 % of all stall samples on the tail walk's
// LDG.256, 16.5 % at the barrier).  Here an unresolved key is pushed into a small per-warp ring (key, next slot); every
// iteration each lane pops at most ONE pending entry and loads its sector TOGETHER with the tile's home-slot gathers, so an
// iteration has exactly one load phase whatever the probe sequences look like.  An entry that is still unresolved after its
// sector goes back into the ring; matches found by the tail are parked with the matches of the tile that is being probed when
// they resolve (row order is unspecified).  When the key column is exhausted the CTA keeps iterating without a tile until
// every ring has drained.  Ring overflow (more than kRingCap pending keys: only adversarial key sets) falls back to walking
// the excess keys to completion on the spot, like the generic kernel does.
constexpr int kRingCap = 64;
struct __align__(16) LeanLpShared {
  uint32_t cnt[2][kPbWarps];
  unsigned long long base[2];
  unsigned long long off[2];
  uint32_t rows[2];
  uint32_t seg_p, seg_lo, seg_hi, seg_total;
  unsigned long long seg_cnt;
  uint64_t ring_key[kPbWarps][kRingCap];
  uint32_t ring_at[kPbWarps][kRingCap];
  // per-warp parking area of the matches of one iteration (at most 128 home matches + 32 tail matches)
  uint64_t stage[kPbWarps][kPbKeysPerThread * 32 + 32];
};

template <class SH>
__device__ __forceinline__ void lean_tile_t(const ProbeArgs &a, SH &sh, unsigned long long g, unsigned long long &off, uint32_t &rows) {
  off = 0;
  rows = 0;
  if (a.seg_parts == 0) {
    const unsigned long long ntiles = (a.n + kPbTile - 1) / kPbTile;
    if (g >= ntiles) return;
    off = g * (unsigned long long) kPbTile;
    rows = (uint32_t) (a.n - off < (unsigned long long) kPbTile ? a.n - off : (unsigned long long) kPbTile);
    return;
  }
  if (g >= sh.seg_total) return;
  uint32_t p = sh.seg_p, lo = sh.seg_lo, hi = sh.seg_hi;
  if ((uint32_t) g >= hi) {
    do {
      ++p;
      lo = hi;
      hi = __ldg(a.seg_prefix + p + 1);
    } while ((uint32_t) g >= hi);
    unsigned long long c = __ldg(a.seg_cursors + seg_region(a, p));
    sh.seg_cnt = c < a.seg_cap ? c : a.seg_cap;
    sh.seg_p = p;
    sh.seg_lo = lo;
    sh.seg_hi = hi;
  }
  const unsigned long long first = (unsigned long long) ((uint32_t) g - lo) * kPbTile;
  const unsigned long long left = sh.seg_cnt - first;
  off = (unsigned long long) seg_region(a, p) * a.seg_cap + first;
  rows = left < (unsigned long long) kPbTile ? (uint32_t) left : (uint32_t) kPbTile;
}

template <int MODE, int OUT>
__global__ void __launch_bounds__(kPbThreads, CCB_LEAN_MIN_BLOCKS) probe_unique_lp_kernel(ProbeArgs a) {
  __shared__ LeanLpShared sh;
  const CachePolicy pol = make_policies();
  if (a.gate && ((*a.gate != 0) != (a.gate_want != 0))) return;  // device-side strategy switch (CTA-uniform)
  const uint32_t mask = (uint32_t) a.mask;
  const unsigned lane = lane_id(), w = threadIdx.x >> 5, lt = lanemask_lt();
  const unsigned outsel = OUT >= 0 ? (unsigned) OUT : ((a.out_key ? 1u : 0u) | (a.out_payload ? 2u : 0u) | (a.out_rowid ? 4u : 0u));
  uint64_t ksum = 0;
  if (threadIdx.x == 0) {
    if (a.seg_parts) {
      unsigned long long c = a.seg_cursors[seg_region(a, 0)];
      sh.seg_cnt = c < a.seg_cap ? c : a.seg_cap;
      sh.seg_p = 0;
      sh.seg_lo = 0;
      sh.seg_hi = a.seg_prefix[1];
      sh.seg_total = a.seg_prefix[a.seg_parts];
    }
    lean_tile_t(a, sh, atomicAdd(a.tile_counter, 1ull), sh.off[0], sh.rows[0]);
    lean_tile_t(a, sh, atomicAdd(a.tile_counter, 1ull), sh.off[1], sh.rows[1]);
  }
  __syncthreads();
  unsigned long long off = sh.off[0], noff = sh.off[1];
  uint32_t rows = sh.rows[0], nrows = sh.rows[1];
  __syncthreads();
  uint64_t kn[kPbKeysPerThread];
  lean_load_keys<MODE>(a, pol, off, rows, kn);
  unsigned par = 0;
  uint32_t prev_woff = 0, prev_cnt = 0;   // this warp's share of the previous iteration's output range
  unsigned long long pending = 0;         // thread 0: output base of the previous iteration (result of its atomicAdd)
  uint32_t rhead = 0, rn = 0;             // this warp's ring: first entry, entries (warp-uniform)
  uint64_t *const rkey = sh.ring_key[w];
  uint32_t *const rat = sh.ring_at[w];
  bool more = rows > 0;
  while (more) {
    unsigned long long g_after = 0;
    if (threadIdx.x == 0) g_after = atomicAdd(a.tile_counter, 1ull);
    uint64_t k[kPbKeysPerThread], v[kPbKeysPerThread];
    uint32_t p[kPbKeysPerThread];
#pragma unroll
    for (int j = 0; j < kPbKeysPerThread; ++j) {
      k[j] = kn[j];
      p[j] = home_slot32(k[j], mask);
    }
    // ---- the ONE load phase of the iteration: the tile's home slots (linear_probing_ht.cpp:45-57) ...
#pragma unroll
    for (int j = 0; j < kPbKeysPerThread; ++j)
      v[j] = (uint32_t) (j * kPbThreads) + threadIdx.x < rows ? ld_table_u64<MODE>(elem_ptr_u64(a.slots, p[j]), pol) : kEmptyU;
    // ... one pending probe sequence per lane (LPScanStructure::Next's advance, linear_probing_ht.cpp:101-109, a sector at a time) ...
    const uint32_t npop = rn < 32u ? rn : 32u;
    const bool popped = lane < npop;
    uint64_t tkey = 0;
    uint32_t tat = 0;
    if (popped) {
      tkey = rkey[(rhead + lane) & (kRingCap - 1)];
      tat = rat[(rhead + lane) & (kRingCap - 1)];
    }
    rhead = (rhead + npop) & (kRingCap - 1);
    rn -= npop;
    uint64_t x[4] = {0, 0, 0, 0};
    if (popped) ld_table_sector<MODE>(a.slots, tat & ~3u, pol, x);
    // ... and the next tile's keys
    lean_load_keys<MODE>(a, pol, noff, nrows, kn);
    __syncwarp();  // every lane holds its popped entry: the ring slots may be rewritten
    // ---- resolve the popped entries
    bool tmatch = false, tcont = false;
    if (popped) {
      unsigned eq = 0, stop = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        eq |= x[i] == tkey ? (1u << i) : 0u;
        stop |= (x[i] == tkey || x[i] == kEmptyU) ? (1u << i) : 0u;
      }
      stop &= 0xFu << (tat & 3u);  // slots before `tat` were examined earlier
      if (stop)
        tmatch = (eq & stop & (0u - stop)) != 0u;  // the first slot that ends the walk: the key, or an empty slot
      else
        tcont = true;
    }
    {
      const unsigned bc = __ballot_sync(0xffffffffu, tcont);
      if (tcont) {
        const uint32_t at = (rhead + rn + __popc(bc & lt)) & (kRingCap - 1);
        rkey[at] = tkey;
        rat[at] = ((tat & ~3u) + 4u) & mask;
      }
      rn += __popc(bc);
    }
    // ---- home slots: match, miss, or a pending probe sequence
    unsigned um = 0, mm = 0;
#pragma unroll
    for (int j = 0; j < kPbKeysPerThread; ++j) {
      const bool hit = v[j] == k[j];
      mm |= hit ? (1u << j) : 0u;
      um |= (!hit && v[j] != kEmptyU) ? (1u << j) : 0u;
    }
#pragma unroll
    for (int j = 0; j < kPbKeysPerThread; ++j) {
      const bool u = (um & (1u << j)) != 0u;
      const unsigned bu = __ballot_sync(0xffffffffu, u);
      if (u) {
        const uint32_t slot = rn + __popc(bu & lt);
        if (slot < (uint32_t) kRingCap) {
          const uint32_t at = (rhead + slot) & (kRingCap - 1);
          rkey[at] = k[j];
          rat[at] = (p[j] + 1u) & mask;
        } else {
          // ring full (adversarial key set): walk this key to the end right here
          uint32_t at = (p[j] + 1u) & mask;
          for (;;) {
            const uint64_t y = ld_table_u64<MODE>(elem_ptr_u64(a.slots, at), pol);
            if (y == k[j]) {
              mm |= 1u << j;
              break;
            }
            if (y == kEmptyU) break;
            at = (at + 1u) & mask;
          }
        }
      }
      const uint32_t added = __popc(bu);
      rn = rn + added < (uint32_t) kRingCap ? rn + added : (uint32_t) kRingCap;
    }
    // ---- compaction, decoupled from the global round trip (see probe_unique_kernel)
    unsigned bal[kPbKeysPerThread];
    uint32_t wtotal = 0;
#pragma unroll
    for (int j = 0; j < kPbKeysPerThread; ++j) {
      bal[j] = __ballot_sync(0xffffffffu, (mm & (1u << j)) != 0u);
      wtotal += __popc(bal[j]);
    }
    const unsigned btail = __ballot_sync(0xffffffffu, tmatch);
    const uint32_t home_total = wtotal;
    wtotal += __popc(btail);
    if (lane == 0) sh.cnt[par][w] = wtotal;
    if (threadIdx.x == 0) {
      lean_tile_t(a, sh, g_after, sh.off[par], sh.rows[par]);
      sh.base[par] = pending;  // output base of the PREVIOUS iteration (reserved one iteration ago)
    }
    const int any_pending = __syncthreads_or(rn > 0u);
    uint32_t woff = 0, total = 0;
    {
      const uint4 c0 = *reinterpret_cast<const uint4 *>(&sh.cnt[par][0]), c1 = *reinterpret_cast<const uint4 *>(&sh.cnt[par][4]);
      const uint32_t c[kPbWarps] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
      for (int i = 0; i < kPbWarps; ++i) {
        woff += (unsigned) i < w ? c[i] : 0u;
        total += c[i];
      }
    }
    lean_flush<MODE, OUT>(a, pol, sh.stage[w], sh.base[par] + prev_woff, prev_cnt, outsel, lane);
    __syncwarp();
    if (OUT != 0) {
      uint32_t run = 0;
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j) {
        if (mm & (1u << j)) {
          ksum += k[j];
          sh.stage[w][run + __popc(bal[j] & lt)] = k[j];
        }
        run += __popc(bal[j]);
      }
      if (tmatch) {
        ksum += tkey;
        sh.stage[w][home_total + __popc(btail & lt)] = tkey;
      }
    } else {
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j) ksum += (mm & (1u << j)) ? k[j] : 0ull;
      ksum += tmatch ? tkey : 0ull;
    }
    __syncwarp();  // parked rows are visible to the flushing lanes
    if (threadIdx.x == 0) pending = total ? atomicAdd((unsigned long long *) &a.res->n_matches, (unsigned long long) total) : 0ull;
    prev_woff = woff;
    prev_cnt = wtotal;
    off = noff;
    rows = nrows;
    noff = sh.off[par];
    nrows = sh.rows[par];
    par ^= 1u;
    more = rows > 0 || any_pending != 0;  // no tile left: keep iterating until every warp's ring has drained
  }
  // flush the rows parked by the last iteration
  if (threadIdx.x == 0) sh.base[par] = pending;
  __syncthreads();
  lean_flush<MODE, OUT>(a, pol, sh.stage[w], sh.base[par] + prev_woff, prev_cnt, outsel, lane);
  ksum = warp_sum_u64(ksum);
  if (lane == 0 && ksum) {
    atomicAdd((unsigned long long *) &a.res->key_sum, (unsigned long long) ksum);
    atomicAdd((unsigned long long *) &a.res->payload_sum, (unsigned long long) ksum);  // payload == matched build key == probe key
  }
}

// region_flag (optional): partition overrun flag of an incremental probe -> overflow bit 1
__global__ void probe_finish_kernel(cc_probe_result *res, size_t cap, const int *region_flag = nullptr) {
  if (threadIdx.x == 0 && blockIdx.x == 0) res->overflow = (res->n_matches > cap ? 1 : 0) | ((region_flag && *region_flag) ? 2 : 0);
}

template <int KIND, bool UNIQUE, int MODE, bool W32, bool PAY = false>
static int launch_probe_w(const ProbeArgs &a, cudaStream_t st) {
  static int blocks_per_sm = 0;
  if (!blocks_per_sm) {
    CC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, probe_batch_kernel<KIND, UNIQUE, MODE, W32, PAY>, kPbThreads, 0));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
  }
  size_t ntiles = a.seg_parts ? (size_t) a.seg_parts * (size_t) (a.seg_cap / kPbTile) : (a.n + kPbTile - 1) / kPbTile;
  size_t grid = (size_t) sm_count() * blocks_per_sm;
  if (grid > ntiles) grid = ntiles;
  if (grid == 0) grid = 1;
  probe_batch_kernel<KIND, UNIQUE, MODE, W32, PAY><<<(unsigned) grid, kPbThreads, 0, st>>>(a);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

template <int KIND, int MODE, int OUT>
static int launch_probe_lean_out(const ProbeArgs &a, cudaStream_t st) {
  static int blocks_per_sm = 0;
  if (!blocks_per_sm) {
    CC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, probe_unique_kernel<KIND, MODE, OUT>, kPbThreads, 0));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
  }
  size_t ntiles = a.seg_parts ? (size_t) a.seg_parts * (size_t) (a.seg_cap / kPbTile) : (a.n + kPbTile - 1) / kPbTile;
  size_t grid = (size_t) sm_count() * blocks_per_sm;
  if (grid > ntiles) grid = ntiles;
  if (grid == 0) grid = 1;
  static const size_t extra_smem = [] {  // experiment knob: unused dynamic shared memory (L1 carve-out sensitivity)
    const char *e = getenv("CCB_PROBE_EXTRA_SMEM");
    return e ? (size_t) atoi(e) : (size_t) 0;
  }();
  probe_unique_kernel<KIND, MODE, OUT><<<(unsigned) grid, kPbThreads, extra_smem, st>>>(a);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

template <int MODE, int OUT>
static int launch_probe_lean_lp_out(const ProbeArgs &a, cudaStream_t st) {
  static int blocks_per_sm = 0;
  if (!blocks_per_sm) {
    CC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, probe_unique_lp_kernel<MODE, OUT>, kPbThreads, 0));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
  }
  size_t ntiles = a.seg_parts ? (size_t) a.seg_parts * (size_t) (a.seg_cap / kPbTile) : (a.n + kPbTile - 1) / kPbTile;
  size_t grid = (size_t) sm_count() * blocks_per_sm;
  if (grid > ntiles) grid = ntiles;
  if (grid == 0) grid = 1;
  probe_unique_lp_kernel<MODE, OUT><<<(unsigned) grid, kPbThreads, 0, st>>>(a);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

// A/B switch for measurements: CCB_LEAN_DEFERRED_TAIL=1 selects probe_unique_lp_kernel (tail walk deferred through the per-warp
// ring).  MEASURED SLOWER than the inline tail walk of probe_unique_kernel on the C4 step (18.45 vs 17.75 ms,
// profiles/r2_probe_deferred_tail_ab.txt): removing the dependent round trips does not pay because the kernel is bound by L1
// wavefront throughput, not by that latency -- so the inline walk stays the default and this kernel stays as the evidence.
static bool deferred_tail_enabled() {
  static const bool on = [] {
    const char *e = getenv("CCB_LEAN_DEFERRED_TAIL");
    return e && e[0] == '1';
  }();
  return on;
}

template <int KIND, int MODE>
static int launch_probe_lean(const ProbeArgs &a, cudaStream_t st) {
  if (KIND == CC_HT_LP && deferred_tail_enabled()) {
    if (a.out_key && a.out_payload && !a.out_rowid) return launch_probe_lean_lp_out<MODE, 3>(a, st);
    if (!a.out_key && !a.out_payload && !a.out_rowid) return launch_probe_lean_lp_out<MODE, 0>(a, st);
    return launch_probe_lean_lp_out<MODE, -1>(a, st);
  }
  if (a.out_key && a.out_payload && !a.out_rowid) return launch_probe_lean_out<KIND, MODE, 3>(a, st);
  if (!a.out_key && !a.out_payload && !a.out_rowid) return launch_probe_lean_out<KIND, MODE, 0>(a, st);
  return launch_probe_lean_out<KIND, MODE, -1>(a, st);
}

// A/B switch for measurements: CCB_PROBE_GENERIC=1 in the environment routes every probe through the generic kernel
static bool lean_enabled() {
  static const bool on = [] {
    const char *e = getenv("CCB_PROBE_GENERIC");
    return !(e && e[0] == '1');
  }();
  return on;
}

template <int KIND, bool UNIQUE, int MODE>
static int launch_probe(const ProbeArgs &a, cudaStream_t st) {
  if (a.n_pay > 0)  // payload columns: the generic kernel knows the table index of every match
    return a.mask <= 0xFFFFFFFFull ? launch_probe_w<KIND, UNIQUE, MODE, true, true>(a, st) : launch_probe_w<KIND, UNIQUE, MODE, false, true>(a, st);
  if (UNIQUE && lean_enabled() && a.mask <= 0xFFFFFFFFull && !a.out_rowid) return launch_probe_lean<KIND, MODE>(a, st);
  return a.mask <= 0xFFFFFFFFull ? launch_probe_w<KIND, UNIQUE, MODE, true>(a, st) : launch_probe_w<KIND, UNIQUE, MODE, false>(a, st);
}

template <int MODE>
static int dispatch_probe_mode(const cc_ht *ht, const ProbeArgs &a, cudaStream_t st) {
  bool unique = !ht->has_duplicates;
  if (ht->kind == CC_HT_LP) return unique ? launch_probe<CC_HT_LP, true, MODE>(a, st) : launch_probe<CC_HT_LP, false, MODE>(a, st);
  return unique ? launch_probe<CC_HT_CHAIN, true, MODE>(a, st) : launch_probe<CC_HT_CHAIN, false, MODE>(a, st);
}

static int dispatch_probe(int mode, const cc_ht *ht, const ProbeArgs &a, cudaStream_t st) {
  return (mode & 2) ? dispatch_probe_mode<2>(ht, a, st) : dispatch_probe_mode<0>(ht, a, st);
}

// ---- strategy ------------------------------------------------------------------------------
// direct      : probe the keys in input order (random gather over the whole table)
// partitioned : first group the probe keys by table slice (partition.cu), then probe slice by
//               slice so that the gather hits L2.  Pays one extra streaming pass over the keys
//               (read 8 B + write 8 B per key, plus the 8 B histogram read) and wins whenever the
//               table is far larger than L2: on B200 random 8-byte gathers over an 8 GiB table
//               are capped at ~37 G/s by HBM (measured, profiles/), an L2-resident slice at ~290 G/s.
static int g_strategy = 0;                // 0 auto, 1 direct, 2 partitioned
static int g_mode_partitioned = 2;        // cache mode of the probe kernel behind the partition pass
static int g_mode_direct = 0;             // cache mode of the direct probe
static int g_single_pass = 1;             // partitioned strategy: single-pass partition with a device-side two-pass fallback
static size_t g_slice_bytes = 32u << 20;  // target table bytes per partition (measured sweep: profiles/r1_sweep_slices.txt)

// optional live phase timing (bench.py): CUDA events on the launching stream around the three kernels
static int g_profile = 0;
static cudaEvent_t g_ev[4] = {nullptr, nullptr, nullptr, nullptr};
static int g_ev_valid = 0;  // 0 nothing recorded, 1 direct (probe only), 2 partitioned (count, scatter, probe)
static void profile_mark(int i, cudaStream_t st) {
  if (!g_profile) return;
  if (!g_ev[i]) cudaEventCreate(&g_ev[i]);
  cudaEventRecord(g_ev[i], st);
}

static int log2_floor(size_t x) {
  int l = 0;
  while ((x >> l) > 1) ++l;
  return l;
}

// bytes a probe may touch: keys (+ bucket directory) and, for a probe that gathers payloads, the payload columns
static size_t probed_table_bytes(const cc_ht *ht, int n_pay) {
  return ht->kind == CC_HT_LP ? ht->n_slots * 8 * (size_t) (1 + n_pay) : ht->n_slots * 8 + ht->n_keys * 8 * (size_t) (1 + n_pay);
}

static bool want_partitioned(const cc_ht *ht, size_t n, const uint64_t *d_out_rowid, int n_pay = 0) {
  if (d_out_rowid) return false;  // row ids refer to input order
  if (g_strategy == 1) return false;
  size_t table_bytes = probed_table_bytes(ht, n_pay);
  if (g_strategy == 2) return table_bytes >= 2 * g_slice_bytes && n > 0;
  // auto: the table must be far beyond L2 AND the batch must revisit each 128-byte table line at
  // least twice on average, otherwise grouping the keys buys no L2 reuse
  return table_bytes >= ((size_t) 96 << 20) && n >= ((size_t) 1 << 22) && n >= table_bytes / 64;
}

// The two-pass partition of a SEGMENTED input writes a dense column whose row count only the device knows (ctl[3], tile
// directory at ctl[4]): the probe reads it as one segment of capacity n.
static void dense_rows_on_device(ProbeArgs &a, unsigned long long *ctl, size_t n) {
  a.seg_prefix = reinterpret_cast<const uint32_t *>(ctl + 4);
  a.seg_cursors = ctl + 3;
  a.seg_cap = n;
  a.seg_parts = 1;
}

// seg (optional): the key column is segmented (SegIn, partition.cuh) -- n is then segments * cap, an upper bound of the row count
// Payload request of a probe: which of the table's payload columns to gather and where to put them
struct PayIO {
  int n = 0;
  int64_t *out[CC_MAX_PAYLOAD_COLS] = {nullptr, nullptr, nullptr, nullptr};
  unsigned long long *col_sum = nullptr;
};

// accumulate: *d_result is NOT cleared -- the rows of this call continue behind those of earlier calls (same output columns, one
// running match count), the way the pieces of an incremental probe do
int probe_batch_device(const cc_ht *ht, const int64_t *d_keys, size_t n, int64_t *d_out_key, int64_t *d_out_payload,
                       uint64_t *d_out_rowid, size_t cap, cc_probe_result *d_result, cudaStream_t st, SegIn seg = SegIn(),
                       PayIO pay = PayIO(), bool accumulate = false) {
  ProbeArgs a;
  a.n_pay = pay.n;
  a.col_sum = pay.col_sum;
  bool any_pay_out = false;
  for (int c = 0; c < pay.n; ++c) {
    a.pay[c] = ht->d_pay[c];
    a.out_pay[c] = pay.out[c];
    any_pay_out |= pay.out[c] != nullptr;
  }
  a.slots = ht->d_slots;
  a.dir = ht->d_dir;
  a.ckeys = ht->d_ckeys;
  a.mask = ht->mask;
  a.keys = d_keys;
  a.n = n;
  a.out_key = d_out_key;
  a.out_payload = d_out_payload;
  a.out_rowid = d_out_rowid;
  a.cap = (d_out_key || d_out_payload || d_out_rowid || any_pay_out) ? cap : 0;
  a.res = d_result;
  a.tile_counter = nullptr;
  a.seg_prefix = nullptr;
  a.seg_cursors = nullptr;
  a.seg_cap = 0;
  a.seg_parts = 0;
  a.gate = nullptr;
  a.gate_want = 0;
  if (!accumulate) CC_CUDA(cudaMemsetAsync(d_result, 0, sizeof(cc_probe_result), st));
  if (n) {
    const bool part = !seg.presliced && want_partitioned(ht, n, d_out_rowid, pay.n);
    size_t table_bytes = probed_table_bytes(ht, pay.n);
    int log2_slots = log2_floor(ht->n_slots);
    int log2p = log2_floor((table_bytes + g_slice_bytes - 1) / g_slice_bytes);
    if ((size_t) 1 << log2p < (table_bytes + g_slice_bytes - 1) / g_slice_bytes) ++log2p;
    if (log2p > log2_floor(kMaxParts)) log2p = log2_floor(kMaxParts);
    if (log2p > log2_slots) log2p = log2_slots;
    if (log2p < 1) log2p = 1;
    const int parts = part ? 1 << log2p : 0;
    // stream-ordered scratch: control words [0] tile counter, [1] tile counter of the fallback probe, [2] overflow flag,
    // [8 ..) cursors | counts | offsets (parts each) | tile prefix (parts + 1 uint32), plus the partitioned keys
    unsigned long long *ctl = nullptr;
    int64_t *scratch = nullptr;
    const size_t ctl_words = 8 + 4 * (size_t) parts + 8 + (size_t) seg.segments;
    CC_CUDA(cudaMallocAsync(&ctl, ctl_words * sizeof(unsigned long long), st));
    CC_CUDA(cudaMemsetAsync(ctl, 0, 8 * sizeof(unsigned long long), st));
    a.tile_counter = ctl;
    int rc = CC_OK;
    if (part) {
      // single-pass partition: fixed regions with 12.5 % slack, no histogram pass; skewed inputs overrun a region, raise
      // the flag, and the gated two-pass sequence below redoes the work -- decided on the device, no host round trip
      const bool single = g_single_pass != 0;
      unsigned long long cap_rows = single ? ((n / parts) + (n / parts) / 8 + 2 * (size_t) kPartTile + kPbTile - 1) / kPbTile * kPbTile : 0;
      const size_t scratch_rows = single ? (size_t) parts * cap_rows : n;
      cudaError_t e = cudaMallocAsync(&scratch, scratch_rows * sizeof(int64_t), st);
      if (e != cudaSuccess) {
        cudaGetLastError();
        cudaFreeAsync(ctl, st);
        set_error("partitioned probe: cannot allocate %zu bytes of scratch: %s", scratch_rows * sizeof(int64_t), cudaGetErrorString(e));
        return CC_ERR_NOMEM;
      }
      const PartFn fn = PartFn::slot_bits(ht->mask, log2_slots, log2p);
      unsigned long long *cursors = ctl + 8, *counts = ctl + 8 + parts, *offsets = ctl + 8 + 2 * parts;
      uint32_t *prefix = reinterpret_cast<uint32_t *>(ctl + 8 + 3 * parts);
      int *flag = reinterpret_cast<int *>(ctl + 2);
      profile_mark(0, st);
      if (single) {
        if (g_profile) profile_mark(1, st);
        rc = partition_single_device(d_keys, n, fn, cap_rows, cursors, flag, kPbTile, prefix, scratch, st, seg);
        profile_mark(2, st);
        if (rc == CC_OK) {
          ProbeArgs b = a;
          b.keys = scratch;
          b.seg_prefix = prefix;
          b.seg_cursors = cursors;
          b.seg_cap = cap_rows;
          b.seg_parts = parts;
          b.gate = flag;
          b.gate_want = 0;
          rc = dispatch_probe(g_mode_partitioned, ht, b, st);
        }
        if (rc == CC_OK) rc = partition_device(d_keys, n, fn, counts, offsets, cursors, scratch, st, nullptr, flag, seg, ctl + 3);  // gated fallback
        if (rc == CC_OK && seg.cap) rc = seg_prefix_device(ctl + 3, 1, n, kPbTile, reinterpret_cast<uint32_t *>(ctl + 4), st);
        if (rc == CC_OK) {
          ProbeArgs b = a;
          b.keys = scratch;
          b.tile_counter = ctl + 1;
          if (seg.cap) dense_rows_on_device(b, ctl, n);
          b.gate = flag;
          b.gate_want = 1;
          rc = dispatch_probe(g_mode_partitioned, ht, b, st);
        }
      } else {
        rc = partition_device(d_keys, n, fn, counts, offsets, cursors, scratch, st, g_profile ? &g_ev[1] : nullptr, nullptr, seg, ctl + 3);
        if (rc == CC_OK && seg.cap) rc = seg_prefix_device(ctl + 3, 1, n, kPbTile, reinterpret_cast<uint32_t *>(ctl + 4), st);
        profile_mark(2, st);
        if (rc == CC_OK) {
          a.keys = scratch;
          if (seg.cap) dense_rows_on_device(a, ctl, n);
          rc = dispatch_probe(g_mode_partitioned, ht, a, st);
        }
      }
      profile_mark(3, st);
      g_ev_valid = g_profile ? 2 : 0;
      cudaFreeAsync(scratch, st);
    } else {
      profile_mark(2, st);
      if (seg.cap) {  // probe the segmented column in place
        uint32_t *prefix = reinterpret_cast<uint32_t *>(ctl + 8);
        rc = seg_prefix_device(seg.counts, seg.segments, seg.cap, kPbTile, prefix, st, seg);
        a.seg_prefix = prefix;
        a.seg_cursors = seg.counts;
        a.seg_cap = seg.cap;
        a.seg_parts = seg.segments;
        a.seg_inner = seg.inner;
        a.seg_outer_stride = seg.outer_stride;
      }
      // regions that already are table slices are probed with the cache hints of the partitioned strategy
      if (rc == CC_OK) rc = dispatch_probe(seg.presliced ? g_mode_partitioned : g_mode_direct, ht, a, st);
      profile_mark(3, st);
      g_ev_valid = g_profile ? 1 : 0;
    }
    cudaFreeAsync(ctl, st);
    CC_TRY(rc);
  }
  if (a.cap || d_out_key || d_out_payload || d_out_rowid || any_pay_out) {
    probe_finish_kernel<<<1, 32, 0, st>>>(d_result, a.cap);
    CC_CHECK_LAUNCH();
  }
  return CC_OK;
}

// entry points for pjoin.cu (partition.cuh)
int probe_segmented_device(const cc_ht *ht, const int64_t *d_keys, SegIn seg, int64_t *d_out_key, int64_t *d_out_payload, size_t cap,
                           cc_probe_result *d_result, cudaStream_t st, bool accumulate) {
  return probe_batch_device(ht, d_keys, (size_t) seg.segments * seg.cap, d_out_key, d_out_payload, nullptr, cap, d_result, st, seg, PayIO(), accumulate);
}

int probe_close_device(cc_probe_result *d_result, size_t cap, const int *d_region_flag, cudaStream_t st) {
  probe_finish_kernel<<<1, 32, 0, st>>>(d_result, cap, d_region_flag);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

}  // namespace ccb

using namespace ccb;

extern "C" {

int cc_probe_set_strategy(int strategy, size_t slice_bytes) {
  CC_REQUIRE(strategy >= 0 && strategy <= 3, "strategy must be 0 (auto), 1 (direct), 2 (partitioned) or 3 (partitioned, two-pass)");
  g_single_pass = strategy != 3;
  g_strategy = strategy == 3 ? 2 : strategy;
  if (slice_bytes) g_slice_bytes = slice_bytes;
  return CC_OK;
}

int cc_probe_set_profiling(int enable) {
  g_profile = enable != 0;
  g_ev_valid = 0;
  return CC_OK;
}

int cc_probe_last_phase_ms(float *ms3) {
  CC_REQUIRE(ms3, "NULL argument");
  ms3[0] = ms3[1] = ms3[2] = 0.f;
  if (!g_ev_valid) return CC_OK;
  CC_CUDA(cudaEventSynchronize(g_ev[3]));
  if (g_ev_valid == 2) {
    CC_CUDA(cudaEventElapsedTime(&ms3[0], g_ev[0], g_ev[1]));
    CC_CUDA(cudaEventElapsedTime(&ms3[1], g_ev[1], g_ev[2]));
  }
  CC_CUDA(cudaEventElapsedTime(&ms3[2], g_ev[2], g_ev[3]));
  return CC_OK;
}

int cc_probe_set_cache_mode(int mode_direct, int mode_partitioned) {
  CC_REQUIRE((mode_direct & ~7) == 0 && (mode_partitioned & ~7) == 0, "cache modes are 0..7");
  g_mode_direct = mode_direct;
  g_mode_partitioned = mode_partitioned;
  return CC_OK;
}

// ---- incremental probe ---------------------------------------------------------------------------------------------
// cc_probe_stream_*: the key column arrives in pieces (the sub-batches of a multi-GPU exchange) but is probed as ONE batch.
// For a table beyond L2 every piece is scattered into the table-slice regions as it arrives (the regions fill up across
// pieces) and the probe runs once at the end -- so the table is streamed from HBM once per batch, not once per piece, while
// the slice partition of piece b still overlaps the exchange of piece b + 1.  For a small table every piece is probed at once.
constexpr int kStreamMaxSegs = kMaxPeers;  // segments per piece of a small-table incremental probe (one per sender)
struct cc_probe_stream {
  const cc_ht *ht = nullptr;
  ProbeArgs a;
  bool part = false;
  int parts = 0;
  PartFn fn;
  unsigned long long cap_rows = 0;
  unsigned long long *ctl = nullptr;
  int64_t *scratch = nullptr;
};

int cc_probe_stream_begin(cc_probe_stream **out, const cc_ht *ht, size_t n_expected, int64_t *d_out_key, int64_t *d_out_payload,
                          size_t out_capacity, cc_probe_result *d_result, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(out && ht && d_result, "NULL argument");
  cudaStream_t st = as_stream(s);
  cc_probe_stream *h = new cc_probe_stream();
  h->ht = ht;
  ProbeArgs &a = h->a;
  a.slots = ht->d_slots;
  a.dir = ht->d_dir;
  a.ckeys = ht->d_ckeys;
  a.mask = ht->mask;
  a.keys = nullptr;
  a.n = 0;
  a.out_key = d_out_key;
  a.out_payload = d_out_payload;
  a.out_rowid = nullptr;
  a.cap = (d_out_key || d_out_payload) ? out_capacity : 0;
  a.res = d_result;
  a.tile_counter = nullptr;
  a.seg_prefix = nullptr;
  a.seg_cursors = nullptr;
  a.seg_cap = 0;
  a.seg_parts = 0;
  a.gate = nullptr;
  a.gate_want = 0;
  cudaError_t e = cudaMemsetAsync(d_result, 0, sizeof(cc_probe_result), st);
  h->part = e == cudaSuccess && want_partitioned(ht, n_expected, nullptr);
  if (e == cudaSuccess && h->part) {
    const size_t table_bytes = ht->kind == CC_HT_LP ? ht->n_slots * 8 : ht->n_slots * 8 + ht->n_keys * 8;
    const int log2_slots = log2_floor(ht->n_slots);
    int log2p = log2_floor((table_bytes + g_slice_bytes - 1) / g_slice_bytes);
    if ((size_t) 1 << log2p < (table_bytes + g_slice_bytes - 1) / g_slice_bytes) ++log2p;
    if (log2p > log2_floor(kMaxParts)) log2p = log2_floor(kMaxParts);
    if (log2p > log2_slots) log2p = log2_slots;
    if (log2p < 1) log2p = 1;
    h->parts = 1 << log2p;
    h->fn = PartFn::slot_bits(ht->mask, log2_slots, log2p);
    const size_t per = n_expected / h->parts;
    h->cap_rows = (per + per / 8 + 2 * (size_t) kPartTile + kPbTile - 1) / kPbTile * kPbTile;
  }
  // control words: [0] tile counter, [2] region overrun flag, [8 ..) cursors (parts) | tile prefix (parts + 1 uint32); a small
  // table (parts == 0) keeps the tile prefix of a segmented piece at [8 ..): kStreamMaxSegs + 1 uint32
  const size_t ctl_words = 8 + 2 * (size_t) h->parts + 8 + (kStreamMaxSegs + 2) / 2;
  if (e == cudaSuccess) e = cudaMallocAsync(&h->ctl, ctl_words * sizeof(unsigned long long), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(h->ctl, 0, ctl_words * sizeof(unsigned long long), st);
  if (e == cudaSuccess && h->part) e = cudaMallocAsync(&h->scratch, (size_t) h->parts * h->cap_rows * sizeof(int64_t), st);
  if (e != cudaSuccess) {
    set_error("cc_probe_stream_begin: %s", cudaGetErrorString(e));
    cudaGetLastError();
    if (h->ctl) cudaFreeAsync(h->ctl, st);
    delete h;
    return e == cudaErrorMemoryAllocation ? CC_ERR_NOMEM : CC_ERR_CUDA;
  }
  *out = h;
  return CC_OK;
}

// one piece: dense (n_segments == 0: d_keys[0 .. n)) or segmented (n ignored: n_segments x segment_capacity rows, counts on the device)
int cc_probe_stream_add(cc_probe_stream *h, const int64_t *d_keys, size_t n, int n_segments, size_t segment_capacity,
                        const uint64_t *d_segment_counts, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(h, "NULL argument");
  cudaStream_t st = as_stream(s);
  SegIn seg;
  if (n_segments) {
    CC_REQUIRE(d_keys && d_segment_counts, "NULL argument");
    CC_REQUIRE(n_segments >= 1 && n_segments <= kMaxParts, "n_segments must be in [1, %d]", kMaxParts);
    CC_REQUIRE(segment_capacity > 0 && segment_capacity % kPartTile == 0, "segment_capacity must be a positive multiple of %d", kPartTile);
    seg.counts = reinterpret_cast<const unsigned long long *>(d_segment_counts);
    seg.cap = segment_capacity;
    seg.segments = n_segments;
    n = (size_t) n_segments * segment_capacity;
  }
  if (n == 0) return CC_OK;
  CC_REQUIRE(d_keys, "d_keys is NULL");
  if (h->part)  // scatter into the slice regions, which keep filling up
    return partition_single_device(d_keys, n, h->fn, h->cap_rows, h->ctl + 8, reinterpret_cast<int *>(h->ctl + 2), 0, nullptr, h->scratch, st, seg,
                                   true);
  // small table: probe the piece right away; the output rows continue behind those of the earlier pieces
  ProbeArgs a = h->a;
  a.keys = d_keys;
  a.n = n;
  a.tile_counter = h->ctl;
  CC_CUDA(cudaMemsetAsync(h->ctl, 0, sizeof(unsigned long long), st));
  if (seg.cap) {
    uint32_t *prefix = reinterpret_cast<uint32_t *>(h->ctl + 8);
    CC_REQUIRE(n_segments <= kStreamMaxSegs, "a small-table incremental probe takes at most %d segments per piece", kStreamMaxSegs);
    CC_TRY(seg_prefix_device(seg.counts, seg.segments, seg.cap, kPbTile, prefix, st));
    a.seg_prefix = prefix;
    a.seg_cursors = seg.counts;
    a.seg_cap = seg.cap;
    a.seg_parts = seg.segments;
  }
  return dispatch_probe(g_mode_direct, h->ht, a, st);
}

// probes what the regions hold, closes the result (overflow bit 0: out_capacity too small, bit 1: a slice region overran --
// heavily skewed keys, use cc_probe_batch) and releases the handle
int cc_probe_stream_finish(cc_probe_stream *h, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(h, "NULL argument");
  cudaStream_t st = as_stream(s);
  int rc = CC_OK;
  if (h->part) {
    unsigned long long *cursors = h->ctl + 8;
    uint32_t *prefix = reinterpret_cast<uint32_t *>(h->ctl + 8 + h->parts);
    rc = seg_prefix_device(cursors, h->parts, h->cap_rows, kPbTile, prefix, st);
    if (rc == CC_OK) {
      ProbeArgs a = h->a;
      a.keys = h->scratch;
      a.n = (size_t) h->parts * h->cap_rows;
      a.tile_counter = h->ctl;
      a.seg_prefix = prefix;
      a.seg_cursors = cursors;
      a.seg_cap = h->cap_rows;
      a.seg_parts = h->parts;
      rc = dispatch_probe(g_mode_partitioned, h->ht, a, st);
    }
  }
  if (rc == CC_OK) {
    probe_finish_kernel<<<1, 32, 0, st>>>(h->a.res, h->a.cap, h->part ? reinterpret_cast<const int *>(h->ctl + 2) : nullptr);
    note_launch();
    if (cudaGetLastError() != cudaSuccess) rc = CC_ERR_CUDA;
  }
  if (h->scratch) cudaFreeAsync(h->scratch, st);
  cudaFreeAsync(h->ctl, st);
  delete h;
  return rc;
}

int cc_probe_batch_segmented(const cc_ht *ht, const int64_t *d_keys, int n_segments, size_t segment_capacity, const uint64_t *d_segment_counts,
                             int64_t *d_out_key, int64_t *d_out_payload, size_t out_capacity, cc_probe_result *d_result, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(ht && d_result && d_keys && d_segment_counts, "NULL argument");
  CC_REQUIRE(n_segments >= 1 && n_segments <= kMaxParts, "n_segments must be in [1, %d]", kMaxParts);
  CC_REQUIRE(segment_capacity > 0 && segment_capacity % kPartTile == 0, "segment_capacity must be a positive multiple of %d", kPartTile);
  SegIn seg;
  seg.counts = reinterpret_cast<const unsigned long long *>(d_segment_counts);
  seg.cap = segment_capacity;
  seg.segments = n_segments;
  return probe_batch_device(ht, d_keys, (size_t) n_segments * segment_capacity, d_out_key, d_out_payload, nullptr, out_capacity, d_result,
                            as_stream(s), seg);
}

int cc_probe_batch_payload(const cc_ht *ht, const int64_t *d_keys, size_t n, int64_t *d_out_key, int64_t *d_out_build_key,
                           int64_t *const *h_out_payload_cols, size_t n_out_cols, uint64_t *d_out_rowid, size_t out_capacity,
                           cc_probe_payload_result *d_result, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(ht && d_result, "NULL argument");
  CC_REQUIRE(n == 0 || d_keys, "d_keys is NULL");
  CC_REQUIRE(ht->n_pay > 0, "the table has no payload columns (cc_ht_attach_payload)");
  CC_REQUIRE(n_out_cols <= (size_t) ht->n_pay, "n_out_cols (%zu) exceeds the table's %d payload columns", n_out_cols, ht->n_pay);
  CC_REQUIRE(n_out_cols == 0 || h_out_payload_cols, "h_out_payload_cols is NULL");
  static_assert(offsetof(cc_probe_payload_result, overflow) == offsetof(cc_probe_result, overflow) &&
                    offsetof(cc_probe_payload_result, col_sum) == sizeof(cc_probe_result),
                "cc_probe_payload_result must start with a cc_probe_result");
  PayIO pay;
  pay.n = ht->n_pay;  // every column is summed; only the requested ones are written
  for (size_t c = 0; c < n_out_cols; ++c) pay.out[c] = h_out_payload_cols[c];
  pay.col_sum = reinterpret_cast<unsigned long long *>(d_result->col_sum);
  CC_CUDA(cudaMemsetAsync(d_result->col_sum, 0, sizeof(d_result->col_sum), as_stream(s)));
  return probe_batch_device(ht, d_keys, n, d_out_key, d_out_build_key, d_out_rowid, out_capacity, reinterpret_cast<cc_probe_result *>(d_result),
                            as_stream(s), SegIn(), pay);
}

int cc_probe_batch(const cc_ht *ht, const int64_t *d_keys, size_t n, int64_t *d_out_key, int64_t *d_out_payload,
                   uint64_t *d_out_rowid, size_t out_capacity, cc_probe_result *d_result, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(ht && d_result, "NULL argument");
  CC_REQUIRE(n == 0 || d_keys, "d_keys is NULL");
  return probe_batch_device(ht, d_keys, n, d_out_key, d_out_payload, d_out_rowid, out_capacity, d_result, as_stream(s));
}

// End-to-end path with HOST buffers: slices of the key column are copied to the device on one
// stream, probed on a second, and the dense result columns copied back on a third, triple-buffered
// so that PCIe traffic in both directions overlaps the kernels.  The device buffers, streams and
// events live in a per-process workspace that is created on first use and reused by later calls
// (allocation used to cost more than the transfers themselves).
namespace {
struct HostWs {
  static constexpr int kBuf = 3;
  struct Buf {
    int64_t *d_keys = nullptr, *d_ok = nullptr, *d_op = nullptr;
    cc_probe_result *d_res = nullptr;
    cc_probe_result *h_res = nullptr;
    cudaEvent_t h2d_done = nullptr, k_done = nullptr, d2h_done = nullptr;
    size_t cap = 0;
  } buf[kBuf];
  cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
  size_t slice = 0;
  int device = -1;
  std::mutex mu;
  void release() {
    for (auto &b : buf) {
      if (b.d_keys) cudaFree(b.d_keys);
      if (b.d_ok) cudaFree(b.d_ok);
      if (b.d_op) cudaFree(b.d_op);
      if (b.d_res) cudaFree(b.d_res);
      if (b.h_res) cudaFreeHost(b.h_res);
      if (b.h2d_done) cudaEventDestroy(b.h2d_done);
      if (b.k_done) cudaEventDestroy(b.k_done);
      if (b.d2h_done) cudaEventDestroy(b.d2h_done);
      b = Buf();
    }
    if (s_in) cudaStreamDestroy(s_in);
    if (s_k) cudaStreamDestroy(s_k);
    if (s_out) cudaStreamDestroy(s_out);
    s_in = s_k = s_out = nullptr;
    slice = 0;
    device = -1;
  }
};
HostWs g_ws;
}  // namespace

#define CC_E2E(expr)                                                                     \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      cudaGetLastError();                                                                \
      return e__ == cudaErrorMemoryAllocation ? CC_ERR_NOMEM : CC_ERR_CUDA;              \
    }                                                                                    \
  } while (0)

static int ws_prepare(HostWs &ws, size_t slice, bool want_out) {
  int dev = 0;
  CC_E2E(cudaGetDevice(&dev));
  if (ws.device != dev || ws.slice < slice) {
    ws.release();
    ws.device = dev;
    ws.slice = slice;
    CC_E2E(cudaStreamCreateWithFlags(&ws.s_in, cudaStreamNonBlocking));
    CC_E2E(cudaStreamCreateWithFlags(&ws.s_k, cudaStreamNonBlocking));
    CC_E2E(cudaStreamCreateWithFlags(&ws.s_out, cudaStreamNonBlocking));
    for (auto &b : ws.buf) {
      CC_E2E(cudaMalloc(&b.d_keys, slice * sizeof(int64_t)));
      CC_E2E(cudaMalloc(&b.d_res, sizeof(cc_probe_result)));
      CC_E2E(cudaMallocHost(&b.h_res, sizeof(cc_probe_result)));
      CC_E2E(cudaEventCreateWithFlags(&b.h2d_done, cudaEventDisableTiming));
      CC_E2E(cudaEventCreateWithFlags(&b.k_done, cudaEventDisableTiming));
      CC_E2E(cudaEventCreateWithFlags(&b.d2h_done, cudaEventDisableTiming));
    }
  }
  if (want_out)
    for (auto &b : ws.buf)
      if (b.cap < slice) {
        if (b.d_ok) cudaFree(b.d_ok);
        if (b.d_op) cudaFree(b.d_op);
        b.d_ok = b.d_op = nullptr;
        b.cap = 0;
        CC_E2E(cudaMalloc(&b.d_ok, slice * sizeof(int64_t)));
        CC_E2E(cudaMalloc(&b.d_op, slice * sizeof(int64_t)));
        b.cap = slice;
      }
  return CC_OK;
}

int cc_probe_batch_host(const cc_ht *ht, const int64_t *h_keys, size_t n, int64_t *h_out_key, int64_t *h_out_payload,
                        size_t out_capacity, cc_probe_result *h_result, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(ht && h_result, "NULL argument");
  CC_REQUIRE(n == 0 || h_keys, "h_keys is NULL");
  (void) s;
  HostWs &ws = g_ws;
  std::lock_guard<std::mutex> lock(ws.mu);
  constexpr int kBuf = HostWs::kBuf;
  const size_t slice = std::min<size_t>(n ? n : 1, (size_t) 1 << 24);  // 16 Mi keys = 128 MiB per slice
  const bool want_out = h_out_key || h_out_payload;
  CC_TRY(ws_prepare(ws, slice, want_out));
  cudaStream_t s_in = ws.s_in, s_k = ws.s_k, s_out = ws.s_out;
  cc_probe_result total = {0, 0, 0, 0};
  const size_t n_slices = (n + slice - 1) / slice;
  auto issue = [&](size_t i) -> int {  // H2D + kernel for slice i
    HostWs::Buf &b = ws.buf[i % kBuf];
    size_t off = i * slice, cnt = std::min(slice, n - off);
    if (i >= (size_t) kBuf) {  // buffer reuse: the D2H of slice i-kBuf must have drained
      CC_E2E(cudaStreamWaitEvent(s_in, b.d2h_done, 0));
      CC_E2E(cudaStreamWaitEvent(s_k, b.d2h_done, 0));
    }
    CC_E2E(cudaMemcpyAsync(b.d_keys, h_keys + off, cnt * sizeof(int64_t), cudaMemcpyHostToDevice, s_in));
    CC_E2E(cudaEventRecord(b.h2d_done, s_in));
    CC_E2E(cudaStreamWaitEvent(s_k, b.h2d_done, 0));
    CC_TRY(probe_batch_device(ht, b.d_keys, cnt, want_out ? b.d_ok : nullptr, want_out ? b.d_op : nullptr, nullptr, b.cap, b.d_res, s_k));
    CC_E2E(cudaMemcpyAsync(b.h_res, b.d_res, sizeof(cc_probe_result), cudaMemcpyDeviceToHost, s_k));
    CC_E2E(cudaEventRecord(b.k_done, s_k));
    return CC_OK;
  };
  for (size_t i = 0; i < std::min<size_t>(n_slices, kBuf - 1); ++i) CC_TRY(issue(i));
  for (size_t i = 0; i < n_slices; ++i) {
    HostWs::Buf &b = ws.buf[i % kBuf];
    if (i + kBuf - 1 < n_slices) CC_TRY(issue(i + kBuf - 1));
    CC_E2E(cudaEventSynchronize(b.k_done));
    cc_probe_result r = *b.h_res;
    if (want_out && r.n_matches > b.cap) {
      // fan-out above the slice buffer: grow this buffer and redo the slice (rare)
      CC_E2E(cudaStreamSynchronize(s_out));
      CC_E2E(cudaStreamSynchronize(s_k));
      cudaFree(b.d_ok);
      cudaFree(b.d_op);
      b.d_ok = b.d_op = nullptr;
      b.cap = 0;
      CC_E2E(cudaMalloc(&b.d_ok, (size_t) r.n_matches * sizeof(int64_t)));
      CC_E2E(cudaMalloc(&b.d_op, (size_t) r.n_matches * sizeof(int64_t)));
      b.cap = (size_t) r.n_matches;
      size_t off = i * slice, cnt = std::min(slice, n - off);
      CC_TRY(probe_batch_device(ht, b.d_keys, cnt, b.d_ok, b.d_op, nullptr, b.cap, b.d_res, s_k));
      CC_E2E(cudaMemcpyAsync(b.h_res, b.d_res, sizeof(cc_probe_result), cudaMemcpyDeviceToHost, s_k));
      CC_E2E(cudaEventRecord(b.k_done, s_k));
      CC_E2E(cudaStreamSynchronize(s_k));
      r = *b.h_res;
    }
    if (want_out) {
      size_t room = total.n_matches < out_capacity ? out_capacity - (size_t) total.n_matches : 0;
      size_t ncopy = std::min<size_t>((size_t) r.n_matches, room);
      if (ncopy) {
        CC_E2E(cudaStreamWaitEvent(s_out, b.k_done, 0));
        if (h_out_key) CC_E2E(cudaMemcpyAsync(h_out_key + total.n_matches, b.d_ok, ncopy * sizeof(int64_t), cudaMemcpyDeviceToHost, s_out));
        if (h_out_payload) CC_E2E(cudaMemcpyAsync(h_out_payload + total.n_matches, b.d_op, ncopy * sizeof(int64_t), cudaMemcpyDeviceToHost, s_out));
      }
      if (ncopy < r.n_matches) total.overflow = 1;
    }
    CC_E2E(cudaEventRecord(b.d2h_done, s_out));
    total.n_matches += r.n_matches;
    total.key_sum += r.key_sum;
    total.payload_sum += r.payload_sum;
  }
  CC_E2E(cudaStreamSynchronize(s_out));
  CC_E2E(cudaStreamSynchronize(s_k));
  *h_result = total;
  return CC_OK;
}

// frees the workspace of cc_probe_batch_host (device buffers, streams, events)
int cc_probe_host_release(void) {
  std::lock_guard<std::mutex> lock(g_ws.mu);
  g_ws.release();
  return CC_OK;
}
#undef CC_E2E

}  // extern "C"
