// partition.cu -- hash partitioning of a key column (new functionality).
//
// One primitive, two users:
//   * multi-GPU partitioned join: partition id = murmurhash64(key) >> (64 - log2 P), the HIGH
//     hash bits, independent of the low bits that pick the slot/bucket inside the owner's table.
//     The exchange itself (all-to-all over NVLink) is issued by the host layer (parallel.py).
//   * single-GPU large tables (probe_batch.cu): partition id = the HIGH bits of the key's home
//     slot / bucket index, so every partition probes one contiguous, L2-sized slice of the table
//     ("radix pre-partition of probe keys into L2-sized slices").
// Both are  id = ((hash & pre_mask) >> shift) & (P - 1).
//
// Two passes: a histogram (per-CTA shared-memory counters, one global atomic per bin per CTA),
// then a scatter in which every CTA ranks its tile with shared-memory counters, reserves ONE
// contiguous range per partition (global atomicAdd), sorts the tile by partition in shared
// memory and writes it out so that consecutive threads store to consecutive addresses of the
// same partition (coalesced runs instead of 8-byte scatters).
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "common.cuh"
#include "partition.cuh"
#include "tma.cuh"

namespace ccb {

// gate (optional): the kernel only runs when *gate != 0 (device-side fallback switch, see partition_single_device)
// rows of tile `tile` (kPartTile rows starting at row tile * kPartTile) that hold keys: plain column -> all but the ragged end;
// segmented column -> what the tile's segment has filled there (0 for tiles in a segment's slack)
__device__ __forceinline__ uint32_t tile_rows(size_t tile, size_t n, const SegIn &seg) {
  const size_t tbase = tile * (size_t) kPartTile;
  if (!seg.cap) return (uint32_t) (n - tbase < (size_t) kPartTile ? n - tbase : (size_t) kPartTile);
  const uint32_t per_seg = (uint32_t) (seg.cap / kPartTile);
  const uint32_t s = (uint32_t) tile / per_seg;
  const unsigned long long first = (unsigned long long) ((uint32_t) tile - s * per_seg) * kPartTile;
  unsigned long long c = __ldg(seg.counts + s);
  if (c > seg.cap) c = seg.cap;
  return c > first ? (uint32_t) (c - first < (unsigned long long) kPartTile ? c - first : (unsigned long long) kPartTile) : 0u;
}

__global__ void __launch_bounds__(kPartThreads) partition_count_kernel(const int64_t *__restrict__ keys, size_t n, PartFn fn,
                                                                       unsigned long long *counts, const int *gate, SegIn seg) {
  __shared__ uint32_t s_cnt[kMaxParts];
  if (gate && *gate == 0) return;
  const int parts = fn.parts();
  for (int i = threadIdx.x; i < parts; i += kPartThreads) s_cnt[i] = 0;
  __syncthreads();
  const size_t ntiles = (n + kPartTile - 1) / kPartTile;
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const size_t tbase = tile * (size_t) kPartTile;
    const uint32_t rows = tile_rows(tile, n, seg);  // CTA-uniform
    const int64_t *src = keys + tbase + threadIdx.x;
    uint64_t k[kPartItems];
#pragma unroll
    for (int j = 0; j < kPartItems; ++j) k[j] = (uint32_t) (j * kPartThreads) + threadIdx.x < rows ? (uint64_t) __ldg(src + j * kPartThreads) : 0;
#pragma unroll
    for (int j = 0; j < kPartItems; ++j)
      if ((uint32_t) (j * kPartThreads) + threadIdx.x < rows) atomicAdd(&s_cnt[fn(k[j])], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < parts; i += kPartThreads)
    if (s_cnt[i]) atomicAdd(counts + i, (unsigned long long) s_cnt[i]);
}

// exclusive scan of the P counters (single CTA), also resets the cursors
__global__ void partition_offsets_kernel(const unsigned long long *__restrict__ counts, int parts, unsigned long long *offsets,
                                         unsigned long long *cursors, const int *gate, unsigned long long *total) {
  __shared__ unsigned long long s[kMaxParts];
  if (gate && *gate == 0) return;
  for (int i = threadIdx.x; i < parts; i += blockDim.x) s[i] = counts[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int i = 0; i < parts; ++i) {
      unsigned long long c = s[i];
      s[i] = run;
      run += c;
    }
    if (total) *total = run;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < parts; i += blockDim.x) {
    offsets[i] = s[i];
    cursors[i] = 0;
  }
}

// PEERS == false: all partitions live in one local buffer `dst.p[0]`.
// PEERS == true : partition p is written into dst.p[p] -- a buffer that may be PEER memory of another
//                 GPU (CUDA IPC mapping over NVLink): the scatter IS the exchange, row runs travel as
//                 coalesced stores straight into the owner's receive buffer.
// s_delta marker of a run that overran its region.  A valid delta is (global row) - (offset in the tile), i.e. anything in
// [-kPartTile, 2^62): -1 is a legitimate value (it used to be the marker, which silently dropped such a run), 2^63 is not.
constexpr unsigned long long kDroppedRun = 1ull << 63;


// TMA == true: the NEXT tile of keys is pulled into shared memory by one cp.async.bulk (UBLKCP) while the
// CTA ranks / sorts / stores the current one, so the key stream never waits behind the tile's barriers.
// (Unlike the probe kernel this kernel does no gathers, so the larger shared-memory carve-out costs nothing.)
#ifndef CCB_SCATTER_ABLATE
#define CCB_SCATTER_ABLATE 0
#endif
template <bool PEERS, bool TMA, bool FUSED>
__global__ void __launch_bounds__(kPartThreads, 2)
    partition_scatter_kernel(const int64_t *__restrict__ keys, size_t n, PartFn fn, const unsigned long long *__restrict__ offsets,
                             unsigned long long *cursors, ScatterDst dst, unsigned long long cap_rows, int *flag, int gated, SegIn seg) {
  // cap_rows > 0 (single-pass mode): partition p owns the fixed region [p * cap_rows, (p + 1) * cap_rows) of the output, no
  //   histogram pass needed; a tile that would overrun a region raises *flag and drops that run (the gated two-pass
  //   fallback then redoes the whole partition).  gated: run only if *flag != 0.
  if (gated && *flag == 0) return;
  extern __shared__ __align__(128) unsigned char s_dyn[];
  uint64_t *s_sorted = reinterpret_cast<uint64_t *>(s_dyn);                  // [kPartTile]
  uint64_t *s_in = s_sorted + kPartTile;                                       // [kPartTile] (TMA only)
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_cnt[kMaxParts];
  __shared__ uint16_t s_off[kMaxParts];  // offsets inside the tile (< kPartTile)
  __shared__ unsigned long long s_delta[kMaxParts];  // global base of the partition's run minus its offset in the tile
  __shared__ uint32_t s_warp[kPartThreads / 32];
  const int parts = fn.parts();
  const size_t ntiles = (n + kPartTile - 1) / kPartTile;
  // complete tiles travel through the TMA buffer (prefetched one tile ahead), ragged ones are loaded directly
  uint32_t phase = 0;
  if (TMA) {
    if (threadIdx.x == 0) {
      mbar_init(&s_bar, 1);
      mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0 && (size_t) blockIdx.x < ntiles && tile_rows(blockIdx.x, n, seg) == (uint32_t) kPartTile) {
      mbar_expect_tx(&s_bar, kPartTile * 8);
      tma_load_1d(s_in, keys + (size_t) blockIdx.x * kPartTile, kPartTile * 8, &s_bar);
    }
  }
  // One tile.  FULL (compile-time): the tile holds kPartTile keys -- all but the ragged last tile of a column and the partly
  // filled tiles of a segmented one -- so no per-key bounds checks, no sentinel partition ids and a fixed-trip output loop
  // (halves the instructions per tile: 960 -> 520 per thread).
  // ABL (CCB_SCATTER_ABLATE, measurement builds only -- results are then WRONG): 1 = no output stores, 2 = no global range
  // reservation, 3 = no shared-memory ranking atomics, 4 = no sort through shared memory.
  auto do_tile = [&](auto full_tag, const size_t tile, const uint32_t tile_n) {
    constexpr bool FULL = decltype(full_tag)::value;
    const size_t tbase = tile * (size_t) kPartTile;
    const bool staged = TMA && FULL;
    const size_t next = tile + gridDim.x;
    const bool next_staged = TMA && next < ntiles && tile_rows(next, n, seg) == (uint32_t) kPartTile;
    for (int i = threadIdx.x; i < parts; i += kPartThreads) s_cnt[i] = 0;
    uint64_t k[kPartItems];
    uint32_t p[kPartItems], r[kPartItems];
    if (staged) {
      mbar_wait(&s_bar, phase);
      phase ^= 1u;
#pragma unroll
      for (int j = 0; j < kPartItems; ++j) k[j] = s_in[j * kPartThreads + threadIdx.x];
    } else {
      const int64_t *src = keys + tbase + threadIdx.x;
#pragma unroll
      for (int j = 0; j < kPartItems; ++j)
        k[j] = (FULL || (uint32_t) (j * kPartThreads) + threadIdx.x < tile_n) ? (uint64_t) __ldg(src + j * kPartThreads) : 0;
    }
    __syncthreads();  // s_cnt is cleared (loop top)
#pragma unroll
    for (int j = 0; j < kPartItems; ++j) {
      if (FULL) {
        p[j] = fn.template id<FUSED>(k[j]);
#if CCB_SCATTER_ABLATE == 3
        r[j] = (uint32_t) j;
        if (j == 0 && threadIdx.x < (unsigned) parts) s_cnt[threadIdx.x] = kPartTile / parts;
#else
        r[j] = atomicAdd(&s_cnt[p[j]], 1u);
#endif
      } else {
        bool ok = (uint32_t) (j * kPartThreads) + threadIdx.x < tile_n;
        p[j] = ok ? fn.template id<FUSED>(k[j]) : 0xFFFFFFFFu;
        r[j] = ok ? atomicAdd(&s_cnt[p[j]], 1u) : 0;
      }
    }
    // Every thread has now CONSUMED its keys (hashed them), so its reads of the staging buffer are complete.
    // The refill is an async-proxy write: it must be ordered after these generic-proxy reads with a proxy fence
    // -- bar.sync alone is not enough (observed: whole 32-key warp slices replaced by the next tile's keys).
    if (TMA) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (next_staged && threadIdx.x == 0) {
      mbar_expect_tx(&s_bar, kPartTile * 8);
      tma_load_1d(s_in, keys + next * (size_t) kPartTile, kPartTile * 8, &s_bar);
    }
    // exclusive scan of the per-partition counts (parts <= kMaxParts = kBins * kPartThreads).  The global
    // range reservations (one atomicAdd per non-empty partition) are ISSUED here but only consumed after the
    // shared-memory sort below, so their L2 round trip overlaps that phase instead of stalling the CTA.
    constexpr int kBins = kMaxParts / kPartThreads;
    unsigned long long gbase[kBins];
    uint32_t first[kBins];
    {
      uint32_t c[kBins], tsum = 0;
#pragma unroll
      for (int q = 0; q < kBins; ++q) {
        int i = threadIdx.x * kBins + q;
        c[q] = i < parts ? s_cnt[i] : 0;
        tsum += c[q];
      }
      uint32_t incl = warp_incl_scan_u32(tsum);
      if (lane_id() == 31) s_warp[threadIdx.x >> 5] = incl;
      __syncthreads();
      uint32_t woff = 0;
      for (unsigned w = 0; w < (threadIdx.x >> 5); ++w) woff += s_warp[w];
      uint32_t run = woff + incl - tsum;
#pragma unroll
      for (int q = 0; q < kBins; ++q) {
        int i = threadIdx.x * kBins + q;
        first[q] = run;
        gbase[q] = 0;
        if (i < parts) {
          s_off[i] = (uint16_t) run;
          if (c[q]) {
#if CCB_SCATTER_ABLATE == 2
            unsigned long long at = (unsigned long long) (tile / gridDim.x) * 24ull;
#else
            unsigned long long at = atomicAdd(cursors + i, (unsigned long long) c[q]);
#endif
            if (cap_rows) {
              if (at + c[q] > cap_rows) {
                atomicOr(flag, 1);
                gbase[q] = ~0ull;  // overrun: this run is dropped
              } else {
                gbase[q] = (unsigned long long) i * cap_rows + at;
              }
            } else {
              gbase[q] = offsets[i] + at;
            }
          }
        }
        run += c[q];
      }
    }
    __syncthreads();
    // the partition id of a sorted key is RE-HASHED in the output loop instead of being parked beside it in shared memory
    // all eight offset loads first, then the eight stores: written as one loop the compiler emits LDS -> add -> STS chains
    // strictly one after the other and pays the shared-memory latency eight times (ncu source page: 16.5 % of the kernel's
    // stall samples sat on them)
    uint32_t slot[kPartItems];
#pragma unroll
    for (int j = 0; j < kPartItems; ++j) {
#if CCB_SCATTER_ABLATE == 4
      slot[j] = (uint32_t) (j * kPartThreads) + threadIdx.x;
#else
      slot[j] = (FULL || p[j] != 0xFFFFFFFFu) ? s_off[p[j]] + r[j] : 0u;
#endif
    }
#pragma unroll
    for (int j = 0; j < kPartItems; ++j)
      if (FULL || p[j] != 0xFFFFFFFFu) s_sorted[slot[j]] = k[j];
#pragma unroll
    for (int q = 0; q < kBins; ++q) {
      int i = threadIdx.x * kBins + q;
      if (i < parts) s_delta[i] = gbase[q] == ~0ull ? kDroppedRun : gbase[q] - first[q];
    }
    __syncthreads();
    if (FULL) {
#pragma unroll
      for (int j = 0; j < kPartItems; ++j) {
        const uint32_t i = (uint32_t) (j * kPartThreads) + threadIdx.x;
        const uint64_t key = s_sorted[i];
        const uint32_t pp = fn.template id<FUSED>(key);
        int64_t *out = PEERS ? dst.p[FUSED ? pp >> fn.sbits : pp] : dst.p[0];
        const unsigned long long d = s_delta[pp];
#if CCB_SCATTER_ABLATE == 1
        if (d == kDroppedRun + 12345 + key) out[d + i] = (int64_t) key;
#else
        if (d != kDroppedRun) out[d + i] = (int64_t) key;
#endif
      }
    } else {
      for (uint32_t i = threadIdx.x; i < tile_n; i += kPartThreads) {
        const uint64_t key = s_sorted[i];
        const uint32_t pp = fn.template id<FUSED>(key);
        int64_t *out = PEERS ? dst.p[FUSED ? pp >> fn.sbits : pp] : dst.p[0];
        unsigned long long d = s_delta[pp];
        if (d != kDroppedRun) out[d + i] = (int64_t) key;
      }
    }
    __syncthreads();
  };
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const uint32_t tile_n = tile_rows(tile, n, seg);  // CTA-uniform
    if (tile_n == (uint32_t) kPartTile) {
      do_tile(std::true_type(), tile, tile_n);
    } else if (tile_n == 0) {
      // slack tile of a segmented input: nothing to scatter.  No bulk copy is in flight (only complete tiles are staged) and
      // the previous iteration ended behind a barrier with every read of the staging buffer fenced, so it can be refilled.
      const size_t next = tile + gridDim.x;
      if (TMA && next < ntiles && tile_rows(next, n, seg) == (uint32_t) kPartTile && threadIdx.x == 0) {
        mbar_expect_tx(&s_bar, kPartTile * 8);
        tma_load_1d(s_in, keys + next * (size_t) kPartTile, kPartTile * 8, &s_bar);
      }
    } else {
      do_tile(std::false_type(), tile, tile_n);
    }
  }
}

template <bool PEERS>
static int launch_scatter(const int64_t *d_keys, size_t n, PartFn fn, const unsigned long long *d_offsets, unsigned long long *d_cursors,
                          const ScatterDst &dst, size_t blocks, cudaStream_t st, unsigned long long cap_rows = 0, int *flag = nullptr,
                          int gated = 0, SegIn seg = SegIn()) {
  static const bool no_tma = [] {  // experiment knob
    const char *e = getenv("CCB_SCATTER_NO_TMA");
    return e && e[0] == '1';
  }();
  const bool tma = !no_tma && (reinterpret_cast<uintptr_t>(d_keys) & 15) == 0;  // bulk copies need 16-byte alignment
  const size_t smem = (tma ? 2 : 1) * (size_t) kPartTile * sizeof(uint64_t);
  auto go = [&](auto kernel) -> int {
    CC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    kernel<<<(unsigned) blocks, kPartThreads, smem, st>>>(d_keys, n, fn, d_offsets, d_cursors, dst, cap_rows, flag, gated, seg);
    return CC_OK;
  };
  const bool fused = fn.obits != 0;  // owner x slice function (pjoin.cu)
  if (tma)
    CC_TRY(fused ? go(partition_scatter_kernel<PEERS, true, true>) : go(partition_scatter_kernel<PEERS, true, false>));
  else
    CC_TRY(fused ? go(partition_scatter_kernel<PEERS, false, true>) : go(partition_scatter_kernel<PEERS, false, false>));
  CC_CHECK_LAUNCH();
  return CC_OK;
}

// tiles of seg_tile rows per partition region: prefix[p] = first tile of segment p (in WALK order), prefix[parts] = total.
// Any number of segments: one CTA, every thread sums a contiguous chunk, one block scan, every thread writes its chunk.
__global__ void __launch_bounds__(1024) partition_seg_prefix_kernel(const unsigned long long *__restrict__ cursors, int parts, unsigned long long cap_rows,
                                                                    uint32_t seg_tile, uint32_t *prefix, SegIn walk) {
  __shared__ uint32_t s_warp[32];
  const int per = (parts + (int) blockDim.x - 1) / (int) blockDim.x;
  const int lo = (int) threadIdx.x * per, hi = lo + per < parts ? lo + per : parts;
  auto tiles_of = [&](int p) -> uint32_t {
    const unsigned long long raw = cursors[walk.region((uint32_t) p)];
    const unsigned long long c = raw < cap_rows ? raw : cap_rows;
    return (uint32_t) ((c + seg_tile - 1) / seg_tile);
  };
  uint32_t mine = 0;
  for (int p = lo; p < hi; ++p) mine += tiles_of(p);
  uint32_t incl = warp_incl_scan_u32(mine);
  if (lane_id() == 31) s_warp[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    const uint32_t x = threadIdx.x < (blockDim.x >> 5) ? s_warp[threadIdx.x] : 0;
    const uint32_t xi = warp_incl_scan_u32(x);
    s_warp[threadIdx.x] = xi - x;
    if (threadIdx.x == 31) prefix[parts] = xi;
  }
  __syncthreads();
  uint32_t run = s_warp[threadIdx.x >> 5] + incl - mine;
  for (int p = lo; p < hi; ++p) {
    prefix[p] = run;
    run += tiles_of(p);
  }
}

int partition_single_device(const int64_t *d_keys, size_t n, PartFn fn, unsigned long long cap_rows, unsigned long long *d_cursors,
                            int *d_flag, uint32_t seg_tile, uint32_t *d_prefix, int64_t *d_out, cudaStream_t st, SegIn seg, bool accumulate,
                            int self_part, int64_t *d_self_out, bool sticky_flag) {
  const int parts = fn.parts();
  if (!accumulate) {
    CC_CUDA(cudaMemsetAsync(d_cursors, 0, parts * sizeof(unsigned long long), st));
    // sticky_flag: the kernel only ever ORs into *d_flag, so an overrun reported by an EARLIER call survives until the caller
    // has looked at it (the copy-engine exchange reuses one flag for many shuffles and checks once per step)
    if (!sticky_flag) CC_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
  }
  size_t blocks = std::min<size_t>((n + kPartTile - 1) / kPartTile, (size_t) sm_count() * 4);
  if (blocks == 0) blocks = 1;
  ScatterDst dst;
  dst.p[0] = d_out;
  if (self_part >= 0 && d_self_out) {
    // one partition goes to a different buffer at the same region offset (the copy-engine exchange: the rows this rank keeps
    // are written straight into its own receive buffer instead of being copied there afterwards)
    const int owners = (fn.obits || fn.sbits) ? 1 << fn.obits : parts;  // plain function: every partition is an "owner"
    CC_REQUIRE(owners <= kMaxPeers && self_part < owners, "a redirected partition needs at most %d destinations", kMaxPeers);
    for (int p = 0; p < owners; ++p) dst.p[p] = p == self_part ? d_self_out : d_out;
    CC_TRY(launch_scatter<true>(d_keys, n, fn, nullptr, d_cursors, dst, blocks, st, cap_rows, d_flag, 0, seg));
  } else {
    CC_TRY(launch_scatter<false>(d_keys, n, fn, nullptr, d_cursors, dst, blocks, st, cap_rows, d_flag, 0, seg));
  }
  if (d_prefix) CC_TRY(seg_prefix_device(d_cursors, parts, cap_rows, seg_tile, d_prefix, st));
  return CC_OK;
}

int partition_single_multi(const int64_t *d_keys, size_t n, PartFn fn, unsigned long long cap_rows, unsigned long long *d_cursors, int *d_flag,
                           int64_t *d_out, const ScatterDst *dsts, cudaStream_t st, bool sticky_flag) {
  const int parts = fn.parts();
  CC_CUDA(cudaMemsetAsync(d_cursors, 0, parts * sizeof(unsigned long long), st));
  if (!sticky_flag) CC_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
  size_t blocks = std::min<size_t>((n + kPartTile - 1) / kPartTile, (size_t) sm_count() * 4);
  if (blocks == 0) blocks = 1;
  const int owners = (fn.obits || fn.sbits) ? 1 << fn.obits : parts;  // a slice-only function (one rank) has ONE owner
  ScatterDst dst;
  bool any = false;
  for (int o = 0; o < kMaxPeers; ++o) {
    dst.p[o] = (o < owners && dsts && dsts->p[o]) ? dsts->p[o] : d_out;
    any = any || dst.p[o] != d_out;
  }
  CC_REQUIRE(!any || owners <= kMaxPeers, "redirected partitions need at most %d destinations", kMaxPeers);
  if (any) return launch_scatter<true>(d_keys, n, fn, nullptr, d_cursors, dst, blocks, st, cap_rows, d_flag, 0, SegIn());
  return launch_scatter<false>(d_keys, n, fn, nullptr, d_cursors, dst, blocks, st, cap_rows, d_flag, 0, SegIn());
}

int seg_prefix_device(const unsigned long long *d_cursors, int parts, unsigned long long cap_rows, uint32_t seg_tile, uint32_t *d_prefix,
                      cudaStream_t st, SegIn walk) {
  partition_seg_prefix_kernel<<<1, parts > 256 ? 1024 : 256, 0, st>>>(d_cursors, parts, cap_rows, seg_tile, d_prefix, walk);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

int partition_device(const int64_t *d_keys, size_t n, PartFn fn, unsigned long long *d_counts, unsigned long long *d_offsets,
                     unsigned long long *d_cursors, int64_t *d_out, cudaStream_t st, cudaEvent_t *after_count, int *gate, SegIn seg,
                     unsigned long long *d_total) {
  const int parts = fn.parts();
  CC_CUDA(cudaMemsetAsync(d_counts, 0, parts * sizeof(unsigned long long), st));
  size_t blocks = std::min<size_t>((n + kPartTile - 1) / kPartTile, (size_t) sm_count() * 4);
  if (blocks == 0) blocks = 1;
  if (n) {
    partition_count_kernel<<<(unsigned) blocks, kPartThreads, 0, st>>>(d_keys, n, fn, d_counts, gate, seg);
    CC_CHECK_LAUNCH();
  }
  partition_offsets_kernel<<<1, 256, 0, st>>>(d_counts, parts, d_offsets, d_cursors, gate, d_total);
  CC_CHECK_LAUNCH();
  if (after_count) {
    if (!*after_count) cudaEventCreate(after_count);
    cudaEventRecord(*after_count, st);
  }
  if (n) {
    ScatterDst dst;
    dst.p[0] = d_out;
    CC_TRY(launch_scatter<false>(d_keys, n, fn, d_offsets, d_cursors, dst, blocks, st, 0, gate, gate ? 1 : 0, seg));
  }
  return CC_OK;
}

}  // namespace ccb

using namespace ccb;

extern "C" {

int cc_partition_count(const int64_t *d_keys, size_t n, int log2_parts, uint64_t *d_counts, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(log2_parts >= 0 && (1 << log2_parts) <= kMaxParts, "log2_parts must be in [0, %d]", 9);
  CC_REQUIRE(d_counts && (n == 0 || d_keys), "NULL argument");
  int parts = 1 << log2_parts;
  CC_CUDA(cudaMemsetAsync(d_counts, 0, parts * sizeof(uint64_t), as_stream(s)));
  if (n == 0) return CC_OK;
  size_t blocks = std::min<size_t>((n + kPartTile - 1) / kPartTile, (size_t) sm_count() * 4);
  partition_count_kernel<<<(unsigned) blocks, kPartThreads, 0, as_stream(s)>>>(d_keys, n, PartFn::high_bits(log2_parts),
                                                                              (unsigned long long *) d_counts, nullptr, SegIn());
  CC_CHECK_LAUNCH();
  return CC_OK;
}

int cc_partition_scatter(const int64_t *d_keys, size_t n, int log2_parts, const uint64_t *d_offsets, uint64_t *d_cursors,
                         int64_t *d_out, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(log2_parts >= 0 && (1 << log2_parts) <= kMaxParts, "log2_parts must be in [0, %d]", 9);
  CC_REQUIRE(d_offsets && d_cursors && (n == 0 || (d_keys && d_out)), "NULL argument");
  int parts = 1 << log2_parts;
  CC_CUDA(cudaMemsetAsync(d_cursors, 0, parts * sizeof(uint64_t), as_stream(s)));
  if (n == 0) return CC_OK;
  size_t blocks = std::min<size_t>((n + kPartTile - 1) / kPartTile, (size_t) sm_count() * 4);
  ScatterDst dst;
  dst.p[0] = d_out;
  return launch_scatter<false>(d_keys, n, PartFn::high_bits(log2_parts), (const unsigned long long *) d_offsets, (unsigned long long *) d_cursors, dst,
                               blocks, as_stream(s));
}

// Single-pass partition by owner (high hash bits) into fixed regions of `region_capacity` rows: no histogram pass and no
// host round trip -- d_counts[p] = rows of partition p, *d_overflow != 0 if a region overran (skewed keys: the result is
// unusable, use the two-pass cc_partition_count + cc_partition_scatter instead).  This is the send side of the copy-engine
// exchange: region p is then copied into peer p's receive buffer as one block.
int cc_partition_single(const int64_t *d_keys, size_t n, int log2_parts, size_t region_capacity, uint64_t *d_counts, int *d_overflow,
                        int64_t *d_out, int self_part, int64_t *d_self_out, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(log2_parts >= 0 && (1 << log2_parts) <= kMaxParts, "log2_parts must be in [0, %d]", 9);
  CC_REQUIRE(d_counts && d_overflow && d_out && (n == 0 || d_keys), "NULL argument");
  CC_REQUIRE(region_capacity > 0, "region_capacity must be positive");
  return partition_single_device(d_keys, n, PartFn::high_bits(log2_parts), region_capacity, (unsigned long long *) d_counts, d_overflow, 0,
                                 nullptr, d_out, as_stream(s), SegIn(), false, self_part, d_self_out, /*sticky_flag=*/true);
}

static int g_peer_blocks = 0;  // 0 = fill the GPU; > 0 = CTA cap of the peer scatter (it is NVLink-bound)

int cc_partition_set_peer_blocks(int blocks) {
  CC_REQUIRE(blocks >= 0, "blocks must be >= 0");
  g_peer_blocks = blocks;
  return CC_OK;
}

int cc_partition_scatter_peers(const int64_t *d_keys, size_t n, int log2_parts, const uint64_t *d_base, uint64_t *d_cursors,
                               int64_t *const *h_peer_bufs, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(log2_parts >= 0 && (1 << log2_parts) <= kMaxPeers, "at most %d peers", kMaxPeers);
  CC_REQUIRE(d_base && d_cursors && h_peer_bufs && (n == 0 || d_keys), "NULL argument");
  int parts = 1 << log2_parts;
  ScatterDst dst;
  for (int p = 0; p < parts; ++p) {
    CC_REQUIRE(h_peer_bufs[p], "NULL peer buffer %d", p);
    dst.p[p] = h_peer_bufs[p];
  }
  CC_CUDA(cudaMemsetAsync(d_cursors, 0, parts * sizeof(uint64_t), as_stream(s)));
  if (n == 0) return CC_OK;
  size_t blocks = std::min<size_t>((n + kPartTile - 1) / kPartTile, (size_t) sm_count() * 4);
  if (g_peer_blocks > 0 && blocks > (size_t) g_peer_blocks) blocks = (size_t) g_peer_blocks;
  return launch_scatter<true>(d_keys, n, PartFn::high_bits(log2_parts), (const unsigned long long *) d_base, (unsigned long long *) d_cursors, dst,
                              blocks, as_stream(s));
}

// ---- CUDA IPC: map another rank's receive buffer into this process (one process per GPU) -------
int cc_ipc_export(void *d_ptr, cc_ipc_handle *out) {
  CC_TRY(require_device());
  CC_REQUIRE(d_ptr && out, "NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(cc_ipc_handle), "handle size");
  cudaIpcMemHandle_t h;
  CC_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
  memset(out, 0, sizeof(*out));
  memcpy(out->bytes, &h, sizeof(h));
  return CC_OK;
}

int cc_ipc_open(const cc_ipc_handle *handle, void **d_ptr) {
  CC_TRY(require_device());
  CC_REQUIRE(handle && d_ptr, "NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle->bytes, sizeof(h));
  CC_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return CC_OK;
}

int cc_ipc_close(void *d_ptr) {
  if (!d_ptr) return CC_OK;
  CC_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return CC_OK;
}

}  // extern "C"
