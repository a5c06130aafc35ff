// partition.cuh -- shared by partition.cu and probe_batch.cu
#pragma once
#include "common.cuh"

namespace ccb {

constexpr int kPartThreads = 512;
constexpr int kPartItems = 8;
constexpr int kPartTile = kPartThreads * kPartItems;  // 4096 keys = 32 KiB of shared-memory staging; measured: 2048-key tiles give 32-byte write runs and run 2x slower
constexpr int kMaxParts = 512;
constexpr int kMaxPeers = 16;  // destination buffers of the peer scatter (GPUs of one NVLink domain)
static_assert(kMaxParts % kPartThreads == 0, "scan assumes a whole number of bins per thread");

// partition id = owner * (pmask + 1) + slice  with  slice = ((murmurhash64(key) & pre_mask) >> shift) & pmask  and
// owner = the high obits bits of the hash (0 when obits == 0): the plain functions have one of the two parts, the FUSED
// function of the partitioned join (pjoin.cu) both -- a sender groups its keys by (owner GPU, table slice of that owner) in
// one pass, so that the owner never has to partition what it receives.
struct PartFn {
  uint64_t pre_mask;
  uint32_t shift;
  uint32_t pmask;
  uint32_t obits = 0;  // log2 of the number of owners
  uint32_t sbits = 0;  // log2 (pmask + 1), only needed when obits > 0
  // FUSED is a compile-time choice in the scatter kernel: the plain form must not pay for the owner bits (measured: +1.2 ms per
  // 2^31 keys on the C4 step when the choice was made per key)
  template <bool FUSED>
  __host__ __device__ __forceinline__ uint32_t id(uint64_t key) const {
    const uint64_t hh = murmurhash64(key);
    const uint64_t h = hh & pre_mask;
    const uint32_t slice = shift >= 64 ? 0u : ((uint32_t) (h >> shift) & pmask);
    if (!FUSED) return slice;
    return (((uint32_t) (hh >> (64 - obits))) << sbits) | slice;
  }
  __host__ __device__ __forceinline__ uint32_t operator()(uint64_t key) const { return obits ? id<true>(key) : id<false>(key); }
  __host__ __device__ __forceinline__ int parts() const { return (int) ((pmask + 1u) << obits); }
  // owner x table slice: owner = high log2_owners hash bits, slice = high log2_slices bits of the home slot / bucket in a table
  // of 2^log2_slots entries (every owner's table has the same size)
  static PartFn owner_and_slice(int log2_owners, uint64_t table_mask, int log2_slots, int log2_slices) {
    PartFn f = log2_slices > 0 ? slot_bits(table_mask, log2_slots, log2_slices) : PartFn{0, 64, 0};
    f.obits = (uint32_t) log2_owners;
    f.sbits = (uint32_t) log2_slices;
    return f;
  }
  // multi-GPU owner: the high log2p bits of the hash
  static PartFn high_bits(int log2p) {
    PartFn f;
    f.pre_mask = ~0ull;
    f.shift = (uint32_t) (64 - log2p);
    f.pmask = (1u << log2p) - 1u;
    return f;
  }
  // table slice: the high log2p bits of the home slot / bucket index (table has 2^log2_slots entries)
  static PartFn slot_bits(uint64_t table_mask, int log2_slots, int log2p) {
    PartFn f;
    f.pre_mask = table_mask;
    f.shift = (uint32_t) (log2_slots - log2p);
    f.pmask = (1u << log2p) - 1u;
    return f;
  }
};

// Segmented INPUT column (the receive buffer of the copy-engine exchange, parallel.py): segment s holds counts[s] valid rows
// at keys + s * cap, cap a multiple of kPartTile; the kernels then walk n = segments * cap rows and skip the slack.
// cap == 0: plain dense input.
// inner > 0: the segments are WALKED in an order that differs from their order in memory -- segment p of the walk is region
//   r = (p % inner) * outer_stride + p / inner   (rows at keys + r * cap, fill counts[r]).  The receive arena of the partitioned
//   join is laid out [piece][sender][slice] and probed slice by slice: inner = pieces * senders, outer_stride = slices allocated.
// presliced: the regions already are table slices (the sender partitioned by owner AND slice): probe them as they are.
struct SegIn {
  const unsigned long long *counts = nullptr;
  unsigned long long cap = 0;
  int segments = 0;
  int inner = 0;
  int outer_stride = 0;
  bool presliced = false;
  __host__ __device__ __forceinline__ uint32_t region(uint32_t p) const {
    return inner ? (p % (uint32_t) inner) * (uint32_t) outer_stride + p / (uint32_t) inner : p;
  }
};

// destination buffers of a scatter, one per OWNER (plain function: per partition): the peer scatter / the redirected partitions
struct ScatterDst {
  int64_t *p[kMaxPeers];
};

// histogram + offsets + scatter on `st`; d_counts/d_offsets/d_cursors hold P entries each
// after_count (optional): an event slot recorded between the histogram and the scatter (phase timing)
// gate (optional): every kernel of the sequence only runs when *gate != 0 (device-side fallback switch)
int partition_device(const int64_t *d_keys, size_t n, PartFn fn, unsigned long long *d_counts, unsigned long long *d_offsets,
                     unsigned long long *d_cursors, int64_t *d_out, cudaStream_t st, cudaEvent_t *after_count = nullptr, int *gate = nullptr,
                     SegIn seg = SegIn(), unsigned long long *d_total = nullptr);  // d_total (optional): rows written, on the device

// Single-pass partition (no histogram): partition p is scattered into the fixed region [p * cap_rows, (p + 1) * cap_rows) of
// d_out; d_cursors[p] = rows written; *d_flag != 0 if some region overran (then the result is unusable and the caller's gated
// two-pass fallback takes over); d_prefix[0 .. parts] = exclusive prefix of the per-partition tile counts (seg_tile rows each).
int partition_single_device(const int64_t *d_keys, size_t n, PartFn fn, unsigned long long cap_rows, unsigned long long *d_cursors,
                            int *d_flag, uint32_t seg_tile, uint32_t *d_prefix, int64_t *d_out, cudaStream_t st, SegIn seg = SegIn(),
                            bool accumulate = false, int self_part = -1, int64_t *d_self_out = nullptr, bool sticky_flag = false);
// (a fused owner x slice function: self_part is the OWNER whose regions go to d_self_out)
// The general form: owner o's regions go to dsts->p[o] (at the same region offsets as in d_out) wherever that is not NULL --
// this rank's own arena, or a PEER's arena mapped over NVLink: the scatter kernel then stores those rows straight into the
// owner's memory and no copy is needed for them.
int partition_single_multi(const int64_t *d_keys, size_t n, PartFn fn, unsigned long long cap_rows, unsigned long long *d_cursors, int *d_flag,
                           int64_t *d_out, const ScatterDst *dsts, cudaStream_t st, bool sticky_flag);
// accumulate: keep cursors / flag of earlier calls (the regions fill up over several inputs)
// d_prefix[0 .. parts] = exclusive prefix of ceil(min(d_cursors[p], cap_rows) / seg_tile) (the probe kernel's tile directory)
int seg_prefix_device(const unsigned long long *d_cursors, int parts, unsigned long long cap_rows, uint32_t seg_tile, uint32_t *d_prefix,
                      cudaStream_t st, SegIn walk = SegIn());  // walk.inner > 0: prefix in walk order (d_cursors indexed by region)

// probe a SEGMENTED key column (probe_batch.cu); accumulate: keep the running match count / output position of earlier calls
int probe_segmented_device(const cc_ht *ht, const int64_t *d_keys, SegIn seg, int64_t *d_out_key, int64_t *d_out_payload, size_t cap,
                           cc_probe_result *d_result, cudaStream_t st, bool accumulate);
// closes a result that was accumulated over several probes: overflow bit 0 = out_capacity too small, bit 1 = *d_region_flag set
int probe_close_device(cc_probe_result *d_result, size_t cap, const int *d_region_flag, cudaStream_t st);

}  // namespace ccb
