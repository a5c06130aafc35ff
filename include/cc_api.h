/*
 * cc_api.h -- C ABI of the B200-native hash-join probe + chunk-compaction
 * engine (libccb200.so).
 *
 * The reference (YimingQiao/Chunk-Compaction-in-Vectorized-Execution-SIMD) has
 * no FFI / plugin surface: its boundary is the C++ class surface that main.cpp
 * and simd_micro_bench.cpp call.  Every entry point below names
 * the reference interface (file:line in the reference tree) it replaces; the
 * C++ facade in  chunk-compaction-in-vectorized-execution-simd_b200/host/
 * simd_compaction.hpp  re-creates those classes on top of this ABI.
 *
 * Conventions
 *   - plain C, opaque handles, no CUDA/torch types in signatures
 *     (cc_stream_t is a cudaStream_t carried as void*; NULL = default stream)
 *   - pointers named d_* are DEVICE pointers owned by the caller, h_* are host
 *   - every function returns CC_OK (0) or a negative cc_status; the message
 *     of the last failure on the calling thread is cc_last_error()
 *   - there is NO CPU fallback: without a usable sm_100 device every compute
 *     entry point fails with CC_ERR_NO_DEVICE
 *   - all arithmetic is integer and bit-exact w.r.t. the reference; the only
 *     floating point is the bandit's double maths (negative_feedback.hpp)
 */
#ifndef CC_API_H
#define CC_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CC_API_VERSION 1

typedef enum {
  CC_OK = 0,
  CC_ERR_INVALID = -1,     /* bad argument                                   */
  CC_ERR_NO_DEVICE = -2,   /* no CUDA device / not sm_100                    */
  CC_ERR_CUDA = -3,        /* CUDA runtime error (see cc_last_error)         */
  CC_ERR_NOMEM = -4,       /* device or host allocation failed               */
  CC_ERR_CAPACITY = -5,    /* output buffer too small (required size stored) */
  CC_ERR_UNSUPPORTED = -6, /* e.g. key -1 in an LP table (empty sentinel)    */
  CC_ERR_STATE = -7        /* call sequence violation (e.g. Next after end)  */
} cc_status;

typedef void *cc_stream_t; /* cudaStream_t */

typedef struct cc_ht cc_ht;               /* HashTable / LPHashTable            */
typedef struct cc_scan cc_scan;           /* ScanStructure / LPScanStructure    */
typedef struct cc_compactor cc_compactor; /* NaiveCompactor (+Binary/Dynamic)   */
typedef struct cc_tuner cc_tuner;         /* CompactTuner                       */
typedef struct cc_chain cc_chain;         /* PipelineState (main.cpp:14-20)     */

/* table kinds */
#define CC_HT_LP 0    /* linear_probing_ht.h:56 LPHashTable */
#define CC_HT_CHAIN 1 /* chaining_ht.h:86 HashTable         */

/* build flags */
#define CC_BUILD_ORDERED 0   /* default: deterministic layout == serial insertion in ascending (unsigned) key
                                order (LP) / insertion order (chain); equals the reference's layout for its
                                own generator (keys are non-decreasing, chaining_ht.cpp:17-26)               */
#define CC_BUILD_UNORDERED 1 /* plain CAS inserts, layout depends on scheduling (same multiset)              */

/* ------------------------------------------------------------------ runtime */
int cc_api_version(void);
const char *cc_last_error(void);

typedef struct {
  int device;
  int sm_major, sm_minor;
  int sm_count;
  size_t l2_bytes;
  size_t total_mem, free_mem;
  char name[128];
} cc_device_info;

int cc_device_init(int device); /* cudaSetDevice + capability check (must be sm_100) */
int cc_device_get_info(cc_device_info *info);

/* thin memory/stream helpers so C / C++ hosts need no CUDA headers */
int cc_malloc(void **d_ptr, size_t bytes);
int cc_free(void *d_ptr);
int cc_host_alloc(void **h_ptr, size_t bytes); /* pinned */
int cc_host_free(void *h_ptr);
int cc_memcpy_h2d(void *d_dst, const void *h_src, size_t bytes, cc_stream_t stream);
int cc_memcpy_d2h(void *h_dst, const void *d_src, size_t bytes, cc_stream_t stream);
int cc_memcpy_d2d(void *d_dst, const void *d_src, size_t bytes, cc_stream_t stream);
int cc_memset(void *d_dst, int byte, size_t bytes, cc_stream_t stream);
int cc_stream_create(cc_stream_t *stream);
int cc_stream_destroy(cc_stream_t stream);
int cc_stream_sync(cc_stream_t stream);
/* Stream-ordered scratch (the partitioned probe's key regions, cudaMallocAsync) is cached in the device's default memory
 * pool between calls -- about 18 GiB after one 2^31-key probe, invisible to any other allocator in the process.
 * cc_scratch_set_retention bounds what stays cached once the work has drained (default: everything), cc_scratch_release
 * synchronises the device and hands all of it back.                                                                    */
int cc_scratch_set_retention(uint64_t bytes);
int cc_scratch_release(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t cc_launch_count(void);

/* ------------------------------------------------------------ hash function */
/* hash_functions.h:8-16 murmurhash64 (K1: mm512_murmurhash64 :18-28)          */
int cc_hash_u64(const uint64_t *d_in, uint64_t *d_out, size_t n, cc_stream_t stream);

/* --------------------------------------------------------------- generators */
/* build keys: chaining_ht.cpp:15-26 == linear_probing_ht.cpp:14-25           */
int cc_gen_build_keys(int64_t *d_keys, size_t n, size_t chunk_factor, cc_stream_t stream);
/* rows [first, first + count) of the same column for a build side of n_total rows (one rank's share of a partitioned join) */
int cc_gen_build_keys_range(int64_t *d_keys, size_t first, size_t count, size_t n_total, size_t chunk_factor, cc_stream_t stream);
/* SURVEY 8d C4/C5 probe keys: d_keys[i] = murmurhash64(seed + first + i) & mask */
int cc_gen_keys_counter(int64_t *d_keys, size_t n, uint64_t seed, uint64_t first, uint64_t mask, cc_stream_t stream);

/* -------------------------------------------------------------- hash tables */
typedef struct {
  int kind;
  size_t n_keys;
  size_t n_slots;    /* LP: slots (pow2 >= 4n); chain: buckets (pow2 >= 2n)            */
  size_t bytes;      /* device bytes held                                              */
  int has_duplicates; /* some key occurs more than once (probe must walk past matches) */
  size_t max_chain;  /* chain: longest bucket; LP: 0                                   */
} cc_ht_info;

/* LPHashTable::LPHashTable (linear_probing_ht.cpp:4-37) / HashTable::HashTable
 * (chaining_ht.cpp:4-36) over caller-supplied keys (the reference generates its
 * own keys; see cc_ht_build_reference).  Key -1 is rejected for LP tables
 * (empty-slot sentinel, linear_probing_ht.cpp:7).                             */
int cc_ht_build(cc_ht **ht, int kind, const int64_t *d_keys, size_t n, int flags, cc_stream_t stream);
/* The same with the slot / bucket count chosen by the caller (a power of two; 0 = the reference's rule above).  The
 * partitioned multi-GPU join sizes every rank's table from the GLOBAL key count (pow2 >= 4 n_total, divided by the number
 * of ranks) so that a hash partition that is 0.005 % above n_total / P does not double the table
 * (linear_probing_ht.cpp:5-6 sizes for one table; the result multiset does not depend on the slot count).  LP: n_slots >= 2n. */
int cc_ht_build_sized(cc_ht **ht, int kind, const int64_t *d_keys, size_t n, size_t n_slots, int flags, cc_stream_t stream);
/* exact equivalent of `HashTable(n_rhs_tuples, chunk_factor)` /
 * `LPHashTable(n_rhs_tuples, chunk_factor)` (chaining_ht.h:88, linear_probing_ht.h:58) */
int cc_ht_build_reference(cc_ht **ht, int kind, size_t n_rhs_tuples, size_t chunk_factor, cc_stream_t stream);
/* import a host-built LP slot array (interop with a CPU engine; n_slots must be pow2) */
int cc_ht_import_lp(cc_ht **ht, const int64_t *h_slots, size_t n_slots, size_t n_keys, cc_stream_t stream);
int cc_ht_get_info(const cc_ht *ht, cc_ht_info *info);
/* export for tests: LP -> slots[n_slots]; chain -> begin[n_buckets] u32, count[n_buckets] u32, keys[n_keys] */
int cc_ht_export_lp(const cc_ht *ht, int64_t *h_slots);
int cc_ht_export_chain(const cc_ht *ht, uint32_t *h_begin, uint32_t *h_count, int64_t *h_keys);
int cc_ht_destroy(cc_ht *ht);

/* ---- real payload columns.  The reference generates `payload = cnt + 10000000` for build row cnt
 * and then DROPS it: only tuple[0] is pushed into the table (chaining_ht.cpp:21-23,34; linear_probing_ht.cpp:20-22,33),
 * so a reference result row carries the matched build key twice.  Here a table can keep up to CC_MAX_PAYLOAD_COLS int64
 * payload columns next to its keys.  Layout: one device array per column, indexed like the keys -- by SLOT for an LP table,
 * by CHAIN POSITION for a chain table -- so a match found at index i reads its payloads at the same index i and the
 * partitioned probe keeps key and payload slices in L2 together.  The key layout (and with it every key-only entry
 * point and its parity) is unchanged.
 *   d_build_keys : the key column the table was built from (LP: locates the slot of every build row; duplicates of
 *                  a key are assigned to that key's slots in unspecified order -- every probe of the key returns all
 *                  of them, so the result multiset does not depend on it); may be NULL for a chain table
 *   h_payload_cols[n_cols] : HOST array of device column pointers, n_keys rows each
 * Attaching again replaces the previous payload.                                                                   */
#define CC_MAX_PAYLOAD_COLS 4
int cc_ht_attach_payload(cc_ht *ht, const int64_t *d_build_keys, const int64_t *const *h_payload_cols, size_t n_cols,
                         cc_stream_t stream);
/* `HashTable(n, cf)` / `LPHashTable(n, cf)` that KEEPS the payload column the reference generates and drops:
 * payload of build row i = i + 10000000 (chaining_ht.cpp:21)                                                      */
int cc_ht_build_reference_payload(cc_ht **ht, int kind, size_t n_rhs_tuples, size_t chunk_factor, cc_stream_t stream);
size_t cc_ht_payload_cols(const cc_ht *ht);
/* export for tests: h_cols[c] receives column c in table order (LP: n_slots rows, rows of empty slots are 0;
 * chain: n_keys rows in chain order)                                                                              */
int cc_ht_export_payload(const cc_ht *ht, int64_t *const *h_cols);

/* ------------------------------------------- chunk-granular probe protocol */
/* HashTable::Probe (chaining_ht.cpp:38-58) / LPHashTable::Probe
 * (linear_probing_ht.cpp:39-60): d_key_col[block] column storage, d_sel[block]
 * selection vector, count active rows.  The scan object keeps a reference to
 * d_sel for chain tables (chaining_ht.h:54) and a private copy for LP tables
 * (linear_probing_ht.h:48).                                                   */
int cc_probe_chunk(const cc_ht *ht, const int64_t *d_key_col, size_t count, const uint32_t *d_sel, size_t block_size,
                   cc_scan **scan, cc_stream_t stream);
/* ScanStructure::HasNext (chaining_ht.h:48) / LPScanStructure::HasNext (:42)  */
int cc_scan_has_next(const cc_scan *scan);
size_t cc_scan_active(const cc_scan *scan);
/* ScanStructure::Next (chaining_ht.cpp:60-80) / LPScanStructure::Next
 * (linear_probing_ht.cpp:62-115).  in_one != 0 selects InOneNext
 * (chaining_ht.cpp:138-173 / linear_probing_ht.cpp:117-153).
 *   d_in_sel      input.selection_vector_ (Slice composes through it, base.cpp:42-46)
 *   d_out_sel     result.selection_vector_[block]: first *out_count entries are the
 *                 composed physical positions, the rest identity (Reset, base.h:96-99)
 *   d_out_payload result.data_[ncol_in + 1] column storage [block]: payload written at
 *                 the physical LHS position (chaining_ht.cpp:132)
 * Synchronous (returns the count like the reference).                         */
int cc_scan_next(cc_scan *scan, int in_one, const int64_t *d_key_col, const uint32_t *d_in_sel, uint32_t *d_out_sel,
                 int64_t *d_out_payload, size_t *out_count, cc_stream_t stream);
int cc_scan_destroy(cc_scan *scan);

/* DataChunk::Append (base.cpp:15-27): for each of ncol columns
 *   dst[c][dst_count + j] = src[c][src_sel[offset + j]],  j < num.
 * d_dst_cols / d_src_cols are HOST arrays of ncol device column pointers.     */
int cc_chunk_append(int64_t *const *h_dst_cols, size_t dst_count, const int64_t *const *h_src_cols,
                    const uint32_t *d_src_sel, size_t num, size_t offset, size_t ncol, cc_stream_t stream);
/* DataChunk::Slice selection composition (base.cpp:42-46, SIMDSlice :49-68):
 *   d_out_sel[i] = d_other_sel[d_sv[i]], i < count                            */
int cc_sel_compose(uint32_t *d_out_sel, const uint32_t *d_other_sel, const uint32_t *d_sv, size_t count,
                   cc_stream_t stream);
/* DataChunk::Reset (base.h:96-99): identity selection vector                  */
int cc_sel_identity(uint32_t *d_sel, size_t block_size, cc_stream_t stream);
/* DataCollection::FetchChunk (data_collection.cpp:23-27) row-major -> columnar
 * and AppendChunk (:10-21) columnar+sel -> row-major, on device.               */
int cc_rows_to_columns(const int64_t *d_rows, size_t n_rows, size_t ncol, int64_t *const *h_cols, cc_stream_t stream);
int cc_columns_to_rows(const int64_t *const *h_cols, const uint32_t *d_sel, size_t count, size_t ncol, int64_t *d_rows,
                       cc_stream_t stream);

/* ------------------------------------------------------ batch probe (fast) */
/* One launch probes a whole key column and writes DENSE (compacted) output
 * columns -- the fused equivalent of the micro-bench loop
 * simd_micro_bench.cpp:83-116 (Probe + while(HasNext) Next) followed by a
 * Compactor on every result chunk.
 *   d_out_key[i], d_out_payload[i] : probe key and matched build key of result
 *       row i (reference result tuple [k, 0, k], SURVEY 8c layout); either may
 *       be NULL (count / checksum only)
 *   d_out_rowid[i] (optional, may be NULL): index of the probe row
 *   out_capacity : rows the output columns can hold
 *   d_result     : device cc_probe_result, written by the kernel
 * Row order in the output is unspecified (multiset semantics).                */
typedef struct {
  uint64_t n_matches;   /* result rows (== #tuples of the micro-bench)             */
  uint64_t key_sum;     /* wrapping sum of probe keys over result rows             */
  uint64_t payload_sum; /* wrapping sum of payloads over result rows               */
  uint64_t overflow;    /* != 0: out_capacity was too small, rows beyond were cut  */
} cc_probe_result;

int cc_probe_batch(const cc_ht *ht, const int64_t *d_keys, size_t n, int64_t *d_out_key, int64_t *d_out_payload,
                   uint64_t *d_out_rowid, size_t out_capacity, cc_probe_result *d_result, cc_stream_t stream);
/* The same probe over a SEGMENTED key column: segment s holds d_segment_counts[s] keys at d_keys + s * segment_capacity
 * (segment_capacity a multiple of 4096, at most 512 segments); the rows in a segment's slack are never read.  The counts
 * live on the DEVICE, so a producer on the same stream -- the copy-engine exchange of the partitioned multi-GPU join, whose
 * receive buffer has one segment per sender -- never has to report them to the host.  Same output contract as cc_probe_batch
 * (new functionality; no row ids).                                                                          */
int cc_probe_batch_segmented(const cc_ht *ht, const int64_t *d_keys, int n_segments, size_t segment_capacity,
                             const uint64_t *d_segment_counts, int64_t *d_out_key, int64_t *d_out_payload, size_t out_capacity,
                             cc_probe_result *d_result, cc_stream_t stream);
/* Batch probe of a table with payload columns (cc_ht_attach_payload): result row i is
 * (d_out_key[i], d_out_build_key[i], h_out_payload_cols[0][i], ...) = probe key, matched build key, payloads of the
 * matched build row -- the reference result tuple [k, 0, k] with the dropped payload restored.  Any output
 * pointer may be NULL; h_out_payload_cols is a HOST array of n_out_cols <= cc_ht_payload_cols(ht) device pointers.
 * Same strategies (direct / partitioned by table slice) and the same dense, unordered output as cc_probe_batch.    */
typedef struct {
  uint64_t n_matches;
  uint64_t key_sum;     /* wrapping sum of probe keys over result rows                    */
  uint64_t payload_sum; /* wrapping sum of matched build keys over result rows            */
  uint64_t overflow;
  uint64_t col_sum[CC_MAX_PAYLOAD_COLS]; /* wrapping sum of every payload column over ALL result rows (also rows cut by overflow) */
} cc_probe_payload_result;
int cc_probe_batch_payload(const cc_ht *ht, const int64_t *d_keys, size_t n, int64_t *d_out_key, int64_t *d_out_build_key,
                           int64_t *const *h_out_payload_cols, size_t n_out_cols, uint64_t *d_out_rowid, size_t out_capacity,
                           cc_probe_payload_result *d_result, cc_stream_t stream);
/* Incremental probe: the key column arrives in pieces (the sub-batches of a multi-GPU exchange) and is probed as ONE batch.
 *   begin : fixes the table, the dense output columns (same contract as cc_probe_batch) and the expected total row count
 *   add   : one piece, dense (n_segments == 0: d_keys[0 .. n)) or segmented (n ignored; layout as in cc_probe_batch_segmented)
 *   finish: completes the probe, writes *d_result and releases the handle.  overflow bit 0: out_capacity too small;
 *           bit 1: heavily skewed keys overran a table-slice region (use cc_probe_batch for such inputs).
 * For a table beyond L2 every piece is scattered into the table-slice regions as it arrives and the table is streamed from
 * HBM once per batch instead of once per piece; a small table is probed piece by piece.  All calls must use one stream
 * (new functionality).                                                                                        */
typedef struct cc_probe_stream cc_probe_stream;
int cc_probe_stream_begin(cc_probe_stream **out, const cc_ht *ht, size_t n_expected, int64_t *d_out_key, int64_t *d_out_payload,
                          size_t out_capacity, cc_probe_result *d_result, cc_stream_t stream);
int cc_probe_stream_add(cc_probe_stream *h, const int64_t *d_keys, size_t n, int n_segments, size_t segment_capacity,
                        const uint64_t *d_segment_counts, cc_stream_t stream);
int cc_probe_stream_finish(cc_probe_stream *h, cc_stream_t stream);
/* Probe strategy for tables far larger than L2 (process-wide):
 *   0 auto (default): partition the probe keys by table slice when the table is >= 96 MiB, the
 *     batch holds >= max(4 Mi, table_bytes / 64) keys (each 128-byte table line is revisited)
 *     and no row ids are requested; 1 always direct; 2 always partitioned; 3 always partitioned with the
 *     two-pass (histogram + scatter) partition instead of the default single-pass one.
 * slice_bytes: target table bytes per partition (0 keeps the current value, default 32 MiB).
 * The partitioned path needs n * 8 bytes of stream-ordered scratch (cudaMallocAsync).
 * THREADING: cc_probe_set_strategy / _cache_mode / _profiling and cc_partition_set_peer_blocks are PROCESS-WIDE measurement
 * settings and cc_probe_last_phase_ms reads process-wide events: set them before the probes they should affect and do not
 * change them while another thread probes.  Everything else is safe to call from several threads on different handles /
 * streams (cc_probe_batch_host serialises on its workspace).                                                              */
int cc_probe_set_strategy(int strategy, size_t slice_bytes);
/* Cache behaviour of the probe kernel's memory operations (tuning knob; results never change):
 *   bit 0: reserved (128-byte L2 line prefetch of table loads: measured harmful, ignored)
 *   bit 1: L2 eviction priorities -- keys / result columns evict_first, table evict_last
 * one mode for the direct probe (default 0), one for the probe behind the partition pass (default 2). */
int cc_probe_set_cache_mode(int mode_direct, int mode_partitioned);
/* Live phase timing of cc_probe_batch (CUDA events on the launching stream): after enabling,
 * cc_probe_last_phase_ms returns {partition histogram, partition scatter, probe kernel} of the
 * most recent call in milliseconds (the first two are 0 for the direct strategy).            */
int cc_probe_set_profiling(int enable);
int cc_probe_last_phase_ms(float *ms3);
/* host-buffer convenience (the end-to-end path bench.py times as `e2e`): copies h_keys to
 * the device in slices, probes, copies the dense result columns back.          */
int cc_probe_batch_host(const cc_ht *ht, const int64_t *h_keys, size_t n, int64_t *h_out_key, int64_t *h_out_payload,
                        size_t out_capacity, cc_probe_result *h_result, cc_stream_t stream);
/* cc_probe_batch_host keeps its device staging buffers / streams in a per-process workspace
 * (created on first use, reused afterwards); this releases it.                              */
int cc_probe_host_release(void);

/* -------------------------------------------------------------- compactor */
/* NaiveCompactor (compactor.h:14-29, compactor.cpp:5-41), result-transparent
 * (deep copies; SURVEY 8c bug 3), plus the Binary/Dynamic threshold of
 * setting.h:21,24 / main.cpp:141,166:  count >= threshold passes through.
 * threshold == block_size  <=> NaiveCompactor.                                */
int cc_compactor_create(cc_compactor **c, size_t ncol, size_t block_size, size_t threshold);
int cc_compactor_set_threshold(cc_compactor *c, size_t threshold); /* main.cpp:141 SetThreshold */
size_t cc_compactor_get_threshold(const cc_compactor *c);          /* main.cpp:166 GetThreshold */
/* Compact(unique_ptr<DataChunk>&): in: chunk columns h_cols[ncol] (device ptrs), d_sel, *count.
 * out: *count == 0 (buffered)  or  the chunk to push downstream: h_out_cols[ncol] receive the
 * device column pointers of the emitted chunk (the compactor's own dense storage, valid until
 * the next Compact/Flush), *d_out_sel its selection vector, *count its row count.
 * A pass-through returns the caller's own pointers.                            */
int cc_compactor_compact(cc_compactor *c, int64_t *const *h_cols, const uint32_t *d_sel, size_t *count,
                         int64_t **h_out_cols, const uint32_t **d_out_sel, cc_stream_t stream);
/* Flush (compactor.h:23): hands out the partial cache.                        */
int cc_compactor_flush(cc_compactor *c, int64_t **h_out_cols, const uint32_t **d_out_sel, size_t *count,
                       cc_stream_t stream);
int cc_compactor_destroy(cc_compactor *c);

/* ------------------------------------------------------ fused join chain */
/* ExecutePipeline + FlushPipelineCache (main.cpp:119-191) for a whole LHS table
 * in ONE persistent kernel: every pipeline instance (one warp; 40 per SM) pulls
 * 64-row chunks of the LHS table, probes level 0, compacts the matches into a
 * shared-memory chunk, runs level 1 on it once it holds enough rows for its
 * threshold, and so on (depth-first, like the reference recursion).  Thresholds
 * are expressed on the CC_CHAIN_WIDTH = 512-row scale whatever the instance's
 * own chunk width W is: it waits for ceil(threshold * W / 512) rows.
 *   h_tables[n_joins]        one table per level
 *   h_lhs_cols[n_joins]      device column pointers of the LHS table (columnar)
 *   thresholds[n_joins]      thresholds[L] = compaction threshold of the compactor that sits
 *                            after join L (main.cpp:155): 0 = no compaction (every Next result
 *                            is pushed downstream as is) .. >= CC_CHAIN_WIDTH = full compaction;
 *                            NULL = full compaction everywhere
 *   d_out_cols (optional)    HOST array of 3*n_joins device column pointers receiving the
 *                            materialised result tuples [k_0..k_{J-1}, 0,k_0, 0,k_1, ...]
 *                            (DataCollection::AppendChunk, data_collection.cpp:10-21);
 *                            NULL = count + checksums only (flag_collect_tuples=false, setting.h:31) */
#define CC_CHAIN_WIDTH 512
#define CC_MAX_JOINS 8
typedef struct {
  uint64_t n_tuples;                /* rows reaching the ResultCollector                      */
  uint64_t digest;                  /* SURVEY 8c digest                                       */
  uint64_t colsum[3 * CC_MAX_JOINS]; /* per-column wrapping sums                               */
  uint64_t level_in[CC_MAX_JOINS];  /* rows entering each level's probe                       */
  uint64_t level_steps[CC_MAX_JOINS]; /* CTA-wide probe steps executed per level ("chunks")    */
  uint64_t level_lanes[CC_MAX_JOINS]; /* sum of active lanes over those steps (density)        */
  uint64_t overflow;
  uint64_t device_ns;               /* device-side duration (globaltimer ns), bandit reward   */
  uint64_t reserved[3];             /* kernel-internal cursors (chunk cursor, first/last timestamp) */
} cc_chain_result;

int cc_chain_execute(const cc_ht *const *h_tables, size_t n_joins, const int64_t *const *h_lhs_cols, size_t n_rows,
                     const uint32_t *thresholds, int64_t *const *h_out_cols, size_t out_capacity,
                     cc_chain_result *d_result, cc_stream_t stream);

/* Chunk-density telemetry (the ZebraProfiler of profiler.h:168-260 keeps a histogram keyed by chunk size; compiled off in the
 * reference, `kEnableProfiling = 0`, :170): per join level, how full the chunks were that the level's Probe received and that
 * its Next rounds ran with, in CC_DENSITY_BINS equal bins of the pipeline instance's chunk width (bin 7 = 87.5 .. 100 %).
 * This is what a compaction threshold moves: bin 7 holds nearly everything under full compaction, the low bins fill up
 * without it.  cc_chain_execute_ex ADDS to *d_telemetry (clear it first); NULL = off.  cc_chain_telemetry_csv writes a HOST
 * copy as CSV (histogram, level, density_from, density_to, chunks).                                                         */
#define CC_DENSITY_BINS 8
typedef struct {
  uint64_t probe_rows_hist[CC_MAX_JOINS][CC_DENSITY_BINS];  /* chunks handed to a level's Probe, by fill              */
  uint64_t round_lanes_hist[CC_MAX_JOINS][CC_DENSITY_BINS]; /* Next rounds of a level, by live lanes                  */
} cc_chain_telemetry;
int cc_chain_execute_ex(const cc_ht *const *h_tables, size_t n_joins, const int64_t *const *h_lhs_cols, size_t n_rows,
                        const uint32_t *thresholds, int64_t *const *h_out_cols, size_t out_capacity,
                        cc_chain_result *d_result, cc_chain_telemetry *d_telemetry, cc_stream_t stream);
int cc_chain_telemetry_csv(const cc_chain_telemetry *h_telemetry, size_t n_joins, const char *path);

/* Dynamic ("negative feedback") compaction (main.cpp:137-167 under flag_dynamic_compact): the LHS table
 * runs through cc_chain_execute in batches of batch_rows; before each batch bandit first_bandit_id + L
 * of `tuner` selects the threshold of join L (SelectArm), afterwards every bandit is rewarded with
 * 2 / seconds / 1e3 (main.cpp:166), seconds = device time of the batch.  The batches are pipelined three
 * deep (a bandit selects with the feedback of all batches but the last two; one deep when rows are
 * materialised); the call returns when all of them are done and writes the accumulated result to the
 * HOST struct *h_result.                                                                             */
int cc_chain_execute_tuned(const cc_ht *const *h_tables, size_t n_joins, const int64_t *const *h_lhs_cols, size_t n_rows,
                           size_t batch_rows, cc_tuner *tuner, size_t first_bandit_id, int64_t *const *h_out_cols,
                           size_t out_capacity, cc_chain_result *h_result, cc_stream_t stream);

/* ------------------------------------------------------ compaction policy */
/* CompactTuner (negative_feedback.hpp:165-260) with one MultiArmedBandit
 * (:20-163) per join; FP64 arithmetic identical to the reference.            */
int cc_tuner_create(cc_tuner **t);
/* Initialize(address, arms = {0,32,64,128,256,384,512,768,1024}) (:172-177); arms NULL => default */
int cc_tuner_initialize(cc_tuner *t, size_t address, const size_t *arms, size_t n_arms);
int cc_tuner_select_arm(cc_tuner *t, size_t id, size_t *arm_value);        /* SelectArm  (:180-186) */
int cc_tuner_update_arm(cc_tuner *t, size_t id, size_t arm_value, double reward); /* UpdateArm (:189-195) */
int64_t cc_tuner_get_id(const cc_tuner *t, size_t address);                /* GetId      (:222-229) */
size_t cc_tuner_bandit_size(const cc_tuner *t);                            /* GetBanditSize (:231)  */
/* introspection for tests: est_rewards[n_arms], n_select[n_arms] of bandit id */
int cc_tuner_state(const cc_tuner *t, size_t id, double *est_rewards, uint64_t *n_select, size_t n_arms);
/* Reset(enable_log) (:197-220): dumps history CSVs under dir when enable_log, clears bandits */
int cc_tuner_reset(cc_tuner *t, int enable_log, const char *log_dir);
int cc_tuner_destroy(cc_tuner *t);

/* ------------------------------------------------------------ multi-GPU */
/* Hash partitioning for the partitioned join (new functionality):
 * partition id = murmurhash64(key) >> (64 - log2 P).  Two-pass, deterministic:
 *   cc_partition_count  : d_counts[P] = rows per partition
 *   cc_partition_scatter: writes keys grouped by partition into d_out (P contiguous
 *                         segments at d_offsets[p]); optional row ids alongside.
 * The exchange itself (NCCL all-to-all over NVLink) is done by the host layer
 * on these buffers.  log2_parts <= 9.                                          */
int cc_partition_count(const int64_t *d_keys, size_t n, int log2_parts, uint64_t *d_counts, cc_stream_t stream);
int cc_partition_scatter(const int64_t *d_keys, size_t n, int log2_parts, const uint64_t *d_offsets,
                         uint64_t *d_cursors, int64_t *d_out, cc_stream_t stream);
/* Single-pass variant (no histogram, no host round trip): partition p is scattered into the fixed region
 * d_out[p * region_capacity ...]; d_counts[p] = its rows; *d_overflow != 0 if a region overran (heavily skewed keys: the
 * output is then unusable and the two-pass pair above must be used).  *d_overflow is STICKY: the call only ever ORs into
 * it and never clears it, so one flag can watch over many calls -- the caller zeroes it (before the first call, and after
 * every look at it).  Send side of the copy-engine exchange: region p is
 * then copied into peer p's receive buffer (cc_ipc_open) with cc_memcpy_d2d, which a copy engine executes over NVLink
 * without occupying an SM.  self_part >= 0 (needs <= 16 partitions): that one partition is written to d_self_out at the
 * same region offset instead -- the rows a rank keeps go straight into its own receive buffer.                                                                                              */
int cc_partition_single(const int64_t *d_keys, size_t n, int log2_parts, size_t region_capacity, uint64_t *d_counts,
                        int *d_overflow, int64_t *d_out, int self_part, int64_t *d_self_out, cc_stream_t stream);
/* Fused scatter + exchange: partition p is written straight into h_peer_bufs[p], which may be the
 * receive buffer of ANOTHER GPU mapped through CUDA IPC (stores travel over NVLink / NVSwitch).
 * d_base[p] = first row of this rank's segment inside peer p's buffer (prefix over the senders of
 * the all-gathered count matrix).  At most 16 peers.  The caller separates the scatter from the
 * readers with a stream-ordered collective (e.g. an all-reduce), see parallel.py.               */
int cc_partition_scatter_peers(const int64_t *d_keys, size_t n, int log2_parts, const uint64_t *d_base,
                               uint64_t *d_cursors, int64_t *const *h_peer_bufs, cc_stream_t stream);
/* CTA cap of cc_partition_scatter_peers (0 = fill the GPU).  The peer scatter is NVLink-bound, so a few dozen
 * CTAs saturate the links and leave the other SMs to a probe kernel running on another stream.  */
int cc_partition_set_peer_blocks(int blocks);
/* CUDA IPC plumbing for one-process-per-GPU peers (the pointer must come from cc_malloc). */
typedef struct {
  unsigned char bytes[64];
} cc_ipc_handle;
int cc_ipc_export(void *d_ptr, cc_ipc_handle *out);
int cc_ipc_open(const cc_ipc_handle *handle, void **d_ptr);
int cc_ipc_close(void *d_ptr);
/* Block copy driven by `blocks` CTAs instead of a copy engine (16-byte loads / stores): the way to measure what the SMs can
 * push into peer memory over NVLink (tools/nvlink_bench.py); d_dst may be a cc_ipc_open mapping.                         */
int cc_peer_copy_sm(void *d_dst, const void *d_src, size_t bytes, int blocks, cc_stream_t stream);

/* ------------------------------------------- partitioned multi-GPU join (C5) */
/* The hash-partitioned join of SURVEY 8e behind the C ABI: one process per GPU, no collective library on the data path.
 * Both sides are partitioned by owner = murmurhash64(key) >> (64 - log2 world); the sub-batches of a probe call travel as
 * copy-engine block copies into CUDA-IPC-mapped peer memory (NVLink 5 / NVSwitch) and are announced / released with
 * device-side flags, so that a call only ENQUEUES work and returns (csrc/pjoin.cu).  Every rank then probes what it owns
 * with the single-GPU kernels -- per rank the semantics of LPHashTable / HashTable::Probe + Next
 * (linear_probing_ht.cpp:39-115, chaining_ht.cpp:38-136); the union of the ranks' result rows is the reference's result.
 *
 * cc_comm is the caller's CONTROL plane, used collectively and only inside cc_pjoin_create / cc_pjoin_destroy:
 *   allgather(user, send, recv, bytes): every rank contributes `bytes` bytes, recv gets world * bytes in rank order;
 *   barrier(user).   Both return 0 on success.  (MPI_Allgather / MPI_Barrier, torch.distributed, or the fork +
 *   shared-memory communicator of host/simd_compaction.hpp.)                                                          */
typedef struct cc_pjoin cc_pjoin;
typedef struct {
  int rank, world; /* world: a power of two, at most 16 (one NVLink domain) */
  int (*allgather)(void *user, const void *send, void *recv, size_t bytes);
  int (*barrier)(void *user);
  void *user;
} cc_comm;
/* Collective.  d_build_keys[n_build_local]: this rank's share of the build side (any split); max_probe_rows: the most probe
 * keys this rank will pass to one cc_pjoin_probe call; n_sub: sub-batches per call (the exchange of sub-batch b + 1 runs
 * underneath the probe of sub-batch b).  Every rank's table is sized from the GLOBAL key count (cc_ht_build_sized).       */
int cc_pjoin_create(cc_pjoin **join, const cc_comm *comm, int kind, const int64_t *d_build_keys, size_t n_build_local,
                    size_t max_probe_rows, int n_sub, cc_stream_t stream);
/* Collective (every rank calls it once per step, with its own share of the probe side; n may differ per rank).  Writes
 * the result rows THIS rank owns densely to d_out_key / d_out_payload (either may be NULL) and one cc_probe_result;
 * overflow bit 0: out_capacity too small, bit 1: an exchange or slice region overran (heavily skewed keys), bit 2: a
 * peer never delivered (bounded wait timed out -- the join is unusable afterwards).  Nothing is synchronised.           */
int cc_pjoin_probe(cc_pjoin *join, const int64_t *d_keys, size_t n, int64_t *d_out_key, int64_t *d_out_payload,
                   size_t out_capacity, cc_probe_result *d_result, cc_stream_t stream);
/* The same in two halves, for software pipelining across steps: _begin enqueues partition + copies of a batch, _end waits
 * for it and probes it.  begin(t + 1) may be called before end(t) (at most two batches in flight): the exchange of batch
 * t + 1 then runs underneath the probe of batch t, and a batch that had that much time to land is probed in one pass.   */
int cc_pjoin_probe_begin(cc_pjoin *join, const int64_t *d_keys, size_t n, cc_stream_t stream);
int cc_pjoin_probe_end(cc_pjoin *join, int64_t *d_out_key, int64_t *d_out_payload, size_t out_capacity,
                       cc_probe_result *d_result, cc_stream_t stream);
int cc_pjoin_table(const cc_pjoin *join, const cc_ht **ht); /* this rank's table (cc_ht_get_info / export) */
int cc_pjoin_destroy(cc_pjoin *join);                       /* collective */

#ifdef __cplusplus
}
#endif
#endif /* CC_API_H */
