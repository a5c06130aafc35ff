/*
 * cc_oracle.h -- CPU restatement of the reference's hash-join probe + chunk
 * compaction path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (chunk-compaction-in-vectorized-execution-simd_b200/) never links, loads
 * or calls it, and has no CPU fallback.
 *
 * Parity status: PINNED.  The restatement is checked against
 *   (1) the known-answer vectors measured from the compiled reference
 *       (7 main.cpp vectors, micro-bench #tuples), and
 *   (2) golden fixtures under tests/golden/ that were produced by the real
 *       reference classes (oracle/ref_driver.cpp links the reference's own
 *       .cpp files from /root/reference; recipe in oracle/Makefile).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference tree).
 */
#ifndef CC_ORACLE_H
#define CC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NIL 0xFFFFFFFFu

/* ---- hash_functions.h:8-16 ------------------------------------------- */
uint64_t orc_murmurhash64(uint64_t x);
void orc_murmurhash64_batch(const uint64_t *in, uint64_t *out, size_t n);

/* ---- build-side key generator: chaining_ht.cpp:15-26 ==
 *      linear_probing_ht.cpp:14-25 (payload is generated and dropped) ----- */
void orc_build_keys(size_t n, size_t chunk_factor, int64_t *out);

/* ---- probe-side generators ---------------------------------------------
 * orc_gen_lhs_main: main.cpp:41-55 -- std::mt19937 gen(2) +
 *   std::uniform_int_distribution<>(0, rhs_size) drawn row-major; restates
 *   libstdc++'s mt19937 and its 32-bit "nearly divisionless" range mapping.
 * orc_gen_keys_rand: simd_micro_bench.cpp:78-79 -- glibc rand() (default
 *   seed) & (mask); uses the host libc's rand(), exactly as the reference.
 * orc_gen_keys_counter: SURVEY 8d C4 generator (not in the reference):
 *   key_i = murmurhash64(seed + i) & mask.                                  */
void orc_gen_lhs_main(size_t rows, size_t n_joins, size_t rhs_size, int64_t *out_rowmajor);
void orc_gen_keys_rand(size_t n, uint64_t mask, int64_t *out);
void orc_gen_keys_counter(size_t n, uint64_t seed, uint64_t first, uint64_t mask, int64_t *out);

/* ---- linear-probing table: linear_probing_ht.cpp:4-37 ------------------- */
typedef struct {
  size_t n_slots; /* min pow2 >= 4n (starts at 1) */
  int64_t *slots; /* -1 == empty */
} orc_lp_table;
orc_lp_table *orc_lp_build(const int64_t *keys, size_t n);
orc_lp_table *orc_lp_build_reference(size_t n, size_t chunk_factor);
void orc_lp_free(orc_lp_table *t);

/* ---- separate-chaining table: chaining_ht.cpp:4-36 ----------------------
 * std::list<int64> buckets restated as index-linked FIFO chains:
 * node i == i-th inserted key; iterator == node index; end() == ORC_NIL.    */
typedef struct {
  size_t n_buckets; /* min pow2 >= 2n (starts at 1) */
  size_t n;
  int64_t *key;   /* [n] node key                      */
  uint32_t *next; /* [n] next node in bucket or NIL    */
  uint32_t *head; /* [n_buckets] first node or NIL     */
  uint32_t *tail; /* [n_buckets] last node (build only)*/
} orc_chain_table;
orc_chain_table *orc_chain_build(const int64_t *keys, size_t n);
orc_chain_table *orc_chain_build_reference(size_t n, size_t chunk_factor);
void orc_chain_free(orc_chain_table *t);

/* ---- DataChunk: base.h:79-100, base.cpp ---------------------------------
 * col[c] points at the storage the chunk currently references (Slice shares
 * the input's storage, base.cpp:40); own[c] is the storage it allocated.    */
typedef struct {
  size_t block; /* kBlockSize at construction */
  size_t ncol;
  size_t count;
  int64_t **col;
  int64_t **own;
  uint32_t *sel;
} orc_chunk;
orc_chunk *orc_chunk_new(size_t ncol, size_t block);
void orc_chunk_free(orc_chunk *c);
void orc_chunk_reset(orc_chunk *c);                                           /* base.h:96-99   */
void orc_chunk_slice(orc_chunk *dst, const orc_chunk *other, const uint32_t *sv, size_t count); /* base.cpp:37-47 */
void orc_chunk_append(orc_chunk *dst, const orc_chunk *src, size_t num, size_t offset);       /* base.cpp:15-27 */

/* ---- scan structures ------------------------------------------------------
 * One object for both table kinds.  Chaining: chaining_ht.h:29-84,
 * chaining_ht.cpp:38-173.  LP: linear_probing_ht.h:24-55,
 * linear_probing_ht.cpp:39-153.                                             */
typedef struct {
  int kind; /* 0 = LP, 1 = chain */
  size_t block;
  size_t count;           /* active lanes                              */
  uint32_t *lane_sel;     /* bucket_sel_vector_ / slot_sel_vector_     */
  uint64_t *pos;          /* LP: slot id; chain: node index (iterator) */
  const uint32_t *key_sel; /* reference to the caller's selection vector
                              (LP copies it: linear_probing_ht.h:48)    */
  uint32_t *key_sel_copy;
  const orc_lp_table *lp;
  const orc_chain_table *ch;
} orc_scan;
orc_scan *orc_lp_probe(const orc_lp_table *t, const int64_t *join_key, size_t count, const uint32_t *sel,
                       size_t block);
orc_scan *orc_chain_probe(const orc_chain_table *t, const int64_t *join_key, size_t count, const uint32_t *sel,
                          size_t block);
int orc_scan_has_next(const orc_scan *s);
/* Next / InOneNext: write result like the reference (Reset, Slice, gather
 * payload into result col [input.ncol + 1] at the physical LHS position).   */
size_t orc_scan_next(orc_scan *s, const int64_t *join_key, const orc_chunk *input, orc_chunk *result);
size_t orc_scan_inone_next(orc_scan *s, const int64_t *join_key, const orc_chunk *input, orc_chunk *result);
void orc_scan_free(orc_scan *s);

/* ---- compactor: compactor.cpp:5-41 with the deep-copy fix of the
 *      reference's own commented line compactor.cpp:36 (SURVEY 8c bug 3) ----
 * threshold semantics (Binary/Dynamic compactor, absent from the reference;
 * SURVEY a19): a chunk with count >= threshold passes through untouched,
 * anything smaller is buffered.  threshold == block  <=> NaiveCompactor.
 * threshold == 0 => never compact.                                          */
typedef struct {
  size_t block, ncol, threshold;
  orc_chunk *cached;
  orc_chunk *temp;
} orc_compactor;
orc_compactor *orc_compactor_new(size_t ncol, size_t block, size_t threshold);
void orc_compactor_free(orc_compactor *c);
/* in/out pointer swap exactly like unique_ptr<DataChunk>& in the reference. */
void orc_compactor_compact(orc_compactor *c, orc_chunk **chunk);
void orc_compactor_flush(orc_compactor *c, orc_chunk **chunk);

/* ---- pipeline: main.cpp:79-102, 119-191 ------------------------------------ */
typedef struct {
  uint64_t n_tuples;      /* rows reaching the ResultCollector                */
  uint64_t digest;        /* sum over tuples of th                  */
  uint64_t colsum[64];    /* per-column wrapping sums (3J columns)             */
  uint64_t probe_tuples;  /* sum over levels of rows entering Probe            */
  uint64_t level_in[16];  /* rows entering Probe per level                     */
  uint64_t level_chunks[16]; /* chunks entering Probe per level                */
  uint64_t next_calls;    /* Next() calls                                      */
} orc_result_stats;

typedef struct {
  size_t n_joins;
  size_t block;         /* kBlockSize                                         */
  int table_kind;       /* 0 LP, 1 chain (main.cpp only uses chain)           */
  int use_inone;        /* 0 Next, 1 InOneNext                                */
  int compaction;       /* 0 none, 1 full (naive), 2 threshold (binary)       */
  size_t threshold;     /* for compaction == 2                                */
  int collect;          /* materialise result tuples                          */
} orc_pipeline_cfg;

/* tables[L]: orc_lp_table* or orc_chain_table* per level.  lhs is row-major
 * rows x n_joins (DataCollection layout, data_collection.h:32).  If
 * cfg->collect, *out_tuples receives a malloc'ed row-major array of
 * n_tuples x 3*n_joins int64 (caller frees with orc_free).                  */
int orc_pipeline(const orc_pipeline_cfg *cfg, void *const *tables, const int64_t *lhs, size_t rows,
                 orc_result_stats *stats, int64_t **out_tuples);

/* Independent second oracle: multiplicity look-ups.
 * Computes count/colsums/digest of the result multiset without running the
 * chunked pipeline.  build_keys[L]/n_build[L] per level.                    */
int orc_multiplicity_oracle(size_t n_joins, const int64_t *const *build_keys, const size_t *n_build,
                            const int64_t *lhs, size_t rows, orc_result_stats *stats);

/* digest of explicit tuples (row-major, ncol columns). */
uint64_t orc_digest_tuples(const int64_t *tuples, size_t n, size_t ncol, uint64_t *colsum /* [ncol] or NULL */);

/* single-join batch helper used by bench cpu baseline: scalar Probe+Next over
 * n keys in chunks of `block`; returns #tuples (simd_micro_bench.cpp loop).
 * checksum (optional) = wrapping sum of matched payloads.                   */
uint64_t orc_microbench_lp(const orc_lp_table *t, const int64_t *keys, size_t n, size_t block, int inone,
                           uint64_t *checksum);
uint64_t orc_microbench_chain(const orc_chain_table *t, const int64_t *keys, size_t n, size_t block, int inone,
                              uint64_t *checksum);

/* ---- payload-keeping single join ------------------------------
 * The reference generates `payload = cnt + 10000000` for build row cnt and pushes
 * only tuple[0] into its tables (chaining_ht.cpp:21-23,34;
 * linear_probing_ht.cpp:20-22,33), so its results never carry a payload.  This is
 * the same build + probe restated with the build ROW kept beside every key: a
 * match with table entry e yields (probe key, key of e, payload[c][row of e]).
 * kind 0: LP insertion (linear_probing_ht.cpp:27-36) and the full walk to the first
 * empty slot (linear_probing_ht.cpp:62-115); kind 1: FIFO chains
 * (chaining_ht.cpp:28-35) walked to their end (chaining_ht.cpp:60-124).
 * Parity status of the PAYLOAD VALUES: unpinned by construction -- the reference
 * cannot produce them; keys, match counts and the row each match comes from follow
 * the pinned table restatements above, and tests cross-check against an independent
 * sort-merge join.
 * out_rows: malloc'ed row-major n_rows x (2 + n_cols) int64 (caller: orc_free).  */
void orc_ref_payload(size_t n, int64_t *out); /* out[i] = i + 10000000 */
int orc_join_payload(int kind, const int64_t *build_keys, size_t n_build, const int64_t *const *payload_cols, size_t n_cols,
                     const int64_t *probe_keys, size_t n_probe, int64_t **out_rows, size_t *n_rows);

/* ---- bandit: negative_feedback.hpp:20-163 (MultiArmedBandit) and
 *      :165-260 (CompactTuner, one bandit per join) ---------------------- */
typedef struct orc_bandit orc_bandit;
orc_bandit *orc_bandit_new(size_t n_arms);
void orc_bandit_free(orc_bandit *b);
size_t orc_bandit_select(orc_bandit *b);
void orc_bandit_update(orc_bandit *b, size_t arm, double reward);
void orc_bandit_state(const orc_bandit *b, double *est_rewards, uint64_t *n_select);

void orc_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
