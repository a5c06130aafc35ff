/*
 * cc_oracle.c -- CPU restatement of the reference's hash-join probe + chunk
 * compaction path (plain C, scalar, single-threaded like the reference).
 *
 * TEST INFRASTRUCTURE ONLY -- see cc_oracle.h for who may load this and for
 * the parity-pinning status.  Citations are file:line in the reference tree.
 */
#include "cc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* hash_functions.h:8-16                                                     */
uint64_t orc_murmurhash64(uint64_t x) {
  x ^= x >> 32;
  x *= 0xd6e8feb86659fd93ULL;
  x ^= x >> 32;
  x *= 0xd6e8feb86659fd93ULL;
  x ^= x >> 32;
  return x;
}

void orc_murmurhash64_batch(const uint64_t *in, uint64_t *out, size_t n) {
  for (size_t i = 0; i < n; ++i) out[i] = orc_murmurhash64(in[i]);
}

/* ------------------------------------------------------------------------ */
/* the join with the payload the reference generates and drops.   */
void orc_ref_payload(size_t n, int64_t *out) {
  for (size_t i = 0; i < n; ++i) out[i] = (int64_t)(i + 10000000); /* chaining_ht.cpp:21 */
}

typedef struct {
  int64_t *rows;
  size_t n, cap, width;
} orc_rowbuf;

static int64_t *rowbuf_next(orc_rowbuf *b) {
  if (b->n == b->cap) {
    b->cap = b->cap ? b->cap * 2 : 1024;
    b->rows = (int64_t *)realloc(b->rows, b->cap * b->width * sizeof(int64_t));
  }
  return b->rows + (b->n++) * b->width;
}

int orc_join_payload(int kind, const int64_t *build_keys, size_t n_build, const int64_t *const *payload_cols, size_t n_cols,
                     const int64_t *probe_keys, size_t n_probe, int64_t **out_rows, size_t *n_rows) {
  if (!out_rows || !n_rows || (kind != 0 && kind != 1)) return -1;
  orc_rowbuf b = {NULL, 0, 0, 2 + n_cols};
  if (kind == 0) {
    /* linear_probing_ht.cpp:4-37, with the build row stored beside the key */
    size_t ns = 1;
    while (ns < (n_build << 2)) ns <<= 1;
    const uint64_t mask = ns - 1;
    int64_t *slots = (int64_t *)malloc(ns * sizeof(int64_t));
    uint32_t *row = (uint32_t *)malloc(ns * sizeof(uint32_t));
    for (size_t i = 0; i < ns; ++i) slots[i] = -1;
    for (size_t i = 0; i < n_build; ++i) {
      uint64_t s = orc_murmurhash64((uint64_t)build_keys[i]) & mask;
      while (slots[s] != -1) s = (s + 1) & mask;
      slots[s] = build_keys[i];
      row[s] = (uint32_t)i;
    }
    /* Probe + Next until the lane retires at an empty slot (linear_probing_ht.cpp:39-115) */
    for (size_t i = 0; i < n_probe; ++i) {
      const int64_t k = probe_keys[i];
      uint64_t s = orc_murmurhash64((uint64_t)k) & mask;
      while (slots[s] != -1) {
        if (slots[s] == k) {
          int64_t *r = rowbuf_next(&b);
          r[0] = k;
          r[1] = slots[s];
          for (size_t c = 0; c < n_cols; ++c) r[2 + c] = payload_cols[c][row[s]];
        }
        s = (s + 1) & mask;
      }
    }
    free(slots);
    free(row);
  } else {
    /* chaining_ht.cpp:4-36: node i == build row i, so the node index IS the payload row */
    orc_chain_table *t = orc_chain_build(build_keys, n_build);
    const uint64_t mask = t->n_buckets - 1;
    for (size_t i = 0; i < n_probe; ++i) {
      const int64_t k = probe_keys[i];
      for (uint32_t e = t->head[orc_murmurhash64((uint64_t)k) & mask]; e != ORC_NIL; e = t->next[e]) {
        if (t->key[e] == k) {
          int64_t *r = rowbuf_next(&b);
          r[0] = k;
          r[1] = t->key[e];
          for (size_t c = 0; c < n_cols; ++c) r[2 + c] = payload_cols[c][e];
        }
      }
    }
    orc_chain_free(t);
  }
  *out_rows = b.rows;
  *n_rows = b.n;
  return 0;
}

void orc_free(void *p) { free(p); }

/* ------------------------------------------------------------------------ */
/* chaining_ht.cpp:15-26 == linear_probing_ht.cpp:14-25.  Note the integer
 * division n / num_unique (cf=3, n=2e6 gives step 2).                       */
void orc_build_keys(size_t n, size_t cf, int64_t *out) {
  if (n == 0) return;
  size_t num_unique = n / cf + (n % cf != 0);
  size_t step = n / num_unique;
  size_t cnt = 0;
  for (size_t i = 0; i < num_unique; ++i) {
    for (size_t j = 0; j < cf && cnt < n; ++j) out[cnt++] = (int64_t)(i * step);
  }
}

/* ------------------------------------------------------------------------ */
/* std::mt19937 (32-bit Mersenne twister, standard parameters).              */
typedef struct {
  uint32_t mt[624];
  int idx;
} mt19937_t;

static void mt_seed(mt19937_t *g, uint32_t seed) {
  g->mt[0] = seed;
  for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}

static uint32_t mt_next(mt19937_t *g) {
  if (g->idx >= 624) {
    for (int i = 0; i < 624; ++i) {
      uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
      g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

/* libstdc++ (GCC >= 11) uniform_int_distribution<int>(0, hi) over a 32-bit
 * URBG: Lemire's nearly-divisionless mapping (bits/uniform_int_dist.h _S_nd).*/
static uint32_t uniform_u32(mt19937_t *g, uint32_t range /* hi + 1, != 0 */) {
  uint64_t product = (uint64_t)mt_next(g) * (uint64_t)range;
  uint32_t low = (uint32_t)product;
  if (low < range) {
    uint32_t threshold = (uint32_t)(-range) % range;
    while (low < threshold) {
      product = (uint64_t)mt_next(g) * (uint64_t)range;
      low = (uint32_t)product;
    }
  }
  return (uint32_t)(product >> 32);
}

/* main.cpp:41-55: gen(2); dist(0, kRHSTupleSize) inclusive; row-major draws. */
void orc_gen_lhs_main(size_t rows, size_t n_joins, size_t rhs_size, int64_t *out) {
  mt19937_t g;
  mt_seed(&g, 2u);
  uint32_t range = (uint32_t)rhs_size + 1u;
  for (size_t i = 0; i < rows; ++i)
    for (size_t j = 0; j < n_joins; ++j) {
      /* range == 0 would mean the full 2^32 span: libstdc++ returns g() */
      out[i * n_joins + j] = range ? (int64_t)uniform_u32(&g, range) : (int64_t)mt_next(&g);
    }
}

/* simd_micro_bench.cpp:78-79: keys[i] = rand() & (kRHSTuples*kHitFreq - 1),
 * glibc rand() with its default seed (the reference never calls srand).     */
void orc_gen_keys_rand(size_t n, uint64_t mask, int64_t *out) {
  srand(1); /* default seed == srand(1), makes the call repeatable */
  for (size_t i = 0; i < n; ++i) out[i] = (int64_t)((uint64_t)rand() & mask);
}

/* SURVEY 8d (C4/C5): counter-based generator, not in the reference.         */
void orc_gen_keys_counter(size_t n, uint64_t seed, uint64_t first, uint64_t mask, int64_t *out) {
  for (size_t i = 0; i < n; ++i) out[i] = (int64_t)(orc_murmurhash64(seed + first + i) & mask);
}

/* ------------------------------------------------------------------------ */
/* linear_probing_ht.cpp:4-37                                                */
orc_lp_table *orc_lp_build(const int64_t *keys, size_t n) {
  orc_lp_table *t = (orc_lp_table *)calloc(1, sizeof(*t));
  size_t ns = 1;
  while (ns < (n << 2)) ns <<= 1;
  t->n_slots = ns;
  t->slots = (int64_t *)malloc(ns * sizeof(int64_t));
  for (size_t i = 0; i < ns; ++i) t->slots[i] = -1;
  uint64_t mask = ns - 1;
  for (size_t i = 0; i < n; ++i) {
    uint64_t s = orc_murmurhash64((uint64_t)keys[i]) & mask;
    while (t->slots[s] != -1) s = (s + 1) & mask;
    t->slots[s] = keys[i];
  }
  return t;
}

orc_lp_table *orc_lp_build_reference(size_t n, size_t cf) {
  int64_t *k = (int64_t *)malloc((n ? n : 1) * sizeof(int64_t));
  orc_build_keys(n, cf, k);
  orc_lp_table *t = orc_lp_build(k, n);
  free(k);
  return t;
}

void orc_lp_free(orc_lp_table *t) {
  if (!t) return;
  free(t->slots);
  free(t);
}

/* chaining_ht.cpp:4-36: push_back => FIFO chains.                           */
orc_chain_table *orc_chain_build(const int64_t *keys, size_t n) {
  orc_chain_table *t = (orc_chain_table *)calloc(1, sizeof(*t));
  size_t nb = 1;
  while (nb < 2 * n) nb *= 2;
  t->n_buckets = nb;
  t->n = n;
  t->key = (int64_t *)malloc((n ? n : 1) * sizeof(int64_t));
  t->next = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
  t->head = (uint32_t *)malloc(nb * sizeof(uint32_t));
  t->tail = (uint32_t *)malloc(nb * sizeof(uint32_t));
  for (size_t b = 0; b < nb; ++b) t->head[b] = t->tail[b] = ORC_NIL;
  uint64_t mask = nb - 1;
  for (size_t i = 0; i < n; ++i) {
    uint64_t b = orc_murmurhash64((uint64_t)keys[i]) & mask;
    t->key[i] = keys[i];
    t->next[i] = ORC_NIL;
    if (t->head[b] == ORC_NIL)
      t->head[b] = (uint32_t)i;
    else
      t->next[t->tail[b]] = (uint32_t)i;
    t->tail[b] = (uint32_t)i;
  }
  return t;
}

orc_chain_table *orc_chain_build_reference(size_t n, size_t cf) {
  int64_t *k = (int64_t *)malloc((n ? n : 1) * sizeof(int64_t));
  orc_build_keys(n, cf, k);
  orc_chain_table *t = orc_chain_build(k, n);
  free(k);
  return t;
}

void orc_chain_free(orc_chain_table *t) {
  if (!t) return;
  free(t->key);
  free(t->next);
  free(t->head);
  free(t->tail);
  free(t);
}

/* ------------------------------------------------------------------------ */
/* DataChunk: base.cpp:10-13                                                 */
orc_chunk *orc_chunk_new(size_t ncol, size_t block) {
  orc_chunk *c = (orc_chunk *)calloc(1, sizeof(*c));
  c->block = block;
  c->ncol = ncol;
  c->count = 0;
  c->col = (int64_t **)calloc(ncol ? ncol : 1, sizeof(int64_t *));
  c->own = (int64_t **)calloc(ncol ? ncol : 1, sizeof(int64_t *));
  for (size_t i = 0; i < ncol; ++i) c->col[i] = c->own[i] = (int64_t *)calloc(block ? block : 1, sizeof(int64_t));
  c->sel = (uint32_t *)malloc((block ? block : 1) * sizeof(uint32_t));
  for (size_t i = 0; i < block; ++i) c->sel[i] = (uint32_t)i;
  return c;
}

void orc_chunk_free(orc_chunk *c) {
  if (!c) return;
  for (size_t i = 0; i < c->ncol; ++i) free(c->own[i]);
  free(c->col);
  free(c->own);
  free(c->sel);
  free(c);
}

/* base.h:96-99 */
void orc_chunk_reset(orc_chunk *c) {
  c->count = 0;
  for (size_t i = 0; i < c->block; ++i) c->sel[i] = (uint32_t)i;
}

/* base.cpp:37-47: zero-copy column sharing + selection-vector composition. */
void orc_chunk_slice(orc_chunk *dst, const orc_chunk *other, const uint32_t *sv, size_t count) {
  dst->count = count;
  for (size_t c = 0; c < other->ncol; ++c) dst->col[c] = other->col[c];
  for (size_t i = 0; i < count; ++i) dst->sel[i] = other->sel[sv[i]];
}

/* base.cpp:15-27: gather through src's selection vector into dst's dense tail.
 * Writes through dst->col[] (whatever storage dst currently references).     */
void orc_chunk_append(orc_chunk *dst, const orc_chunk *src, size_t num, size_t offset) {
  for (size_t i = 0; i < dst->ncol; ++i)
    for (size_t j = 0; j < num; ++j) dst->col[i][dst->count + j] = src->col[i][src->sel[j + offset]];
  dst->count += num;
}

/* ------------------------------------------------------------------------ */
static orc_scan *scan_alloc(int kind, size_t block) {
  orc_scan *s = (orc_scan *)calloc(1, sizeof(*s));
  s->kind = kind;
  s->block = block;
  s->lane_sel = (uint32_t *)calloc(block ? block : 1, sizeof(uint32_t));
  s->pos = (uint64_t *)calloc(block ? block : 1, sizeof(uint64_t));
  return s;
}

/* linear_probing_ht.cpp:39-60 */
orc_scan *orc_lp_probe(const orc_lp_table *t, const int64_t *join_key, size_t count, const uint32_t *sel,
                       size_t block) {
  orc_scan *s = scan_alloc(0, block);
  s->lp = t;
  uint64_t mask = t->n_slots - 1;
  for (size_t i = 0; i < count; ++i) s->pos[i] = orc_murmurhash64((uint64_t)join_key[sel[i]]) & mask;
  size_t valid = 0;
  for (size_t i = 0; i < count; ++i)
    if (t->slots[s->pos[i]] != -1) s->lane_sel[valid++] = (uint32_t)i;
  s->count = valid;
  /* LPScanStructure keeps a COPY of the key selection vector
   * (linear_probing_ht.h:48 is a value member).                             */
  s->key_sel_copy = (uint32_t *)malloc((block ? block : 1) * sizeof(uint32_t));
  memcpy(s->key_sel_copy, sel, block * sizeof(uint32_t));
  s->key_sel = s->key_sel_copy;
  return s;
}

/* chaining_ht.cpp:38-58 + ScanStructure ctor chaining_ht.h:31-42 */
orc_scan *orc_chain_probe(const orc_chain_table *t, const int64_t *join_key, size_t count, const uint32_t *sel,
                          size_t block) {
  orc_scan *s = scan_alloc(1, block);
  s->ch = t;
  uint64_t mask = t->n_buckets - 1;
  size_t n_non_empty = 0;
  for (size_t i = 0; i < count; ++i) {
    uint64_t b = orc_murmurhash64((uint64_t)join_key[sel[i]]) & mask;
    s->pos[i] = t->head[b]; /* iterator = begin(); NIL if the bucket is empty */
  }
  for (size_t i = 0; i < count; ++i)
    if (s->pos[i] != ORC_NIL) s->lane_sel[n_non_empty++] = (uint32_t)i;
  s->count = n_non_empty;
  s->key_sel = sel; /* ScanStructure holds a reference (chaining_ht.h:54) */
  return s;
}

int orc_scan_has_next(const orc_scan *s) { return s->count > 0; }

void orc_scan_free(orc_scan *s) {
  if (!s) return;
  free(s->lane_sel);
  free(s->pos);
  free(s->key_sel_copy);
  free(s);
}

/* chaining_ht.cpp:109-124 */
static void chain_advance(orc_scan *s) {
  size_t new_count = 0;
  const orc_chain_table *t = s->ch;
  for (size_t i = 0; i < s->count; ++i) {
    uint32_t idx = s->lane_sel[i];
    s->lane_sel[new_count] = idx;
    s->pos[idx] = t->next[s->pos[idx]];
    new_count += (s->pos[idx] != ORC_NIL);
  }
  s->count = new_count;
}

/* chaining_ht.cpp:82-107 */
static size_t chain_scan_inner_join(orc_scan *s, const int64_t *join_key, uint32_t *result_vector) {
  const orc_chain_table *t = s->ch;
  for (;;) {
    size_t rc = 0;
    for (size_t i = 0; i < s->count; ++i) {
      uint32_t idx = s->lane_sel[i];
      int64_t l_key = join_key[s->key_sel[idx]];
      int64_t r_key = t->key[s->pos[idx]];
      result_vector[rc] = idx;
      rc += (l_key == r_key);
    }
    if (rc > 0) return rc;
    chain_advance(s);
    if (s->count == 0) return 0;
  }
}

/* chaining_ht.cpp:60-80 (chain), linear_probing_ht.cpp:62-115 (LP) */
size_t orc_scan_next(orc_scan *s, const int64_t *join_key, const orc_chunk *input, orc_chunk *result) {
  orc_chunk_reset(result);
  uint32_t *rv = (uint32_t *)malloc((s->block ? s->block : 1) * sizeof(uint32_t));
  size_t rc = 0;
  if (s->kind == 1) {
    if (s->count == 0) {
      free(rv);
      return 0;
    }
    rc = chain_scan_inner_join(s, join_key, rv);
    if (rc > 0) {
      orc_chunk_slice(result, input, rv, rc);
      int64_t *col1 = result->col[input->ncol + 1];
      for (size_t i = 0; i < rc; ++i) { /* GatherResult chaining_ht.cpp:126-136 */
        uint32_t idx = rv[i];
        col1[s->key_sel[idx]] = s->ch->key[s->pos[idx]];
      }
    }
    chain_advance(s);
  } else {
    const orc_lp_table *t = s->lp;
    uint64_t mask = t->n_slots - 1;
    for (size_t i = 0; i < s->count; ++i) { /* match :72-80 */
      uint32_t idx = s->lane_sel[i];
      rv[rc] = idx;
      rc += (join_key[s->key_sel[idx]] == t->slots[s->pos[idx]]);
    }
    orc_chunk_slice(result, input, rv, rc); /* :85 (also for rc == 0) */
    int64_t *col1 = result->col[input->ncol + 1];
    for (size_t i = 0; i < rc; ++i) { /* gather :90-94 */
      uint32_t idx = rv[i];
      col1[s->key_sel[idx]] = t->slots[s->pos[idx]];
    }
    size_t new_count = 0; /* advance :100-110 */
    for (size_t i = 0; i < s->count; ++i) {
      uint32_t idx = s->lane_sel[i];
      uint64_t id = (s->pos[idx] + 1) & mask;
      s->pos[idx] = id;
      s->lane_sel[new_count] = idx;
      new_count += (t->slots[id] != -1);
    }
    s->count = new_count;
  }
  free(rv);
  return rc;
}

/* chaining_ht.cpp:138-173 (chain), linear_probing_ht.cpp:117-153 (LP):
 * fused match+gather+advance; payload written for ALL active lanes.        */
size_t orc_scan_inone_next(orc_scan *s, const int64_t *join_key, const orc_chunk *input, orc_chunk *result) {
  orc_chunk_reset(result);
  uint32_t *rv = (uint32_t *)malloc((s->block ? s->block : 1) * sizeof(uint32_t));
  size_t rc = 0, new_count = 0;
  /* cols[] are taken BEFORE Slice in the reference; result's RHS columns are
   * never re-pointed by Slice, so the pointer is the same either way.       */
  int64_t *col1 = result->col[input->ncol + 1];
  if (s->kind == 1) {
    const orc_chain_table *t = s->ch;
    for (size_t i = 0; i < s->count; ++i) {
      uint32_t idx = s->lane_sel[i];
      int64_t l_key = join_key[s->key_sel[idx]];
      int64_t r_key = t->key[s->pos[idx]];
      col1[s->key_sel[idx]] = r_key;
      rv[rc] = idx;
      rc += (l_key == r_key);
      s->lane_sel[new_count] = idx;
      s->pos[idx] = t->next[s->pos[idx]];
      new_count += (s->pos[idx] != ORC_NIL);
    }
  } else {
    const orc_lp_table *t = s->lp;
    uint64_t mask = t->n_slots - 1;
    for (size_t i = 0; i < s->count; ++i) {
      uint32_t idx = s->lane_sel[i];
      int64_t l_key = join_key[s->key_sel[idx]];
      int64_t r_key = t->slots[s->pos[idx]];
      col1[s->key_sel[idx]] = r_key;
      rv[rc] = idx;
      rc += (l_key == r_key);
      uint64_t id = (s->pos[idx] + 1) & mask;
      s->pos[idx] = id;
      s->lane_sel[new_count] = idx;
      new_count += (t->slots[id] != -1);
    }
  }
  orc_chunk_slice(result, input, rv, rc);
  s->count = new_count;
  free(rv);
  return rc;
}

/* ------------------------------------------------------------------------ */
/* compactor.cpp:5-41 with compactor.cpp:36 (fresh temp chunk) enabled.      */
orc_compactor *orc_compactor_new(size_t ncol, size_t block, size_t threshold) {
  orc_compactor *c = (orc_compactor *)calloc(1, sizeof(*c));
  c->block = block;
  c->ncol = ncol;
  c->threshold = threshold;
  c->cached = orc_chunk_new(ncol, block);
  c->temp = orc_chunk_new(ncol, block);
  return c;
}

void orc_compactor_free(orc_compactor *c) {
  if (!c) return;
  orc_chunk_free(c->cached);
  orc_chunk_free(c->temp);
  free(c);
}

void orc_compactor_compact(orc_compactor *c, orc_chunk **chunk) {
  orc_chunk *ch = *chunk;
  if (ch->count == c->block) return;              /* compactor.cpp:6 */
  if (ch->count >= c->threshold) return;          /* Binary/Dynamic; never true for naive */
  if (ch->count <= c->block - c->cached->count) { /* :12-19 */
    orc_chunk_append(c->cached, ch, ch->count, 0);
    orc_chunk_reset(ch);
    return;
  }
  size_t n_move = c->block - c->cached->count; /* :22-24 */
  orc_chunk_append(c->cached, ch, n_move, 0);
  orc_chunk_append(c->temp, ch, ch->count - n_move, n_move);
  /* :33-36 swap; then a FRESH temp chunk (the reference's commented fix)   */
  orc_chunk *old_chunk = ch;
  *chunk = c->cached;
  c->cached = c->temp;
  orc_chunk_free(old_chunk);
  c->temp = orc_chunk_new(c->ncol, c->block);
}

/* compactor.h:23: Flush moves the cache out; the compactor is dead after.  */
void orc_compactor_flush(orc_compactor *c, orc_chunk **chunk) {
  orc_chunk_free(*chunk);
  *chunk = c->cached;
  c->cached = NULL;
}

/* ------------------------------------------------------------------------ */
uint64_t orc_digest_tuples(const int64_t *tuples, size_t n, size_t ncol, uint64_t *colsum) {
  uint64_t h = 0;
  if (colsum) memset(colsum, 0, ncol * sizeof(uint64_t));
  for (size_t i = 0; i < n; ++i) {
    uint64_t th = 0x9e3779b97f4a7c15ULL;
    for (size_t j = 0; j < ncol; ++j) {
      uint64_t v = (uint64_t)tuples[i * ncol + j];
      th = orc_murmurhash64(th ^ v) + j;
      if (colsum) colsum[j] += v;
    }
    h += th;
  }
  return h;
}

typedef struct {
  const orc_pipeline_cfg *cfg;
  void *const *tables;
  orc_chunk **intermediates;
  orc_compactor **compactors;
  orc_result_stats *st;
  int64_t *tuples;
  size_t cap;
  size_t ncol_out;
} pipe_state;

/* ResultCollector: main.cpp:125-128 + data_collection.cpp:10-21 */
static void collect(pipe_state *ps, const orc_chunk *in) {
  size_t nc = ps->ncol_out;
  orc_result_stats *st = ps->st;
  for (size_t i = 0; i < in->count; ++i) {
    uint32_t idx = in->sel[i];
    uint64_t th = 0x9e3779b97f4a7c15ULL;
    for (size_t j = 0; j < nc; ++j) {
      uint64_t v = (uint64_t)in->col[j][idx];
      th = orc_murmurhash64(th ^ v) + j;
      st->colsum[j] += v;
    }
    st->digest += th;
    if (ps->cfg->collect) {
      if (st->n_tuples + 1 > ps->cap) {
        ps->cap = ps->cap ? ps->cap * 2 : 4096;
        ps->tuples = (int64_t *)realloc(ps->tuples, ps->cap * nc * sizeof(int64_t));
      }
      for (size_t j = 0; j < nc; ++j) ps->tuples[st->n_tuples * nc + j] = in->col[j][idx];
    }
    st->n_tuples++;
  }
}

/* main.cpp:119-170 */
static void execute_pipeline(pipe_state *ps, orc_chunk *input, size_t level) {
  const orc_pipeline_cfg *cfg = ps->cfg;
  if (level == cfg->n_joins) {
    collect(ps, input);
    return;
  }
  const int64_t *join_key = input->col[level];
  orc_result_stats *st = ps->st;
  st->level_in[level] += input->count;
  st->level_chunks[level] += 1;
  st->probe_tuples += input->count;
  orc_scan *ss = cfg->table_kind == 0
                     ? orc_lp_probe((const orc_lp_table *)ps->tables[level], join_key, input->count, input->sel, cfg->block)
                     : orc_chain_probe((const orc_chain_table *)ps->tables[level], join_key, input->count, input->sel,
                                       cfg->block);
  while (orc_scan_has_next(ss)) {
    orc_chunk **result = &ps->intermediates[level];
    st->next_calls++;
    if (cfg->use_inone)
      orc_scan_inone_next(ss, join_key, input, *result);
    else
      orc_scan_next(ss, join_key, input, *result);
    if (cfg->compaction) { /* main.cpp:153-157 */
      orc_compactor_compact(ps->compactors[level], result);
      if ((*result)->count == 0) continue;
    }
    execute_pipeline(ps, *result, level + 1);
  }
  orc_scan_free(ss);
}

/* main.cpp:172-191 */
static void flush_pipeline_cache(pipe_state *ps, size_t level) {
  if (level == ps->cfg->n_joins) return;
  orc_compactor_flush(ps->compactors[level], &ps->intermediates[level]);
  execute_pipeline(ps, ps->intermediates[level], level + 1);
  flush_pipeline_cache(ps, level + 1);
}

int orc_pipeline(const orc_pipeline_cfg *cfg, void *const *tables, const int64_t *lhs, size_t rows,
                 orc_result_stats *stats, int64_t **out_tuples) {
  size_t J = cfg->n_joins, B = cfg->block;
  if (J == 0 || J > 16 || B == 0) return -1;
  memset(stats, 0, sizeof(*stats));
  pipe_state ps;
  memset(&ps, 0, sizeof(ps));
  ps.cfg = cfg;
  ps.tables = tables;
  ps.st = stats;
  ps.ncol_out = 3 * J;
  ps.intermediates = (orc_chunk **)calloc(J, sizeof(orc_chunk *));
  ps.compactors = (orc_compactor **)calloc(J, sizeof(orc_compactor *));
  for (size_t i = 0; i < J; ++i) { /* main.cpp:62-68 */
    size_t ncol = J + 2 * (i + 1);
    ps.intermediates[i] = orc_chunk_new(ncol, B);
    if (cfg->compaction) ps.compactors[i] = orc_compactor_new(ncol, B, cfg->compaction == 1 ? B : cfg->threshold);
  }
  /* main.cpp:81-95: FetchChunk transposes rows [start,end) into a fresh chunk */
  size_t start = 0, end;
  do {
    end = start + B < rows ? start + B : rows;
    orc_chunk *chunk = orc_chunk_new(J, B);
    for (size_t i = start; i < end; ++i) { /* data_collection.cpp:23-27, base.cpp:29-35 */
      for (size_t j = 0; j < J; ++j) chunk->col[j][chunk->count] = lhs[i * J + j];
      chunk->count++;
    }
    start = end;
    execute_pipeline(&ps, chunk, 0);
    orc_chunk_free(chunk);
  } while (end < rows);
  if (cfg->compaction) flush_pipeline_cache(&ps, 0);
  for (size_t i = 0; i < J; ++i) {
    orc_chunk_free(ps.intermediates[i]);
    if (ps.compactors[i]) orc_compactor_free(ps.compactors[i]);
  }
  free(ps.intermediates);
  free(ps.compactors);
  if (out_tuples)
    *out_tuples = ps.tuples;
  else
    free(ps.tuples);
  return 0;
}

/* ------------------------------------------------------------------------ */
/* the result multiset is, per LHS row r, the tuple
 * [T[r][0..J), 0, T[r][0], 0, T[r][1], ...] with multiplicity
 * prod_L mult_L(T[r][L]).  Multiplicities via a private open-addressing
 * counter table (independent of the LP/chain code above).                   */
typedef struct {
  size_t cap;
  int64_t *k;
  uint32_t *c;
} mult_map;

static void mm_build(mult_map *m, const int64_t *keys, size_t n) {
  size_t cap = 16;
  while (cap < 2 * n + 1) cap <<= 1;
  m->cap = cap;
  m->k = (int64_t *)malloc(cap * sizeof(int64_t));
  m->c = (uint32_t *)calloc(cap, sizeof(uint32_t));
  for (size_t i = 0; i < n; ++i) {
    size_t s = (size_t)(((uint64_t)keys[i] * 0x9E3779B97F4A7C15ULL) >> 20) & (cap - 1);
    while (m->c[s] && m->k[s] != keys[i]) s = (s + 1) & (cap - 1);
    m->k[s] = keys[i];
    m->c[s]++;
  }
}

static uint32_t mm_get(const mult_map *m, int64_t key) {
  size_t s = (size_t)(((uint64_t)key * 0x9E3779B97F4A7C15ULL) >> 20) & (m->cap - 1);
  while (m->c[s]) {
    if (m->k[s] == key) return m->c[s];
    s = (s + 1) & (m->cap - 1);
  }
  return 0;
}

int orc_multiplicity_oracle(size_t J, const int64_t *const *build_keys, const size_t *n_build, const int64_t *lhs,
                            size_t rows, orc_result_stats *st) {
  if (J == 0 || J > 16) return -1;
  memset(st, 0, sizeof(*st));
  mult_map *maps = (mult_map *)calloc(J, sizeof(mult_map));
  for (size_t l = 0; l < J; ++l) mm_build(&maps[l], build_keys[l], n_build[l]);
  size_t nc = 3 * J;
  for (size_t r = 0; r < rows; ++r) {
    uint64_t mult = 1;
    for (size_t l = 0; l < J; ++l) {
      st->level_in[l] += mult;
      st->probe_tuples += mult;
      mult *= mm_get(&maps[l], lhs[r * J + l]);
      if (!mult) break;
    }
    if (!mult) continue;
    uint64_t th = 0x9e3779b97f4a7c15ULL;
    for (size_t j = 0; j < nc; ++j) {
      uint64_t v;
      if (j < J)
        v = (uint64_t)lhs[r * J + j];
      else
        v = ((j - J) & 1) ? (uint64_t)lhs[r * J + (j - J) / 2] : 0;
      th = orc_murmurhash64(th ^ v) + j;
      st->colsum[j] += v * mult;
    }
    st->digest += th * mult;
    st->n_tuples += mult;
  }
  for (size_t l = 0; l < J; ++l) {
    free(maps[l].k);
    free(maps[l].c);
  }
  free(maps);
  return 0;
}

/* ------------------------------------------------------------------------ */
/* simd_micro_bench.cpp:226-256 (LP scalar) / :122-152 (chain scalar) loops. */
static uint64_t microbench(int kind, const void *table, const int64_t *keys, size_t n, size_t block, int inone,
                           uint64_t *checksum) {
  orc_chunk *input = orc_chunk_new(1, block);
  orc_chunk *output = orc_chunk_new(3, block);
  uint32_t *sel = (uint32_t *)malloc(block * sizeof(uint32_t));
  for (size_t i = 0; i < block; ++i) sel[i] = (uint32_t)i;
  uint64_t n_tuples = 0, sum = 0;
  for (size_t k = 0; k < n; k += block) {
    size_t fill = block < n - k ? block : n - k;
    memcpy(input->col[0], keys + k, fill * sizeof(int64_t)); /* :95 load one block */
    input->count = fill;
    orc_scan *ss = kind == 0 ? orc_lp_probe((const orc_lp_table *)table, input->col[0], fill, sel, block)
                             : orc_chain_probe((const orc_chain_table *)table, input->col[0], fill, sel, block);
    while (orc_scan_has_next(ss)) {
      size_t rc = inone ? orc_scan_inone_next(ss, input->col[0], input, output)
                        : orc_scan_next(ss, input->col[0], input, output);
      n_tuples += rc;
      if (checksum)
        for (size_t i = 0; i < rc; ++i) sum += (uint64_t)output->col[2][output->sel[i]];
    }
    orc_scan_free(ss);
  }
  if (checksum) *checksum = sum;
  free(sel);
  orc_chunk_free(input);
  orc_chunk_free(output);
  return n_tuples;
}

uint64_t orc_microbench_lp(const orc_lp_table *t, const int64_t *keys, size_t n, size_t block, int inone,
                           uint64_t *checksum) {
  return microbench(0, t, keys, n, block, inone, checksum);
}

uint64_t orc_microbench_chain(const orc_chain_table *t, const int64_t *keys, size_t n, size_t block, int inone,
                              uint64_t *checksum) {
  return microbench(1, t, keys, n, block, inone, checksum);
}

/* ------------------------------------------------------------------------ */
/* negative_feedback.hpp:20-163                                              */
struct orc_bandit {
  size_t k_arms;
  double k_epsilon;        /* :133 */
  size_t k_start_sampling; /* :134 */
  size_t k_heart;          /* :154 */
  size_t select_times;
  size_t *n_select;
  double *est_rewards;
  double *est_square_rewards;
  size_t stage_update_times;
  size_t *stage_n_update;
  size_t n_start_sampling;
  double *r_means;
  int r_means_set;
  size_t history_len;
};

orc_bandit *orc_bandit_new(size_t n_arms) {
  orc_bandit *b = (orc_bandit *)calloc(1, sizeof(*b));
  b->k_arms = n_arms;
  b->k_epsilon = 0.1;
  b->k_start_sampling = 4;
  b->k_heart = 256;
  b->n_select = (size_t *)calloc(n_arms, sizeof(size_t));
  b->est_rewards = (double *)calloc(n_arms, sizeof(double));
  b->est_square_rewards = (double *)calloc(n_arms, sizeof(double));
  b->stage_n_update = (size_t *)calloc(n_arms, sizeof(size_t));
  b->r_means = (double *)calloc(n_arms, sizeof(double));
  return b;
}

void orc_bandit_free(orc_bandit *b) {
  if (!b) return;
  free(b->n_select);
  free(b->est_rewards);
  free(b->est_square_rewards);
  free(b->stage_n_update);
  free(b->r_means);
  free(b);
}

/* :123-127 */
static double ucb_tuned(const orc_bandit *b, size_t arm) {
  double lt = log((double)b->stage_update_times);
  double denom = (double)b->stage_n_update[arm] + b->k_epsilon;
  double ucb_var = b->est_square_rewards[arm] - b->est_rewards[arm] * b->est_rewards[arm] + sqrt(2 * lt / denom);
  return sqrt(lt / denom * fmin(0.25, ucb_var));
}

/* :34-61 */
size_t orc_bandit_select(orc_bandit *b) {
  if (b->n_start_sampling < b->k_arms * b->k_start_sampling) {
    size_t arm = b->n_start_sampling % b->k_arms;
    b->n_start_sampling++;
    b->select_times++;
    b->n_select[arm]++;
    return arm;
  }
  double max_value = -1;
  size_t max_arm = 0;
  for (size_t i = 0; i < b->k_arms; ++i) {
    double value = b->est_rewards[i] + ucb_tuned(b, i);
    if (value > max_value) {
      max_value = value;
      max_arm = i;
    }
  }
  b->select_times++;
  b->n_select[max_arm]++;
  return max_arm;
}

/* :64-91 */
void orc_bandit_update(orc_bandit *b, size_t arm, double reward) {
  if (b->select_times % b->k_heart == 0 && b->n_start_sampling >= b->k_arms * b->k_start_sampling) {
    b->history_len++;
    if (!b->r_means_set) {
      memcpy(b->r_means, b->est_rewards, b->k_arms * sizeof(double));
      b->r_means_set = 1;
    }
    int detected = b->est_rewards[arm] > b->r_means[arm] * 2 || b->est_rewards[arm] < b->r_means[arm] / 2;
    memcpy(b->r_means, b->est_rewards, b->k_arms * sizeof(double));
    if (detected) {
      b->n_start_sampling = 0;
      for (size_t i = 0; i < b->k_arms; ++i) {
        b->est_rewards[i] = 0;
        b->est_square_rewards[i] = 0;
        b->stage_n_update[i] = 0;
      }
      b->stage_update_times = 0;
    }
  }
  size_t update_factor = b->stage_n_update[arm] < 15 ? b->stage_n_update[arm] : 15;
  double ratio = (double)update_factor / ((double)update_factor + 1.0);
  b->est_rewards[arm] = b->est_rewards[arm] * ratio + reward * (1 - ratio);
  b->est_square_rewards[arm] = b->est_square_rewards[arm] * ratio + reward * reward * (1 - ratio);
  b->stage_update_times++;
  b->stage_n_update[arm]++;
}

void orc_bandit_state(const orc_bandit *b, double *est_rewards, uint64_t *n_select) {
  for (size_t i = 0; i < b->k_arms; ++i) {
    if (est_rewards) est_rewards[i] = b->est_rewards[i];
    if (n_select) n_select[i] = b->n_select[i];
  }
}
