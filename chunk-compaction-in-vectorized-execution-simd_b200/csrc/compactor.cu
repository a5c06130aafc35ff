// compactor.cu -- standalone chunk compactor with the reference's call protocol.
//
// NaiveCompactor::Compact / Flush (compactor.cpp:5-41, compactor.h:23) on device
// chunks, result-transparent: the cache always owns its storage (the reference's
// stock build aliases upstream storage through the recycled temp_chunk_, SURVEY 8c
// bug 3; its own commented line compactor.cpp:36 is the intended behaviour).
// The Binary/Dynamic compactors named in setting.h:21,24 are absent from the
// reference; their contract here: a chunk holding >= threshold rows is
// passed through untouched, anything smaller is buffered.  threshold == block_size
// is exactly NaiveCompactor, threshold == 0 never compacts.
//
// The copy itself is DataChunk::Append (base.cpp:15-27) as one gather kernel over all
// columns (cc_chunk_append).  The in-kernel compaction used by the throughput paths
// lives in probe_batch.cu / chain_fused.cu.
#include <vector>

#include "common.cuh"

struct cc_compactor {
  size_t ncol, block, threshold;
  // three rotating dense chunks: [cached], [temp], [emitted]; every chunk owns ncol columns
  std::vector<int64_t *> cols[3];
  size_t count[3];
  int cached, temp, emitted;
  uint32_t *d_identity;  // identity selection vector shared by the dense chunks
};

using namespace ccb;

extern "C" {

int cc_compactor_create(cc_compactor **out, size_t ncol, size_t block, size_t threshold) {
  CC_REQUIRE(out, "compactor is NULL");
  *out = nullptr;
  CC_TRY(require_device());
  CC_REQUIRE(ncol > 0 && ncol <= 32 && block > 0, "need 0 < ncol <= 32 and block_size > 0");
  cc_compactor *c = new cc_compactor();
  c->ncol = ncol;
  c->block = block;
  c->threshold = threshold;
  c->cached = 0;
  c->temp = 1;
  c->emitted = 2;
  c->d_identity = nullptr;
  for (int i = 0; i < 3; ++i) {
    c->count[i] = 0;
    c->cols[i].assign(ncol, nullptr);
  }
  cudaError_t e = cudaMalloc(&c->d_identity, block * sizeof(uint32_t));
  for (int i = 0; i < 3 && e == cudaSuccess; ++i)
    for (size_t j = 0; j < ncol && e == cudaSuccess; ++j) {
      e = cudaMalloc(&c->cols[i][j], block * sizeof(int64_t));
      if (e == cudaSuccess) e = cudaMemset(c->cols[i][j], 0, block * sizeof(int64_t));
    }
  if (e != cudaSuccess) {
    set_error("cc_compactor_create: %s", cudaGetErrorString(e));
    cc_compactor_destroy(c);
    return e == cudaErrorMemoryAllocation ? CC_ERR_NOMEM : CC_ERR_CUDA;
  }
  int rc = cc_sel_identity(c->d_identity, block, nullptr);
  if (rc == CC_OK && cudaDeviceSynchronize() != cudaSuccess) rc = CC_ERR_CUDA;
  if (rc != CC_OK) {
    cc_compactor_destroy(c);
    return rc;
  }
  *out = c;
  return CC_OK;
}

int cc_compactor_set_threshold(cc_compactor *c, size_t threshold) {
  CC_REQUIRE(c, "compactor is NULL");
  c->threshold = threshold;
  return CC_OK;
}

size_t cc_compactor_get_threshold(const cc_compactor *c) { return c ? c->threshold : 0; }

int cc_compactor_compact(cc_compactor *c, int64_t *const *h_cols, const uint32_t *d_sel, size_t *count, int64_t **h_out_cols,
                         const uint32_t **d_out_sel, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(c && h_cols && d_sel && count && h_out_cols && d_out_sel, "NULL argument");
  CC_REQUIRE(*count <= c->block, "chunk holds %zu rows, block_size is %zu", *count, c->block);
  const size_t n = *count;
  // pass-through: full chunk (compactor.cpp:6) or at/above the threshold
  if (n == c->block || n >= c->threshold) {
    for (size_t j = 0; j < c->ncol; ++j) h_out_cols[j] = h_cols[j];
    *d_out_sel = d_sel;
    return CC_OK;
  }
  const int64_t *const *src = const_cast<const int64_t *const *>(h_cols);
  size_t &cc = c->count[c->cached];
  if (n <= c->block - cc) {  // compactor.cpp:12-19
    CC_TRY(cc_chunk_append(c->cols[c->cached].data(), cc, src, d_sel, n, 0, c->ncol, s));
    cc += n;
    *count = 0;
    for (size_t j = 0; j < c->ncol; ++j) h_out_cols[j] = h_cols[j];
    *d_out_sel = d_sel;
    return CC_OK;
  }
  // compactor.cpp:22-36: top the cache up, spill the rest into temp, emit the cache
  size_t n_move = c->block - cc;
  CC_TRY(cc_chunk_append(c->cols[c->cached].data(), cc, src, d_sel, n_move, 0, c->ncol, s));
  CC_TRY(cc_chunk_append(c->cols[c->temp].data(), 0, src, d_sel, n - n_move, n_move, c->ncol, s));
  int full = c->cached;
  c->cached = c->temp;
  c->count[c->cached] = n - n_move;
  c->temp = c->emitted;  // storage of the chunk emitted last time is recycled now
  c->count[c->temp] = 0;
  c->emitted = full;
  c->count[full] = c->block;
  for (size_t j = 0; j < c->ncol; ++j) h_out_cols[j] = c->cols[full][j];
  *d_out_sel = c->d_identity;
  *count = c->block;
  return CC_OK;
}

int cc_compactor_flush(cc_compactor *c, int64_t **h_out_cols, const uint32_t **d_out_sel, size_t *count, cc_stream_t s) {
  (void) s;
  CC_TRY(require_device());
  CC_REQUIRE(c && h_out_cols && d_out_sel && count, "NULL argument");
  int part = c->cached;
  for (size_t j = 0; j < c->ncol; ++j) h_out_cols[j] = c->cols[part][j];
  *d_out_sel = c->d_identity;
  *count = c->count[part];
  // the reference moves the cache out and the compactor is dead afterwards (compactor.h:23);
  // here it simply starts over with an empty cache
  c->cached = c->temp;
  c->count[c->cached] = 0;
  c->temp = c->emitted;
  c->count[c->temp] = 0;
  c->emitted = part;
  return CC_OK;
}

int cc_compactor_destroy(cc_compactor *c) {
  if (!c) return CC_OK;
  for (int i = 0; i < 3; ++i)
    for (auto p : c->cols[i])
      if (p) cudaFree(p);
  if (c->d_identity) cudaFree(c->d_identity);
  delete c;
  return CC_OK;
}

}  // extern "C"
