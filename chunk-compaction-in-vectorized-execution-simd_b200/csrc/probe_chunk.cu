// probe_chunk.cu -- the reference's chunk-granular operator protocol on the GPU:
//   Probe -> ScanStructure ; while (HasNext) Next / InOneNext
// (chaining_ht.cpp:38-173, linear_probing_ht.cpp:39-153) plus the DataChunk
// primitives Slice / Append / Reset (base.cpp:15-47, base.h:96-99) and the
// DataCollection row<->column transposes (data_collection.cpp:10-27).
//
// One CTA of 1024 threads walks the chunk in 1024-lane tiles and keeps every
// compaction STABLE (ballot + warp-count scan), so lane order -- and with it every
// selection vector -- is identical to the scalar reference, call by call.
// This path exists for drop-in parity and interop with a chunk-at-a-time engine;
// the throughput paths are probe_batch.cu and chain_fused.cu.
#include "common.cuh"

struct cc_scan {
  const cc_ht *ht;
  size_t block;
  size_t count;  // host mirror of the active-lane count
  uint32_t *d_lane_sel;
  uint64_t *d_pos;
  uint32_t *d_end;
  uint32_t *d_rv;
  const uint32_t *d_key_sel;
  uint32_t *d_key_sel_copy;
  uint32_t *d_counts;  // [0] active lanes, [1] result rows of the last Next
  uint32_t *h_counts;  // pinned mirror
};

namespace ccb {

constexpr int kChunkThreads = 1024;

struct TableView {
  const uint64_t *slots;
  const uint2 *dir;
  const int64_t *ckeys;
  uint64_t mask;
};

static TableView view_of(const cc_ht *ht) {
  TableView v;
  v.slots = ht->d_slots;
  v.dir = ht->d_dir;
  v.ckeys = ht->d_ckeys;
  v.mask = ht->mask;
  return v;
}

// stable rank of `flag` among the CTA's threads; total = number of set flags.
__device__ __forceinline__ uint32_t block_rank(bool flag, uint32_t &total, uint32_t *s_warp /* [33] */) {
  unsigned b = __ballot_sync(0xffffffffu, flag);
  unsigned w = threadIdx.x >> 5;
  if (lane_id() == 0) s_warp[w] = __popc(b);
  __syncthreads();
  if (w == 0) {
    uint32_t v = lane_id() < (blockDim.x >> 5) ? s_warp[lane_id()] : 0;
    uint32_t incl = warp_incl_scan_u32(v);
    s_warp[lane_id()] = incl - v;
    if (lane_id() == 31) s_warp[32] = incl;
  }
  __syncthreads();
  uint32_t off = s_warp[w] + __popc(b & lanemask_lt());
  total = s_warp[32];
  __syncthreads();
  return off;
}

template <int KIND>
__global__ void __launch_bounds__(kChunkThreads)
    chunk_probe_kernel(TableView t, const int64_t *__restrict__ key_col, const uint32_t *__restrict__ sel, uint32_t count,
                       uint32_t *lane_sel, uint64_t *pos, uint32_t *end, uint32_t *counts) {
  __shared__ uint32_t s_warp[33];
  uint32_t base = 0;
  for (uint32_t t0 = 0; t0 < count; t0 += kChunkThreads) {
    uint32_t i = t0 + threadIdx.x;
    bool valid = false;
    if (i < count) {
      uint64_t k = (uint64_t) key_col[sel[i]];
      uint64_t h = murmurhash64(k) & t.mask;
      if (KIND == CC_HT_LP) {
        pos[i] = h;
        valid = t.slots[h] != kEmptyU;  // linear_probing_ht.cpp:53-57
      } else {
        uint2 d = t.dir[h];
        pos[i] = d.x;
        end[i] = d.x + d.y;
        valid = d.y != 0;  // chaining_ht.cpp:52-55
      }
    }
    uint32_t total;
    uint32_t off = block_rank(valid, total, s_warp);
    if (valid) lane_sel[base + off] = i;
    base += total;
  }
  if (threadIdx.x == 0) {
    counts[0] = base;
    counts[1] = 0;
  }
}

template <int KIND>
__device__ __forceinline__ uint64_t entry_at(const TableView &t, uint64_t p) {
  return KIND == CC_HT_LP ? t.slots[p] : (uint64_t) t.ckeys[p];
}

// advance all active lanes one step and drop finished ones (stable, in place).
// chaining_ht.cpp:109-124 / linear_probing_ht.cpp:100-110
template <int KIND>
__device__ __forceinline__ uint32_t advance_lanes(const TableView &t, uint32_t cnt, uint32_t *lane_sel, uint64_t *pos,
                                                  const uint32_t *end, uint32_t *s_warp) {
  uint32_t nc = 0;
  for (uint32_t t0 = 0; t0 < cnt; t0 += kChunkThreads) {
    uint32_t a = t0 + threadIdx.x;
    bool keep = false;
    uint32_t idx = 0;
    if (a < cnt) {
      idx = lane_sel[a];
      if (KIND == CC_HT_LP) {
        uint64_t id = (pos[idx] + 1) & t.mask;
        pos[idx] = id;
        keep = t.slots[id] != kEmptyU;
      } else {
        uint64_t p = pos[idx] + 1;
        pos[idx] = p;
        keep = p != end[idx];
      }
    }
    uint32_t total;
    uint32_t off = block_rank(keep, total, s_warp);
    if (keep) lane_sel[nc + off] = idx;
    nc += total;
  }
  __syncthreads();  // lane_sel / pos writes visible to the next phase
  return nc;
}

template <int KIND>
__device__ __forceinline__ uint32_t match_lanes(const TableView &t, uint32_t cnt, const uint32_t *lane_sel, const uint64_t *pos,
                                                const int64_t *key_col, const uint32_t *key_sel, uint32_t *rv, uint32_t *s_warp) {
  uint32_t rc = 0;
  for (uint32_t t0 = 0; t0 < cnt; t0 += kChunkThreads) {
    uint32_t a = t0 + threadIdx.x;
    bool m = false;
    uint32_t idx = 0;
    if (a < cnt) {
      idx = lane_sel[a];
      uint64_t l = (uint64_t) key_col[key_sel[idx]];
      m = (l == entry_at<KIND>(t, pos[idx]));
    }
    uint32_t total;
    uint32_t off = block_rank(m, total, s_warp);
    if (m) rv[rc + off] = idx;
    rc += total;
  }
  __syncthreads();
  return rc;
}

// Next (in_one == false) / InOneNext (in_one == true)
template <int KIND, bool IN_ONE>
__global__ void __launch_bounds__(kChunkThreads)
    chunk_next_kernel(TableView t, uint32_t block, const int64_t *__restrict__ key_col, const uint32_t *key_sel,
                      const uint32_t *in_sel, uint32_t *lane_sel, uint64_t *pos, const uint32_t *end, uint32_t *rv,
                      uint32_t *out_sel, int64_t *out_payload, uint32_t *counts) {
  __shared__ uint32_t s_warp[33];
  uint32_t cnt = counts[0];
  uint32_t rc = 0;
  __syncthreads();
  if (IN_ONE) {
    // fused match + gather + advance; payload written for ALL active lanes
    // (chaining_ht.cpp:150-165, linear_probing_ht.cpp:129-145)
    uint32_t nc = 0;
    for (uint32_t t0 = 0; t0 < cnt; t0 += kChunkThreads) {
      uint32_t a = t0 + threadIdx.x;
      bool m = false, keep = false;
      uint32_t idx = 0;
      if (a < cnt) {
        idx = lane_sel[a];
        uint32_t phys = key_sel[idx];
        uint64_t l = (uint64_t) key_col[phys];
        uint64_t r = entry_at<KIND>(t, pos[idx]);
        out_payload[phys] = (int64_t) r;
        m = (l == r);
        if (KIND == CC_HT_LP) {
          uint64_t id = (pos[idx] + 1) & t.mask;
          pos[idx] = id;
          keep = t.slots[id] != kEmptyU;
        } else {
          uint64_t p = pos[idx] + 1;
          pos[idx] = p;
          keep = p != end[idx];
        }
      }
      uint32_t total;
      uint32_t off = block_rank(m, total, s_warp);
      if (m) rv[rc + off] = idx;
      rc += total;
      off = block_rank(keep, total, s_warp);
      if (keep) lane_sel[nc + off] = idx;
      nc += total;
    }
    cnt = nc;
  } else if (KIND == CC_HT_CHAIN) {
    // ScanInnerJoin: retry until >= 1 match or all chains end (chaining_ht.cpp:82-107)
    for (;;) {
      rc = match_lanes<KIND>(t, cnt, lane_sel, pos, key_col, key_sel, rv, s_warp);
      if (rc > 0) break;
      cnt = advance_lanes<KIND>(t, cnt, lane_sel, pos, end, s_warp);
      if (cnt == 0) break;
    }
  } else {
    rc = match_lanes<KIND>(t, cnt, lane_sel, pos, key_col, key_sel, rv, s_warp);
  }
  __syncthreads();
  // Reset + Slice (base.h:96-99, base.cpp:42-46) and GatherResult (chaining_ht.cpp:126-136)
  for (uint32_t j = threadIdx.x; j < block; j += kChunkThreads) {
    if (j < rc) {
      uint32_t idx = rv[j];
      out_sel[j] = in_sel[idx];
      if (!IN_ONE) out_payload[key_sel[idx]] = (int64_t) entry_at<KIND>(t, pos[idx]);
    } else {
      out_sel[j] = j;
    }
  }
  __syncthreads();
  if (!IN_ONE) cnt = advance_lanes<KIND>(t, cnt, lane_sel, pos, end, s_warp);
  if (threadIdx.x == 0) {
    counts[0] = cnt;
    counts[1] = rc;
  }
}

// ---- DataChunk primitives ------------------------------------------------------
struct ColPtrs {
  int64_t *dst[32];
  const int64_t *src[32];
};

__global__ void chunk_append_kernel(ColPtrs p, int ncol, size_t dst_count, const uint32_t *__restrict__ src_sel, size_t num,
                                    size_t offset) {
  size_t total = num * (size_t) ncol;
  size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int c = (int) (i / num);
    size_t j = i - (size_t) c * num;
    p.dst[c][dst_count + j] = p.src[c][src_sel[offset + j]];
  }
}

__global__ void sel_compose_kernel(uint32_t *out, const uint32_t *__restrict__ other, const uint32_t *__restrict__ sv, size_t count) {
  size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) out[i] = other[sv[i]];
}

__global__ void sel_identity_kernel(uint32_t *sel, size_t n) {
  size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) sel[i] = (uint32_t) i;
}

__global__ void rows_to_columns_kernel(const int64_t *__restrict__ rows, size_t n_rows, int ncol, ColPtrs p) {
  size_t total = n_rows * (size_t) ncol;
  size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    size_t r = i / ncol;
    int c = (int) (i - r * ncol);
    p.dst[c][r] = rows[i];
  }
}

__global__ void columns_to_rows_kernel(ColPtrs p, const uint32_t *__restrict__ sel, size_t count, int ncol, int64_t *rows) {
  size_t total = count * (size_t) ncol;
  size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    size_t r = i / ncol;
    int c = (int) (i - r * ncol);
    size_t phys = sel ? sel[r] : r;
    rows[i] = p.src[c][phys];
  }
}

static int small_grid(size_t n) {
  size_t b = (n + 255) / 256;
  size_t cap = (size_t) sm_count() * 8;
  if (b > cap) b = cap;
  return (int) (b ? b : 1);
}

}  // namespace ccb

using namespace ccb;

extern "C" {

int cc_probe_chunk(const cc_ht *ht, const int64_t *d_key_col, size_t count, const uint32_t *d_sel, size_t block,
                   cc_scan **out, cc_stream_t s) {
  CC_REQUIRE(out, "scan is NULL");
  *out = nullptr;
  CC_TRY(require_device());
  CC_REQUIRE(ht && d_key_col && d_sel, "NULL argument");
  CC_REQUIRE(block > 0 && block < 0xFFFFFFFFull && count <= block, "count %zu must be <= block_size %zu", count, block);
  cudaStream_t st = as_stream(s);
  cc_scan *sc = new cc_scan();
  sc->ht = ht;
  sc->block = block;
  cudaError_t e = cudaMalloc(&sc->d_lane_sel, block * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&sc->d_pos, block * sizeof(uint64_t));
  if (e == cudaSuccess) e = cudaMalloc(&sc->d_end, block * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&sc->d_rv, block * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&sc->d_counts, 2 * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMallocHost(&sc->h_counts, 2 * sizeof(uint32_t));
  if (e == cudaSuccess && ht->kind == CC_HT_LP) {
    // LPScanStructure owns a COPY of the key selection vector (linear_probing_ht.h:48)
    e = cudaMalloc(&sc->d_key_sel_copy, block * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemcpyAsync(sc->d_key_sel_copy, d_sel, block * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st);
    sc->d_key_sel = sc->d_key_sel_copy;
  } else {
    sc->d_key_sel = d_sel;  // ScanStructure keeps a reference (chaining_ht.h:54)
  }
  if (e != cudaSuccess) {
    set_error("cc_probe_chunk: %s", cudaGetErrorString(e));
    cc_scan_destroy(sc);
    return e == cudaErrorMemoryAllocation ? CC_ERR_NOMEM : CC_ERR_CUDA;
  }
  TableView tv = view_of(ht);
  if (ht->kind == CC_HT_LP)
    chunk_probe_kernel<CC_HT_LP><<<1, kChunkThreads, 0, st>>>(tv, d_key_col, d_sel, (uint32_t) count, sc->d_lane_sel, sc->d_pos, sc->d_end, sc->d_counts);
  else
    chunk_probe_kernel<CC_HT_CHAIN><<<1, kChunkThreads, 0, st>>>(tv, d_key_col, d_sel, (uint32_t) count, sc->d_lane_sel, sc->d_pos, sc->d_end, sc->d_counts);
  note_launch();
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(sc->h_counts, sc->d_counts, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    set_error("cc_probe_chunk: %s", cudaGetErrorString(e));
    cc_scan_destroy(sc);
    return CC_ERR_CUDA;
  }
  sc->count = sc->h_counts[0];
  *out = sc;
  return CC_OK;
}

int cc_scan_has_next(const cc_scan *sc) { return sc && sc->count > 0; }
size_t cc_scan_active(const cc_scan *sc) { return sc ? sc->count : 0; }

int cc_scan_next(cc_scan *sc, int in_one, const int64_t *d_key_col, const uint32_t *d_in_sel, uint32_t *d_out_sel,
                 int64_t *d_out_payload, size_t *out_count, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(sc && d_key_col && d_in_sel && d_out_sel && d_out_payload && out_count, "NULL argument");
  cudaStream_t st = as_stream(s);
  TableView tv = view_of(sc->ht);
  uint32_t block = (uint32_t) sc->block;
#define CC_NEXT_ARGS tv, block, d_key_col, sc->d_key_sel, d_in_sel, sc->d_lane_sel, sc->d_pos, sc->d_end, sc->d_rv, d_out_sel, d_out_payload, sc->d_counts
  if (sc->ht->kind == CC_HT_LP) {
    if (in_one)
      chunk_next_kernel<CC_HT_LP, true><<<1, kChunkThreads, 0, st>>>(CC_NEXT_ARGS);
    else
      chunk_next_kernel<CC_HT_LP, false><<<1, kChunkThreads, 0, st>>>(CC_NEXT_ARGS);
  } else {
    if (in_one)
      chunk_next_kernel<CC_HT_CHAIN, true><<<1, kChunkThreads, 0, st>>>(CC_NEXT_ARGS);
    else
      chunk_next_kernel<CC_HT_CHAIN, false><<<1, kChunkThreads, 0, st>>>(CC_NEXT_ARGS);
  }
#undef CC_NEXT_ARGS
  CC_CHECK_LAUNCH();
  CC_CUDA(cudaMemcpyAsync(sc->h_counts, sc->d_counts, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CC_CUDA(cudaStreamSynchronize(st));
  sc->count = sc->h_counts[0];
  *out_count = sc->h_counts[1];
  return CC_OK;
}

int cc_scan_destroy(cc_scan *sc) {
  if (!sc) return CC_OK;
  if (sc->d_lane_sel) cudaFree(sc->d_lane_sel);
  if (sc->d_pos) cudaFree(sc->d_pos);
  if (sc->d_end) cudaFree(sc->d_end);
  if (sc->d_rv) cudaFree(sc->d_rv);
  if (sc->d_key_sel_copy) cudaFree(sc->d_key_sel_copy);
  if (sc->d_counts) cudaFree(sc->d_counts);
  if (sc->h_counts) cudaFreeHost(sc->h_counts);
  delete sc;
  return CC_OK;
}

int cc_chunk_append(int64_t *const *h_dst_cols, size_t dst_count, const int64_t *const *h_src_cols, const uint32_t *d_src_sel,
                    size_t num, size_t offset, size_t ncol, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(ncol <= 32, "at most 32 columns per chunk (got %zu)", ncol);
  if (num == 0 || ncol == 0) return CC_OK;
  CC_REQUIRE(h_dst_cols && h_src_cols && d_src_sel, "NULL argument");
  ColPtrs p;
  for (size_t c = 0; c < ncol; ++c) {
    p.dst[c] = h_dst_cols[c];
    p.src[c] = h_src_cols[c];
  }
  chunk_append_kernel<<<small_grid(num * ncol), 256, 0, as_stream(s)>>>(p, (int) ncol, dst_count, d_src_sel, num, offset);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

int cc_sel_compose(uint32_t *d_out_sel, const uint32_t *d_other_sel, const uint32_t *d_sv, size_t count, cc_stream_t s) {
  CC_TRY(require_device());
  if (count == 0) return CC_OK;
  CC_REQUIRE(d_out_sel && d_other_sel && d_sv, "NULL argument");
  sel_compose_kernel<<<small_grid(count), 256, 0, as_stream(s)>>>(d_out_sel, d_other_sel, d_sv, count);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

int cc_sel_identity(uint32_t *d_sel, size_t block, cc_stream_t s) {
  CC_TRY(require_device());
  if (block == 0) return CC_OK;
  CC_REQUIRE(d_sel, "NULL argument");
  sel_identity_kernel<<<small_grid(block), 256, 0, as_stream(s)>>>(d_sel, block);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

int cc_rows_to_columns(const int64_t *d_rows, size_t n_rows, size_t ncol, int64_t *const *h_cols, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(ncol <= 32, "at most 32 columns (got %zu)", ncol);
  if (n_rows == 0 || ncol == 0) return CC_OK;
  CC_REQUIRE(d_rows && h_cols, "NULL argument");
  ColPtrs p;
  for (size_t c = 0; c < ncol; ++c) p.dst[c] = h_cols[c];
  rows_to_columns_kernel<<<small_grid(n_rows * ncol), 256, 0, as_stream(s)>>>(d_rows, n_rows, (int) ncol, p);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

int cc_columns_to_rows(const int64_t *const *h_cols, const uint32_t *d_sel, size_t count, size_t ncol, int64_t *d_rows,
                       cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(ncol <= 32, "at most 32 columns (got %zu)", ncol);
  if (count == 0 || ncol == 0) return CC_OK;
  CC_REQUIRE(d_rows && h_cols, "NULL argument");
  ColPtrs p;
  for (size_t c = 0; c < ncol; ++c) p.src[c] = h_cols[c];
  columns_to_rows_kernel<<<small_grid(count * ncol), 256, 0, as_stream(s)>>>(p, d_sel, count, (int) ncol, d_rows);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

}  // extern "C"
