// probe_batch.cu -- the throughput path for ONE join: a persistent kernel probes a
// whole key column and writes dense (compacted) result columns.
//
// It is the fused GPU form of the micro-bench loop (simd_micro_bench.cpp:83-116:
// Probe + while(HasNext) Next over every 2048-row chunk) followed by a Compactor on
// every sparse result chunk (compactor.cpp:5-41):
//   * a CTA iteration owns one chunk of kTile probe rows, kKeysPerThread per thread,
//     loaded with coalesced strided loads so that all table accesses of a thread are
//     independent and in flight together (memory-level parallelism for the gather);
//   * a "round" is one Next(): every active lane compares its current slot / chain
//     entry (ScanInnerJoin), matches are ranked with warp ballots + a 32-entry scan of
//     the (key-slice, warp) counts, the CTA reserves a contiguous range of the global
//     output with ONE atomicAdd, and the lanes store key and payload at consecutive
//     positions (coalesced) -- the compaction step; then lanes advance
//     (AdvancePointers) and finished lanes retire;
//   * tables known to hold no duplicate key retire a lane at its first match (the
//     walk past a match only ever finds duplicates, linear_probing_ht.cpp:101-109).
// Output row order is unspecified; the multiset equals the reference's.
#include "common.cuh"

namespace ccb {

constexpr int kPbThreads = 256;
constexpr int kPbKeysPerThread = 4;
constexpr int kPbWarps = kPbThreads / 32;
constexpr int kPbTile = kPbThreads * kPbKeysPerThread;
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
};

template <int KIND, bool UNIQUE>
__global__ void __launch_bounds__(kPbThreads, 4) probe_batch_kernel(ProbeArgs a) {
  __shared__ uint32_t s_cnt[32];
  __shared__ unsigned long long s_base;
  const unsigned w = threadIdx.x >> 5;
  const unsigned lt = lanemask_lt();
  uint64_t ksum = 0, psum = 0;
  const size_t ntiles = (a.n + kPbTile - 1) / kPbTile;
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const size_t tbase = tile * (size_t) kPbTile;
    uint64_t k[kPbKeysPerThread], v[kPbKeysPerThread], pos[kPbKeysPerThread];
    uint32_t end[kPbKeysPerThread];
    bool act[kPbKeysPerThread];
    // ---- Probe (chaining_ht.cpp:44-55 / linear_probing_ht.cpp:45-57)
#pragma unroll
    for (int j = 0; j < kPbKeysPerThread; ++j) {
      size_t idx = tbase + (size_t) j * kPbThreads + threadIdx.x;
      act[j] = idx < a.n;
      k[j] = act[j] ? (uint64_t) __ldg(a.keys + idx) : 0;
      pos[j] = murmurhash64(k[j]) & a.mask;
    }
    if (KIND == CC_HT_LP) {
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j) v[j] = act[j] ? ld_nc_u64(a.slots + pos[j]) : kEmptyU;
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j) act[j] = v[j] != kEmptyU;
    } else {
      uint2 d[kPbKeysPerThread];
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j) d[j] = act[j] ? __ldg(a.dir + pos[j]) : make_uint2(0, 0);
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j) {
        pos[j] = d[j].x;
        end[j] = d[j].x + d[j].y;
        act[j] = d[j].y != 0;
        v[j] = act[j] ? (uint64_t) __ldg(a.ckeys + pos[j]) : 0;
      }
    }
    // ---- rounds: one Next() each
    bool any;
    do {
      bool m[kPbKeysPerThread];
      unsigned bal[kPbKeysPerThread];
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j) {
        m[j] = act[j] && (v[j] == k[j]);
        bal[j] = __ballot_sync(0xffffffffu, m[j]);
        if (lane_id() == 0) s_cnt[j * kPbWarps + w] = __popc(bal[j]);
      }
      __syncthreads();
      if (w == 0) {
        uint32_t c = s_cnt[lane_id()];
        uint32_t incl = warp_incl_scan_u32(c);
        s_cnt[lane_id()] = incl - c;
        if (lane_id() == 31) s_base = incl ? atomicAdd((unsigned long long *) &a.res->n_matches, (unsigned long long) incl) : 0ull;
      }
      __syncthreads();
      const uint64_t base = s_base;
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j) {
        if (m[j]) {
          uint64_t dst = base + s_cnt[j * kPbWarps + w] + __popc(bal[j] & lt);
          ksum += k[j];
          psum += v[j];
          if (dst < a.cap) {
            if (a.out_key) a.out_key[dst] = (int64_t) k[j];
            if (a.out_payload) a.out_payload[dst] = (int64_t) v[j];
            if (a.out_rowid) a.out_rowid[dst] = tbase + (size_t) j * kPbThreads + threadIdx.x;
          }
        }
      }
      // ---- AdvancePointers (chaining_ht.cpp:109-124 / linear_probing_ht.cpp:100-110)
      bool mine = false;
#pragma unroll
      for (int j = 0; j < kPbKeysPerThread; ++j) {
        if (act[j]) {
          if (UNIQUE && m[j]) {
            act[j] = false;
          } else if (KIND == CC_HT_LP) {
            pos[j] = (pos[j] + 1) & a.mask;
            v[j] = ld_nc_u64(a.slots + pos[j]);
            act[j] = v[j] != kEmptyU;
          } else {
            pos[j] += 1;
            act[j] = pos[j] != end[j];
            if (act[j]) v[j] = (uint64_t) __ldg(a.ckeys + pos[j]);
          }
        }
        mine |= act[j];
      }
      any = __syncthreads_or(mine);
    } while (any);
  }
  ksum = warp_sum_u64(ksum);
  psum = warp_sum_u64(psum);
  if (lane_id() == 0 && (ksum | psum)) {
    atomicAdd((unsigned long long *) &a.res->key_sum, (unsigned long long) ksum);
    atomicAdd((unsigned long long *) &a.res->payload_sum, (unsigned long long) psum);
  }
}

__global__ void probe_finish_kernel(cc_probe_result *res, size_t cap) {
  if (threadIdx.x == 0 && blockIdx.x == 0) res->overflow = res->n_matches > cap ? 1 : 0;
}

template <int KIND, bool UNIQUE>
static int launch_probe(const ProbeArgs &a, cudaStream_t st) {
  static int blocks_per_sm = 0;
  if (!blocks_per_sm) {
    CC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, probe_batch_kernel<KIND, UNIQUE>, kPbThreads, 0));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
  }
  size_t ntiles = (a.n + kPbTile - 1) / kPbTile;
  size_t grid = (size_t) sm_count() * blocks_per_sm;
  if (grid > ntiles) grid = ntiles;
  if (grid == 0) grid = 1;
  probe_batch_kernel<KIND, UNIQUE><<<(unsigned) grid, kPbThreads, 0, st>>>(a);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

int probe_batch_device(const cc_ht *ht, const int64_t *d_keys, size_t n, int64_t *d_out_key, int64_t *d_out_payload,
                       uint64_t *d_out_rowid, size_t cap, cc_probe_result *d_result, cudaStream_t st) {
  ProbeArgs a;
  a.slots = ht->d_slots;
  a.dir = ht->d_dir;
  a.ckeys = ht->d_ckeys;
  a.mask = ht->mask;
  a.keys = d_keys;
  a.n = n;
  a.out_key = d_out_key;
  a.out_payload = d_out_payload;
  a.out_rowid = d_out_rowid;
  a.cap = (d_out_key || d_out_payload || d_out_rowid) ? cap : 0;
  a.res = d_result;
  CC_CUDA(cudaMemsetAsync(d_result, 0, sizeof(cc_probe_result), st));
  if (n) {
    bool unique = !ht->has_duplicates;
    if (ht->kind == CC_HT_LP)
      CC_TRY(unique ? (launch_probe<CC_HT_LP, true>(a, st)) : (launch_probe<CC_HT_LP, false>(a, st)));
    else
      CC_TRY(unique ? (launch_probe<CC_HT_CHAIN, true>(a, st)) : (launch_probe<CC_HT_CHAIN, false>(a, st)));
  }
  if (a.cap || d_out_key || d_out_payload || d_out_rowid) {
    probe_finish_kernel<<<1, 32, 0, st>>>(d_result, a.cap);
    CC_CHECK_LAUNCH();
  }
  return CC_OK;
}

}  // namespace ccb

using namespace ccb;

extern "C" {

int cc_probe_batch(const cc_ht *ht, const int64_t *d_keys, size_t n, int64_t *d_out_key, int64_t *d_out_payload,
                   uint64_t *d_out_rowid, size_t out_capacity, cc_probe_result *d_result, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(ht && d_result, "NULL argument");
  CC_REQUIRE(n == 0 || d_keys, "d_keys is NULL");
  return probe_batch_device(ht, d_keys, n, d_out_key, d_out_payload, d_out_rowid, out_capacity, d_result, as_stream(s));
}

// End-to-end convenience path with HOST buffers: slices of the key column are copied to the
// device on one stream, probed on a second, and the dense result columns copied back on a
// third, double-buffered so that PCIe traffic in both directions overlaps the kernel.
int cc_probe_batch_host(const cc_ht *ht, const int64_t *h_keys, size_t n, int64_t *h_out_key, int64_t *h_out_payload,
                        size_t out_capacity, cc_probe_result *h_result, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(ht && h_result, "NULL argument");
  CC_REQUIRE(n == 0 || h_keys, "h_keys is NULL");
  (void) s;
  constexpr int kBuf = 2;
  const size_t slice = std::min<size_t>(n ? n : 1, (size_t) 1 << 25);  // 32 Mi keys = 256 MiB per slice
  size_t out_slice_cap = slice * 2;
  struct Buf {
    int64_t *d_keys = nullptr, *d_ok = nullptr, *d_op = nullptr;
    cc_probe_result *d_res = nullptr;
    cc_probe_result *h_res = nullptr;
    cudaEvent_t h2d_done = nullptr, k_done = nullptr, d2h_done = nullptr;
    size_t cap = 0;
  } buf[kBuf];
  cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
  int rc = CC_OK;
  auto cleanup = [&]() {
    for (auto &b : buf) {
      if (b.d_keys) cudaFree(b.d_keys);
      if (b.d_ok) cudaFree(b.d_ok);
      if (b.d_op) cudaFree(b.d_op);
      if (b.d_res) cudaFree(b.d_res);
      if (b.h_res) cudaFreeHost(b.h_res);
      if (b.h2d_done) cudaEventDestroy(b.h2d_done);
      if (b.k_done) cudaEventDestroy(b.k_done);
      if (b.d2h_done) cudaEventDestroy(b.d2h_done);
    }
    if (s_in) cudaStreamDestroy(s_in);
    if (s_k) cudaStreamDestroy(s_k);
    if (s_out) cudaStreamDestroy(s_out);
  };
#define CC_E2E(expr)                                                                    \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      cleanup();                                                                        \
      return e__ == cudaErrorMemoryAllocation ? CC_ERR_NOMEM : CC_ERR_CUDA;             \
    }                                                                                   \
  } while (0)
  CC_E2E(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
  CC_E2E(cudaStreamCreateWithFlags(&s_k, cudaStreamNonBlocking));
  CC_E2E(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
  const bool want_out = h_out_key || h_out_payload;
  for (auto &b : buf) {
    CC_E2E(cudaMalloc(&b.d_keys, slice * sizeof(int64_t)));
    if (want_out) {
      CC_E2E(cudaMalloc(&b.d_ok, out_slice_cap * sizeof(int64_t)));
      CC_E2E(cudaMalloc(&b.d_op, out_slice_cap * sizeof(int64_t)));
      b.cap = out_slice_cap;
    }
    CC_E2E(cudaMalloc(&b.d_res, sizeof(cc_probe_result)));
    CC_E2E(cudaMallocHost(&b.h_res, sizeof(cc_probe_result)));
    CC_E2E(cudaEventCreateWithFlags(&b.h2d_done, cudaEventDisableTiming));
    CC_E2E(cudaEventCreateWithFlags(&b.k_done, cudaEventDisableTiming));
    CC_E2E(cudaEventCreateWithFlags(&b.d2h_done, cudaEventDisableTiming));
  }
  cc_probe_result total = {0, 0, 0, 0};
  size_t n_slices = (n + slice - 1) / slice;
  auto issue = [&](size_t i) -> int {  // H2D + kernel for slice i
    Buf &b = buf[i % kBuf];
    size_t off = i * slice, cnt = std::min(slice, n - off);
    // buffer reuse: the D2H of slice i-kBuf must have drained
    if (i >= kBuf) {
      cudaError_t e = cudaStreamWaitEvent(s_in, b.d2h_done, 0);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(s_k, b.d2h_done, 0);
      if (e != cudaSuccess) return CC_ERR_CUDA;
    }
    if (cudaMemcpyAsync(b.d_keys, h_keys + off, cnt * sizeof(int64_t), cudaMemcpyHostToDevice, s_in) != cudaSuccess) return CC_ERR_CUDA;
    if (cudaEventRecord(b.h2d_done, s_in) != cudaSuccess) return CC_ERR_CUDA;
    if (cudaStreamWaitEvent(s_k, b.h2d_done, 0) != cudaSuccess) return CC_ERR_CUDA;
    int r = probe_batch_device(ht, b.d_keys, cnt, b.d_ok, b.d_op, nullptr, b.cap, b.d_res, s_k);
    if (r != CC_OK) return r;
    if (cudaMemcpyAsync(b.h_res, b.d_res, sizeof(cc_probe_result), cudaMemcpyDeviceToHost, s_k) != cudaSuccess) return CC_ERR_CUDA;
    if (cudaEventRecord(b.k_done, s_k) != cudaSuccess) return CC_ERR_CUDA;
    return CC_OK;
  };
  if (n_slices) rc = issue(0);
  for (size_t i = 0; i < n_slices && rc == CC_OK; ++i) {
    Buf &b = buf[i % kBuf];
    if (i + 1 < n_slices) rc = issue(i + 1);
    if (rc != CC_OK) break;
    CC_E2E(cudaEventSynchronize(b.k_done));
    cc_probe_result r = *b.h_res;
    if (want_out && r.n_matches > b.cap) {
      // fan-out above the slice buffer: grow this buffer and redo the slice (rare)
      CC_E2E(cudaStreamSynchronize(s_out));
      cudaFree(b.d_ok);
      cudaFree(b.d_op);
      b.d_ok = b.d_op = nullptr;
      b.cap = (size_t) r.n_matches;
      CC_E2E(cudaMalloc(&b.d_ok, b.cap * sizeof(int64_t)));
      CC_E2E(cudaMalloc(&b.d_op, b.cap * sizeof(int64_t)));
      size_t off = i * slice, cnt = std::min(slice, n - off);
      rc = probe_batch_device(ht, b.d_keys, cnt, b.d_ok, b.d_op, nullptr, b.cap, b.d_res, s_k);
      if (rc != CC_OK) break;
      CC_E2E(cudaMemcpyAsync(b.h_res, b.d_res, sizeof(cc_probe_result), cudaMemcpyDeviceToHost, s_k));
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
  if (rc == CC_OK) {
    CC_E2E(cudaStreamSynchronize(s_out));
    CC_E2E(cudaStreamSynchronize(s_k));
  }
#undef CC_E2E
  cleanup();
  if (rc != CC_OK) return rc;
  *h_result = total;
  return CC_OK;
}

}  // extern "C"
