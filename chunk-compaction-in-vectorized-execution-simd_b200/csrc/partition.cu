// partition.cu -- hash partitioning for the multi-GPU partitioned join (SURVEY 8e;
// new functionality, the reference is single-process).
//
// partition id = murmurhash64(key) >> (64 - log2 P): the HIGH hash bits, so the id is
// independent of the low bits that pick the slot/bucket inside the owning GPU's table.
// cc_partition_count builds the P-bin histogram with per-CTA shared-memory counters;
// cc_partition_scatter groups the keys into P contiguous segments.  Each CTA ranks its
// tile locally (shared-memory counters), reserves one contiguous range per partition
// with a single global atomicAdd, and writes rows of the same partition at consecutive
// addresses.  The exchange itself (all-to-all over NVLink) is issued by the host layer
// (torch.distributed / NCCL) directly on the segment buffers.
#include "common.cuh"

namespace ccb {

constexpr int kPartThreads = 256;
constexpr int kPartItems = 8;
constexpr int kPartTile = kPartThreads * kPartItems;
constexpr int kMaxParts = 256;

__global__ void __launch_bounds__(kPartThreads) partition_count_kernel(const int64_t *__restrict__ keys, size_t n, int shift,
                                                                       int parts, unsigned long long *counts) {
  __shared__ uint32_t s_cnt[kMaxParts];
  for (int i = threadIdx.x; i < parts; i += kPartThreads) s_cnt[i] = 0;
  __syncthreads();
  size_t stride = (size_t) gridDim.x * kPartThreads;
  for (size_t i = (size_t) blockIdx.x * kPartThreads + threadIdx.x; i < n; i += stride) {
    uint32_t p = shift >= 64 ? 0u : (uint32_t) (murmurhash64((uint64_t) keys[i]) >> shift);
    atomicAdd(&s_cnt[p], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < parts; i += kPartThreads)
    if (s_cnt[i]) atomicAdd(counts + i, (unsigned long long) s_cnt[i]);
}

__global__ void __launch_bounds__(kPartThreads)
    partition_scatter_kernel(const int64_t *__restrict__ keys, size_t n, int shift, int parts, const unsigned long long *__restrict__ offsets,
                             unsigned long long *cursors, int64_t *out) {
  __shared__ uint32_t s_cnt[kMaxParts];
  __shared__ unsigned long long s_base[kMaxParts];
  const size_t ntiles = (n + kPartTile - 1) / kPartTile;
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    for (int i = threadIdx.x; i < parts; i += kPartThreads) s_cnt[i] = 0;
    __syncthreads();
    uint64_t k[kPartItems];
    uint32_t p[kPartItems], r[kPartItems];
    const size_t tbase = tile * (size_t) kPartTile;
#pragma unroll
    for (int j = 0; j < kPartItems; ++j) {
      size_t idx = tbase + (size_t) j * kPartThreads + threadIdx.x;
      bool ok = idx < n;
      k[j] = ok ? (uint64_t) __ldg(keys + idx) : 0;
      p[j] = ok ? (shift >= 64 ? 0u : (uint32_t) (murmurhash64(k[j]) >> shift)) : 0xFFFFFFFFu;
      r[j] = ok ? atomicAdd(&s_cnt[p[j]], 1u) : 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < parts; i += kPartThreads)
      s_base[i] = s_cnt[i] ? offsets[i] + atomicAdd(cursors + i, (unsigned long long) s_cnt[i]) : 0ull;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kPartItems; ++j)
      if (p[j] != 0xFFFFFFFFu) out[s_base[p[j]] + r[j]] = (int64_t) k[j];
    __syncthreads();
  }
}

}  // namespace ccb

using namespace ccb;

extern "C" {

int cc_partition_count(const int64_t *d_keys, size_t n, int log2_parts, uint64_t *d_counts, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(log2_parts >= 0 && (1 << log2_parts) <= kMaxParts, "log2_parts must be in [0, 8]");
  CC_REQUIRE(d_counts && (n == 0 || d_keys), "NULL argument");
  int parts = 1 << log2_parts;
  CC_CUDA(cudaMemsetAsync(d_counts, 0, parts * sizeof(uint64_t), as_stream(s)));
  if (n == 0) return CC_OK;
  size_t blocks = std::min<size_t>((n + kPartTile - 1) / kPartTile, (size_t) sm_count() * 8);
  partition_count_kernel<<<(unsigned) blocks, kPartThreads, 0, as_stream(s)>>>(d_keys, n, 64 - log2_parts, parts,
                                                                              (unsigned long long *) d_counts);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

int cc_partition_scatter(const int64_t *d_keys, size_t n, int log2_parts, const uint64_t *d_offsets, uint64_t *d_cursors,
                         int64_t *d_out, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(log2_parts >= 0 && (1 << log2_parts) <= kMaxParts, "log2_parts must be in [0, 8]");
  CC_REQUIRE(d_offsets && d_cursors && (n == 0 || (d_keys && d_out)), "NULL argument");
  int parts = 1 << log2_parts;
  CC_CUDA(cudaMemsetAsync(d_cursors, 0, parts * sizeof(uint64_t), as_stream(s)));
  if (n == 0) return CC_OK;
  size_t blocks = std::min<size_t>((n + kPartTile - 1) / kPartTile, (size_t) sm_count() * 8);
  partition_scatter_kernel<<<(unsigned) blocks, kPartThreads, 0, as_stream(s)>>>(
      d_keys, n, 64 - log2_parts, parts, (const unsigned long long *) d_offsets, (unsigned long long *) d_cursors, d_out);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

}  // extern "C"
