// tables.cu -- GPU-resident hash tables.
//
//  * LP table  (replaces LPHashTable, linear_probing_ht.cpp:4-37): int64 key-only
//    slots in HBM, -1 == empty, pow2 >= 4n slots.  Built by concurrent CAS inserts.
//    The default ORDERED build uses priority swaps (a thread that meets a slot holding
//    a larger key CAS-swaps itself in and carries the evicted key on), which converges
//    to the unique layout that serial insertion in ascending key order produces --
//    i.e. exactly the reference's layout for its own (non-decreasing) key generator,
//    and a scheduling-independent layout for any input.
//  * chain table (replaces HashTable, chaining_ht.cpp:4-36): the
//    vector<unique_ptr<std::list<Key>>> becomes a bucket directory dir[b] = (begin, count)
//    plus one contiguous key array holding every bucket's chain in insertion order.
//    Built with atomic bucket counters + an exclusive scan + atomic slot claims, then
//    each chain is put into FIFO order (row id order) so that the chunk-granular
//    Next() protocol emits matches in the same call as the reference.
#include <algorithm>
#include <vector>

#include <cstdlib>

#include "common.cuh"
#include "partition.cuh"

namespace ccb {

// ---------------------------------------------------------------- LP build
__global__ void lp_insert_ordered_kernel(const int64_t *__restrict__ keys, size_t n, uint64_t *slots, uint64_t mask,
                                         int *flags) {
  size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint64_t k = (uint64_t) keys[i];
    if (k == kEmptyU) {  // -1 is the empty sentinel (linear_probing_ht.cpp:7): unrepresentable
      atomicOr(flags, 1);
      continue;
    }
    uint64_t s = murmurhash64(k) & mask;
    for (;;) {
      uint64_t cur = ld_cg_u64(slots + s);
      if (cur > k) {  // empty (max) or lower priority: swap ourselves in
        uint64_t old = atomicCAS((unsigned long long *) (slots + s), (unsigned long long) cur, (unsigned long long) k);
        if (old == cur) {
          if (cur == kEmptyU) break;
          k = cur;  // carry the evicted key forward
          s = (s + 1) & mask;
        }
      } else {
        s = (s + 1) & mask;
      }
    }
  }
}

__global__ void lp_insert_unordered_kernel(const int64_t *__restrict__ keys, size_t n, uint64_t *slots, uint64_t mask,
                                           int *flags) {
  size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint64_t k = (uint64_t) keys[i];
    if (k == kEmptyU) {
      atomicOr(flags, 1);
      continue;
    }
    uint64_t s = murmurhash64(k) & mask;
    for (;;) {
      uint64_t cur = ld_cg_u64(slots + s);
      if (cur == kEmptyU) {
        uint64_t old = atomicCAS((unsigned long long *) (slots + s), (unsigned long long) kEmptyU, (unsigned long long) k);
        if (old == kEmptyU) break;
      } else {
        s = (s + 1) & mask;
      }
    }
  }
}

// post-build audit: every key must be reachable from its home slot; a key seen more than
// once marks the table as holding duplicates (probes must then walk past a match).
__global__ void lp_audit_kernel(const int64_t *__restrict__ keys, size_t n, const uint64_t *__restrict__ slots,
                                uint64_t mask, int *flags) {
  size_t stride = (size_t) gridDim.x * blockDim.x;
  int local = 0;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint64_t k = (uint64_t) keys[i];
    if (k == kEmptyU) continue;
    uint64_t s = murmurhash64(k) & mask;
    int found = 0;
    for (;;) {
      uint64_t cur = slots[s];
      if (cur == kEmptyU) break;
      found += (cur == k);
      s = (s + 1) & mask;
    }
    if (found == 0) local |= 4;  // internal error: lost key
    if (found > 1) local |= 2;   // duplicates present
  }
  if (local) atomicOr(flags, local);
}

// ------------------------------------------------------------- chain build
__global__ void chain_count_kernel(const int64_t *__restrict__ keys, size_t n, uint64_t mask, uint32_t *counts) {
  size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint64_t b = murmurhash64((uint64_t) keys[i]) & mask;
    // warp-cooperative: lanes that hit the same bucket (adjacent duplicate keys, cf > 1)
    // elect one leader that adds the whole group's count
    unsigned active = __activemask();
    unsigned peers = __match_any_sync(active, b);
    if ((int) lane_id() == __ffs(peers) - 1) atomicAdd(counts + b, (uint32_t) __popc(peers));
  }
}

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void scan_block_sums_kernel(const uint32_t *__restrict__ counts, size_t n, uint64_t *block_sums) {
  __shared__ uint64_t s_part[kScanThreads / 32];
  size_t base = (size_t) blockIdx.x * kScanTile;
  uint64_t sum = 0;
  for (int j = 0; j < kScanItems; ++j) {
    size_t i = base + (size_t) j * kScanThreads + threadIdx.x;
    if (i < n) sum += counts[i];
  }
  sum = warp_sum_u64(sum);
  if (lane_id() == 0) s_part[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t t = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) t += s_part[w];
    block_sums[blockIdx.x] = t;
  }
}

// single block: exclusive scan of block sums in place; total written to block_sums[nblocks]
__global__ void scan_of_sums_kernel(uint64_t *block_sums, size_t nblocks) {
  __shared__ uint64_t s_warp[32];
  __shared__ uint64_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (size_t base = 0; base < nblocks; base += blockDim.x) {
    size_t i = base + threadIdx.x;
    uint64_t v = i < nblocks ? block_sums[i] : 0;
    uint64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint64_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane_id() >= (unsigned) o) incl += t;
    }
    if (lane_id() == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint64_t w = threadIdx.x < (blockDim.x >> 5) ? s_warp[threadIdx.x] : 0;
      uint64_t wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane_id() >= (unsigned) o) wi += t;
      }
      s_warp[threadIdx.x] = wi - w;  // exclusive
    }
    __syncthreads();
    uint64_t excl = s_carry + s_warp[threadIdx.x >> 5] + (incl - v);
    if (i < nblocks) block_sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) block_sums[nblocks] = s_carry;
}

// writes dir[b] = (begin, count); each thread owns kScanItems consecutive buckets
__global__ void scan_write_dir_kernel(const uint32_t *__restrict__ counts, size_t n, const uint64_t *__restrict__ block_sums,
                                      uint2 *dir) {
  __shared__ uint32_t s_warp[kScanThreads / 32];
  size_t base = (size_t) blockIdx.x * kScanTile + (size_t) threadIdx.x * kScanItems;
  uint32_t c[kScanItems];
  uint32_t tsum = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    c[j] = (base + j < n) ? counts[base + j] : 0;
    tsum += c[j];
  }
  uint32_t incl = warp_incl_scan_u32(tsum);
  if (lane_id() == 31) s_warp[threadIdx.x >> 5] = incl;
  __syncthreads();
  uint32_t woff = 0;
  for (int w = 0; w < (int) (threadIdx.x >> 5); ++w) woff += s_warp[w];
  uint64_t run = block_sums[blockIdx.x] + woff + (incl - tsum);
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    if (base + j < n) dir[base + j] = make_uint2((uint32_t) run, c[j]);
    run += c[j];
  }
}

__global__ void chain_scatter_kernel(const int64_t *__restrict__ keys, size_t n, uint64_t mask, const uint2 *__restrict__ dir,
                                     uint32_t *fill, uint32_t *rowid) {
  size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint64_t b = murmurhash64((uint64_t) keys[i]) & mask;
    unsigned active = __activemask();
    unsigned peers = __match_any_sync(active, b);
    int leader = __ffs(peers) - 1;
    uint32_t basepos = 0;
    if ((int) lane_id() == leader) basepos = atomicAdd(fill + b, (uint32_t) __popc(peers));
    basepos = __shfl_sync(peers, basepos, leader);
    uint32_t rank = __popc(peers & lanemask_lt());
    rowid[dir[b].x + basepos + rank] = (uint32_t) i;
  }
}

constexpr uint32_t kChainSortLimit = 4096;  // longer buckets keep claim order (documented)

// one thread per bucket: order the chain by build row id (== std::list push_back order,
// chaining_ht.cpp:34), materialise the chain keys, audit duplicates / longest chain.
__global__ void chain_finalize_kernel(const int64_t *__restrict__ keys, size_t n_buckets, const uint2 *__restrict__ dir,
                                      uint32_t *rowid, int64_t *ckeys, int *flags, unsigned long long *max_chain) {
  size_t stride = (size_t) gridDim.x * blockDim.x;
  uint32_t local_max = 0;
  int local_flags = 0;
  for (size_t b = (size_t) blockIdx.x * blockDim.x + threadIdx.x; b < n_buckets; b += stride) {
    uint2 d = dir[b];
    uint32_t c = d.y;
    if (c == 0) continue;
    uint32_t *r = rowid + d.x;
    if (c > local_max) local_max = c;
    if (c > 1 && c <= kChainSortLimit) {
      if (c <= 32) {  // insertion sort
        for (uint32_t i = 1; i < c; ++i) {
          uint32_t v = r[i];
          uint32_t j = i;
          while (j > 0 && r[j - 1] > v) {
            r[j] = r[j - 1];
            --j;
          }
          r[j] = v;
        }
      } else {  // heap sort
        for (uint32_t start = c / 2; start-- > 0;) {
          uint32_t root = start;
          for (;;) {
            uint32_t child = 2 * root + 1;
            if (child >= c) break;
            if (child + 1 < c && r[child] < r[child + 1]) ++child;
            if (r[root] >= r[child]) break;
            uint32_t t = r[root];
            r[root] = r[child];
            r[child] = t;
            root = child;
          }
        }
        for (uint32_t end = c - 1; end > 0; --end) {
          uint32_t t = r[0];
          r[0] = r[end];
          r[end] = t;
          uint32_t root = 0;
          for (;;) {
            uint32_t child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && r[child] < r[child + 1]) ++child;
            if (r[root] >= r[child]) break;
            uint32_t t2 = r[root];
            r[root] = r[child];
            r[child] = t2;
            root = child;
          }
        }
      }
    }
    int64_t *ck = ckeys + d.x;
    for (uint32_t i = 0; i < c; ++i) ck[i] = keys[r[i]];
    if (c > 1) {
      if (c <= 64) {
        for (uint32_t i = 1; i < c && !(local_flags & 2); ++i)
          for (uint32_t j = 0; j < i; ++j)
            if (ck[i] == ck[j]) {
              local_flags |= 2;
              break;
            }
      } else {
        local_flags |= 2;  // > 64 entries in one bucket at load <= 0.5: duplicates
      }
    }
  }
  if (local_flags) atomicOr(flags, local_flags);
  if (local_max) atomicMax(max_chain, (unsigned long long) local_max);
}

// ------------------------------------------------------------- payload columns
struct PayCols {
  const int64_t *src[CC_MAX_PAYLOAD_COLS];
  int64_t *dst[CC_MAX_PAYLOAD_COLS];
  int n;
};

// LP: build row i walks the probe sequence of its key and claims the first slot that holds the key and that no other
// row has claimed yet (duplicates of a key own as many slots as there are rows with that key, so every row finds one),
// then drops its payloads at the slot's index.
__global__ void lp_place_payload_kernel(const int64_t *__restrict__ keys, size_t n, const uint64_t *__restrict__ slots, uint64_t mask,
                                        uint32_t *claimed, PayCols p, int *flags) {
  size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t k = (uint64_t) keys[i];
    uint64_t s = murmurhash64(k) & mask;
    bool placed = false;
    for (uint64_t step = 0; step <= mask; ++step) {
      const uint64_t cur = slots[s];
      if (cur == kEmptyU) break;
      if (cur == k && atomicCAS(claimed + s, 0u, 1u) == 0u) {
        for (int c = 0; c < p.n; ++c) p.dst[c][s] = p.src[c][i];
        placed = true;
        break;
      }
      s = (s + 1) & mask;
    }
    if (!placed) atomicOr(flags, 1);  // d_build_keys is not the column this table was built from
  }
}

// chain: entry at chain position q came from build row rowid[q]
__global__ void chain_place_payload_kernel(const uint32_t *__restrict__ rowid, size_t n, PayCols p) {
  size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t q = (size_t) blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    const uint32_t r = rowid[q];
    for (int c = 0; c < p.n; ++c) p.dst[c][q] = p.src[c][r];
  }
}

__global__ void gen_ref_payload_kernel(int64_t *__restrict__ pay, size_t n) {  // chaining_ht.cpp:21: payload = cnt + 10000000
  size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) pay[i] = (int64_t) (i + 10000000);
}

static int launch_grid(size_t n, int threads, int per_sm = 8) {
  size_t blocks = (n + threads - 1) / threads;
  size_t cap = (size_t) sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  return (int) blocks;
}

// want_slots == 0: the reference's sizing rule; otherwise the caller's power of two (cc_ht_build_sized)
static int build_lp(cc_ht *ht, const int64_t *d_keys, size_t n, int flags, cudaStream_t st, size_t want_slots = 0) {
  size_t ns = 1;
  while (ns < (n << 2)) ns <<= 1;  // linear_probing_ht.cpp:5-6
  if (want_slots) ns = want_slots;
  ht->n_slots = ns;
  ht->mask = ns - 1;
  CC_CUDA(cudaMalloc(&ht->d_slots, ns * sizeof(uint64_t)));
  ht->bytes = ns * sizeof(uint64_t);
  CC_CUDA(cudaMemsetAsync(ht->d_slots, 0xFF, ns * sizeof(uint64_t), st));  // all slots = -1 (:7)
  int *d_flags = nullptr;
  CC_CUDA(cudaMalloc(&d_flags, sizeof(int)));
  CC_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int), st));
  int64_t *scratch = nullptr;
  unsigned long long *pctl = nullptr;
  if (n) {
    // STREAMING BUILD for tables beyond L2: the build keys are first grouped by table slice (the probe side's radix
    // partition, partition.cu), so that the inserts -- CAS on random slots of a 4..8 GiB table otherwise: ncu measured 144 B of
    // DRAM reads per key -- hit one L2-sized slice after the other.  The grid-stride loops below keep every thread of the grid
    // inside one moving window of consecutive keys, i.e. of one or two slices.  The ordered build converges to the same unique
    // layout whatever the insertion order (see lp_insert_ordered_kernel), so the table is unchanged slot for slot.
    // Measurement switch: CCB_BUILD_DIRECT=1 inserts in input order.
    static const bool direct = [] {
      const char *e = getenv("CCB_BUILD_DIRECT");
      return e && e[0] == '1';
    }();
    const int64_t *keys_in = d_keys;
    const size_t table_bytes = ns * sizeof(uint64_t), slice_bytes = (size_t) 32 << 20;
    if (!direct && table_bytes >= ((size_t) 96 << 20) && n >= ((size_t) 1 << 22)) {
      int log2_slots = 0, log2p = 0;
      while (((size_t) 1 << log2_slots) < ns) ++log2_slots;
      while (((size_t) slice_bytes << log2p) < table_bytes && (1 << (log2p + 1)) <= kMaxParts) ++log2p;
      if (log2p >= 1) {
        const int parts = 1 << log2p;
        cudaError_t e = cudaMalloc(&scratch, n * sizeof(int64_t));
        if (e == cudaSuccess) e = cudaMalloc(&pctl, 3 * (size_t) parts * sizeof(unsigned long long));
        if (e == cudaSuccess) {
          int rc = partition_device(d_keys, n, PartFn::slot_bits(ht->mask, log2_slots, log2p), pctl, pctl + parts, pctl + 2 * parts, scratch, st);
          if (rc != CC_OK) {
            cudaFree(scratch);
            cudaFree(pctl);
            cudaFree(d_flags);
            return rc;
          }
          keys_in = scratch;
        } else {  // no room for the partitioned copy: build in input order
          cudaGetLastError();
          if (scratch) cudaFree(scratch);
          scratch = nullptr;
          if (pctl) cudaFree(pctl);
          pctl = nullptr;
        }
      }
    }
    // streaming build: only as many CTAs as are resident at once (8 x 256 threads per SM) -- a second wave of CTAs would walk all
    // the slices a second time
    int grid = launch_grid(n, 256, scratch ? 8 : 16);
    if (flags & CC_BUILD_UNORDERED)
      lp_insert_unordered_kernel<<<grid, 256, 0, st>>>(keys_in, n, ht->d_slots, ht->mask, d_flags);
    else
      lp_insert_ordered_kernel<<<grid, 256, 0, st>>>(keys_in, n, ht->d_slots, ht->mask, d_flags);
    CC_CHECK_LAUNCH();
    lp_audit_kernel<<<grid, 256, 0, st>>>(keys_in, n, ht->d_slots, ht->mask, d_flags);
    CC_CHECK_LAUNCH();
  }
  int h_flags = 0;
  CC_CUDA(cudaMemcpyAsync(&h_flags, d_flags, sizeof(int), cudaMemcpyDeviceToHost, st));
  CC_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_flags);
  if (scratch) cudaFree(scratch);
  if (pctl) cudaFree(pctl);
  if (h_flags & 1) {
    set_error("LP table cannot hold key -1 (empty-slot sentinel, linear_probing_ht.cpp:7)");
    return CC_ERR_UNSUPPORTED;
  }
  if (h_flags & 4) {
    set_error("internal error: LP build lost a key");
    return CC_ERR_CUDA;
  }
  ht->has_duplicates = (h_flags & 2) ? 1 : 0;
  return CC_OK;
}

static int build_chain(cc_ht *ht, const int64_t *d_keys, size_t n, cudaStream_t st, size_t want_slots = 0) {
  CC_REQUIRE(n < 0xFFFFFFFFull, "chain table supports < 2^32 keys per table (got %zu)", n);
  size_t nb = 1;
  while (nb < 2 * n) nb *= 2;  // chaining_ht.cpp:5-6
  if (want_slots) nb = want_slots;
  ht->n_slots = nb;
  ht->mask = nb - 1;
  size_t nalloc = n ? n : 1;
  CC_CUDA(cudaMalloc(&ht->d_dir, nb * sizeof(uint2)));
  CC_CUDA(cudaMalloc(&ht->d_ckeys, nalloc * sizeof(int64_t)));
  CC_CUDA(cudaMalloc(&ht->d_rowid, nalloc * sizeof(uint32_t)));
  ht->bytes = nb * sizeof(uint2) + nalloc * (sizeof(int64_t) + sizeof(uint32_t));
  uint32_t *d_counts = nullptr;
  uint64_t *d_sums = nullptr;
  int *d_flags = nullptr;
  unsigned long long *d_max = nullptr;
  size_t nblocks = (nb + kScanTile - 1) / kScanTile;
  CC_CUDA(cudaMalloc(&d_counts, nb * sizeof(uint32_t)));
  CC_CUDA(cudaMalloc(&d_sums, (nblocks + 1) * sizeof(uint64_t)));
  CC_CUDA(cudaMalloc(&d_flags, sizeof(int)));
  CC_CUDA(cudaMalloc(&d_max, sizeof(unsigned long long)));
  CC_CUDA(cudaMemsetAsync(d_counts, 0, nb * sizeof(uint32_t), st));
  CC_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int), st));
  CC_CUDA(cudaMemsetAsync(d_max, 0, sizeof(unsigned long long), st));
  int grid = launch_grid(n, 256, 16);
  if (n) {
    chain_count_kernel<<<grid, 256, 0, st>>>(d_keys, n, ht->mask, d_counts);
    CC_CHECK_LAUNCH();
  }
  scan_block_sums_kernel<<<(unsigned) nblocks, kScanThreads, 0, st>>>(d_counts, nb, d_sums);
  CC_CHECK_LAUNCH();
  scan_of_sums_kernel<<<1, 1024, 0, st>>>(d_sums, nblocks);
  CC_CHECK_LAUNCH();
  scan_write_dir_kernel<<<(unsigned) nblocks, kScanThreads, 0, st>>>(d_counts, nb, d_sums, ht->d_dir);
  CC_CHECK_LAUNCH();
  if (n) {
    CC_CUDA(cudaMemsetAsync(d_counts, 0, nb * sizeof(uint32_t), st));
    chain_scatter_kernel<<<grid, 256, 0, st>>>(d_keys, n, ht->mask, ht->d_dir, d_counts, ht->d_rowid);
    CC_CHECK_LAUNCH();
    chain_finalize_kernel<<<launch_grid(nb, 128, 16), 128, 0, st>>>(d_keys, nb, ht->d_dir, ht->d_rowid, ht->d_ckeys, d_flags,
                                                                   d_max);
    CC_CHECK_LAUNCH();
  }
  int h_flags = 0;
  unsigned long long h_max = 0;
  CC_CUDA(cudaMemcpyAsync(&h_flags, d_flags, sizeof(int), cudaMemcpyDeviceToHost, st));
  CC_CUDA(cudaMemcpyAsync(&h_max, d_max, sizeof(h_max), cudaMemcpyDeviceToHost, st));
  CC_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_counts);
  cudaFree(d_sums);
  cudaFree(d_flags);
  cudaFree(d_max);
  ht->has_duplicates = (h_flags & 2) ? 1 : 0;
  ht->max_chain = (size_t) h_max;
  return CC_OK;
}

// one bit per bucket / slot (see cc_ht::d_occ)
__global__ void occupancy_kernel(const uint64_t *__restrict__ slots, const uint2 *__restrict__ dir, size_t n, uint32_t *__restrict__ occ) {
  const size_t words = (n + 31) / 32;
  const size_t warps = ((size_t) gridDim.x * blockDim.x) >> 5;
  for (size_t wi = ((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi < words; wi += warps) {
    const size_t i = wi * 32 + (threadIdx.x & 31);
    bool bit = false;
    if (i < n) bit = slots ? slots[i] != kEmptyU : dir[i].y != 0u;
    const unsigned w = __ballot_sync(0xffffffffu, bit);
    if ((threadIdx.x & 31) == 0) occ[wi] = w;
  }
}

static int build_occupancy(cc_ht *ht, cudaStream_t st) {
  const size_t words = (ht->n_slots + 31) / 32;
  if (ht->d_occ) cudaFree(ht->d_occ);
  ht->d_occ = nullptr;
  CC_CUDA(cudaMalloc(&ht->d_occ, words * sizeof(uint32_t)));
  ht->bytes += words * sizeof(uint32_t);
  size_t blocks = (words * 32 + 255) / 256;
  const size_t cap = (size_t) sm_count() * 16;
  if (blocks > cap) blocks = cap;
  occupancy_kernel<<<(unsigned) blocks, 256, 0, st>>>(ht->kind == CC_HT_LP ? ht->d_slots : nullptr, ht->d_dir, ht->n_slots, ht->d_occ);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

static void free_table(cc_ht *ht) {
  if (!ht) return;
  if (ht->d_occ) cudaFree(ht->d_occ);
  if (ht->d_slots) cudaFree(ht->d_slots);
  if (ht->d_dir) cudaFree(ht->d_dir);
  if (ht->d_ckeys) cudaFree(ht->d_ckeys);
  if (ht->d_rowid) cudaFree(ht->d_rowid);
  for (int c = 0; c < CC_MAX_PAYLOAD_COLS; ++c)
    if (ht->d_pay[c]) cudaFree(ht->d_pay[c]);
  delete ht;
}

}  // namespace ccb

using namespace ccb;

extern "C" {

int cc_ht_build(cc_ht **out, int kind, const int64_t *d_keys, size_t n, int flags, cc_stream_t s) {
  return cc_ht_build_sized(out, kind, d_keys, n, 0, flags, s);
}

int cc_ht_build_sized(cc_ht **out, int kind, const int64_t *d_keys, size_t n, size_t n_slots, int flags, cc_stream_t s) {
  CC_REQUIRE(out, "ht is NULL");
  *out = nullptr;
  CC_TRY(require_device());
  CC_REQUIRE(kind == CC_HT_LP || kind == CC_HT_CHAIN, "unknown table kind %d", kind);
  CC_REQUIRE(n == 0 || d_keys, "d_keys is NULL");
  CC_REQUIRE(n <= (1ull << 40), "n too large");
  CC_REQUIRE((n_slots & (n_slots - 1)) == 0, "n_slots (%zu) must be 0 or a power of two", n_slots);
  // an LP table needs empty slots to end its probe sequences: at most half full
  CC_REQUIRE(n_slots == 0 || kind != CC_HT_LP || n_slots >= 2 * n, "an LP table of %zu keys needs at least %zu slots (got %zu)", n, 2 * n, n_slots);
  cc_ht *ht = new cc_ht();
  ht->kind = kind;
  ht->n_keys = n;
  cudaGetDevice(&ht->device);
  int rc = kind == CC_HT_LP ? build_lp(ht, d_keys, n, flags, as_stream(s), n_slots) : build_chain(ht, d_keys, n, as_stream(s), n_slots);
  if (rc == CC_OK) rc = build_occupancy(ht, as_stream(s));
  if (rc != CC_OK) {
    free_table(ht);
    return rc;
  }
  *out = ht;
  return CC_OK;
}

int cc_ht_build_reference(cc_ht **out, int kind, size_t n, size_t cf, cc_stream_t s) {
  CC_REQUIRE(out, "ht is NULL");
  *out = nullptr;
  CC_TRY(require_device());
  CC_REQUIRE(cf > 0, "chunk_factor must be > 0");
  int64_t *d_keys = nullptr;
  CC_CUDA(cudaMalloc(&d_keys, (n ? n : 1) * sizeof(int64_t)));
  int rc = cc_gen_build_keys(d_keys, n, cf, s);
  if (rc == CC_OK) rc = cc_ht_build(out, kind, d_keys, n, CC_BUILD_ORDERED, s);
  cudaStreamSynchronize(as_stream(s));
  cudaFree(d_keys);
  return rc;
}

int cc_ht_attach_payload(cc_ht *ht, const int64_t *d_build_keys, const int64_t *const *h_payload_cols, size_t n_cols, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(ht, "ht is NULL");
  CC_REQUIRE(n_cols >= 1 && n_cols <= CC_MAX_PAYLOAD_COLS, "n_cols must be in [1, %d]", CC_MAX_PAYLOAD_COLS);
  CC_REQUIRE(h_payload_cols, "h_payload_cols is NULL");
  CC_REQUIRE(ht->kind == CC_HT_CHAIN || d_build_keys || ht->n_keys == 0, "an LP table needs d_build_keys to place the payloads");
  CC_REQUIRE(ht->kind == CC_HT_LP || ht->d_rowid, "chain table without row ids");
  for (size_t c = 0; c < n_cols; ++c) CC_REQUIRE(h_payload_cols[c] || ht->n_keys == 0, "payload column %zu is NULL", c);
  cudaStream_t st = as_stream(s);
  for (int c = 0; c < CC_MAX_PAYLOAD_COLS; ++c)
    if (ht->d_pay[c]) {
      cudaFree(ht->d_pay[c]);
      ht->d_pay[c] = nullptr;
      ht->bytes -= (ht->kind == CC_HT_LP ? ht->n_slots : (ht->n_keys ? ht->n_keys : 1)) * sizeof(int64_t);
    }
  ht->n_pay = 0;
  const size_t rows = ht->kind == CC_HT_LP ? ht->n_slots : (ht->n_keys ? ht->n_keys : 1);
  PayCols p;
  p.n = (int) n_cols;
  for (int c = 0; c < CC_MAX_PAYLOAD_COLS; ++c) p.src[c] = nullptr, p.dst[c] = nullptr;
  for (size_t c = 0; c < n_cols; ++c) {
    cudaError_t e = cudaMalloc(&ht->d_pay[c], rows * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMemsetAsync(ht->d_pay[c], 0, rows * sizeof(int64_t), st);
    if (e != cudaSuccess) {
      set_error("cc_ht_attach_payload: %s", cudaGetErrorString(e));
      cudaGetLastError();
      for (int d = 0; d < CC_MAX_PAYLOAD_COLS; ++d)
        if (ht->d_pay[d]) {
          cudaFree(ht->d_pay[d]);
          ht->d_pay[d] = nullptr;
        }
      return e == cudaErrorMemoryAllocation ? CC_ERR_NOMEM : CC_ERR_CUDA;
    }
    p.src[c] = h_payload_cols[c];
    p.dst[c] = ht->d_pay[c];
  }
  int h_flags = 0;
  if (ht->n_keys) {
    const int grid = launch_grid(ht->n_keys, 256, 16);
    if (ht->kind == CC_HT_LP) {
      uint32_t *d_claimed = nullptr;
      int *d_flags = nullptr;
      CC_CUDA(cudaMalloc(&d_claimed, ht->n_slots * sizeof(uint32_t)));
      CC_CUDA(cudaMalloc(&d_flags, sizeof(int)));
      CC_CUDA(cudaMemsetAsync(d_claimed, 0, ht->n_slots * sizeof(uint32_t), st));
      CC_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int), st));
      lp_place_payload_kernel<<<grid, 256, 0, st>>>(d_build_keys, ht->n_keys, ht->d_slots, ht->mask, d_claimed, p, d_flags);
      note_launch();
      cudaError_t e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaMemcpyAsync(&h_flags, d_flags, sizeof(int), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      cudaFree(d_claimed);
      cudaFree(d_flags);
      CC_CUDA(e);
    } else {
      chain_place_payload_kernel<<<grid, 256, 0, st>>>(ht->d_rowid, ht->n_keys, p);
      CC_CHECK_LAUNCH();
      CC_CUDA(cudaStreamSynchronize(st));
    }
  }
  if (h_flags) {
    for (int c = 0; c < CC_MAX_PAYLOAD_COLS; ++c)
      if (ht->d_pay[c]) {
        cudaFree(ht->d_pay[c]);
        ht->d_pay[c] = nullptr;
      }
    set_error("cc_ht_attach_payload: d_build_keys is not the key column this table was built from");
    return CC_ERR_INVALID;
  }
  ht->n_pay = (int) n_cols;
  ht->bytes += n_cols * rows * sizeof(int64_t);
  return CC_OK;
}

int cc_ht_build_reference_payload(cc_ht **out, int kind, size_t n, size_t cf, cc_stream_t s) {
  CC_REQUIRE(out, "ht is NULL");
  *out = nullptr;
  CC_TRY(require_device());
  CC_REQUIRE(cf > 0, "chunk_factor must be > 0");
  cudaStream_t st = as_stream(s);
  int64_t *d_keys = nullptr, *d_pay = nullptr;
  CC_CUDA(cudaMalloc(&d_keys, (n ? n : 1) * sizeof(int64_t)));
  cudaError_t e = cudaMalloc(&d_pay, (n ? n : 1) * sizeof(int64_t));
  if (e != cudaSuccess) {
    cudaFree(d_keys);
    CC_CUDA(e);
  }
  int rc = cc_gen_build_keys(d_keys, n, cf, s);
  if (rc == CC_OK && n) {
    gen_ref_payload_kernel<<<launch_grid(n, 256, 8), 256, 0, st>>>(d_pay, n);
    note_launch();
    if (cudaGetLastError() != cudaSuccess) rc = CC_ERR_CUDA;
  }
  if (rc == CC_OK) rc = cc_ht_build(out, kind, d_keys, n, CC_BUILD_ORDERED, s);
  if (rc == CC_OK) {
    const int64_t *cols[1] = {d_pay};
    rc = cc_ht_attach_payload(*out, d_keys, cols, 1, s);
    if (rc != CC_OK) {
      free_table(*out);
      *out = nullptr;
    }
  }
  cudaStreamSynchronize(st);
  cudaFree(d_keys);
  cudaFree(d_pay);
  return rc;
}

size_t cc_ht_payload_cols(const cc_ht *ht) { return ht ? (size_t) ht->n_pay : 0; }

int cc_ht_export_payload(const cc_ht *ht, int64_t *const *h_cols) {
  CC_REQUIRE(ht && h_cols, "NULL argument");
  const size_t rows = ht->kind == CC_HT_LP ? ht->n_slots : ht->n_keys;
  for (int c = 0; c < ht->n_pay; ++c) {
    CC_REQUIRE(h_cols[c], "h_cols[%d] is NULL", c);
    if (rows) CC_CUDA(cudaMemcpy(h_cols[c], ht->d_pay[c], rows * sizeof(int64_t), cudaMemcpyDeviceToHost));
  }
  return CC_OK;
}

int cc_ht_import_lp(cc_ht **out, const int64_t *h_slots, size_t n_slots, size_t n_keys, cc_stream_t s) {
  CC_REQUIRE(out && h_slots, "NULL argument");
  *out = nullptr;
  CC_TRY(require_device());
  CC_REQUIRE(n_slots > 0 && (n_slots & (n_slots - 1)) == 0, "n_slots must be a power of two");
  cc_ht *ht = new cc_ht();
  ht->kind = CC_HT_LP;
  ht->n_keys = n_keys;
  ht->n_slots = n_slots;
  ht->mask = n_slots - 1;
  ht->has_duplicates = 1;  // unknown: be conservative (walk past matches like the reference)
  cudaGetDevice(&ht->device);
  cudaError_t e = cudaMalloc(&ht->d_slots, n_slots * sizeof(uint64_t));
  if (e == cudaSuccess) e = cudaMemcpyAsync(ht->d_slots, h_slots, n_slots * sizeof(uint64_t), cudaMemcpyHostToDevice, as_stream(s));
  if (e == cudaSuccess) e = cudaStreamSynchronize(as_stream(s));
  if (e != cudaSuccess) {
    set_error("cc_ht_import_lp: %s", cudaGetErrorString(e));
    free_table(ht);
    return CC_ERR_CUDA;
  }
  ht->bytes = n_slots * sizeof(uint64_t);
  int rc = build_occupancy(ht, as_stream(s));
  if (rc == CC_OK && cudaStreamSynchronize(as_stream(s)) != cudaSuccess) rc = CC_ERR_CUDA;
  if (rc != CC_OK) {
    free_table(ht);
    return rc;
  }
  *out = ht;
  return CC_OK;
}

int cc_ht_get_info(const cc_ht *ht, cc_ht_info *info) {
  CC_REQUIRE(ht && info, "NULL argument");
  info->kind = ht->kind;
  info->n_keys = ht->n_keys;
  info->n_slots = ht->n_slots;
  info->bytes = ht->bytes;
  info->has_duplicates = ht->has_duplicates;
  info->max_chain = ht->max_chain;
  return CC_OK;
}

int cc_ht_export_lp(const cc_ht *ht, int64_t *h_slots) {
  CC_REQUIRE(ht && h_slots && ht->kind == CC_HT_LP, "not an LP table");
  CC_CUDA(cudaMemcpy(h_slots, ht->d_slots, ht->n_slots * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return CC_OK;
}

int cc_ht_export_chain(const cc_ht *ht, uint32_t *h_begin, uint32_t *h_count, int64_t *h_keys) {
  CC_REQUIRE(ht && ht->kind == CC_HT_CHAIN, "not a chain table");
  std::vector<uint2> dir(ht->n_slots);
  CC_CUDA(cudaMemcpy(dir.data(), ht->d_dir, ht->n_slots * sizeof(uint2), cudaMemcpyDeviceToHost));
  for (size_t b = 0; b < ht->n_slots; ++b) {
    if (h_begin) h_begin[b] = dir[b].x;
    if (h_count) h_count[b] = dir[b].y;
  }
  if (h_keys && ht->n_keys) CC_CUDA(cudaMemcpy(h_keys, ht->d_ckeys, ht->n_keys * sizeof(int64_t), cudaMemcpyDeviceToHost));
  return CC_OK;
}

int cc_ht_destroy(cc_ht *ht) {
  free_table(ht);
  return CC_OK;
}

}  // extern "C"
