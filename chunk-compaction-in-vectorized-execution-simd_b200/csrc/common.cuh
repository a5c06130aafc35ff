// common.cuh -- shared declarations of libccb200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "cc_api.h"

namespace ccb {

// ---- error plumbing ---------------------------------------------------------
void set_error(const char *fmt, ...);
int require_device();  // CC_OK or CC_ERR_NO_DEVICE (no CPU fallback anywhere)
void note_launch();    // counts kernel launches (cc_launch_count)

#define CC_CUDA(expr)                                                                 \
  do {                                                                                \
    cudaError_t e__ = (expr);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      ccb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      return e__ == cudaErrorMemoryAllocation ? CC_ERR_NOMEM : CC_ERR_CUDA;           \
    }                                                                                 \
  } while (0)

#define CC_REQUIRE(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      ccb::set_error(__VA_ARGS__);   \
      return CC_ERR_INVALID;         \
    }                                \
  } while (0)

#define CC_CHECK_LAUNCH()            \
  do {                               \
    ccb::note_launch();              \
    CC_CUDA(cudaGetLastError());     \
  } while (0)

#define CC_TRY(expr)                 \
  do {                               \
    int rc__ = (expr);               \
    if (rc__ != CC_OK) return rc__;  \
  } while (0)

inline cudaStream_t as_stream(cc_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- constants ----------------------------------------------------------------
constexpr uint64_t kEmptyU = 0xFFFFFFFFFFFFFFFFULL;  // LP empty slot == int64 -1 (linear_probing_ht.cpp:7)
constexpr int kSmCountFallback = 148;

int sm_count();

// ---- hash_functions.h:8-16 (bit-exact) ------------------------------------------
__host__ __device__ __forceinline__ uint64_t murmurhash64(uint64_t x) {
  x ^= x >> 32;
  x *= 0xd6e8feb86659fd93ULL;
  x ^= x >> 32;
  x *= 0xd6e8feb86659fd93ULL;
  x ^= x >> 32;
  return x;
}

// ---- device helpers -------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
__device__ __forceinline__ uint64_t ld_cg_u64(const uint64_t *p) {  // L2-coherent load (no L1)
  uint64_t v;
  asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint64_t ld_nc_u64(const uint64_t *p) { return __ldg(reinterpret_cast<const unsigned long long *>(p)); }
// one aligned 32-byte HBM sector == 4 table slots in a single request (LDG.E.ENL2.256)
__device__ __forceinline__ void ld_sector_u64x4(const uint64_t *p, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d) {
  asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane_id() >= (unsigned) o) v += t;
  }
  return v;
}
#endif

}  // namespace ccb

// ---- handle layouts ---------------------------------------------------------------
struct cc_ht {
  int kind;
  int device;
  size_t n_keys;
  size_t n_slots;  // LP: slots, chain: buckets (power of two)
  uint64_t mask;
  // LP: open addressing, key-only slots, -1 == empty (linear_probing_ht.h:65)
  uint64_t *d_slots;
  // chain: bucket directory + contiguous chains.  dir[b] = (begin, count); the bucket's
  // keys are d_ckeys[begin .. begin+count) in insertion (FIFO) order, i.e. the std::list
  // of chaining_ht.h:97 laid out flat: iterator == position, end() == begin+count.
  uint2 *d_dir;
  int64_t *d_ckeys;
  uint32_t *d_rowid;  // build-side row id of each chain entry (payload hook)
  // occupancy bitmap: bit i set <=> bucket i is non-empty (chain) / slot i holds a key (LP).  It answers the reference's
  // "drop lanes with an empty bucket / first slot" (chaining_ht.cpp:52-55, linear_probing_ht.cpp:53-57) from 1 bit per
  // entry -- 64x smaller than the directory, so it stays L2-resident where the directories of a join chain do not.
  uint32_t *d_occ;
  int has_duplicates;
  size_t max_chain;
  size_t bytes;
  // payload columns: d_pay[c][i] belongs to the key at index i (LP: slot, chain: chain position)
  int n_pay;
  int64_t *d_pay[CC_MAX_PAYLOAD_COLS];
};
