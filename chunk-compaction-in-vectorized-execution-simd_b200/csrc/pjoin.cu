// pjoin.cu -- the hash-partitioned multi-GPU join behind the C ABI (new functionality, SURVEY 8e / 8b "cc_partition_exchange"):
// one process per GPU, NO collective library on the data path.
//
// Both sides of an equi-join are partitioned by owner = murmurhash64(key) >> (64 - log2 P) (high hash bits, independent of the
// low bits that address the owner's table) and every rank builds / probes its own table with the single-GPU kernels
// (linear_probing_ht.cpp:4-115 / chaining_ht.cpp:4-136 per rank).
//
// The probe side is partitioned ONCE, on the sender: partition_scatter_kernel groups a piece of the probe keys by
// (owner GPU, table slice of that owner) -- every rank's table has the same size, so a sender knows the slice a key will fall
// into on its owner.  What an owner receives is therefore already grouped by L2-sized table slices and is probed as it lies:
// per key the SMs do one scatter and one probe, exactly what a single GPU does for the same keys (the round-1 pipeline did an
// owner partition on the sender AND a slice partition on the receiver: 2x the SM work per key, scaling efficiency 0.53).
// The exchange itself is made of device-side pieces only:
//   * copy engines: the regions of owner o travel as ONE cudaMemcpyAsync into slot [piece][this rank] of o's receive arena
//     (CUDA-IPC mapping of peer memory over NVLink 5 / NVSwitch), on several copy streams, without occupying an SM; the rows a
//     rank keeps are written straight into its own arena by the scatter kernel;
//   * signal: behind the copies in stream order, the regions' row counts (a small copy) and an epoch flag are written into the
//     owner's control block over NVLink ("piece b of this step from sender s has landed");
//   * wait: the owner's stream waits for the flags of the pieces it is about to probe, then a probe walks their part of the
//     arena slice by slice (all pieces and senders of slice 0, then slice 1, ...); a `consumed` flag then tells every sender
//     that the arena may be refilled (two arenas alternate between steps, so the copies of step t + 1 land while step t is
//     still being probed).
//   Signals and waits are STREAM MEMORY OPERATIONS (cuStreamWriteValue64 / cuStreamWaitValue64, resolved through
//   cudaGetDriverEntryPoint): they need no SM.  Small kernels were measured first and starve: the scatter and probe kernels
//   fill every SM's register file, so a one-CTA signal kernel on another stream only runs at the next kernel boundary -- and the
//   copies queued behind it wait with it (profiles/r2_pjoin_timeline_n2.txt).  The kernels remain as the fallback for drivers
//   without stream memory operations on peer memory (bounded spins, error bit instead of a hang).
// The host never learns a count and never blocks: a probe call enqueues  F(0) F(1) .. F(B-1) W PROBE C  and returns.  The host's
// only job is the control plane at create / destroy time (exchange of the IPC handles and sizes through the caller's cc_comm
// callbacks: MPI, torch.distributed, or the fork + shared-memory communicator of host/simd_compaction.hpp).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include <cuda.h>  // types of the stream memory operations only: the entry points are resolved at run time, libcuda is not linked

#include "common.cuh"
#include "partition.cuh"

namespace ccb {

constexpr int kPjArenas = 2;        // receive arenas, alternating between steps
constexpr int kPjMaxCopyStreams = 8; // one stream drives one copy engine at a time (CCB_PJ_COPY_STREAMS, default 4)
constexpr int kPjMaxPieces = 16;    // pieces per batch; send slots = 2 x pieces: a batch is partitioned without waiting for the copies of the one before
constexpr unsigned long long kPjSpinNs = 20ull * 1000 * 1000 * 1000;  // a wait gives up after 20 s (a peer died): error bit, no hang
constexpr size_t kPjSliceBytes = 32u << 20;                           // target table bytes per slice (profiles/r1_sweep_slices.txt)
constexpr size_t kPjSliceMinTable = 96u << 20;                        // smaller tables are probed directly (they live in L2 anyway)

// Control block at the head of every rank's exchange allocation, written by PEERS over NVLink.  Geometry-independent layout
// (strides are the ALLOCATED piece / slice counts): flag index (arena * Ba + piece) * P + sender, count index that * Sa + slice.
struct PjLayout {
  int P = 1, Ba = 1, Sa = 1;      // ranks, pieces allocated, slices allocated
  unsigned long long cap = 0;     // rows per region
  size_t ready_off = 0, consumed_off = 0, counts_off = 0, data_off = 0, total = 0;
  __host__ __device__ size_t flag_index(int arena, int piece, int sender) const { return ((size_t) arena * Ba + piece) * P + sender; }
  __host__ __device__ size_t region_index(int piece, int sender, int slice) const { return ((size_t) piece * P + sender) * Sa + slice; }
  size_t arena_rows() const { return (size_t) Ba * P * Sa * cap; }
  void finish() {
    auto up = [](size_t x) { return (x + 4095) / 4096 * 4096; };
    ready_off = 0;
    consumed_off = up(ready_off + (size_t) kPjArenas * Ba * P * 8);
    counts_off = up(consumed_off + (size_t) (kPjArenas + 1) * kMaxPeers * 8);  // (+ one row of scratch words for the memory-operation self-test)
    data_off = up(counts_off + (size_t) kPjArenas * Ba * P * Sa * 8);
    total = data_off + (size_t) kPjArenas * arena_rows() * 8;
  }
};

struct PjPeers {
  unsigned char *base[kMaxPeers];
};

static int log2_floor_pj(size_t x) {
  int l = 0;
  while ((x >> l) > 1) ++l;
  return l;
}

// stream memory operations of the driver API, resolved at run time (no link dependency on libcuda)
struct PjMemOps {
  typedef CUresult (*write_fn)(CUstream, CUdeviceptr, cuuint64_t, unsigned int);
  typedef CUresult (*wait_fn)(CUstream, CUdeviceptr, cuuint64_t, unsigned int);
  write_fn write = nullptr;
  wait_fn wait = nullptr;
  bool ok() const { return write && wait; }
  static PjMemOps resolve() {
    PjMemOps m;
    const char *off = getenv("CCB_PJ_NO_MEMOPS");  // measurement / fallback switch: signal and wait with kernels
    if (off && off[0] == '1') return m;
    void *w = nullptr, *q = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue64", &w, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess &&
        cudaGetDriverEntryPoint("cuStreamWaitValue64", &q, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess) {
      m.write = reinterpret_cast<write_fn>(w);
      m.wait = reinterpret_cast<wait_fn>(q);
    }
    cudaGetLastError();
    return m;
  }
};

__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// block o < P: tell owner o that this rank's regions of (arena, piece) have landed -- S row counts, then the epoch flag
__global__ void pj_signal_kernel(PjPeers peers, PjLayout lay, int rank, int arena, int piece, int S, const unsigned long long *__restrict__ d_counts,
                                 unsigned long long epoch) {
  const int o = blockIdx.x;
  unsigned long long *counts = reinterpret_cast<unsigned long long *>(peers.base[o] + lay.counts_off) +
                               ((size_t) arena * lay.Ba * lay.P * lay.Sa + lay.region_index(piece, rank, 0));
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    unsigned long long c = d_counts[(size_t) o * S + s];
    counts[s] = c < lay.cap ? c : lay.cap;  // an overrun region was clamped by the scatter kernel (and flagged)
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0)
    st_sys_u64(reinterpret_cast<unsigned long long *>(peers.base[o] + lay.ready_off) + lay.flag_index(arena, piece, rank), epoch);
}

// thread i < n: wait until flags[i] >= epoch (the flags live in THIS rank's memory, peers write them); bounded
__global__ void pj_wait_kernel(const unsigned long long *flags, int n, unsigned long long epoch, int *d_err) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned long long t0 = globaltimer_ns();
    while (ld_sys_u64(flags + i) < epoch) {
      __nanosleep(200);
      if (globaltimer_ns() - t0 > kPjSpinNs) {
        atomicOr(d_err, 1);
        return;
      }
    }
  }
}

// thread s < P: tell sender s that this rank has finished reading its arena `arena` (use `epoch`)
__global__ void pj_consumed_kernel(PjPeers peers, PjLayout lay, int rank, int arena, unsigned long long epoch) {
  const int s = threadIdx.x;
  if (s >= lay.P) return;
  st_sys_u64(reinterpret_cast<unsigned long long *>(peers.base[s] + lay.consumed_off) + (size_t) arena * kMaxPeers + rank, epoch);
}

// append the valid rows of `segments` regions (region q: counts[q * count_stride] rows at src + q * row_stride) to
// dst[*cursor ...]; rows beyond dst_cap are dropped (the cursor still counts them: the host sees the overflow)
__global__ void pj_compact_kernel(const int64_t *__restrict__ src, const unsigned long long *__restrict__ counts, unsigned long long cap, int segments,
                                  size_t row_stride, size_t count_stride, int64_t *__restrict__ dst, unsigned long long dst_cap,
                                  const unsigned long long *__restrict__ cursor) {
  const unsigned long long base0 = *cursor;
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  unsigned long long before = base0;
  for (int q = 0; q < segments; ++q) {
    const unsigned long long raw = counts[(size_t) q * count_stride];
    const unsigned long long c = raw < cap ? raw : cap;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < c; i += stride)
      if (before + i < dst_cap) dst[before + i] = src[(size_t) q * row_stride + i];
    before += c;
  }
}
__global__ void pj_advance_kernel(const unsigned long long *__restrict__ counts, unsigned long long cap, int segments, size_t count_stride,
                                  unsigned long long *cursor) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long t = 0;
    for (int q = 0; q < segments; ++q) {
      const unsigned long long raw = counts[(size_t) q * count_stride];
      t += raw < cap ? raw : cap;
    }
    *cursor += t;
  }
}

// SM-driven block copy (the alternative to a copy engine for peer memory): `blocks` CTAs stream src -> dst with 16-byte loads and
// stores, 8 per thread in flight.  Measurement tool for the NVLink store rate of the SMs (tools/nvlink_bench.py).
__global__ void __launch_bounds__(256) pj_sm_copy_kernel(uint4 *__restrict__ dst, const uint4 *__restrict__ src, size_t n16) {
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 7 * stride < n16; i += 8 * stride) {
    uint4 v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = __ldg(src + i + q * stride);
#pragma unroll
    for (int q = 0; q < 8; ++q) dst[i + q * stride] = v[q];
  }
  for (; i < n16; i += stride) dst[i] = __ldg(src + i);
}

// The SM-driven tail of a piece's block copies (see cc_pjoin::sm_pct): up to kMaxPeers jobs, all CTAs of the grid stream through
// the concatenation of the jobs with 16-byte loads / stores.
struct PjCopyJobs {
  uint4 *dst[kMaxPeers];
  const uint4 *src[kMaxPeers];
  unsigned long long end16[kMaxPeers];  // running end of job q in 16-byte units (prefix sum)
  int n;
};
__global__ void __launch_bounds__(256) pj_sm_copy_jobs_kernel(PjCopyJobs jobs) {
  const unsigned long long total = jobs.n ? jobs.end16[jobs.n - 1] : 0;
  const unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int q = 0;
    while (i >= jobs.end16[q]) ++q;
    const unsigned long long at = i - (q ? jobs.end16[q - 1] : 0ull);
    jobs.dst[q][at] = __ldg(jobs.src[q] + at);
  }
}

__global__ void pj_close_kernel(cc_probe_result *res, size_t cap, int *region_flag, int *err) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    res->overflow = (res->n_matches > cap ? 1 : 0) | (*region_flag ? 2 : 0) | (*err ? 4 : 0);
    *region_flag = 0;  // the next probe starts clean; a wait timeout (*err) stays: the join is unusable after it
  }
}

}  // namespace ccb

using namespace ccb;

struct cc_pjoin {
  cc_comm comm;
  int world = 1, rank = 0, log2p = 0, kind = CC_HT_LP, device = 0;
  int pieces = 1;                 // pieces (sub-batches) of one probe call
  int slices = 1, log2s = 0;      // table slices the probe side is grouped by on the sender (1: the table lives in L2)
  size_t max_rows = 0;            // rows per piece
  PjLayout lay;
  unsigned char *block = nullptr;              // own exchange allocation: flags | counts | kPjArenas arenas
  unsigned char *peer_block[kMaxPeers] = {};   // every rank's allocation in this address space (own pointer at [rank])
  int n_slots = 1;                             // send slots (= pieces: the partition pass of a batch never waits for its copies)
  int n_cs = 4;                                // copy streams in use
  int n_direct = 0;                            // remote owners whose regions the scatter kernel stores straight into their arena
  int sm_pct = 0;                              // per cent of every block copy that an SM kernel moves behind the partition pass (see create)
  struct Deferred {                            // pieces whose signal waits for the SM copy at the end of cc_pjoin_probe_begin
    int arena, piece, S, slot;
    unsigned long long epoch;
  };
  std::vector<Deferred> deferred;
  std::vector<PjCopyJobs> sm_jobs;             // one job list per piece of the batch being begun
  cudaEvent_t sm_done = nullptr;
  int64_t *send[2 * kPjMaxPieces] = {};        // [owner][slice][cap]
  unsigned long long *d_counts = nullptr;      // [n_slots][P * Sa] fill counts of the scatter
  int *d_flag = nullptr, *d_err = nullptr;     // sticky region-overrun flag, wait-timeout flag
  cudaStream_t cs[kPjMaxCopyStreams] = {};
  cudaEvent_t parted[2 * kPjMaxPieces] = {}, copied[2 * kPjMaxPieces] = {}, gate = nullptr, joined[kPjMaxCopyStreams] = {};
  unsigned long long sends = 0;                // pieces sent so far (send slot = sends % n_slots)
  unsigned long long uses = 0;                 // arena uses so far (arena = uses % kPjArenas, its epoch = uses / kPjArenas + 1)
  cc_ht *table = nullptr;
  size_t n_build_total = 0, table_slots = 0;
  PjMemOps memops;                             // stream memory operations (empty: signal / wait with kernels)
  struct Batch {
    int arena;
    unsigned long long epoch, call;
  };
  std::vector<Batch> begun;                    // probe batches whose exchange is under way (cc_pjoin_probe_begin .. _end)
  unsigned long long calls = 0;
  std::vector<std::pair<cudaEvent_t, std::string>> marks;  // CCB_PJ_TRACE

  int64_t *arena(int r, int a) const { return reinterpret_cast<int64_t *>(peer_block[r] + lay.data_off) + (size_t) a * lay.arena_rows(); }
  unsigned long long *counts(int a) const {
    return reinterpret_cast<unsigned long long *>(block + lay.counts_off) + (size_t) a * lay.Ba * lay.P * lay.Sa;
  }
};

namespace {

#define PJ_CUDA(expr)                                                                     \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      cudaGetLastError();                                                                 \
      return e__ == cudaErrorMemoryAllocation ? CC_ERR_NOMEM : CC_ERR_CUDA;               \
    }                                                                                     \
  } while (0)

PjPeers peers_of(const cc_pjoin *j) {
  PjPeers p;
  for (int r = 0; r < kMaxPeers; ++r) p.base[r] = r < j->world ? j->peer_block[r] : nullptr;
  return p;
}

unsigned long long *ready_flag(const cc_pjoin *j, int r, int arena, int piece, int sender) {
  return reinterpret_cast<unsigned long long *>(j->peer_block[r] + j->lay.ready_off) + j->lay.flag_index(arena, piece, sender);
}
unsigned long long *consumed_flag(const cc_pjoin *j, int r, int arena, int receiver) {
  return reinterpret_cast<unsigned long long *>(j->peer_block[r] + j->lay.consumed_off) + (size_t) arena * kMaxPeers + receiver;
}

#define PJ_DRV(expr)                                                                    \
  do {                                                                                  \
    CUresult r__ = (expr);                                                              \
    if (r__ != CUDA_SUCCESS) {                                                          \
      set_error("%s:%d: %s failed: CUresult %d", __FILE__, __LINE__, #expr, (int) r__); \
      return CC_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

// `n` consecutive flags of THIS rank's control block must reach `epoch` before the work behind it on `st` runs
int wait_flags(cc_pjoin *j, unsigned long long *flags, int n, unsigned long long epoch, cudaStream_t st) {
  if (j->memops.ok()) {
    for (int i = 0; i < n; ++i) PJ_DRV(j->memops.wait((CUstream) st, (CUdeviceptr) (flags + i), epoch, CU_STREAM_WAIT_VALUE_GEQ));
    return CC_OK;
  }
  pj_wait_kernel<<<1, 256, 0, st>>>(flags, n, epoch, j->d_err);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

// tells every owner (in stream order on `st`) that this rank's regions of (arena, piece) have landed, with their row counts
int signal_piece(cc_pjoin *j, int arena, int piece, int S, const unsigned long long *d_counts, unsigned long long epoch, cudaStream_t st) {
  const PjLayout &lay = j->lay;
  if (!j->memops.ok()) {
    pj_signal_kernel<<<j->world, 128, 0, st>>>(peers_of(j), lay, j->rank, arena, piece, S, d_counts, epoch);
    CC_CHECK_LAUNCH();
    return CC_OK;
  }
  for (int i = 0; i < j->world; ++i) {
    const int o = (j->rank + i) % j->world;
    unsigned long long *counts = reinterpret_cast<unsigned long long *>(j->peer_block[o] + lay.counts_off) +
                                 ((size_t) arena * lay.Ba * lay.P * lay.Sa + lay.region_index(piece, j->rank, 0));
    // raw fill counts (an overrun region was flagged by the scatter kernel; every reader clamps to the region capacity)
    PJ_CUDA(cudaMemcpyAsync(counts, d_counts + (size_t) o * S, (size_t) S * 8, cudaMemcpyDeviceToDevice, st));
    PJ_DRV(j->memops.write((CUstream) st, (CUdeviceptr) ready_flag(j, o, arena, piece, j->rank), epoch, CU_STREAM_WRITE_VALUE_DEFAULT));
  }
  return CC_OK;
}

// tells every sender that this rank has finished reading its arena `arena` (use `epoch`)
int signal_consumed(cc_pjoin *j, int arena, unsigned long long epoch, cudaStream_t st) {
  if (!j->memops.ok()) {
    pj_consumed_kernel<<<1, 32, 0, st>>>(peers_of(j), j->lay, j->rank, arena, epoch);
    CC_CHECK_LAUNCH();
    return CC_OK;
  }
  for (int s = 0; s < j->world; ++s)
    PJ_DRV(j->memops.write((CUstream) st, (CUdeviceptr) consumed_flag(j, s, arena, j->rank), epoch, CU_STREAM_WRITE_VALUE_DEFAULT));
  return CC_OK;
}

// Sends one piece: the keys are grouped by (owner, slice) with `S` slices per owner (S == 1: by owner only) into a send slot,
// owner o's regions are copied into [piece][this rank] of o's arena `arena`, then o is signalled with epoch `epoch`.
// first_of_use: the arena is about to be refilled for the first time in this use -- wait until every owner has consumed its
// previous use.
int send_piece(cc_pjoin *j, const int64_t *d_keys, size_t n, int S, int log2s, int arena, int piece, unsigned long long epoch, bool first_of_use,
               cudaStream_t st, bool sm_share = false) {
  const int P = j->world;
  const PjLayout &lay = j->lay;
  const unsigned long long k = j->sends++;
  const int slot = (int) (k % j->n_slots);
  unsigned long long *counts = j->d_counts + (size_t) slot * P * lay.Sa;
  if (k >= (unsigned long long) j->n_slots) PJ_CUDA(cudaStreamWaitEvent(st, j->copied[slot], 0));  // the copies of piece k - n_slots have left the slot
  const PartFn fn = S > 1 ? PartFn::owner_and_slice(j->log2p, j->table_slots - 1, log2_floor_pj(j->table_slots), log2s)
                          : PartFn::high_bits(j->log2p);
  // Region q = owner * S + slice sits at q * cap in the send slot.  In owner o's arena the same region lives at
  // region_index(piece, rank, slice) * cap, so a pointer into an arena is shifted accordingly.  In place go: the rows this rank
  // keeps, and the rows of the first n_direct remote owners (the scatter kernel stores them over NVLink itself -- SM stores and
  // copy engines then share the exchange).  A geometry with fewer slices than allocated (the build side: S == 1) cannot be
  // written in place: the arena's slice stride is the allocated one.
  const bool in_place = S == lay.Sa && P > 1;
  ScatterDst dsts;
  bool direct[kMaxPeers];
  for (int o = 0; o < kMaxPeers; ++o) {
    dsts.p[o] = nullptr;
    direct[o] = false;
  }
  if (in_place) {
    for (int i = 0; i <= j->n_direct && i < P; ++i) {
      const int o = (j->rank + i) % P;
      dsts.p[o] = j->arena(o, arena) + lay.region_index(piece, j->rank, 0) * lay.cap - (size_t) o * S * lay.cap;
      direct[o] = true;
    }
    // stores into a peer's arena: that owner must have consumed the arena's previous use BEFORE the kernel runs
    if (first_of_use && epoch > 1 && j->n_direct > 0) CC_TRY(wait_flags(j, consumed_flag(j, j->rank, arena, 0), P, epoch - 1, st));
  }
  CC_TRY(partition_single_multi(d_keys, n, fn, lay.cap, counts, j->d_flag, j->send[slot], in_place ? &dsts : nullptr, st, /*sticky_flag=*/true));
  PJ_CUDA(cudaEventRecord(j->parted[slot], st));
  cudaStream_t c0 = j->cs[0];
  PJ_CUDA(cudaStreamWaitEvent(c0, j->parted[slot], 0));
  if (first_of_use && epoch > 1) CC_TRY(wait_flags(j, consumed_flag(j, j->rank, arena, 0), P, epoch - 1, c0));
  PJ_CUDA(cudaEventRecord(j->gate, c0));
  for (int s = 1; s < j->n_cs; ++s) PJ_CUDA(cudaStreamWaitEvent(j->cs[s], j->gate, 0));
  // one stream drives one copy engine at a time and a single engine does not fill an NVLink 5 port (measured at P = 2 with one
  // copy per piece: 450 GB/s): with fewer destinations than copy streams every block is cut into chunks dealt over the streams
  const size_t rows = (size_t) S * lay.cap;
  int dests = 0;
  for (int o = 0; o < P; ++o) dests += direct[o] ? 0 : 1;
  const int chunks = (dests == 0 || dests >= j->n_cs) ? 1 : (j->n_cs + dests - 1) / dests;
  const size_t chunk_rows = ((rows + chunks - 1) / chunks + 511) / 512 * 512;
  int n_copy = 0;
  // sm_share: the last rows_sm rows of every block are left to an SM copy kernel that cc_pjoin_probe_begin launches behind the
  // partition pass of the whole batch -- in the steady state of a copy-bound step the SMs would otherwise idle there, waiting
  // for the copy engines (which run at half their idle rate under the kernels)
  const size_t rows_sm = (sm_share && in_place) ? ((rows * (size_t) j->sm_pct / 100) / 512 * 512) : 0;
  const size_t rows_ce = rows - rows_sm;
  PjCopyJobs jobs;
  jobs.n = 0;
  for (int i = 0; i < P; ++i) {
    const int o = (j->rank + i) % P;  // stagger the destinations so that the ranks do not all hit the same peer at once
    if (direct[o]) continue;          // written in place by the scatter kernel
    int64_t *dst = j->arena(o, arena) + lay.region_index(piece, j->rank, 0) * lay.cap;
    const int64_t *src = j->send[slot] + (size_t) o * rows;
    for (size_t at = 0; at < rows_ce; at += chunk_rows)
      PJ_CUDA(cudaMemcpyAsync(dst + at, src + at, std::min(chunk_rows, rows_ce - at) * 8, cudaMemcpyDeviceToDevice, j->cs[n_copy++ % j->n_cs]));
    if (rows_sm) {
      jobs.dst[jobs.n] = reinterpret_cast<uint4 *>(dst + rows_ce);
      jobs.src[jobs.n] = reinterpret_cast<const uint4 *>(src + rows_ce);
      jobs.end16[jobs.n] = (jobs.n ? jobs.end16[jobs.n - 1] : 0ull) + rows_sm / 2;
      ++jobs.n;
    }
  }
  for (int s = 1; s < j->n_cs; ++s) {
    PJ_CUDA(cudaEventRecord(j->joined[s], j->cs[s]));
    PJ_CUDA(cudaStreamWaitEvent(c0, j->joined[s], 0));
  }
  if (rows_sm) {  // signalled by cc_pjoin_probe_begin once the SM copy has run
    j->sm_jobs.push_back(jobs);
    j->deferred.push_back({arena, piece, S, slot, epoch});
    return CC_OK;
  }
  CC_TRY(signal_piece(j, arena, piece, S, counts, epoch, c0));
  PJ_CUDA(cudaEventRecord(j->copied[slot], c0));
  return CC_OK;
}

void pj_release(cc_pjoin *j) {
  if (!j) return;
  for (int r = 0; r < j->world; ++r)
    if (r != j->rank && j->peer_block[r]) cudaIpcCloseMemHandle(j->peer_block[r]);
  if (j->block) cudaFree(j->block);
  for (auto &p : j->send)
    if (p) cudaFree(p);
  if (j->d_counts) cudaFree(j->d_counts);
  if (j->d_flag) cudaFree(j->d_flag);
  if (j->d_err) cudaFree(j->d_err);
  for (auto &s : j->cs)
    if (s) cudaStreamDestroy(s);
  for (auto &e : j->parted)
    if (e) cudaEventDestroy(e);
  for (auto &e : j->copied)
    if (e) cudaEventDestroy(e);
  for (auto &e : j->joined)
    if (e) cudaEventDestroy(e);
  if (j->gate) cudaEventDestroy(j->gate);
  if (j->sm_done) cudaEventDestroy(j->sm_done);
  if (j->table) cc_ht_destroy(j->table);
  delete j;
}

}  // namespace

extern "C" {

int cc_pjoin_create(cc_pjoin **out, const cc_comm *comm, int kind, const int64_t *d_build_keys, size_t n_build_local, size_t max_probe_rows,
                    int n_sub, cc_stream_t s) {
  CC_REQUIRE(out, "pjoin is NULL");
  *out = nullptr;
  CC_TRY(require_device());
  CC_REQUIRE(comm && comm->allgather && comm->barrier, "cc_comm needs allgather and barrier callbacks");
  CC_REQUIRE(comm->world >= 1 && comm->world <= kMaxPeers && (comm->world & (comm->world - 1)) == 0, "world size %d must be a power of two <= %d",
             comm->world, kMaxPeers);
  CC_REQUIRE(comm->rank >= 0 && comm->rank < comm->world, "rank %d out of range", comm->rank);
  CC_REQUIRE(kind == CC_HT_LP || kind == CC_HT_CHAIN, "unknown table kind %d", kind);
  CC_REQUIRE(n_build_local == 0 || d_build_keys, "d_build_keys is NULL");
  CC_REQUIRE(n_sub >= 1 && n_sub <= kPjMaxPieces, "n_sub must be in [1, %d]", kPjMaxPieces);
  cudaStream_t st = as_stream(s);
  cc_pjoin *j = new cc_pjoin();
  j->comm = *comm;
  j->world = comm->world;
  j->rank = comm->rank;
  j->kind = kind;
  j->pieces = n_sub;
  while ((1 << j->log2p) < j->world) ++j->log2p;
  cudaGetDevice(&j->device);
  const int P = j->world;
  int rc = CC_OK;
  auto fail = [&](int code) {
    pj_release(j);
    return code;
  };
  // ---- sizes: every rank learns every rank's build rows and probe rows per call
  unsigned long long mine[2] = {(unsigned long long) n_build_local, (unsigned long long) max_probe_rows};
  std::vector<unsigned long long> all(2 * (size_t) P, 0);
  if (comm->allgather(comm->user, mine, all.data(), sizeof(mine)) != 0) {
    set_error("cc_pjoin_create: allgather callback failed");
    return fail(CC_ERR_INVALID);
  }
  unsigned long long n_total = 0, max_local = 0, max_probe = 0;
  for (int r = 0; r < P; ++r) {
    n_total += all[2 * r];
    max_local = std::max(max_local, all[2 * r]);
    max_probe = std::max(max_probe, all[2 * r + 1]);
  }
  j->n_build_total = (size_t) n_total;
  j->max_rows = std::max<size_t>(1, (size_t) ((max_probe + n_sub - 1) / n_sub));
  // ---- every rank's table gets the SAME size: the reference's rule on the GLOBAL key count (linear_probing_ht.cpp:5-6: pow2 >=
  // 4n; chaining_ht.cpp:5-6: pow2 >= 2n) divided by the ranks, with headroom for an uneven hash partition (LP: at most half full
  // as long as no rank owns more than twice its share) -- a sender must know the slice a key falls into on its owner
  size_t slots = 1;
  const size_t per_key = kind == CC_HT_LP ? 4 : 2;
  while (slots < per_key * (size_t) n_total) slots <<= 1;
  slots = std::max<size_t>(1, slots / P);  // (one rank: exactly the reference's rule)
  j->table_slots = slots;
  // table slices: L2-sized, a power of two, at most kMaxParts / P (the scatter kernel ranks P * S partitions per tile)
  const size_t table_bytes = kind == CC_HT_LP ? slots * 8 : slots * 8 + (size_t) (n_total / P) * 8;
  // test switch CCB_PJ_SLICE_BYTES=<bytes>: slice size (and half the smallest table that is sliced), so that small tables take
  // the fused owner x slice path too
  size_t slice_bytes = kPjSliceBytes, min_table = kPjSliceMinTable;
  if (const char *e = getenv("CCB_PJ_SLICE_BYTES")) {
    const long long v = atoll(e);
    if (v >= 1024) {
      slice_bytes = (size_t) v;
      min_table = 2 * slice_bytes;
    }
  }
  int S = 1, log2s = 0;
  if (table_bytes >= min_table && slots <= (1ull << 32)) {
    while ((size_t) S * slice_bytes < table_bytes && S * 2 * P <= kMaxParts && (size_t) S * 2 <= slots) {
      S *= 2;
      ++log2s;
    }
  }
  j->slices = S;
  j->log2s = log2s;
  // ---- exchange geometry
  PjLayout &lay = j->lay;
  lay.P = P;
  lay.Ba = n_sub;
  lay.Sa = S;
  const unsigned long long per = (j->max_rows + (size_t) P * S - 1) / ((size_t) P * S);
  lay.cap = (per + per / 32 + 2 * (unsigned long long) kPartTile + kPartTile - 1) / kPartTile * kPartTile;
  lay.finish();
  cudaError_t e = cudaMalloc(&j->block, lay.total);
  if (e == cudaSuccess) e = cudaMemset(j->block, 0, lay.data_off);
  // two batches' worth of send slots: with the probe split in begin / end a batch is partitioned while the copies of the batch
  // before it are still draining (copy engines run at about half their idle rate under the kernels: profiles/r2_nvlink_bench_n2.txt)
  j->n_slots = 2 * n_sub;
  auto env_int = [](const char *name, int dflt, int lo, int hi) {
    const char *v = getenv(name);
    if (!v) return dflt;
    const int x = atoi(v);
    return x < lo ? lo : (x > hi ? hi : x);
  };
  j->n_cs = env_int("CCB_PJ_COPY_STREAMS", 4, 1, kPjMaxCopyStreams);      // measurement switches, see DESIGN.md
  j->n_direct = env_int("CCB_PJ_DIRECT", 0, 0, P - 1);                    // remote owners served by the scatter kernel's own NVLink stores
  // share of every block copy that an SM kernel moves: 0 unless the step is copy-bound (many ranks: 7/8 of the keys travel)
  j->sm_pct = env_int("CCB_PJ_SM_COPY_PCT", 0, 0, 90);
  for (int b = 0; b < j->n_slots && e == cudaSuccess; ++b) e = cudaMalloc(&j->send[b], (size_t) P * S * lay.cap * 8);
  if (e == cudaSuccess) e = cudaMalloc(&j->d_counts, (size_t) j->n_slots * P * S * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMalloc(&j->d_flag, sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&j->d_err, sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(j->d_flag, 0, sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(j->d_err, 0, sizeof(int));
  for (int i = 0; i < j->n_cs && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&j->cs[i], cudaStreamNonBlocking);
  for (int i = 0; i < j->n_slots && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&j->parted[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&j->copied[i], cudaEventDisableTiming);
  }
  for (int i = 0; i < j->n_cs && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&j->joined[i], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&j->gate, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&j->sm_done, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("cc_pjoin_create: %s (exchange memory: %zu bytes)", cudaGetErrorString(e), lay.total);
    cudaGetLastError();
    return fail(e == cudaErrorMemoryAllocation ? CC_ERR_NOMEM : CC_ERR_CUDA);
  }
  j->peer_block[j->rank] = j->block;
  j->memops = PjMemOps::resolve();
  if (P > 1) {
    cc_ipc_handle h_mine;
    rc = cc_ipc_export(j->block, &h_mine);
    if (rc != CC_OK) return fail(rc);
    std::vector<cc_ipc_handle> handles(P);
    if (comm->allgather(comm->user, &h_mine, handles.data(), sizeof(cc_ipc_handle)) != 0) {
      set_error("cc_pjoin_create: allgather callback failed");
      return fail(CC_ERR_INVALID);
    }
    for (int r = 0; r < P; ++r) {
      if (r == j->rank) continue;
      void *q = nullptr;
      rc = cc_ipc_open(&handles[r], &q);
      if (rc != CC_OK) return fail(rc);
      j->peer_block[r] = static_cast<unsigned char *>(q);
    }
  }
  if (comm->barrier(comm->user) != 0) return fail(CC_ERR_INVALID);  // every control block is zeroed and mapped before anyone signals
  if (j->memops.ok()) {
    // self-test: a stream memory operation on every peer's block (a scratch word per sender); a driver that refuses them on
    // peer memory gets the kernel protocol instead (same flags, so ranks that decide differently still understand each other)
    bool works = true;
    for (int r = 0; r < P && works; ++r)
      works = j->memops.write((CUstream) st, (CUdeviceptr) consumed_flag(j, r, kPjArenas, j->rank), 1, CU_STREAM_WRITE_VALUE_DEFAULT) == CUDA_SUCCESS;
    if (works) works = j->memops.wait((CUstream) st, (CUdeviceptr) consumed_flag(j, j->rank, kPjArenas, j->rank), 1, CU_STREAM_WAIT_VALUE_GEQ) == CUDA_SUCCESS;
    if (works) works = cudaStreamSynchronize(st) == cudaSuccess;
    if (!works) {
      cudaGetLastError();
      j->memops = PjMemOps();
    }
  }
  // ---- build side: exchanged piece by piece by owner only (one region per owner), appended to a dense column
  const size_t piece_rows = std::max<size_t>(1, (size_t) ((lay.cap - 2 * (unsigned long long) kPartTile) * 32 / 33) * P);  // its owner shares fit a region
  const size_t pieces = (size_t) ((max_local + piece_rows - 1) / piece_rows);
  const size_t build_cap = (size_t) (n_total / P + n_total / P / 4 + (1u << 16));
  int64_t *d_build = nullptr;
  unsigned long long *d_cursor = nullptr;
  e = cudaMalloc(&d_build, std::max<size_t>(build_cap, 1) * 8);
  if (e == cudaSuccess) e = cudaMalloc(&d_cursor, sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), st);
  if (e != cudaSuccess) {
    if (d_build) cudaFree(d_build);
    set_error("cc_pjoin_create: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return fail(CC_ERR_NOMEM);
  }
  const size_t saved_max = j->max_rows;
  j->max_rows = piece_rows;
  for (size_t piece = 0; piece < pieces && rc == CC_OK; ++piece) {
    const size_t off = std::min(n_build_local, piece * piece_rows), cnt = std::min(piece_rows, n_build_local - off);
    const int arena = (int) (j->uses % kPjArenas);
    const unsigned long long epoch = j->uses / kPjArenas + 1;
    ++j->uses;
    rc = send_piece(j, cnt ? d_build_keys + off : nullptr, cnt, /*S=*/1, 0, arena, /*piece=*/0, epoch, true, st);
    if (rc != CC_OK) break;
    rc = wait_flags(j, ready_flag(j, j->rank, arena, 0, 0), P, epoch, st);
    if (rc != CC_OK) break;
    // one region per sender: region_index(0, sender, 0) = sender * Sa
    pj_compact_kernel<<<sm_count() * 4, 256, 0, st>>>(j->arena(j->rank, arena), j->counts(arena), lay.cap, P, (size_t) lay.Sa * lay.cap, (size_t) lay.Sa,
                                                    d_build, build_cap, d_cursor);
    note_launch();
    pj_advance_kernel<<<1, 32, 0, st>>>(j->counts(arena), lay.cap, P, (size_t) lay.Sa, d_cursor);
    note_launch();
    if (cudaGetLastError() != cudaSuccess) rc = CC_ERR_CUDA;
    if (rc == CC_OK) rc = signal_consumed(j, arena, epoch, st);
  }
  j->max_rows = saved_max;
  unsigned long long n_owned = 0;
  int h_flag = 0, h_err = 0;
  if (rc == CC_OK) {
    e = cudaMemcpyAsync(&n_owned, d_cursor, sizeof(n_owned), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_flag, j->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_err, j->d_err, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      set_error("cc_pjoin_create: %s", cudaGetErrorString(e));
      cudaGetLastError();
      rc = CC_ERR_CUDA;
    }
  }
  // every rank learns whether EVERY rank's build exchange went through, so that all of them fail (or none)
  const bool fits = kind != CC_HT_LP || 2 * (size_t) n_owned <= slots;
  unsigned long long ok_mine = (rc == CC_OK && !h_flag && !h_err && n_owned <= build_cap && fits) ? 1 : 0;
  std::vector<unsigned long long> ok_all(P, 0);
  if (comm->allgather(comm->user, &ok_mine, ok_all.data(), sizeof(ok_mine)) != 0) rc = CC_ERR_INVALID;
  bool all_ok = rc == CC_OK;
  for (int r = 0; r < P; ++r) all_ok = all_ok && ok_all[r] == 1;
  if (!all_ok) {
    cudaFree(d_build);
    cudaFree(d_cursor);
    if (rc == CC_OK) {
      set_error("cc_pjoin_create: the build-side exchange failed on some rank (here: region overrun %d, wait timeout %d, %llu rows owned, room for %zu, "
                "table of %zu slots) -- heavily skewed build keys",
                h_flag, h_err, n_owned, build_cap, slots);
      rc = CC_ERR_UNSUPPORTED;
    }
    return fail(rc);
  }
  cudaMemsetAsync(j->d_flag, 0, sizeof(int), st);
  rc = cc_ht_build_sized(&j->table, kind, d_build, (size_t) n_owned, slots, CC_BUILD_ORDERED, s);
  cudaFree(d_build);
  cudaFree(d_cursor);
  if (rc != CC_OK) return fail(rc);
  if (j->table->n_slots != slots) {
    set_error("cc_pjoin_create: internal error: table of %zu slots, %zu expected", j->table->n_slots, slots);
    return fail(CC_ERR_INVALID);
  }
  *out = j;
  return CC_OK;
}

// evidence switch CCB_PJ_TRACE=1: CUDA-event timeline of the calls on their stream (synchronises; printed by every rank)
static bool pj_trace() {
  static const bool on = [] {
    const char *e = getenv("CCB_PJ_TRACE");
    return e && e[0] == '1';
  }();
  return on;
}
static void pj_mark(cc_pjoin *j, cudaStream_t st, const char *name, int idx = -1) {
  if (!pj_trace()) return;
  cudaEvent_t ev;
  cudaEventCreate(&ev);
  cudaEventRecord(ev, st);
  j->marks.push_back({ev, idx >= 0 ? std::string(name) + std::to_string(idx) : std::string(name)});
}
static void pj_print_trace(cc_pjoin *j, cudaStream_t st) {
  if (!pj_trace() || j->marks.empty()) return;
  cudaStreamSynchronize(st);
  std::string line = "pjoin timeline rank " + std::to_string(j->rank) + " (ms since " + j->marks[0].second + ", " + std::to_string(j->pieces) + " pieces, " +
                     std::to_string(j->slices) + " slices):";
  for (size_t i = 1; i < j->marks.size(); ++i) {
    float ms = 0;
    cudaEventElapsedTime(&ms, j->marks[0].first, j->marks[i].first);
    char buf[64];
    snprintf(buf, sizeof(buf), " %s=%.2f", j->marks[i].second.c_str(), ms);
    line += buf;
  }
  fprintf(stderr, "%s\n", line.c_str());
  for (auto &m : j->marks) cudaEventDestroy(m.first);
  j->marks.clear();
}

int cc_pjoin_probe_begin(cc_pjoin *j, const int64_t *d_keys, size_t n, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(j, "NULL argument");
  CC_REQUIRE(n == 0 || d_keys, "d_keys is NULL");
  CC_REQUIRE(n <= j->max_rows * (size_t) j->pieces, "%zu probe rows exceed the %zu this join was sized for", n, j->max_rows * (size_t) j->pieces);
  CC_REQUIRE(j->begun.size() < (size_t) kPjArenas, "at most %d probe batches can be in flight: call cc_pjoin_probe_end first", kPjArenas);
  cudaStream_t st = as_stream(s);
  const int B = j->pieces;
  const int arena = (int) (j->uses % kPjArenas);
  const unsigned long long epoch = j->uses / kPjArenas + 1;
  ++j->uses;
  pj_mark(j, st, "begin");
  // EVERY rank sends B pieces per batch, whatever its own row count: the piece boundaries only depend on B
  const size_t per = (n + B - 1) / B;
  for (int b = 0; b < B; ++b) {
    const size_t off = std::min(n, (size_t) b * per), cnt = std::min(per, n - off);
    CC_TRY(send_piece(j, cnt ? d_keys + off : nullptr, cnt, j->slices, j->log2s, arena, b, epoch, b == 0, st, j->sm_pct > 0));
    pj_mark(j, st, "F", b);
  }
  if (!j->deferred.empty()) {
    // the SM share of the block copies, behind the partition pass of the whole batch; then the pieces are signalled
    // (the kernel stores into the owners' arenas: they must have consumed the arena's previous use -- long ago in steady state)
    if (epoch > 1) CC_TRY(wait_flags(j, consumed_flag(j, j->rank, arena, 0), j->world, epoch - 1, st));
    for (const PjCopyJobs &jobs : j->sm_jobs) {
      pj_sm_copy_jobs_kernel<<<sm_count() * 4, 256, 0, st>>>(jobs);
      CC_CHECK_LAUNCH();
    }
    pj_mark(j, st, "SMCOPY");
    PJ_CUDA(cudaEventRecord(j->sm_done, st));
    cudaStream_t c0 = j->cs[0];
    PJ_CUDA(cudaStreamWaitEvent(c0, j->sm_done, 0));
    for (const cc_pjoin::Deferred &d : j->deferred) {
      CC_TRY(signal_piece(j, d.arena, d.piece, d.S, j->d_counts + (size_t) d.slot * j->world * j->lay.Sa, d.epoch, c0));
      PJ_CUDA(cudaEventRecord(j->copied[d.slot], c0));
    }
    j->deferred.clear();
    j->sm_jobs.clear();
  }
  j->begun.push_back({arena, epoch, j->calls++});
  return CC_OK;
}

int cc_pjoin_probe_end(cc_pjoin *j, int64_t *d_out_key, int64_t *d_out_payload, size_t out_capacity, cc_probe_result *d_result, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(j && d_result, "NULL argument");
  CC_REQUIRE(!j->begun.empty(), "cc_pjoin_probe_end without a batch begun by cc_pjoin_probe_begin");
  cudaStream_t st = as_stream(s);
  const PjLayout &lay = j->lay;
  const int P = j->world, B = j->pieces, S = j->slices;
  const size_t cap = (d_out_key || d_out_payload) ? out_capacity : 0;
  const cc_pjoin::Batch batch = j->begun.front();
  j->begun.erase(j->begun.begin());
  const int arena = batch.arena;
  const unsigned long long epoch = batch.epoch;
  // A batch whose exchange had a whole later cc_pjoin_probe_begin to complete (software pipelining across steps) has landed: ONE
  // probe, the table is streamed once.  A batch that was begun just now is probed in two groups of pieces: the first half while
  // the second half is still on its way -- the copy chain of the last pieces hides under the first probe at the price of
  // streaming the table twice.  Within a group the arena [piece][sender][slice] is walked slice by slice: all pieces and
  // senders of one table slice, then the next slice.
  const bool pipelined = !j->begun.empty();
  PJ_CUDA(cudaMemsetAsync(d_result, 0, sizeof(cc_probe_result), st));
  const int groups = (!pipelined && B >= 4 && P > 1) ? 2 : 1;
  for (int g = 0; g < groups; ++g) {
    const int b0 = g * B / groups, b1 = (g + 1) * B / groups;
    CC_TRY(wait_flags(j, ready_flag(j, j->rank, arena, b0, 0), (b1 - b0) * P, epoch, st));
    pj_mark(j, st, "WAIT", g);
    SegIn seg;
    seg.counts = j->counts(arena) + lay.region_index(b0, 0, 0);
    seg.cap = lay.cap;
    seg.segments = (b1 - b0) * P * S;
    seg.inner = (b1 - b0) * P;
    seg.outer_stride = lay.Sa;
    seg.presliced = S > 1;
    CC_TRY(probe_segmented_device(j->table, j->arena(j->rank, arena) + lay.region_index(b0, 0, 0) * lay.cap, seg, d_out_key, d_out_payload, cap, d_result, st,
                                  /*accumulate=*/true));
    pj_mark(j, st, "PROBE", g);
  }
  pj_close_kernel<<<1, 32, 0, st>>>(d_result, cap, j->d_flag, j->d_err);
  CC_CHECK_LAUNCH();
  CC_TRY(signal_consumed(j, arena, epoch, st));
  pj_mark(j, st, "END");
  pj_print_trace(j, st);
  return CC_OK;
}

int cc_pjoin_probe(cc_pjoin *j, const int64_t *d_keys, size_t n, int64_t *d_out_key, int64_t *d_out_payload, size_t out_capacity,
                   cc_probe_result *d_result, cc_stream_t s) {
  CC_REQUIRE(j && j->begun.empty(), "cc_pjoin_probe with a batch in flight (cc_pjoin_probe_begin without cc_pjoin_probe_end)");
  CC_TRY(cc_pjoin_probe_begin(j, d_keys, n, s));
  return cc_pjoin_probe_end(j, d_out_key, d_out_payload, out_capacity, d_result, s);
}

int cc_peer_copy_sm(void *d_dst, const void *d_src, size_t bytes, int blocks, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(d_dst && d_src && bytes % 16 == 0 && blocks >= 1, "NULL pointer, size not a multiple of 16, or no blocks");
  CC_REQUIRE(((uintptr_t) d_dst & 15) == 0 && ((uintptr_t) d_src & 15) == 0, "pointers must be 16-byte aligned");
  pj_sm_copy_kernel<<<blocks, 256, 0, as_stream(s)>>>(static_cast<uint4 *>(d_dst), static_cast<const uint4 *>(d_src), bytes / 16);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

int cc_pjoin_table(const cc_pjoin *j, const cc_ht **ht) {
  CC_REQUIRE(j && ht, "NULL argument");
  *ht = j->table;
  return CC_OK;
}

int cc_pjoin_destroy(cc_pjoin *j) {
  if (!j) return CC_OK;
  cudaDeviceSynchronize();
  int rc = CC_OK;
  if (j->comm.barrier && j->comm.barrier(j->comm.user) != 0) rc = CC_ERR_INVALID;  // nobody unmaps memory a peer may still write
  pj_release(j);
  return rc;
}

}  // extern "C"
