// pjoin.cu -- the hash-partitioned multi-GPU join behind the C ABI (new functionality, SURVEY 8e / 8b "cc_partition_exchange"):
// one process per GPU, NO collective library on the data path.
//
// Both sides of an equi-join are partitioned by owner = murmurhash64(key) >> (64 - log2 P) (high hash bits, independent of the
// low bits that address the owner's table) and every rank builds / probes its own table with the single-GPU kernels
// (linear_probing_ht.cpp:4-115 / chaining_ht.cpp:4-136 per rank).  The exchange is made of four device-side pieces:
//   * partition_scatter_kernel (single pass, fixed regions): groups a sub-batch by owner in a local send buffer; the rows this
//     rank keeps go straight into its own receive buffer;
//   * copy engines: region p travels as ONE cudaMemcpyAsync into slot [this rank] of owner p's receive buffer (CUDA-IPC mapping of
//     peer memory over NVLink 5 / NVSwitch), on several copy streams, without occupying an SM;
//   * pj_signal_kernel: stores the region's row count and an epoch flag into the owner's control block over NVLink, behind the
//     copies in stream order ("rows of shuffle k from sender s have landed");
//   * pj_wait_kernel: the owner's main stream spins (bounded) on its own flags until every sender's epoch has arrived; after the
//     probe pj_consumed_kernel tells every sender that the buffer may be refilled (a sender waits for that before shuffle k + 3).
// So the host never learns a count and never blocks: a probe call enqueues  P(0) P(1) W(0) L(0) C(0) P(2) W(1) L(1) C(1) ...
// (P = owner partition + copies, W = wait, L = slice partition + probe of the received sub-batch, C = consumed) and returns.
// The host's only job is the control plane at create / destroy time (exchange of the IPC handles and sizes through the
// caller's cc_comm callbacks: MPI, torch.distributed, or the fork + shared-memory communicator of host/simd_compaction.hpp).
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "partition.cuh"

namespace ccb {

constexpr int kPjBuffers = 3;       // rotating receive / send buffers
constexpr int kPjCopyStreams = 4;   // one stream drives one copy engine at a time
constexpr unsigned long long kPjSpinNs = 20ull * 1000 * 1000 * 1000;  // a wait gives up after 20 s (a peer died): error bit, no hang

// control block at the head of every rank's exchange allocation; written by PEERS over NVLink
struct PjCtrl {
  unsigned long long counts[kPjBuffers][kMaxPeers];    // [b][s]: rows sender s delivered into buffer b
  unsigned long long ready[kPjBuffers][kMaxPeers];     // [b][s]: epoch of the last shuffle sender s delivered into buffer b
  unsigned long long consumed[kPjBuffers][kMaxPeers];  // [b][r]: epoch of the last shuffle receiver r has finished reading from ITS buffer b (of OUR rows)
};
constexpr size_t kPjCtrlBytes = (sizeof(PjCtrl) + 4095) / 4096 * 4096;

struct PjPeers {
  PjCtrl *ctrl[kMaxPeers];
};

__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// thread p < world: tell owner p that this rank's region of shuffle `epoch` (buffer b) has landed, with its row count
__global__ void pj_signal_kernel(PjPeers peers, int world, int rank, int b, const unsigned long long *__restrict__ d_counts, unsigned long long cap,
                                 unsigned long long epoch) {
  const int p = threadIdx.x;
  if (p >= world) return;
  unsigned long long c = d_counts[p];
  if (c > cap) c = cap;  // an overrun region was clamped by the scatter kernel (and flagged)
  peers.ctrl[p]->counts[b][rank] = c;
  __threadfence_system();
  st_sys_u64(&peers.ctrl[p]->ready[b][rank], epoch);
}

// thread s < world: wait until flags[s] >= epoch (flags live in THIS rank's memory, peers write them); bounded
__global__ void pj_wait_kernel(const unsigned long long *flags, int world, unsigned long long epoch, int *d_err) {
  const int s = threadIdx.x;
  if (s >= world) return;
  const unsigned long long t0 = globaltimer_ns();
  while (ld_sys_u64(flags + s) < epoch) {
    __nanosleep(200);
    if (globaltimer_ns() - t0 > kPjSpinNs) {
      atomicOr(d_err, 1);
      return;
    }
  }
}

// thread s < world: tell sender s that this rank has finished reading buffer b of shuffle `epoch`
__global__ void pj_consumed_kernel(PjPeers peers, int world, int rank, int b, unsigned long long epoch) {
  const int s = threadIdx.x;
  if (s >= world) return;
  st_sys_u64(&peers.ctrl[s]->consumed[b][rank], epoch);
}

// append the valid rows of a segmented column (segment s: counts[s] rows at src + s * cap) to dst[*cursor ...]; rows beyond
// dst_cap are dropped (the cursor still counts them: the host sees the overflow)
__global__ void pj_compact_kernel(const int64_t *__restrict__ src, const unsigned long long *__restrict__ counts, unsigned long long cap, int segments,
                                  int64_t *__restrict__ dst, unsigned long long dst_cap, const unsigned long long *__restrict__ cursor) {
  const unsigned long long base0 = *cursor;
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (int s = 0; s < segments; ++s) {
    unsigned long long before = base0;
    for (int t = 0; t < s; ++t) before += counts[t] < cap ? counts[t] : cap;
    const unsigned long long c = counts[s] < cap ? counts[s] : cap;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < c; i += stride)
      if (before + i < dst_cap) dst[before + i] = src[(size_t) s * cap + i];
  }
}
__global__ void pj_advance_kernel(const unsigned long long *__restrict__ counts, unsigned long long cap, int segments, unsigned long long *cursor) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long t = 0;
    for (int s = 0; s < segments; ++s) t += counts[s] < cap ? counts[s] : cap;
    *cursor += t;
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
  int world = 1, rank = 0, log2p = 0, kind = CC_HT_LP, n_sub = 1, device = 0;
  size_t max_rows = 0;            // rows per shuffle (one sub-batch)
  unsigned long long cap = 0;     // rows per (sender, owner) region
  unsigned char *block = nullptr;                // own exchange allocation: PjCtrl | kPjBuffers x world x cap rows
  unsigned char *peer_block[kMaxPeers] = {};     // every rank's allocation in this address space (own pointer at [rank])
  int64_t *send[kPjBuffers] = {};
  unsigned long long *d_counts = nullptr;        // [kPjBuffers][kMaxParts... world] cursors of the owner partition
  int *d_flag = nullptr, *d_err = nullptr;       // sticky region-overrun flag, wait-timeout flag
  cudaStream_t cs[kPjCopyStreams] = {};
  cudaEvent_t parted[kPjBuffers] = {}, copied[kPjBuffers] = {}, gate = nullptr, joined[kPjCopyStreams] = {};
  unsigned long long shuffles = 0;               // global shuffle counter: epoch of shuffle k is k + 1, its buffer k % kPjBuffers
  cc_ht *table = nullptr;
  size_t n_build_total = 0;

  PjCtrl *ctrl(int r) const { return reinterpret_cast<PjCtrl *>(peer_block[r]); }
  int64_t *data(int r, int b) const { return reinterpret_cast<int64_t *>(peer_block[r] + kPjCtrlBytes) + (size_t) b * world * cap; }
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
  for (int r = 0; r < kMaxPeers; ++r) p.ctrl[r] = r < j->world ? j->ctrl(r) : nullptr;
  return p;
}

// first half of shuffle k: owner partition on `st`, block copies + signal on the copy streams.  Returns k.
int shuffle_start(cc_pjoin *j, const int64_t *d_keys, size_t n, cudaStream_t st, unsigned long long *out_k) {
  CC_REQUIRE(n <= j->max_rows, "%zu rows exceed the %zu rows per shuffle this join was sized for", n, j->max_rows);
  const unsigned long long k = j->shuffles++;
  const int b = (int) (k % kPjBuffers);
  const int P = j->world;
  unsigned long long *counts = j->d_counts + (size_t) b * kMaxPeers;
  if (k >= (unsigned long long) kPjBuffers) PJ_CUDA(cudaStreamWaitEvent(st, j->copied[b], 0));  // the copies of shuffle k - 3 have left send[b]
  // the rows this rank keeps go straight into slot [rank] of its own receive buffer (same region offset rank * cap)
  CC_TRY(partition_single_device(d_keys, n, PartFn::high_bits(j->log2p), j->cap, counts, j->d_flag, 0, nullptr, j->send[b], st, SegIn(), false,
                                 P > 1 ? j->rank : -1, P > 1 ? j->data(j->rank, b) : nullptr, /*sticky_flag=*/true));
  if (P == 1) {  // a single rank: the send buffer IS the receive column
    PJ_CUDA(cudaMemcpyAsync(j->data(0, b), j->send[b], (size_t) j->cap * 8, cudaMemcpyDeviceToDevice, st));
  }
  PJ_CUDA(cudaEventRecord(j->parted[b], st));
  cudaStream_t c0 = j->cs[0];
  PJ_CUDA(cudaStreamWaitEvent(c0, j->parted[b], 0));
  if (k >= (unsigned long long) kPjBuffers) {
    // every owner must have finished reading what shuffle k - 3 put into its buffer b before it is refilled
    pj_wait_kernel<<<1, 32, 0, c0>>>(&j->ctrl(j->rank)->consumed[b][0], P, k - kPjBuffers + 1, j->d_err);
    CC_CHECK_LAUNCH();
  }
  PJ_CUDA(cudaEventRecord(j->gate, c0));
  for (int s = 1; s < kPjCopyStreams; ++s) PJ_CUDA(cudaStreamWaitEvent(j->cs[s], j->gate, 0));
  const size_t bytes = (size_t) j->cap * 8;
  for (int i = 1; i < P; ++i) {
    const int p = (j->rank + i) % P;  // stagger the destinations so that the ranks do not all hit the same peer at once
    PJ_CUDA(cudaMemcpyAsync(j->data(p, b) + (size_t) j->rank * j->cap, j->send[b] + (size_t) p * j->cap, bytes, cudaMemcpyDeviceToDevice,
                            j->cs[(i - 1) % kPjCopyStreams]));
  }
  for (int s = 1; s < kPjCopyStreams; ++s) {
    PJ_CUDA(cudaEventRecord(j->joined[s], j->cs[s]));
    PJ_CUDA(cudaStreamWaitEvent(c0, j->joined[s], 0));
  }
  pj_signal_kernel<<<1, 32, 0, c0>>>(peers_of(j), P, j->rank, b, counts, j->cap, k + 1);
  CC_CHECK_LAUNCH();
  PJ_CUDA(cudaEventRecord(j->copied[b], c0));
  *out_k = k;
  return CC_OK;
}

// second half: `st` waits until every sender's rows of shuffle k have landed; the receive column is then data(rank, b) with
// world segments of cap rows and the counts in the control block
int shuffle_finish(cc_pjoin *j, unsigned long long k, cudaStream_t st, const int64_t **col, SegIn *seg) {
  const int b = (int) (k % kPjBuffers);
  pj_wait_kernel<<<1, 32, 0, st>>>(&j->ctrl(j->rank)->ready[b][0], j->world, k + 1, j->d_err);
  CC_CHECK_LAUNCH();
  *col = j->data(j->rank, b);
  seg->counts = &j->ctrl(j->rank)->counts[b][0];
  seg->cap = j->cap;
  seg->segments = j->world;
  return CC_OK;
}

int shuffle_consumed(cc_pjoin *j, unsigned long long k, cudaStream_t st) {
  pj_consumed_kernel<<<1, 32, 0, st>>>(peers_of(j), j->world, j->rank, (int) (k % kPjBuffers), k + 1);
  CC_CHECK_LAUNCH();
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
  CC_REQUIRE(n_sub >= 1 && n_sub <= 64, "n_sub must be in [1, 64]");
  cudaStream_t st = as_stream(s);
  cc_pjoin *j = new cc_pjoin();
  j->comm = *comm;
  j->world = comm->world;
  j->rank = comm->rank;
  j->kind = kind;
  j->n_sub = n_sub;
  while ((1 << j->log2p) < j->world) ++j->log2p;
  cudaGetDevice(&j->device);
  const int P = j->world;
  int rc = CC_OK;
  auto fail = [&](int code) {
    pj_release(j);
    return code;
  };
  // ---- sizes: every rank learns every rank's build rows; one shuffle moves at most max_rows rows per rank
  std::vector<unsigned long long> sizes(P, 0);
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
  const unsigned long long per = (j->max_rows + P - 1) / P;
  j->cap = (per + per / 32 + 2 * (unsigned long long) kPartTile + kPartTile - 1) / kPartTile * kPartTile;
  // ---- exchange memory + IPC mapping of every peer's block
  const size_t data_bytes = (size_t) kPjBuffers * P * j->cap * 8;
  cudaError_t e = cudaMalloc(&j->block, kPjCtrlBytes + data_bytes);
  if (e == cudaSuccess) e = cudaMemset(j->block, 0, kPjCtrlBytes);
  for (int b = 0; b < kPjBuffers && e == cudaSuccess; ++b) e = cudaMalloc(&j->send[b], (size_t) P * j->cap * 8);
  if (e == cudaSuccess) e = cudaMalloc(&j->d_counts, (size_t) kPjBuffers * kMaxPeers * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMalloc(&j->d_flag, sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&j->d_err, sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(j->d_flag, 0, sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(j->d_err, 0, sizeof(int));
  for (int i = 0; i < kPjCopyStreams && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&j->cs[i], cudaStreamNonBlocking);
  for (int i = 0; i < kPjBuffers && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&j->parted[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&j->copied[i], cudaEventDisableTiming);
  }
  for (int i = 0; i < kPjCopyStreams && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&j->joined[i], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&j->gate, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("cc_pjoin_create: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return fail(e == cudaErrorMemoryAllocation ? CC_ERR_NOMEM : CC_ERR_CUDA);
  }
  j->peer_block[j->rank] = j->block;
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
  // ---- build side: shuffle it piece by piece, append what arrives to a dense column, build the local table
  const size_t pieces = (size_t) ((max_local + j->max_rows - 1) / j->max_rows);
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
  for (size_t piece = 0; piece < pieces && rc == CC_OK; ++piece) {
    const size_t off = std::min(n_build_local, piece * j->max_rows), cnt = std::min(j->max_rows, n_build_local - off);
    unsigned long long k = 0;
    rc = shuffle_start(j, cnt ? d_build_keys + off : nullptr, cnt, st, &k);
    const int64_t *col = nullptr;
    SegIn seg;
    if (rc == CC_OK) rc = shuffle_finish(j, k, st, &col, &seg);
    if (rc == CC_OK) {
      pj_compact_kernel<<<sm_count() * 4, 256, 0, st>>>(col, seg.counts, seg.cap, seg.segments, d_build, build_cap, d_cursor);
      note_launch();
      pj_advance_kernel<<<1, 32, 0, st>>>(seg.counts, seg.cap, seg.segments, d_cursor);
      note_launch();
      if (cudaGetLastError() != cudaSuccess) rc = CC_ERR_CUDA;
    }
    if (rc == CC_OK) rc = shuffle_consumed(j, k, st);
  }
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
  unsigned long long ok_mine = (rc == CC_OK && !h_flag && !h_err && n_owned <= build_cap) ? 1 : 0;
  std::vector<unsigned long long> ok_all(P, 0);
  if (comm->allgather(comm->user, &ok_mine, ok_all.data(), sizeof(ok_mine)) != 0) rc = CC_ERR_INVALID;
  bool all_ok = rc == CC_OK;
  for (int r = 0; r < P; ++r) all_ok = all_ok && ok_all[r] == 1;
  if (!all_ok) {
    cudaFree(d_build);
    cudaFree(d_cursor);
    if (rc == CC_OK) {
      set_error("cc_pjoin_create: the build-side exchange failed on some rank (region overrun %d, wait timeout %d, %llu rows owned of %zu)", h_flag,
                h_err, n_owned, build_cap);
      rc = CC_ERR_UNSUPPORTED;
    }
    return fail(rc);
  }
  cudaMemsetAsync(j->d_flag, 0, sizeof(int), st);
  // the reference's sizing rule on the GLOBAL key count, divided by the ranks (see cc_ht_build_sized)
  size_t slots = 1;
  const size_t per_key = kind == CC_HT_LP ? 4 : 2;
  while (slots < per_key * (size_t) n_total) slots <<= 1;
  slots = std::max<size_t>(1, slots / P);
  while (kind == CC_HT_LP && slots < 2 * (size_t) n_owned) slots <<= 1;
  rc = cc_ht_build_sized(&j->table, kind, d_build, (size_t) n_owned, P > 1 ? slots : 0, CC_BUILD_ORDERED, s);
  cudaFree(d_build);
  cudaFree(d_cursor);
  if (rc != CC_OK) return fail(rc);
  *out = j;
  return CC_OK;
}

int cc_pjoin_probe(cc_pjoin *j, const int64_t *d_keys, size_t n, int64_t *d_out_key, int64_t *d_out_payload, size_t out_capacity,
                   cc_probe_result *d_result, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(j && d_result, "NULL argument");
  CC_REQUIRE(n == 0 || d_keys, "d_keys is NULL");
  CC_REQUIRE(n <= j->max_rows * (size_t) j->n_sub, "%zu probe rows exceed the %zu this join was sized for", n, j->max_rows * (size_t) j->n_sub);
  cudaStream_t st = as_stream(s);
  const size_t cap = (d_out_key || d_out_payload) ? out_capacity : 0;
  PJ_CUDA(cudaMemsetAsync(d_result, 0, sizeof(cc_probe_result), st));
  // EVERY rank runs n_sub shuffles per call, whatever its own row count: the sub-batch boundaries only depend on n_sub
  const int n_sub = j->n_sub;
  const size_t per = (n + n_sub - 1) / n_sub;
  auto piece = [&](int b, const int64_t **p, size_t *cnt) {
    const size_t off = std::min(n, (size_t) b * per);
    *cnt = std::min(per, n - off);
    *p = *cnt ? d_keys + off : nullptr;
  };
  std::vector<unsigned long long> ks(n_sub);
  const int64_t *p = nullptr;
  size_t cnt = 0;
  piece(0, &p, &cnt);
  CC_TRY(shuffle_start(j, p, cnt, st, &ks[0]));
  for (int b = 0; b < n_sub; ++b) {
    if (b + 1 < n_sub) {
      piece(b + 1, &p, &cnt);
      CC_TRY(shuffle_start(j, p, cnt, st, &ks[b + 1]));  // its copies run underneath the probe of sub-batch b
    }
    const int64_t *col = nullptr;
    SegIn seg;
    CC_TRY(shuffle_finish(j, ks[b], st, &col, &seg));
    CC_TRY(probe_segmented_device(j->table, col, seg, d_out_key, d_out_payload, cap, d_result, st, /*accumulate=*/true));
    CC_TRY(shuffle_consumed(j, ks[b], st));
  }
  pj_close_kernel<<<1, 32, 0, st>>>(d_result, cap, j->d_flag, j->d_err);
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
