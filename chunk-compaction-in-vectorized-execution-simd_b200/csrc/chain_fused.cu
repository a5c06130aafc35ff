// chain_fused.cu -- a whole chain of hash joins in ONE persistent kernel with
// in-kernel chunk compaction between the joins.
//
// Reference protocol (main.cpp:119-191):
//   ExecutePipeline(chunk, L): ss = hts[L]->Probe(...); while (ss.HasNext()) {
//       ss.Next(...); compactors[L]->Compact(result); if (result->count_ == 0) continue;
//       ExecutePipeline(result, L + 1); }            -- depth first, one live chunk per level
//   FlushPipelineCache: drain every compactor top-down at the end.
//
// GPU form: a CTA is one pipeline instance.  All per-level operator state lives in
// shared memory, so nothing between two joins ever touches HBM:
//   * scan[L]   = the ScanStructure of level L: (row, pos, end) for kW lanes
//   * chunk[L]  = the Compactor cache in front of level L: up to (kS+1)*kW LHS row ids
// A CTA-uniform state machine replays the reference's recursion iteratively:
//   - if chunk[cur+1] holds >= threshold rows -> Probe the next join with them (descend)
//   - else if scan[cur] still has lanes        -> one round (Next) at level cur
//   - else                                     -> return to the parent level (ascend)
//   - at the top: pull the next kW-row LHS chunk; when the table is exhausted, flush
//     the caches top-down (FlushPipelineCache).
// A round lets every active lane inspect up to kS consecutive chain entries / slots
// (one 32-byte sector), ranks the matches with a block-wide exclusive scan and appends
// the matching rows DENSELY to chunk[cur+1] -- the compaction step (K3/K10/K11 of
// SURVEY 2.1).  With threshold == kW downstream joins only ever see full chunks
// (NaiveCompactor); threshold == 0 pushes every Next result down as is (no compaction).
//
// In the reference's key-only tables the payload of a match IS the probe key
// (chaining_ht.cpp:34 drops the payload column), so an intermediate row is fully
// described by its LHS row id; the result tuple [k_0..k_{J-1}, 0,k_0, 0,k_1, ...]
//  is rebuilt from the LHS columns at the ResultCollector.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace ccb {

constexpr int kW = CC_CHAIN_WIDTH;  // chunk width == threads per CTA
constexpr int kS = 4;               // entries inspected per lane per round (one sector)
constexpr int kBufCap = (kS + 1) * kW;
constexpr uint32_t kNoRow = 0xFFFFFFFFu;
constexpr int kWarps = kW / 32;

struct ChainLevel {
  const uint64_t *slots;
  const uint2 *dir;
  const int64_t *ckeys;
  const uint32_t *occ;  // occupancy bitmap of the buckets / slots (cc_ht::d_occ)
  uint64_t mask;
  const int64_t *col;  // LHS join-key column of this level
  uint32_t need;       // rows that must be cached before the next join runs (1..kW)
  int kind;
  int unique;
};

struct ChainArgs {
  ChainLevel lv[CC_MAX_JOINS];
  int64_t *out[3 * CC_MAX_JOINS];
  int n_joins;
  int materialize;
  size_t n_rows;
  size_t cap;
  cc_chain_result *res;
  cc_chain_telemetry *tel;  // optional: chunk-density histograms (the ZebraProfiler analogue, profiler.h:168-260)
};

struct ChainShared {
  unsigned long long cs[CC_MAX_JOINS];
  unsigned long long digest;
  unsigned long long base;
  unsigned long long level_in[CC_MAX_JOINS], steps[CC_MAX_JOINS], lanes[CC_MAX_JOINS];
  unsigned long long tile;
  uint32_t bufcnt[CC_MAX_JOINS + 1];
  uint32_t active[CC_MAX_JOINS];
  uint32_t scan[33];
};

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t &total, uint32_t *s_w) {
  uint32_t incl = warp_incl_scan_u32(v);
  unsigned w = threadIdx.x >> 5;
  if (lane_id() == 31) s_w[w] = incl;
  __syncthreads();
  if (w == 0) {
    uint32_t x = lane_id() < kWarps ? s_w[lane_id()] : 0;
    uint32_t xi = warp_incl_scan_u32(x);
    s_w[lane_id()] = xi - x;
    if (lane_id() == 31) s_w[32] = xi;
  }
  __syncthreads();
  uint32_t off = s_w[w] + incl - v;
  total = s_w[32];
  __syncthreads();
  return off;
}

// scan[L] holds up to kScanCap live lanes (a probe step may add kW lanes to kW-1 waiting ones)
constexpr int kScanCap = 2 * kW;

__global__ void __launch_bounds__(kW, 2) chain_fused_kernel(ChainArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ChainShared &S = *reinterpret_cast<ChainShared *>(smem_raw);
  uint32_t *sc_row = reinterpret_cast<uint32_t *>(smem_raw + sizeof(ChainShared));  // [J][kScanCap]
  uint32_t *sc_pos = sc_row + a.n_joins * kScanCap;
  uint32_t *sc_end = sc_pos + a.n_joins * kScanCap;
  uint32_t *bufs = sc_end + a.n_joins * kScanCap;  // chunk[L] for L = 1 .. J-1 at bufs + (L-1)*kBufCap
  const int J = a.n_joins;
  const unsigned tid = threadIdx.x;

  if (tid == 0) {
    for (int j = 0; j < CC_MAX_JOINS; ++j) S.cs[j] = S.level_in[j] = S.steps[j] = S.lanes[j] = 0, S.active[j] = 0;
    for (int j = 0; j <= CC_MAX_JOINS; ++j) S.bufcnt[j] = 0;
    S.digest = 0;
    atomicMax((unsigned long long *) &a.res->reserved[1], ~(unsigned long long) globaltimer_ns());  // ~(earliest start): 0-initialised
  }
  __syncthreads();

  const size_t ntiles = (a.n_rows + kW - 1) / kW;
  int cur = 0;
  bool exhausted = false;  // no LHS tiles left (this CTA)
  int flush_upto = 0;      // levels <= flush_upto receive no more input: their caches are drained regardless of threshold

  // Probe (chaining_ht.cpp:38-58 / linear_probing_ht.cpp:39-60) for up to kW rows handed to level L.  Only lanes
  // whose bucket / first slot is non-empty enter the ScanStructure, and they are APPENDED densely to the lanes already
  // waiting there (lane refill): rounds then always run on a full set of live lanes.
  auto probe_step = [&](int L, uint32_t row) {
    const ChainLevel &lv = a.lv[L];
    bool act = false;
    uint32_t pos = 0, end = 0;
    if (row != kNoRow) {
      uint64_t key = (uint64_t) __ldg(lv.col + row);
      uint64_t h = murmurhash64(key) & lv.mask;
      if (lv.kind == CC_HT_CHAIN) {
        uint2 d = __ldg(lv.dir + h);
        pos = d.x;
        end = d.x + d.y;
        act = d.y != 0;
      } else {
        pos = (uint32_t) h;
        act = ld_nc_u64(lv.slots + h) != kEmptyU;
      }
    }
    int n_valid = __syncthreads_count(row != kNoRow);
    uint32_t total;
    uint32_t off = block_excl_scan(act ? 1u : 0u, total, S.scan);
    uint32_t base = S.active[L];
    if (act) {
      sc_row[L * kScanCap + base + off] = row;
      sc_pos[L * kScanCap + base + off] = pos;
      sc_end[L * kScanCap + base + off] = end;
    }
    __syncthreads();
    if (tid == 0) {
      S.active[L] = base + total;
      S.level_in[L] += (unsigned long long) n_valid;
    }
    __syncthreads();
  };

  for (;;) {
    // ---- 1. descend: the compactor in front of level cur+1 holds a chunk (or is being flushed)
    if (cur + 1 < J) {
      uint32_t cnt = S.bufcnt[cur + 1];
      const uint32_t need = (cur + 1 <= flush_upto) ? 1u : a.lv[cur].need;  // flushing: drain regardless of the threshold
      if (cnt >= need) {
        ++cur;
        continue;
      }
    }
    const uint32_t n_act = S.active[cur];
    // does level cur still have input to probe?  level 0: LHS tiles; level L: its cached chunk
    const bool has_input = cur == 0 ? !exhausted : S.bufcnt[cur] > 0;
    // live lanes wanted before a round: the threshold of the compactor feeding this level (1 .. kW), so that
    // threshold 0 reproduces the uncompacted pipeline (every Next result and every probe runs as is)
    const uint32_t want = a.lv[cur > 0 ? cur - 1 : 0].need;
    const bool closed = cur == 0 ? exhausted : cur <= flush_upto;  // this level will receive no further input
    // ---- 2. round (Next) when enough live lanes wait, or when nothing more can ever be added.  Too few lanes
    //         and an open upstream: the lanes simply wait in the scan (step 4 ascends) until more rows arrive.
    if (n_act >= want || (n_act > 0 && !has_input && closed)) {
      const int L = cur;
      const ChainLevel &lv = a.lv[L];
      const uint32_t first = n_act > (uint32_t) kW ? n_act - kW : 0;  // the last <= kW lanes are processed
      const uint32_t lanes = n_act - first;
      uint32_t row = kNoRow, p = 0, e = 0;
      if (tid < lanes) {
        row = sc_row[L * kScanCap + first + tid];
        p = sc_pos[L * kScanCap + first + tid];
        e = sc_end[L * kScanCap + first + tid];
      }
      uint32_t m = 0;
      bool still = false;
      if (row != kNoRow) {
        uint64_t key = (uint64_t) __ldg(lv.col + row);
        if (lv.kind == CC_HT_CHAIN) {
          uint64_t v[kS];
#pragma unroll
          for (int s = 0; s < kS; ++s) v[s] = (p + s < e) ? (uint64_t) __ldg(lv.ckeys + p + s) : ~key;
#pragma unroll
          for (int s = 0; s < kS; ++s) m += (p + s < e) && (v[s] == key);
          p = (e - p > (uint32_t) kS) ? p + kS : e;
          still = p != e;
        } else {
          uint64_t v[kS];
#pragma unroll
          for (int s = 0; s < kS; ++s) v[s] = ld_nc_u64(lv.slots + ((uint64_t) (p + s) & lv.mask));
          still = true;
#pragma unroll
          for (int s = 0; s < kS; ++s) {
            if (still) {
              if (v[s] == kEmptyU)
                still = false;  // walk ends at the first empty slot (linear_probing_ht.cpp:104-108)
              else
                m += (v[s] == key);
            }
          }
          p = (uint32_t) ((uint64_t) (p + kS) & lv.mask);
        }
        if (lv.unique && m) still = false;
      }
      __syncthreads();  // every lane has read its scan entry: the tail may be rewritten
      // AdvancePointers: surviving lanes stay in the scan, compacted in place at the tail
      uint32_t n_still;
      uint32_t soff = block_excl_scan(still ? 1u : 0u, n_still, S.scan);
      if (still) {
        sc_row[L * kScanCap + first + soff] = row;
        sc_pos[L * kScanCap + first + soff] = p;
        sc_end[L * kScanCap + first + soff] = e;
      }
      uint32_t total;
      uint32_t off = block_excl_scan(m, total, S.scan);
      if (L + 1 < J) {
        // Compact: append the matching rows densely to the next level's cached chunk
        uint32_t cnt0 = S.bufcnt[L + 1];
        uint32_t *dst = bufs + L * kBufCap + cnt0 + off;
        for (uint32_t q = 0; q < m; ++q) dst[q] = row;
        __syncthreads();
        if (tid == 0) S.bufcnt[L + 1] = cnt0 + total;
      } else if (total) {
        // ResultCollector (main.cpp:125-128, data_collection.cpp:10-21)
        if (tid == 0) S.base = atomicAdd((unsigned long long *) &a.res->n_tuples, (unsigned long long) total);
        uint64_t th = 0;
        if (m) {
          th = 0x9e3779b97f4a7c15ULL;
          for (int j = 0; j < 3 * J; ++j) {
            uint64_t v = j < J ? (uint64_t) __ldg(a.lv[j].col + row) : (((j - J) & 1) ? (uint64_t) __ldg(a.lv[(j - J) >> 1].col + row) : 0ull);
            th = murmurhash64(th ^ v) + (uint64_t) j;
          }
          th *= (uint64_t) m;
        }
        th = warp_sum_u64(th);
        if (lane_id() == 0 && th) atomicAdd(&S.digest, (unsigned long long) th);
        for (int j = 0; j < J; ++j) {
          uint64_t v = m ? (uint64_t) __ldg(a.lv[j].col + row) * (uint64_t) m : 0ull;
          v = warp_sum_u64(v);
          if (lane_id() == 0 && v) atomicAdd(&S.cs[j], (unsigned long long) v);
        }
        __syncthreads();
        if (a.materialize && m) {
          uint64_t base = S.base + off;
          for (int j = 0; j < 3 * J; ++j) {
            int64_t v = j < J ? __ldg(a.lv[j].col + row) : (((j - J) & 1) ? __ldg(a.lv[(j - J) >> 1].col + row) : 0ll);
            for (uint32_t q = 0; q < m; ++q)
              if (base + q < a.cap) a.out[j][base + q] = v;
          }
        }
      }
      __syncthreads();
      if (tid == 0) {
        S.active[L] = first + n_still;
        S.steps[L] += 1;
        S.lanes[L] += (unsigned long long) lanes;
      }
      __syncthreads();
      continue;
    }
    // ---- 3. probe step: refill the scan of level cur from its input
    if (has_input) {
      uint32_t row = kNoRow;
      if (cur == 0) {
        if (tid == 0) S.tile = atomicAdd((unsigned long long *) &a.res->reserved[0], 1ull);
        __syncthreads();
        size_t tile = (size_t) S.tile;
        __syncthreads();
        if (tile < ntiles) {
          size_t r = tile * (size_t) kW + tid;
          row = r < a.n_rows ? (uint32_t) r : kNoRow;
        } else {
          exhausted = true;
          continue;
        }
      } else {
        uint32_t cnt = S.bufcnt[cur];
        uint32_t take = cnt < (uint32_t) kW ? cnt : (uint32_t) kW;
        row = tid < take ? bufs[(cur - 1) * kBufCap + (cnt - take) + tid] : kNoRow;
        __syncthreads();
        if (tid == 0) S.bufcnt[cur] = cnt - take;
      }
      probe_step(cur, row);
      continue;
    }
    // ---- 4. level cur has no input (and no lanes, or too few while its upstream can still deliver)
    if (cur > 0) {
      --cur;  // ascend: back to the producer of this level's input
      continue;
    }
    // level 0 idle and the table exhausted: FlushPipelineCache (main.cpp:172-191) -- drain the caches top-down
    {
      int next = 0;
      for (int l = 1; l < J; ++l)
        if (S.bufcnt[l] > 0 || S.active[l] > 0) {
          next = l;
          break;
        }
      if (next == 0) break;  // everything drained
      flush_upto = next;     // levels <= next get no new input any more
      cur = next;
    }
  }

  __syncthreads();
  if (tid < (unsigned) J) {
    unsigned long long v = S.cs[tid];
    if (v) {
      atomicAdd((unsigned long long *) &a.res->colsum[tid], v);
      atomicAdd((unsigned long long *) &a.res->colsum[J + 2 * tid + 1], v);
    }
    atomicAdd((unsigned long long *) &a.res->level_in[tid], S.level_in[tid]);
    atomicAdd((unsigned long long *) &a.res->level_steps[tid], S.steps[tid]);
    atomicAdd((unsigned long long *) &a.res->level_lanes[tid], S.lanes[tid]);
  }
  if (tid == 0) {
    if (S.digest) atomicAdd((unsigned long long *) &a.res->digest, S.digest);
    atomicMax((unsigned long long *) &a.res->reserved[2], (unsigned long long) globaltimer_ns());
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Warp-granular form of the same state machine (the default).  ONE WARP is one pipeline instance: a chunk is kR rows per
// lane (W = 32 * kR rows), every __syncthreads() of the CTA-wide kernel becomes a __syncwarp(), the three block-wide scans
// of a round become ballots / one packed warp scan, and warps never wait for each other -- while one instance sits on a
// table gather the others rank, compact and probe.  (The CTA-wide kernel spent its time at 6-10 barriers per step with
// every thread of the SM waiting for the same loads at the same moment: 2.8 us per step, profiles/r1_chain_c3_thresholds.txt.)
// A lane keeps kR independent gathers in flight in a probe step and kR * kS in a round; checksums are accumulated in
// registers over the whole kernel instead of being warp-reduced into shared memory every round.
// Thresholds keep their meaning on the CC_CHAIN_WIDTH = 512-row scale: need = ceil(threshold * W / 512), clamped to [1, W].
struct WarpShared {
  unsigned long long level_in[CC_MAX_JOINS], steps[CC_MAX_JOINS], lanes[CC_MAX_JOINS];
  uint32_t bufcnt[CC_MAX_JOINS + 1];
  uint32_t active[CC_MAX_JOINS];
};
// optional tail of a warp's slice (only when telemetry is requested: shared memory is what limits the resident warps)
struct WarpHist {
  uint32_t hist[2][CC_MAX_JOINS][CC_DENSITY_BINS];  // [0]: rows handed to a Probe, [1]: live lanes of a Next round, as a fraction of W
};

constexpr int kWarpsPerCta = 4;  // independent pipeline instances per CTA (a CTA is only a container: 32 CTAs per SM would cap the warps)
constexpr int kChunkMul = 3;  // chunk[L] holds kChunkMul * W row ids: fewer than W waiting + what a round may add (rounds emit only what fits)
__host__ __device__ constexpr size_t chain_warp_smem(int n_joins, int W, bool telemetry) {
  return (sizeof(WarpShared) + (size_t) n_joins * 3 * (2 * W) * sizeof(uint32_t) + (size_t) (n_joins - 1) * (kChunkMul * W) * sizeof(uint32_t) +
          (telemetry ? sizeof(WarpHist) : 0) + 15) & ~(size_t) 15;
}

#ifndef CCB_CHAIN_MINB_R2
#define CCB_CHAIN_MINB_R2 6  // resident CTAs per SM the compiler must allow for R = 2 (A/B: tools/build_variant.sh -DCCB_CHAIN_MINB_R2=8)
#endif
template <int R>
__global__ void __launch_bounds__(32 * kWarpsPerCta, R == 1 ? 10 : (R == 2 ? CCB_CHAIN_MINB_R2 : 3)) chain_warp_kernel(ChainArgs a) {
  static_assert(R >= 1 && R <= 4, "the match ranks of a round are packed into four 16-bit fields");
  constexpr int W = 32 * R;       // rows per chunk of this pipeline instance
  constexpr int SC = 2 * W;       // scan[L]: a probe step may add W lanes to W - 1 waiting ones
  constexpr int BC = kChunkMul * W;  // chunk[L]: a round emits only the matches that fit, the other lanes wait (see `room`)
  constexpr int KSC = 8;          // chain entries inspected per lane and round (two sectors: a 5-entry chain ends in one round)
  constexpr int KSL = 4;          // LP slots inspected per lane and round
  extern __shared__ __align__(16) unsigned char smem_all[];
  unsigned char *smem_raw = smem_all + (threadIdx.x >> 5) * chain_warp_smem(a.n_joins, W, a.tel != nullptr);  // this warp's private slice
  WarpShared &S = *reinterpret_cast<WarpShared *>(smem_raw);
  uint32_t *sc_row = reinterpret_cast<uint32_t *>(smem_raw + sizeof(WarpShared));  // [J][SC]
  uint32_t *sc_pos = sc_row + a.n_joins * SC;
  uint32_t *sc_end = sc_pos + a.n_joins * SC;
  uint32_t *bufs = sc_end + a.n_joins * SC;  // chunk[L] for L = 1 .. J-1 at bufs + (L-1)*BC
  WarpHist *H = a.tel ? reinterpret_cast<WarpHist *>(bufs + (a.n_joins - 1) * BC) : nullptr;
  const int J = a.n_joins;
  const unsigned lane = threadIdx.x & 31u, lt = lanemask_lt();
  const auto need_of = [&](int l) -> uint32_t {  // threshold of the compactor behind join l, on this instance's chunk width
    const uint32_t t = (a.lv[l].need * (uint32_t) W + (uint32_t) kW - 1u) / (uint32_t) kW;
    return t < 1u ? 1u : (t > (uint32_t) W ? (uint32_t) W : t);
  };

  if (lane == 0) {
    for (int j = 0; j < CC_MAX_JOINS; ++j) S.level_in[j] = S.steps[j] = S.lanes[j] = 0, S.active[j] = 0;
    for (int j = 0; j <= CC_MAX_JOINS; ++j) S.bufcnt[j] = 0;
    if (H)
      for (int k = 0; k < 2; ++k)
        for (int j = 0; j < CC_MAX_JOINS; ++j)
          for (int q = 0; q < CC_DENSITY_BINS; ++q) H->hist[k][j][q] = 0;
    atomicMax((unsigned long long *) &a.res->reserved[1], ~(unsigned long long) globaltimer_ns());  // ~(earliest start): 0-initialised
  }
  __syncwarp();
  // density bin of a chunk of n rows on this instance's width: CC_DENSITY_BINS equal bins over (0, W], full chunks in the last
  const auto bin_of = [](uint32_t n) -> uint32_t { return n == 0 ? 0u : ((n - 1u) * (uint32_t) CC_DENSITY_BINS) / (uint32_t) W; };

  uint64_t cs_acc[CC_MAX_JOINS];  // per-lane partial column sums / digest of the result rows, reduced once at the end
#pragma unroll
  for (int j = 0; j < CC_MAX_JOINS; ++j) cs_acc[j] = 0;
  uint64_t digest_acc = 0;

  const size_t ntiles = (a.n_rows + W - 1) / W;
  int cur = 0;
  bool exhausted = false;
  int flush_upto = 0;

  for (;;) {
    // ---- 1. descend: the compactor in front of level cur+1 holds a chunk (or is being flushed)
    if (cur + 1 < J) {
      const uint32_t cnt = S.bufcnt[cur + 1];
      const uint32_t need = (cur + 1 <= flush_upto) ? 1u : need_of(cur);
      if (cnt >= need) {
        ++cur;
        continue;
      }
    }
    const uint32_t n_act = S.active[cur];
    const bool has_input = cur == 0 ? !exhausted : S.bufcnt[cur] > 0;
    const uint32_t want = need_of(cur > 0 ? cur - 1 : 0);
    const bool closed = cur == 0 ? exhausted : cur <= flush_upto;
    // ---- 2. round (Next): ScanInnerJoin + GatherResult + AdvancePointers for up to W waiting lanes
    if (n_act >= want || (n_act > 0 && !has_input && closed)) {
      const int L = cur;
      const ChainLevel &lv = a.lv[L];
      const bool chain = lv.kind == CC_HT_CHAIN;
      const uint32_t ks = chain ? (uint32_t) KSC : (uint32_t) KSL;
      (void) ks;
      const uint32_t lanes = n_act < (uint32_t) W ? n_act : (uint32_t) W;
      const uint32_t first = n_act - lanes;  // the LAST `lanes` entries of the scan are processed
      uint32_t row[R], p[R], p0[R], e[R], m[R];
      bool still[R];
      uint64_t key[R];
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const uint32_t idx = (uint32_t) i * 32u + lane;
        row[i] = kNoRow;
        p[i] = e[i] = m[i] = 0;
        still[i] = false;
        if (idx < lanes) {
          row[i] = sc_row[L * SC + first + idx];
          p[i] = sc_pos[L * SC + first + idx];
          e[i] = sc_end[L * SC + first + idx];
        }
        p0[i] = p[i];
      }
#pragma unroll
      for (int i = 0; i < R; ++i) key[i] = row[i] != kNoRow ? (uint64_t) __ldg(lv.col + row[i]) : 0ull;
      if (chain) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
          if (row[i] != kNoRow) {
            uint64_t v[KSC];
#pragma unroll
            for (int q = 0; q < KSC; ++q) v[q] = (p[i] + q < e[i]) ? (uint64_t) __ldg(lv.ckeys + p[i] + q) : ~key[i];
#pragma unroll
            for (int q = 0; q < KSC; ++q) m[i] += (p[i] + q < e[i]) && (v[q] == key[i]);
            p[i] = (e[i] - p[i] > (uint32_t) KSC) ? p[i] + KSC : e[i];
            still[i] = p[i] != e[i];
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < R; ++i) {
          if (row[i] != kNoRow) {
            uint64_t v[KSL];
#pragma unroll
            for (int q = 0; q < KSL; ++q) v[q] = ld_nc_u64(lv.slots + ((uint64_t) (p[i] + q) & lv.mask));
            bool open = true;
#pragma unroll
            for (int q = 0; q < KSL; ++q) {
              open = open && v[q] != kEmptyU;  // the walk ends at the first empty slot (linear_probing_ht.cpp:104-108)
              m[i] += (open && v[q] == key[i]) ? 1u : 0u;
            }
            still[i] = open;
            p[i] = (uint32_t) ((uint64_t) (p[i] + KSL) & lv.mask);
          }
        }
      }
      if (lv.unique) {
#pragma unroll
        for (int i = 0; i < R; ++i)
          if (m[i]) still[i] = false;
      }
      // rank the matches: ONE warp scan over the kR per-lane counts packed into 16-bit fields (a field sums to <= 32 * 8)
      uint64_t packed = 0;
#pragma unroll
      for (int i = 0; i < R; ++i) packed |= (uint64_t) m[i] << (16 * i);
      uint64_t incl = packed;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint64_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned) o) incl += t;
      }
      const uint64_t tot = __shfl_sync(0xffffffffu, incl, 31);
      const uint64_t excl = incl - packed;
      uint32_t off[R], total = 0;
#pragma unroll
      for (int i = 0; i < R; ++i) {
        off[i] = total + (uint32_t) ((excl >> (16 * i)) & 0xFFFFu);
        total += (uint32_t) ((tot >> (16 * i)) & 0xFFFFu);
      }
      if (L + 1 < J) {
        // The next level's chunk takes what fits.  Ranks grow with the entry index, so the entries that fit are a prefix; an
        // entry beyond it is put back as it was (position restored, still waiting) and emits nothing in this round.  At least
        // one entry always fits: the chunk holds fewer than W rows here (else the state machine had descended) and an entry
        // emits at most KSC <= BC - W matches (BC = 3 W >= 96).
        const uint32_t room = (uint32_t) BC - S.bufcnt[L + 1];
        if (total > room) {
          uint32_t fit_end = 0;
#pragma unroll
          for (int i = 0; i < R; ++i) {
            if (off[i] + m[i] > room) {
              if (row[i] != kNoRow) {
                p[i] = p0[i];
                still[i] = true;
              }
              m[i] = 0;
            } else {
              fit_end = off[i] + m[i] > fit_end ? off[i] + m[i] : fit_end;
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const uint32_t t = __shfl_xor_sync(0xffffffffu, fit_end, o);
            fit_end = t > fit_end ? t : fit_end;
          }
          total = fit_end;
        }
      }
      __syncwarp();  // every lane has read its scan entries: the tail may be rewritten
      // AdvancePointers: surviving lanes stay in the scan, compacted in place at the tail
      uint32_t sbase = first;
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const unsigned bal = __ballot_sync(0xffffffffu, still[i]);
        if (still[i]) {
          const uint32_t o = sbase + __popc(bal & lt);
          sc_row[L * SC + o] = row[i];
          sc_pos[L * SC + o] = p[i];
          sc_end[L * SC + o] = e[i];
        }
        sbase += __popc(bal);
      }
      if (L + 1 < J) {
        // Compact: append the matching rows densely to the next level's cached chunk
        const uint32_t cnt0 = S.bufcnt[L + 1];
        uint32_t *dst = bufs + L * BC + cnt0;
#pragma unroll
        for (int i = 0; i < R; ++i) {
          if (m[i]) {
            // the row will be probed at the next level soon: pull its key there into L2 now, off the critical path
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.lv[L + 1].col + row[i]));
            for (uint32_t q = 0; q < m[i]; ++q) dst[off[i] + q] = row[i];
          }
        }
        __syncwarp();
        if (lane == 0) S.bufcnt[L + 1] = cnt0 + total;
      } else if (total) {
        // ResultCollector (main.cpp:125-128, data_collection.cpp:10-21)
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd((unsigned long long *) &a.res->n_tuples, (unsigned long long) total);
        base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
        for (int i = 0; i < R; ++i) {
          if (m[i]) {
            uint64_t th = 0x9e3779b97f4a7c15ULL;
            for (int j = 0; j < 3 * J; ++j) {
              const uint64_t v = j < J ? (uint64_t) __ldg(a.lv[j].col + row[i]) : (((j - J) & 1) ? (uint64_t) __ldg(a.lv[(j - J) >> 1].col + row[i]) : 0ull);
              th = murmurhash64(th ^ v) + (uint64_t) j;
            }
            digest_acc += th * (uint64_t) m[i];
#pragma unroll
            for (int j = 0; j < CC_MAX_JOINS; ++j)
              if (j < J) cs_acc[j] += (uint64_t) __ldg(a.lv[j].col + row[i]) * (uint64_t) m[i];
            if (a.materialize) {
              const uint64_t at = base + off[i];
              for (int j = 0; j < 3 * J; ++j) {
                const int64_t v = j < J ? __ldg(a.lv[j].col + row[i]) : (((j - J) & 1) ? __ldg(a.lv[(j - J) >> 1].col + row[i]) : 0ll);
                for (uint32_t q = 0; q < m[i]; ++q)
                  if (at + q < a.cap) a.out[j][at + q] = v;
              }
            }
          }
        }
      }
      if (lane == 0) {
        S.active[L] = sbase;
        S.steps[L] += 1;
        S.lanes[L] += (unsigned long long) lanes;
        if (H) H->hist[1][L][bin_of(lanes)] += 1;
      }
      __syncwarp();
      continue;
    }
    // ---- 3. probe step (Probe, chaining_ht.cpp:38-58 / linear_probing_ht.cpp:39-60): refill the scan of level cur; only
    //         lanes with a non-empty bucket / first slot enter it, appended densely behind the lanes already waiting
    if (has_input) {
      const int L = cur;
      const ChainLevel &lv = a.lv[L];
      uint32_t row[R];
      if (L == 0) {
        unsigned long long tile = 0;
        if (lane == 0) tile = atomicAdd((unsigned long long *) &a.res->reserved[0], 1ull);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= (unsigned long long) ntiles) {
          exhausted = true;
          continue;
        }
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const size_t r = (size_t) tile * W + (size_t) i * 32 + lane;
          row[i] = r < a.n_rows ? (uint32_t) r : kNoRow;
        }
      } else {
        const uint32_t cnt = S.bufcnt[L];
        const uint32_t take = cnt < (uint32_t) W ? cnt : (uint32_t) W;
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const uint32_t idx = (uint32_t) i * 32u + lane;
          row[i] = idx < take ? bufs[(L - 1) * BC + (cnt - take) + idx] : kNoRow;
        }
        __syncwarp();
        if (lane == 0) S.bufcnt[L] = cnt - take;
      }
      uint64_t h[R];
#pragma unroll
      for (int i = 0; i < R; ++i) h[i] = row[i] != kNoRow ? (murmurhash64((uint64_t) __ldg(lv.col + row[i])) & lv.mask) : 0ull;
      uint32_t pos[R], end[R];
      bool act[R];
      // empty bucket / empty first slot (chaining_ht.cpp:52-55, linear_probing_ht.cpp:53-57): answered by the occupancy bitmap,
      // which is small enough to stay in L2 -- only the lanes that pass touch the bucket directory
      uint32_t ow[R];
#pragma unroll
      for (int i = 0; i < R; ++i) ow[i] = row[i] != kNoRow ? __ldg(lv.occ + (h[i] >> 5)) : 0u;
#pragma unroll
      for (int i = 0; i < R; ++i) act[i] = ((ow[i] >> ((uint32_t) h[i] & 31u)) & 1u) != 0u;
      if (lv.kind == CC_HT_CHAIN) {
        uint2 d[R];
#pragma unroll
        for (int i = 0; i < R; ++i) d[i] = act[i] ? __ldg(lv.dir + h[i]) : make_uint2(0u, 0u);
#pragma unroll
        for (int i = 0; i < R; ++i) {
          pos[i] = d[i].x;
          end[i] = d[i].x + d[i].y;
        }
      } else {
#pragma unroll
        for (int i = 0; i < R; ++i) {
          pos[i] = (uint32_t) h[i];
          end[i] = 0;
        }
      }
      uint32_t n_valid = 0;
      uint32_t base = S.active[L];
#pragma unroll
      for (int i = 0; i < R; ++i) {
        n_valid += __popc(__ballot_sync(0xffffffffu, row[i] != kNoRow));
        const unsigned bal = __ballot_sync(0xffffffffu, act[i]);
        if (act[i]) {
          const uint32_t o = base + __popc(bal & lt);
          sc_row[L * SC + o] = row[i];
          sc_pos[L * SC + o] = pos[i];
          sc_end[L * SC + o] = end[i];
        }
        base += __popc(bal);
      }
      __syncwarp();
      if (lane == 0) {
        S.active[L] = base;
        S.level_in[L] += (unsigned long long) n_valid;
        if (H) H->hist[0][L][bin_of(n_valid)] += 1;
      }
      __syncwarp();
      continue;
    }
    // ---- 4. level cur has no input (and no lanes, or too few while its upstream can still deliver)
    if (cur > 0) {
      --cur;
      continue;
    }
    // level 0 idle and the table exhausted: FlushPipelineCache (main.cpp:172-191) -- drain the caches top-down
    {
      int next = 0;
      for (int l = 1; l < J; ++l)
        if (S.bufcnt[l] > 0 || S.active[l] > 0) {
          next = l;
          break;
        }
      if (next == 0) break;
      flush_upto = next;
      cur = next;
    }
  }

  __syncwarp();
  digest_acc = warp_sum_u64(digest_acc);
#pragma unroll
  for (int j = 0; j < CC_MAX_JOINS; ++j) cs_acc[j] = j < J ? warp_sum_u64(cs_acc[j]) : 0ull;
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < CC_MAX_JOINS; ++j) {
      if (j < J) {
        if (cs_acc[j]) {
          atomicAdd((unsigned long long *) &a.res->colsum[j], (unsigned long long) cs_acc[j]);
          atomicAdd((unsigned long long *) &a.res->colsum[J + 2 * j + 1], (unsigned long long) cs_acc[j]);
        }
        if (S.level_in[j]) atomicAdd((unsigned long long *) &a.res->level_in[j], S.level_in[j]);
        if (S.steps[j]) atomicAdd((unsigned long long *) &a.res->level_steps[j], S.steps[j]);
        if (S.lanes[j]) atomicAdd((unsigned long long *) &a.res->level_lanes[j], S.lanes[j]);
      }
    }
    if (digest_acc) atomicAdd((unsigned long long *) &a.res->digest, (unsigned long long) digest_acc);
    atomicMax((unsigned long long *) &a.res->reserved[2], (unsigned long long) globaltimer_ns());
  }
  if (a.tel) {  // one atomic per non-empty (histogram, level, bin) and warp, lanes share the bins
    for (int idx = (int) lane; idx < 2 * J * CC_DENSITY_BINS; idx += 32) {
      const int k = idx / (J * CC_DENSITY_BINS), j = (idx / CC_DENSITY_BINS) % J, q = idx % CC_DENSITY_BINS;
      const uint32_t v = H->hist[k][j][q];
      if (v) atomicAdd((unsigned long long *) (k == 0 ? &a.tel->probe_rows_hist[j][q] : &a.tel->round_lanes_hist[j][q]), (unsigned long long) v);
    }
  }
}

template <int R>
static int launch_chain_warp(const ChainArgs &a, size_t n_joins, cudaStream_t st) {
  constexpr int W = 32 * R;
  const size_t smem = chain_warp_smem((int) n_joins, W, a.tel != nullptr) * kWarpsPerCta;
  CC_CUDA(cudaFuncSetAttribute(chain_warp_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  int per_sm = 0;
  CC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chain_warp_kernel<R>, 32 * kWarpsPerCta, smem));
  if (per_sm < 1) per_sm = 1;
  const size_t ntiles = (a.n_rows + W - 1) / W;
  const size_t grid = std::min<size_t>((ntiles + kWarpsPerCta - 1) / kWarpsPerCta, (size_t) sm_count() * per_sm);
  chain_warp_kernel<R><<<(unsigned) grid, 32 * kWarpsPerCta, smem, st>>>(a);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

__global__ void chain_finish_kernel(cc_chain_result *res, size_t cap, int materialize) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    res->overflow = (materialize && res->n_tuples > cap) ? 1 : 0;
    const unsigned long long first = ~res->reserved[1];  // the kernels keep ~(earliest start) there, so a cleared result needs no init pass
    res->device_ns = (res->reserved[1] != 0 && res->reserved[2] >= first) ? res->reserved[2] - first : 0;
  }
}


}  // namespace ccb

using namespace ccb;

extern "C" int cc_chain_execute(const cc_ht *const *h_tables, size_t n_joins, const int64_t *const *h_lhs_cols, size_t n_rows,
                                const uint32_t *thresholds, int64_t *const *h_out_cols, size_t out_capacity,
                                cc_chain_result *d_result, cc_stream_t s) {
  return cc_chain_execute_ex(h_tables, n_joins, h_lhs_cols, n_rows, thresholds, h_out_cols, out_capacity, d_result, nullptr, s);
}

extern "C" int cc_chain_execute_ex(const cc_ht *const *h_tables, size_t n_joins, const int64_t *const *h_lhs_cols, size_t n_rows,
                                   const uint32_t *thresholds, int64_t *const *h_out_cols, size_t out_capacity,
                                   cc_chain_result *d_result, cc_chain_telemetry *d_telemetry, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(h_tables && h_lhs_cols && d_result, "NULL argument");
  CC_REQUIRE(n_joins >= 1 && n_joins <= CC_MAX_JOINS, "n_joins must be in [1, %d]", CC_MAX_JOINS);
  CC_REQUIRE(n_rows < 0xFFFFFFFFull, "at most 2^32-2 LHS rows per call (got %zu)", n_rows);
  cudaStream_t st = as_stream(s);
  ChainArgs a;
  memset(&a, 0, sizeof(a));
  a.n_joins = (int) n_joins;
  a.n_rows = n_rows;
  a.res = d_result;
  a.tel = d_telemetry;  // accumulated into (the caller clears it): histograms of several calls add up
  a.materialize = h_out_cols != nullptr;
  a.cap = a.materialize ? out_capacity : 0;
  for (size_t l = 0; l < n_joins; ++l) {
    const cc_ht *t = h_tables[l];
    CC_REQUIRE(t && (h_lhs_cols[l] || n_rows == 0), "NULL table or column at level %zu", l);
    CC_REQUIRE(t->n_slots <= (1ull << 32), "table at level %zu has more than 2^32 slots", l);
    ChainLevel &lv = a.lv[l];
    lv.slots = t->d_slots;
    lv.dir = t->d_dir;
    lv.ckeys = t->d_ckeys;
    lv.occ = t->d_occ;
    lv.mask = t->mask;
    lv.col = h_lhs_cols[l];
    lv.kind = t->kind;
    lv.unique = !t->has_duplicates;
    uint32_t thr = thresholds ? thresholds[l] : (uint32_t) kW;
    lv.need = thr < 1 ? 1u : (thr > (uint32_t) kW ? (uint32_t) kW : thr);
  }
  if (a.materialize)
    for (size_t j = 0; j < 3 * n_joins; ++j) {
      CC_REQUIRE(h_out_cols[j], "NULL output column %zu", j);
      a.out[j] = h_out_cols[j];
    }
  CC_CUDA(cudaMemsetAsync(d_result, 0, sizeof(cc_chain_result), st));
  // measurement switch: CCB_CHAIN_IMPL = cta (the round-1 CTA-wide kernel), w1 / w2 / w4 (warp pipelines, rows per lane); default w2 (measured: profiles/r2_chain_variants.txt)
  static const int impl = [] {
    const char *e = getenv("CCB_CHAIN_IMPL");
    if (!e) return 2;
    if (!strcmp(e, "cta")) return 0;
    if (!strcmp(e, "w1")) return 1;
    if (!strcmp(e, "w4")) return 4;
    return 2;
  }();
  if (n_rows && impl != 0) {
    CC_TRY(impl == 1 ? launch_chain_warp<1>(a, n_joins, st) : impl == 2 ? launch_chain_warp<2>(a, n_joins, st) : launch_chain_warp<4>(a, n_joins, st));
  } else if (n_rows) {
    size_t smem = sizeof(ChainShared) + n_joins * 3 * (size_t) kScanCap * sizeof(uint32_t) + (n_joins - 1) * (size_t) kBufCap * sizeof(uint32_t);
    CC_CUDA(cudaFuncSetAttribute(chain_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    int per_sm = 0;
    CC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chain_fused_kernel, kW, smem));
    if (per_sm < 1) per_sm = 1;
    size_t ntiles = (n_rows + kW - 1) / kW;
    size_t grid = std::min<size_t>(ntiles, (size_t) sm_count() * per_sm);
    chain_fused_kernel<<<(unsigned) grid, kW, smem, st>>>(a);
    CC_CHECK_LAUNCH();
  }
  chain_finish_kernel<<<1, 32, 0, st>>>(d_result, a.cap, a.materialize);
  CC_CHECK_LAUNCH();
  return CC_OK;
}


// ZebraProfiler-style dump (profiler.h:168-260 writes one CSV row per observed chunk size): one row per histogram, level and bin
extern "C" int cc_chain_telemetry_csv(const cc_chain_telemetry *h_telemetry, size_t n_joins, const char *path) {
  CC_REQUIRE(h_telemetry && path, "NULL argument");
  CC_REQUIRE(n_joins >= 1 && n_joins <= CC_MAX_JOINS, "n_joins must be in [1, %d]", CC_MAX_JOINS);
  FILE *f = fopen(path, "w");
  CC_REQUIRE(f, "Unable to open file %s", path);  // negative_feedback.hpp:103-105 / profiler.h throw the same message
  fprintf(f, "histogram,level,density_from,density_to,chunks\n");
  for (int k = 0; k < 2; ++k)
    for (size_t l = 0; l < n_joins; ++l)
      for (int q = 0; q < CC_DENSITY_BINS; ++q)
        fprintf(f, "%s,%zu,%.4f,%.4f,%llu\n", k == 0 ? "probe_rows" : "round_lanes", l, (double) q / CC_DENSITY_BINS, (double) (q + 1) / CC_DENSITY_BINS,
                (unsigned long long) (k == 0 ? h_telemetry->probe_rows_hist[l][q] : h_telemetry->round_lanes_hist[l][q]));
  fclose(f);
  return CC_OK;
}

// Dynamic ("negative feedback") compaction: main.cpp:137-167 under flag_dynamic_compact.
// The LHS table is processed in batches; before every batch each join's bandit picks a compaction
// threshold (CompactTuner::SelectArm), the batch runs through the fused kernel, and every bandit is
// rewarded with the reference's formula  2 / seconds / 1e3  (main.cpp:166) where seconds is the
// DEVICE time of the batch (globaltimer, cc_chain_result.device_ns) -- a wall-clock reward would be
// dominated by launch jitter.  Results are accumulated over the batches into *h_result.
// The batches are PIPELINED kTunedDepth deep: batch i is launched while the results of batches i-1 .. i-kTunedDepth+1 are
// still on their way back (pinned host records + events), so the GPU never idles for a host round trip between batches; a
// bandit therefore selects with the feedback of all batches up to i - kTunedDepth (its arithmetic is unchanged: UpdateArm is
// told which arm a reward belongs to, negative_feedback.hpp:189-195).
extern "C" int cc_chain_execute_tuned(const cc_ht *const *h_tables, size_t n_joins, const int64_t *const *h_lhs_cols, size_t n_rows,
                                      size_t batch_rows, cc_tuner *tuner, size_t first_bandit_id, int64_t *const *h_out_cols,
                                      size_t out_capacity, cc_chain_result *h_result, cc_stream_t s) {
  CC_TRY(require_device());
  CC_REQUIRE(h_tables && h_lhs_cols && h_result && tuner, "NULL argument");
  CC_REQUIRE(n_joins >= 1 && n_joins <= CC_MAX_JOINS, "n_joins must be in [1, %d]", CC_MAX_JOINS);
  CC_REQUIRE(batch_rows > 0, "batch_rows must be > 0");
  CC_REQUIRE(cc_tuner_bandit_size(tuner) >= first_bandit_id + n_joins, "tuner holds %zu bandits, need %zu", cc_tuner_bandit_size(tuner),
             first_bandit_id + n_joins);
  cudaStream_t st = as_stream(s);
  // materialised output: every batch appends behind the rows of the batches before it, so its first row must be known when it
  // is launched -- that needs the previous batch's count, i.e. no overlap
  const int depth = h_out_cols ? 1 : 3;
  constexpr int kMaxDepth = 3;
  cc_chain_result *d_res = nullptr, *h_pin = nullptr;
  cudaEvent_t ev[kMaxDepth] = {nullptr, nullptr, nullptr};
  size_t arms[kMaxDepth][CC_MAX_JOINS];
  CC_CUDA(cudaMalloc(&d_res, kMaxDepth * sizeof(cc_chain_result)));
  cudaError_t e = cudaMallocHost(&h_pin, kMaxDepth * sizeof(cc_chain_result));
  for (int i = 0; i < kMaxDepth && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
  if (e != cudaSuccess) {
    cudaFree(d_res);
    if (h_pin) cudaFreeHost(h_pin);
    for (auto &x : ev)
      if (x) cudaEventDestroy(x);
    set_error("cc_chain_execute_tuned: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return CC_ERR_NOMEM;
  }
  cc_chain_result total;
  memset(&total, 0, sizeof(total));
  int rc = CC_OK;
  const size_t n_batches = n_rows ? (n_rows + batch_rows - 1) / batch_rows : 1;
  // takes the finished batch b off the pipeline: rewards its arms, adds its counters
  auto retire = [&](size_t b) -> int {
    const int slot = (int) (b % depth);
    if (cudaEventSynchronize(ev[slot]) != cudaSuccess) {
      set_error("cc_chain_execute_tuned: %s", cudaGetErrorString(cudaGetLastError()));
      return CC_ERR_CUDA;
    }
    const cc_chain_result &r = h_pin[slot];
    const double seconds = (double) r.device_ns * 1e-9;
    if (seconds > 0)
      for (size_t l = 0; l < n_joins; ++l) cc_tuner_update_arm(tuner, first_bandit_id + l, arms[slot][l], 2 / seconds / 1e3);  // main.cpp:166
    total.n_tuples += r.n_tuples;
    total.digest += r.digest;
    for (size_t j = 0; j < 3 * n_joins; ++j) total.colsum[j] += r.colsum[j];
    for (size_t l = 0; l < n_joins; ++l) {
      total.level_in[l] += r.level_in[l];
      total.level_steps[l] += r.level_steps[l];
      total.level_lanes[l] += r.level_lanes[l];
    }
    total.device_ns += r.device_ns;
    if (h_out_cols) total.overflow = total.n_tuples > out_capacity ? 1 : 0;
    return CC_OK;
  };
  size_t retired = 0;
  for (size_t b = 0; b < n_batches && rc == CC_OK; ++b) {
    if (b >= (size_t) depth) {  // the slot's previous batch must be off the pipeline before the slot is reused
      rc = retire(retired++);
      if (rc != CC_OK) break;
    }
    const int slot = (int) (b % depth);
    const size_t start = b * batch_rows, cnt = n_rows - start < batch_rows ? n_rows - start : batch_rows;
    const int64_t *cols[CC_MAX_JOINS];
    int64_t *outs[3 * CC_MAX_JOINS];
    uint32_t thr[CC_MAX_JOINS];
    for (size_t l = 0; l < n_joins; ++l) {
      cols[l] = h_lhs_cols[l] ? h_lhs_cols[l] + start : nullptr;
      rc = cc_tuner_select_arm(tuner, first_bandit_id + l, &arms[slot][l]);  // main.cpp:140
      if (rc != CC_OK) break;
      thr[l] = (uint32_t) (arms[slot][l] > 0xFFFFFFFFull ? 0xFFFFFFFFull : arms[slot][l]);
    }
    if (rc != CC_OK) break;
    const size_t used = (size_t) total.n_tuples;  // depth == 1 whenever rows are materialised: every earlier batch is retired
    const size_t room = h_out_cols && out_capacity > used ? out_capacity - used : 0;
    if (h_out_cols)
      for (size_t j = 0; j < 3 * n_joins; ++j) outs[j] = h_out_cols[j] + (used < out_capacity ? used : out_capacity);
    rc = cc_chain_execute(h_tables, n_joins, cols, cnt, thr, (h_out_cols && room) ? outs : nullptr, room, d_res + slot, s);
    if (rc != CC_OK) break;
    if (cudaMemcpyAsync(h_pin + slot, d_res + slot, sizeof(cc_chain_result), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaEventRecord(ev[slot], st) != cudaSuccess) {
      set_error("cc_chain_execute_tuned: %s", cudaGetErrorString(cudaGetLastError()));
      rc = CC_ERR_CUDA;
      break;
    }
  }
  while (rc == CC_OK && retired < n_batches) rc = retire(retired++);
  if (rc != CC_OK) cudaStreamSynchronize(st);
  cudaFree(d_res);
  cudaFreeHost(h_pin);
  for (auto &x : ev) cudaEventDestroy(x);
  CC_TRY(rc);
  *h_result = total;
  return CC_OK;
}
