// ref_driver.cpp -- drives the UNMODIFIED reference classes (HashTable,
// LPHashTable, ScanStructure, DataChunk, NaiveCompactor) with explicit inputs.
//
// TEST / BASELINE INFRASTRUCTURE ONLY.  This file contains no reference code:
// it #includes the reference headers from /root/reference and is linked with
// the reference's own base.cpp / chaining_ht.cpp / linear_probing_ht.cpp /
// compactor.cpp, compiled where they lie (recipe: oracle/Makefile, outputs in
// oracle/_ref/).  It exists because the reference's two drivers cannot be
// used as-is: simd_micro_bench.cpp has UB in ParseParameters
// (:35-73, missing return) and main.cpp discards results
// (setting.h:31 flag_collect_tuples=false) and takes no input arrays.
//
// Sub-commands (all print one JSON line on stdout):
//   main  J cf lhs rhs block compact [dump.bin]
//         == main.cpp:37-115 with std::mt19937 gen(2) LHS, chaining table;
//         compact: 0 none, 1 NaiveCompactor (built with the compactor.cpp:36 fix)
//   pipe  lhs.bin rows J cf rhs block kind compact [dump.bin]
//         same pipeline over an explicit row-major int64 LHS file; kind 0 LP, 1 chain
//   nextdump kind n cf block keys.bin nkeys inone out.bin
//         per-Next golden records of the chunk-granular protocol
//   micro kind variant n cf block keys.bin nkeys procs [share [reps]]
//         times one micro-bench variant (simd_micro_bench.cpp:83-361 loop shape); share = 1: one table built before the
//         fork and shared by the workers; reps: timed repetitions over the same table ("seconds" = the best one)
//         variant: 0 scalar Probe+Next, 1 SIMD Probe+Next, 2 scalar InOneNext, 3 SIMD InOneNext
//         procs > 1 forks that many workers over disjoint key ranges sharing the
//         read-only table (the reference itself is single-threaded; its
//         CycleProfiler singleton is not thread-safe, profiler.h:262-290)
//   bandit steps
//         drives the reference CompactTuner / MultiArmedBandit (negative_feedback.hpp) with a
//         deterministic synthetic reward stream and prints every selected threshold
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>

#include "base.h"
#include "chaining_ht.h"
#include "compactor.h"
#include "hash_functions.h"
#include "linear_probing_ht.h"
#include "negative_feedback.hpp"

using namespace simd_compaction;

namespace {

struct Collector {
  uint64_t n = 0, digest = 0;
  std::vector<uint64_t> colsum;
  std::vector<int64_t> tuples;
  bool keep = false;
  void Add(DataChunk &chunk) {  // same access pattern as data_collection.cpp:10-21
    size_t nc = chunk.data_.size();
    if (colsum.size() < nc) colsum.resize(nc, 0);
    for (size_t i = 0; i < chunk.count_; ++i) {
      auto idx = chunk.selection_vector_[i];
      uint64_t th = 0x9e3779b97f4a7c15ULL;
      for (size_t j = 0; j < nc; ++j) {
        uint64_t v = (uint64_t) chunk.data_[j].GetValue(idx);
        th = murmurhash64(th ^ v) + j;
        colsum[j] += v;
        if (keep) tuples.push_back((int64_t) v);
      }
      digest += th;
      ++n;
    }
  }
};

template <class HT>
struct Pipe {
  std::vector<std::unique_ptr<HT>> hts;
  std::vector<std::unique_ptr<DataChunk>> intermediates;
  std::vector<std::unique_ptr<NaiveCompactor>> compactors;
  Collector out;
  bool compact = false;
  uint64_t probe_tuples = 0, next_calls = 0;
  std::vector<uint64_t> level_in, level_chunks;

  void Execute(DataChunk &input, size_t level) {  // protocol of main.cpp:119-170
    if (level == hts.size()) {
      out.Add(input);
      return;
    }
    level_in[level] += input.count_;
    level_chunks[level] += 1;
    probe_tuples += input.count_;
    auto &join_key = input.data_[level];
    auto &result = intermediates[level];
    auto ss = hts[level]->Probe(join_key, input.count_, input.selection_vector_);
    while (ss.HasNext()) {
      ++next_calls;
      ss.Next(join_key, input, *result);
      if (compact) {
        compactors[level]->Compact(result);
        if (result->count_ == 0) continue;
      }
      Execute(*result, level + 1);
    }
  }
  void Flush(size_t level) {  // main.cpp:172-191
    if (level == hts.size()) return;
    compactors[level]->Flush(intermediates[level]);
    Execute(*intermediates[level], level + 1);
    Flush(level + 1);
  }
};

std::vector<int64_t> ReadFile(const char *path, size_t n) {
  std::vector<int64_t> v(n);
  FILE *f = fopen(path, "rb");
  if (!f || fread(v.data(), 8, n, f) != n) {
    fprintf(stderr, "cannot read %zu int64 from %s\n", n, path);
    exit(2);
  }
  fclose(f);
  return v;
}

template <class HT>
int RunPipe(const std::vector<int64_t> &lhs, size_t rows, size_t J, size_t cf, size_t rhs, bool compact,
            const char *dump) {
  std::vector<AttributeType> types(J, AttributeType::INTEGER);
  Pipe<HT> p;
  p.compact = compact;
  p.out.keep = dump != nullptr;
  p.level_in.assign(J, 0);
  p.level_chunks.assign(J, 0);
  for (size_t i = 0; i < J; ++i) {  // main.cpp:62-68
    p.hts.push_back(std::make_unique<HT>(rhs, cf));
    types.push_back(AttributeType::INTEGER);
    types.push_back(AttributeType::INTEGER);
    p.intermediates.push_back(std::make_unique<DataChunk>(types));
    p.compactors.push_back(std::make_unique<NaiveCompactor>(types));
  }
  std::vector<AttributeType> in_types(J, AttributeType::INTEGER);
  std::vector<Attribute> tuple(J);
  auto t0 = std::chrono::steady_clock::now();
  double secs = 0;
  size_t start = 0, end;
  do {  // main.cpp:81-95
    end = std::min(start + kBlockSize, rows);
    DataChunk chunk(in_types);
    for (size_t i = start; i < end; ++i) {
      for (size_t j = 0; j < J; ++j) tuple[j] = lhs[i * J + j];
      chunk.AppendTuple(tuple);
    }
    start = end;
    t0 = std::chrono::steady_clock::now();
    p.Execute(chunk, 0);
    secs += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  } while (end < rows);
  if (compact) {
    t0 = std::chrono::steady_clock::now();
    p.Flush(0);
    secs += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  printf("{\"n_tuples\": %llu, \"digest\": %llu, \"probe_tuples\": %llu, \"next_calls\": %llu, \"seconds\": %.6f, \"colsum\": [",
         (unsigned long long) p.out.n, (unsigned long long) p.out.digest, (unsigned long long) p.probe_tuples,
         (unsigned long long) p.next_calls, secs);
  for (size_t j = 0; j < 3 * J; ++j)
    printf("%s%llu", j ? ", " : "", (unsigned long long) (j < p.out.colsum.size() ? p.out.colsum[j] : 0));
  printf("], \"level_in\": [");
  for (size_t j = 0; j < J; ++j) printf("%s%llu", j ? ", " : "", (unsigned long long) p.level_in[j]);
  printf("], \"level_chunks\": [");
  for (size_t j = 0; j < J; ++j) printf("%s%llu", j ? ", " : "", (unsigned long long) p.level_chunks[j]);
  printf("]}\n");
  if (dump) {
    FILE *f = fopen(dump, "wb");
    fwrite(p.out.tuples.data(), 8, p.out.tuples.size(), f);
    fclose(f);
  }
  return 0;
}

// per-Next records: u32 count, then count x (u32 physical position, i64 payload at that position)
template <class HT>
int NextDump(size_t n, size_t cf, const std::vector<int64_t> &keys, bool inone, const char *out) {
  HT table(n, cf);
  std::vector<uint32_t> sel(kBlockSize);
  for (uint32_t i = 0; i < kBlockSize; ++i) sel[i] = i;
  DataChunk input(std::vector<AttributeType>{AttributeType::INTEGER});
  DataChunk output(std::vector<AttributeType>{AttributeType::INTEGER, AttributeType::INTEGER, AttributeType::INTEGER});
  Vector block(AttributeType::INTEGER);
  FILE *f = fopen(out, "wb");
  uint64_t n_tuples = 0, n_calls = 0;
  for (size_t k = 0; k < keys.size(); k += kBlockSize) {
    size_t fill = std::min(kBlockSize, keys.size() - k);
    for (size_t i = 0; i < fill; ++i) block.GetValue(i) = keys[k + i];
    input.data_[0] = block;
    input.count_ = fill;
    auto ss = table.Probe(block, fill, sel);
    while (ss.HasNext()) {
      size_t rc = inone ? ss.InOneNext(block, input, output) : ss.Next(block, input, output);
      uint32_t c32 = (uint32_t) rc;
      fwrite(&c32, 4, 1, f);
      for (size_t i = 0; i < rc; ++i) {
        uint32_t pos = output.selection_vector_[i];
        int64_t payload = output.data_[2].GetValue(pos);
        fwrite(&pos, 4, 1, f);
        fwrite(&payload, 8, 1, f);
      }
      n_tuples += rc;
      ++n_calls;
    }
    uint32_t end_marker = 0xFFFFFFFFu;  // end of this input chunk
    fwrite(&end_marker, 4, 1, f);
  }
  fclose(f);
  printf("{\"n_tuples\": %llu, \"next_calls\": %llu}\n", (unsigned long long) n_tuples, (unsigned long long) n_calls);
  return 0;
}

template <class HT>
uint64_t MicroLoop(HT &table, const int64_t *keys, size_t nkeys, int variant) {
  std::vector<uint32_t> sel(kBlockSize);
  for (uint32_t i = 0; i < kBlockSize; ++i) sel[i] = i;
  DataChunk input(std::vector<AttributeType>{AttributeType::INTEGER});
  DataChunk output(std::vector<AttributeType>{AttributeType::INTEGER, AttributeType::INTEGER, AttributeType::INTEGER});
  Vector block(AttributeType::INTEGER);
  uint64_t n_tuples = 0;
  for (size_t k = 0; k < nkeys; k += kBlockSize) {
    size_t fill = std::min(kBlockSize, nkeys - k);
    for (size_t i = 0; i < fill; ++i) block.GetValue(i) = keys[k + i];
    input.data_[0] = block;
    input.count_ = fill;
    switch (variant) {
      case 0: {
        auto ss = table.Probe(block, fill, sel);
        while (ss.HasNext()) n_tuples += ss.Next(block, input, output);
        break;
      }
      case 1: {
        auto ss = table.SIMDProbe(block, fill, sel);
        while (ss.HasNext()) n_tuples += ss.SIMDNext(block, input, output);
        break;
      }
      case 2: {
        auto ss = table.Probe(block, fill, sel);
        while (ss.HasNext()) n_tuples += ss.InOneNext(block, input, output);
        break;
      }
      default: {
        auto ss = table.SIMDProbe(block, fill, sel);
        while (ss.HasNext()) n_tuples += ss.SIMDInOneNext(block, input, output);
        break;
      }
    }
  }
  return n_tuples;
}

template <class HT>
int Micro(size_t n, size_t cf, const std::vector<int64_t> &keys, int variant, int procs, int share, int reps) {
  // procs > 1: workers are forked processes (the reference is single-threaded and its CycleProfiler singleton is not
  // thread-safe, profiler.h:262-290) over disjoint key ranges; all of them start every repetition together and a
  // repetition's time is its slowest worker's.  share == 0: fork FIRST, every worker builds its own private table;
  // share == 1: ONE table is built before the fork and shared copy-on-write (read-only afterwards) -- the only way a
  // 2^28-key table (8 GiB) fits beside 16 workers.  The table is built once, then `reps` timed repetitions run.
  if (reps < 1) reps = 1;
  std::vector<double> rep_secs(reps, 0.0);
  uint64_t total = 0;
  double build_s = 0;
  if (procs <= 1) {
    auto b0 = std::chrono::steady_clock::now();
    HT table(n, cf);
    build_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - b0).count();
    for (int r = 0; r < reps; ++r) {
      auto t0 = std::chrono::steady_clock::now();
      total = MicroLoop(table, keys.data(), keys.size(), variant);
      rep_secs[r] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
  } else {
    std::unique_ptr<HT> shared;
    if (share) {
      auto b0 = std::chrono::steady_clock::now();
      shared = std::make_unique<HT>(n, cf);
      build_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - b0).count();
    }
    std::vector<int> res_fd(procs), go_fd(procs), ready_fd(procs);
    std::vector<pid_t> pids(procs);
    size_t per = (keys.size() / procs / kBlockSize) * kBlockSize;
    if (per == 0) per = keys.size();
    for (int p = 0; p < procs; ++p) {
      int res[2], go[2], ready[2];
      if (pipe(res) || pipe(go) || pipe(ready)) return 3;
      pid_t pid = fork();
      if (pid == 0) {
        close(res[0]);
        close(go[1]);
        close(ready[0]);
        auto b0 = std::chrono::steady_clock::now();
        std::unique_ptr<HT> own;
        if (!share) own = std::make_unique<HT>(n, cf);
        HT &table = share ? *shared : *own;
        double bs = std::chrono::duration<double>(std::chrono::steady_clock::now() - b0).count();
        char c = 1;
        if (write(ready[1], &c, 1) != 1) _exit(4);
        size_t lo = std::min(keys.size(), (size_t) p * per);
        size_t hi = p == procs - 1 ? keys.size() : std::min(keys.size(), lo + per);
        for (int r = 0; r < reps; ++r) {
          if (read(go[0], &c, 1) != 1) _exit(4);
          auto c0 = std::chrono::steady_clock::now();
          uint64_t cnt = MicroLoop(table, keys.data() + lo, hi - lo, variant);
          double out[3] = {std::chrono::duration<double>(std::chrono::steady_clock::now() - c0).count(), (double) cnt, bs};
          if (write(res[1], out, sizeof(out)) != (ssize_t) sizeof(out)) _exit(4);
        }
        _exit(0);
      }
      close(res[1]);
      close(go[0]);
      close(ready[1]);
      res_fd[p] = res[0];
      go_fd[p] = go[1];
      ready_fd[p] = ready[0];
      pids[p] = pid;
    }
    char c = 0;
    for (int p = 0; p < procs; ++p)
      if (read(ready_fd[p], &c, 1) != 1) return 5;
    for (int r = 0; r < reps; ++r) {
      for (int p = 0; p < procs; ++p)
        if (write(go_fd[p], &c, 1) != 1) return 5;
      total = 0;
      for (int p = 0; p < procs; ++p) {
        double out[3] = {0, 0, 0};
        if (read(res_fd[p], out, sizeof(out)) != (ssize_t) sizeof(out)) return 5;
        rep_secs[r] = std::max(rep_secs[r], out[0]);
        total += (uint64_t) out[1];
        if (!share) build_s = std::max(build_s, out[2]);
      }
    }
    for (int p = 0; p < procs; ++p) {
      int st;
      waitpid(pids[p], &st, 0);
    }
  }
  double best = rep_secs[0];
  for (double x : rep_secs) best = std::min(best, x);
  printf("{\"n_tuples\": %llu, \"seconds\": %.6f, \"build_seconds\": %.3f, \"probe_keys\": %zu, \"procs\": %d, \"variant\": %d, \"shared_table\": %d, \"rep_seconds\": [",
         (unsigned long long) total, best, build_s, keys.size(), procs > 1 ? procs : 1, variant, share ? 1 : 0);
  for (int r = 0; r < reps; ++r) printf("%s%.6f", r ? ", " : "", rep_secs[r]);
  printf("]}\n");
  return 0;
}

}  // namespace

int main(int argc, char **argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: ref_driver main|pipe|nextdump|micro ...\n");
    return 1;
  }
  std::string cmd = argv[1];
  if (cmd == "main" && argc >= 8) {
    size_t J = atol(argv[2]), cf = atol(argv[3]), lhs_n = atol(argv[4]), rhs = atol(argv[5]);
    kBlockSize = atol(argv[6]);
    bool compact = atoi(argv[7]) != 0;
    kJoins = J;
    std::mt19937 gen(2);  // main.cpp:41-43
    std::uniform_int_distribution<> dist(0, rhs);
    std::vector<int64_t> lhs(lhs_n * J);
    for (size_t i = 0; i < lhs_n; ++i)
      for (size_t j = 0; j < J; ++j) lhs[i * J + j] = size_t(dist(gen));
    return RunPipe<HashTable>(lhs, lhs_n, J, cf, rhs, compact, argc > 8 ? argv[8] : nullptr);
  }
  if (cmd == "pipe" && argc >= 10) {
    size_t rows = atol(argv[3]), J = atol(argv[4]), cf = atol(argv[5]), rhs = atol(argv[6]);
    kBlockSize = atol(argv[7]);
    int kind = atoi(argv[8]);
    bool compact = atoi(argv[9]) != 0;
    kJoins = J;
    auto lhs = ReadFile(argv[2], rows * J);
    const char *dump = argc > 10 ? argv[10] : nullptr;
    return kind == 0 ? RunPipe<LPHashTable>(lhs, rows, J, cf, rhs, compact, dump)
                     : RunPipe<HashTable>(lhs, rows, J, cf, rhs, compact, dump);
  }
  if (cmd == "nextdump" && argc >= 10) {
    int kind = atoi(argv[2]);
    size_t n = atol(argv[3]), cf = atol(argv[4]);
    kBlockSize = atol(argv[5]);
    auto keys = ReadFile(argv[6], atol(argv[7]));
    bool inone = atoi(argv[8]) != 0;
    return kind == 0 ? NextDump<LPHashTable>(n, cf, keys, inone, argv[9]) : NextDump<HashTable>(n, cf, keys, inone, argv[9]);
  }
  if (cmd == "micro" && argc >= 10) {
    int kind = atoi(argv[2]), variant = atoi(argv[3]);
    size_t n = atol(argv[4]), cf = atol(argv[5]);
    kBlockSize = atol(argv[6]);
    auto keys = ReadFile(argv[7], atol(argv[8]));
    int procs = atoi(argv[9]);
    int share = argc > 10 ? atoi(argv[10]) : 0, reps = argc > 11 ? atoi(argv[11]) : 1;
    return kind == 0 ? Micro<LPHashTable>(n, cf, keys, variant, procs, share, reps)
                     : Micro<HashTable>(n, cf, keys, variant, procs, share, reps);
  }
  if (cmd == "bandit" && argc >= 3) {
    size_t steps = atol(argv[2]);
    auto &tuner = CompactTuner::Get();
    tuner.Initialize(0x1234);
    uint64_t lcg = 88172645463325252ULL;
    printf("{\"arms\": [");
    for (size_t i = 0; i < steps; ++i) {
      size_t thr = tuner.SelectArm(0);
      // synthetic reward: best threshold 256 in the first half, 768 in the second (forces a drift restart)
      lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL;
      double noise = double((lcg >> 33) % 1000) / 1000.0;
      double best = i < steps / 2 ? 256.0 : 768.0;
      double scale = i < steps / 2 ? 1.0 : 4.0;
      double reward = scale * (2.0 - std::fabs(double(thr) - best) / 1024.0) + 0.05 * noise;
      tuner.UpdateArm(0, thr, reward);
      printf("%s%zu", i ? ", " : "", thr);
    }
    printf("]}\n");
    return 0;
  }
  fprintf(stderr, "bad arguments\n");
  return 1;
}
