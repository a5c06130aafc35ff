// pipeline_main.cpp -- the reference's main.cpp driver (main.cpp:37-243) re-written against the
// facade (simd_compaction.hpp): same CLI, same LHS generator (std::mt19937 gen(2),
// uniform_int_distribution<>(0, rhs) drawn row-major), same operator protocol
//     ExecutePipeline: Probe -> while (HasNext) { Next -> Compact -> recurse }   (main.cpp:119-170)
//     FlushPipelineCache                                                          (main.cpp:172-191)
// plus `--mode fused`, which hands the whole LHS table to cc_chain_execute (one persistent kernel,
// in-kernel compaction).  Prints one JSON line: tuple count, per-column sums and the SURVEY 8c digest.
//
//   pipeline_main --join-num 4 --chunk-factor 5 --lhs-size 300000 --rhs-size 20000
//                 [--block 256] [--compact none|full|<threshold>] [--mode chunk|fused|payload] [--table chain|lp]
#include <chrono>
#include <cstring>
#include <random>

#include "simd_compaction.hpp"

using namespace simd_compaction;

namespace {

uint64_t murmur(uint64_t x) {  // hash_functions.h:8-16 (host copy for the digest only)
  x ^= x >> 32;
  x *= 0xd6e8feb86659fd93ULL;
  x ^= x >> 32;
  x *= 0xd6e8feb86659fd93ULL;
  x ^= x >> 32;
  return x;
}

template <class HT>
struct PipelineState {  // main.cpp:14-20
  vector<unique_ptr<HT>> hts;
  vector<unique_ptr<DataChunk>> intermediates;
  vector<unique_ptr<NaiveCompactor>> compactors;
};

template <class HT>
void ExecutePipeline(DataChunk &input, PipelineState<HT> &state, DataCollection &result_table, size_t level, bool compact) {
  if (level == state.hts.size()) {  // ResultCollector (flag_collect_tuples == true here)
    result_table.AppendChunk(input);
    return;
  }
  auto &join_key = input.data_[level];
  auto &result = state.intermediates[level];
  auto ss = state.hts[level]->Probe(join_key, input.count_, *input.selection_vector_);
  while (ss.HasNext()) {
    ss.Next(join_key, input, *result);
    if (compact) {
      state.compactors[level]->Compact(result);
      if (result->count_ == 0) continue;
    }
    ExecutePipeline(*result, state, result_table, level + 1, compact);
  }
}

template <class HT>
void FlushPipelineCache(PipelineState<HT> &state, DataCollection &result_table, size_t level) {
  if (level == state.hts.size()) return;
  auto &result = state.intermediates[level];
  state.compactors[level]->Flush(result);
  ExecutePipeline(*result, state, result_table, level + 1, true);
  FlushPipelineCache(state, result_table, level + 1);
}

void Report(const char *mode, size_t n_cols, const vector<Attribute> &rows, double seconds) {
  size_t n = n_cols ? rows.size() / n_cols : 0;
  vector<uint64_t> colsum(n_cols, 0);
  uint64_t digest = 0;
  for (size_t i = 0; i < n; ++i) {
    uint64_t th = 0x9e3779b97f4a7c15ULL;
    for (size_t j = 0; j < n_cols; ++j) {
      uint64_t v = (uint64_t) rows[i * n_cols + j];
      th = murmur(th ^ v) + j;
      colsum[j] += v;
    }
    digest += th;
  }
  printf("{\"mode\": \"%s\", \"n_tuples\": %zu, \"digest\": %llu, \"seconds\": %.6f, \"colsum\": [", mode, n, (unsigned long long) digest, seconds);
  for (size_t j = 0; j < n_cols; ++j) printf("%s%llu", j ? ", " : "", (unsigned long long) colsum[j]);
  printf("]}\n");
}

template <class HT>
int Run(bool fused, bool compact, size_t threshold) {
  std::mt19937 gen(2);  // main.cpp:41-43
  std::uniform_int_distribution<> dist(0, kRHSTupleSize);
  vector<AttributeType> types(kJoins, AttributeType::INTEGER);
  DataCollection table(types);
  vector<Attribute> tuple(kJoins);
  for (size_t i = 0; i < kLHSTupleSize; ++i) {
    for (size_t j = 0; j < kJoins; ++j) tuple[j] = size_t(dist(gen));
    table.AppendTuple(tuple);
  }
  PipelineState<HT> state;
  for (size_t i = 0; i < kJoins; ++i) {  // main.cpp:62-68
    state.hts.push_back(std::make_unique<HT>(kRHSTupleSize, kChunkFactor));
    types.push_back(AttributeType::INTEGER);
    types.push_back(AttributeType::INTEGER);
    state.intermediates.push_back(std::make_unique<DataChunk>(types));
    state.compactors.push_back(std::make_unique<NaiveCompactor>(types));
    state.compactors.back()->SetThreshold(threshold);
  }
  DataCollection result_table(types);
  auto t0 = std::chrono::steady_clock::now();
  if (!fused) {
    size_t start = 0, end;
    do {  // main.cpp:86-95
      end = std::min(start + kBlockSize, kLHSTupleSize);
      DataChunk chunk = table.FetchChunk(start, end);
      start = end;
      ExecutePipeline(chunk, state, result_table, 0, compact);
    } while (end < kLHSTupleSize);
    if (compact) FlushPipelineCache(state, result_table, 0);
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    Report("chunk", 3 * kJoins, result_table.Rows(), secs);
    return 0;
  }
  // fused: upload the LHS columns once, one kernel for the whole chain, materialised result columns
  size_t J = kJoins, n = kLHSTupleSize;
  vector<unique_ptr<DeviceArray<Attribute>>> cols;
  vector<const int64_t *> col_ptrs;
  vector<const cc_ht *> tabs;
  const auto &rows = table.Rows();
  vector<Attribute> col(n);
  for (size_t j = 0; j < J; ++j) {
    for (size_t i = 0; i < n; ++i) col[i] = rows[i * J + j];
    cols.push_back(std::make_unique<DeviceArray<Attribute>>(n ? n : 1, false));
    cols.back()->FromHost(col.data(), n);
    col_ptrs.push_back(cols.back()->data());
    tabs.push_back(state.hts[j]->Handle());
  }
  vector<uint32_t> thr(J, (uint32_t) (compact ? threshold : 0));
  DeviceArray<cc_chain_result> d_res(1);
  cc_chain_result res;
  t0 = std::chrono::steady_clock::now();
  Check(cc_chain_execute(tabs.data(), J, col_ptrs.data(), n, thr.data(), nullptr, 0, d_res.data(), nullptr));  // count first
  res = d_res.ToHost()[0];
  size_t cap = res.n_tuples ? res.n_tuples : 1;
  vector<unique_ptr<DeviceArray<Attribute>>> outs;
  vector<int64_t *> out_ptrs;
  for (size_t j = 0; j < 3 * J; ++j) {
    outs.push_back(std::make_unique<DeviceArray<Attribute>>(cap, false));
    out_ptrs.push_back(outs.back()->data());
  }
  Check(cc_chain_execute(tabs.data(), J, col_ptrs.data(), n, thr.data(), out_ptrs.data(), cap, d_res.data(), nullptr));
  res = d_res.ToHost()[0];
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  vector<Attribute> out_rows(res.n_tuples * 3 * J);
  for (size_t j = 0; j < 3 * J; ++j) {
    auto h = outs[j]->ToHost(res.n_tuples);
    for (size_t i = 0; i < res.n_tuples; ++i) out_rows[i * 3 * J + j] = h[i];
  }
  Report("fused", 3 * J, out_rows, secs);
  return 0;
}

// `--mode payload`: ONE join over the first LHS column against a table that keeps the payload the reference
// generates and drops (chaining_ht.cpp:21-23,34); result rows (probe key, build key, payload) through
// HT::ProbeBatchPayload.  Reported like the other modes (3 columns).
template <class HT>
int RunPayload() {
  std::mt19937 gen(2);  // main.cpp:41-43, first column only
  std::uniform_int_distribution<> dist(0, kRHSTupleSize);
  vector<Attribute> keys(kLHSTupleSize);
  for (size_t i = 0; i < kLHSTupleSize; ++i) {
    keys[i] = size_t(dist(gen));
    for (size_t j = 1; j < kJoins; ++j) (void) dist(gen);  // row-major draw order of the reference
  }
  HT ht(kRHSTupleSize, kChunkFactor, /*keep_payload=*/true);
  auto t0 = std::chrono::steady_clock::now();
  auto cols = ht.ProbeBatchPayload(keys, kLHSTupleSize * kChunkFactor);
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  const size_t n = cols.empty() ? 0 : cols[0].size();
  vector<Attribute> rows(n * cols.size());
  for (size_t c = 0; c < cols.size(); ++c)
    for (size_t i = 0; i < n; ++i) rows[i * cols.size() + c] = cols[c][i];
  Report("payload", cols.size(), rows, secs);
  return 0;
}

}  // namespace

int main(int argc, char **argv) {
  bool fused = false, compact = false, lp = false, payload = false;
  size_t threshold = 0;
  kBlockSize = 256;
  for (int i = 1; i + 1 < argc; i += 2) {
    std::string a = argv[i], v = argv[i + 1];
    if (a == "--join-num") kJoins = std::stoi(v);
    else if (a == "--chunk-factor") kChunkFactor = std::stoi(v);
    else if (a == "--lhs-size") kLHSTupleSize = std::stoi(v);
    else if (a == "--rhs-size") kRHSTupleSize = std::stoi(v);
    else if (a == "--block") kBlockSize = std::stoi(v);
    else if (a == "--mode") fused = v == "fused", payload = v == "payload";
    else if (a == "--table") lp = v == "lp";
    else if (a == "--compact") {
      compact = v != "none";
      threshold = v == "full" ? SIZE_MAX : (compact ? std::stoul(v) : 0);
    }
  }
  if (threshold == SIZE_MAX) threshold = fused ? CC_CHAIN_WIDTH : kBlockSize;
  try {
    Check(cc_device_init(0));
    if (payload) return lp ? RunPayload<LPHashTable>() : RunPayload<HashTable>();
    return lp ? Run<LPHashTable>(fused, compact, threshold) : Run<HashTable>(fused, compact, threshold);
  } catch (const std::exception &e) {
    fprintf(stderr, "pipeline_main: %s\n", e.what());
    return 2;
  }
}
