// micro_bench_main.cpp -- the reference's second driver, simd_micro_bench.cpp (:35-360), on the B200 path.
//
// Same CLI (--scale --hit-frequency --chunk-factor), same inputs (glibc rand() & (kRHSTuples * kHitFreq - 1) over
// kLHSTuples = 2^27 keys, identity selection vector, kBlockSize = 256 << scale, kRHSTuples = 128 << scale), same
// self-check: every variant prints `#tuples` and a human (or tests/test_gpu_facade_cpp.py) compares them.
// The reference's 8 variants are {chaining, linear probing} x {scalar, AVX-512} x {Next, InOneNext}; the AVX-512 twins
// compute the same results, so here there are 4 chunk-protocol variants (one C-ABI call per Probe / Next,
// the parity path) plus the fused whole-column probe (cc_probe_batch, the throughput path) per table kind.
// Cycle counts per phase (CycleProfiler) have no GPU equivalent -- the fused kernels have no separable phases; the
// wall time per variant is printed instead.
//   micro_bench_main --scale 3 --hit-frequency 2 --chunk-factor 1 [--lhs-tuples N] [--variants chunk|batch|all]
#include <chrono>
#include <cstdlib>
#include <cstring>

#include "simd_compaction.hpp"

using namespace simd_compaction;

namespace {

struct Outcome {
  const char *name;
  uint64_t n_tuples;
  double seconds;
};

template <class HT>
Outcome RunChunked(const char *name, const vector<int64_t> &keys, bool in_one) {
  HT hash_table(kRHSTuples, kChunkFactor);
  DataChunk input(vector<AttributeType>{AttributeType::INTEGER});
  DataChunk output(vector<AttributeType>{AttributeType::INTEGER, AttributeType::INTEGER, AttributeType::INTEGER});
  Vector keys_block(AttributeType::INTEGER);
  SelectionVector sel_vector(kBlockSize, false);
  Check(cc_sel_identity(sel_vector.data(), kBlockSize, nullptr));
  uint64_t n_tuples = 0;
  auto t0 = std::chrono::steady_clock::now();
  for (size_t k = 0; k < kLHSTuples; k += kBlockSize) {  // simd_micro_bench.cpp:89-104
    const size_t n_filling = std::min(kBlockSize, kLHSTuples - k);
    keys_block.data_->FromHost(keys.data() + k, n_filling);  // "load one block"
    input.data_[0] = keys_block;
    input.count_ = n_filling;
    auto scan_structure = hash_table.Probe(keys_block, n_filling, sel_vector);
    while (scan_structure.HasNext())
      n_tuples += in_one ? scan_structure.InOneNext(keys_block, input, output) : scan_structure.Next(keys_block, input, output);
  }
  Check(cc_stream_sync(nullptr));
  return {name, n_tuples, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()};
}

template <class HT>
Outcome RunBatch(const char *name, const vector<int64_t> &keys) {
  HT hash_table(kRHSTuples, kChunkFactor);
  DeviceArray<Attribute> d_keys(keys.size() ? keys.size() : 1, false);
  d_keys.FromHost(keys.data(), keys.size());
  const size_t cap = keys.size() * kChunkFactor + 1;  // every probe key matches at most chunk_factor build rows
  DeviceArray<Attribute> out_key(cap, false), out_payload(cap, false);
  DeviceArray<uint64_t> res(sizeof(cc_probe_result) / sizeof(uint64_t));
  Check(cc_stream_sync(nullptr));
  auto t0 = std::chrono::steady_clock::now();
  Check(cc_probe_batch(hash_table.Handle(), d_keys.data(), keys.size(), out_key.data(), out_payload.data(), nullptr, cap,
                       reinterpret_cast<cc_probe_result *>(res.data()), nullptr));
  Check(cc_stream_sync(nullptr));
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  auto r = res.ToHost();
  if (r[3]) throw std::runtime_error("output capacity overflow");
  return {name, r[0], secs};
}

}  // namespace

int main(int argc, char **argv) {
  std::string variants = "all";
  for (int i = 1; i + 1 < argc; i += 2) {  // simd_micro_bench.cpp:35-60
    std::string a = argv[i], v = argv[i + 1];
    if (a == "--scale") kScale = std::stoi(v);
    else if (a == "--hit-frequency") kHitFreq = std::stoi(v);
    else if (a == "--chunk-factor") kChunkFactor = std::stoi(v);
    else if (a == "--lhs-tuples") kLHSTuples = std::stoull(v);
    else if (a == "--variants") variants = v;
  }
  kBlockSize = 256 << kScale;  // :62-63
  kRHSTuples = 128 << kScale;
  try {
    // keys first: nothing may draw from rand() before the reference's own sequence (:78-79, default seed)
    vector<int64_t> keys(kLHSTuples);
    for (uint64_t i = 0; i < kLHSTuples; ++i) keys[i] = rand() & (kRHSTuples * kHitFreq - 1);
    Check(cc_device_init(0));
    vector<Outcome> out;
    if (variants == "all" || variants == "chunk") {
      out.push_back(RunChunked<HashTable>("chaining Probe+Next", keys, false));
      out.push_back(RunChunked<HashTable>("chaining Probe+InOneNext", keys, true));
      out.push_back(RunChunked<LPHashTable>("linear-probing Probe+Next", keys, false));
      out.push_back(RunChunked<LPHashTable>("linear-probing Probe+InOneNext", keys, true));
    }
    if (variants == "all" || variants == "batch") {
      out.push_back(RunBatch<HashTable>("chaining fused batch probe", keys));
      out.push_back(RunBatch<LPHashTable>("linear-probing fused batch probe", keys));
    }
    for (auto &o : out)
      fprintf(stderr, "--------------- %s ---------------\n%.3f s  %.1f M probe tuples/s\n#tuples: %llu\n", o.name, o.seconds,
              kLHSTuples / o.seconds / 1e6, (unsigned long long) o.n_tuples);
    printf("{\"scale\": %zu, \"hit_frequency\": %zu, \"chunk_factor\": %zu, \"lhs_tuples\": %zu, \"variants\": [", kScale, kHitFreq,
           kChunkFactor, kLHSTuples);
    for (size_t i = 0; i < out.size(); ++i)
      printf("%s{\"name\": \"%s\", \"n_tuples\": %llu, \"seconds\": %.6f}", i ? ", " : "", out[i].name, (unsigned long long) out[i].n_tuples,
             out[i].seconds);
    printf("]}\n");
  } catch (const std::exception &e) {
    fprintf(stderr, "micro_bench_main: %s\n", e.what());
    return 2;
  }
  return 0;
}
