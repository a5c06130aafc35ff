// simd_compaction.hpp -- C++ facade: the reference's class surface (namespace simd_compaction,
// SURVEY 8b) re-created on top of the C ABI of libccb200 (include/cc_api.h).
//
// A driver written against the reference (main.cpp / simd_micro_bench.cpp) ports by swapping its
// includes for this header and linking libccb200.so: same class names, same method names, same
// argument meaning, same call protocol
//     ss = ht.Probe(join_key, count, sel);  while (ss.HasNext()) { ss.Next(join_key, input, result); ... }
// The one deliberate difference: columns and selection vectors are DEVICE resident (HBM), so
// `Vector::data_` / `DataChunk::selection_vector_` are device buffers; GetValue()/ToHost() copy to the
// host for inspection.  Errors of the C ABI surface as simd_compaction::Error (the reference only asserts).
//
// Header-only; needs C++17.  Build:  g++ -std=c++17 -Iinclude -I<pkg>/host driver.cpp -L<pkg> -lccb200
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "cc_api.h"

namespace simd_compaction {

using std::shared_ptr;
using std::string;
using std::unique_ptr;
using std::vector;
using idx_t = size_t;

struct Error : std::runtime_error {
  explicit Error(int code) : std::runtime_error("libccb200 error " + std::to_string(code) + ": " + cc_last_error()), code_(code) {}
  int code_;
};
inline void Check(int rc) {
  if (rc != CC_OK) throw Error(rc);
}

// ---- base.h:37-51 ----------------------------------------------------------------------------
inline size_t kScale = 0;
inline size_t kBlockSize = 256 << kScale;
inline size_t kRHSTuples = 128 << kScale;
inline size_t kLHSTuples = 1024 << 17;
inline size_t kHitFreq = 1;
inline size_t kJoins = 3;
inline size_t kLHSTupleSize = 2e7;
inline size_t kRHSTupleSize = 2e6;
inline size_t kChunkFactor = 1;

using Attribute = int64_t;
enum class AttributeType : uint8_t { INTEGER = 0, INVALID = 3 };

// device buffer, owned (cc_malloc / cc_free) or borrowed (library-owned storage)
template <class T>
class DeviceArray {
 public:
  DeviceArray() = default;
  explicit DeviceArray(size_t n, bool zero = true) : n_(n), owned_(true) {
    void *p = nullptr;
    Check(cc_malloc(&p, n * sizeof(T)));
    ptr_ = static_cast<T *>(p);
    if (zero) Check(cc_memset(ptr_, 0, n * sizeof(T), nullptr));
  }
  DeviceArray(T *borrowed, size_t n) : ptr_(borrowed), n_(n), owned_(false) {}
  DeviceArray(const DeviceArray &) = delete;
  DeviceArray &operator=(const DeviceArray &) = delete;
  ~DeviceArray() {
    if (owned_ && ptr_) cc_free(ptr_);
  }
  T *data() const { return ptr_; }
  size_t size() const { return n_; }
  vector<T> ToHost(size_t count = SIZE_MAX) const {
    vector<T> h(count < n_ ? count : n_);
    if (!h.empty()) {
      Check(cc_memcpy_d2h(h.data(), ptr_, h.size() * sizeof(T), nullptr));
      Check(cc_stream_sync(nullptr));
    }
    return h;
  }
  void FromHost(const T *h, size_t count, size_t offset = 0) {
    if (count) {
      Check(cc_memcpy_h2d(ptr_ + offset, h, count * sizeof(T), nullptr));
      Check(cc_stream_sync(nullptr));
    }
  }

 private:
  T *ptr_ = nullptr;
  size_t n_ = 0;
  bool owned_ = false;
};

using SelectionVector = DeviceArray<uint32_t>;  // vector<uint32_t> of base.h:84, device resident

// base.h:59-76
class Vector {
 public:
  AttributeType type_;
  shared_ptr<DeviceArray<Attribute>> data_;

  Vector() : type_(AttributeType::INTEGER), data_(std::make_shared<DeviceArray<Attribute>>(kBlockSize)) {}
  explicit Vector(AttributeType type) : type_(type), data_(std::make_shared<DeviceArray<Attribute>>(kBlockSize)) {}
  Vector(AttributeType type, shared_ptr<DeviceArray<Attribute>> data) : type_(type), data_(std::move(data)) {}

  void Reference(Vector &other) { data_ = other.data_; }  // base.cpp:5-8
  Attribute GetValue(size_t idx) const {                  // host read of one element (inspection only)
    Attribute v = 0;
    Check(cc_memcpy_d2h(&v, data_->data() + idx, sizeof(v), nullptr));
    Check(cc_stream_sync(nullptr));
    return v;
  }
  void SetValue(size_t idx, Attribute v) { data_->FromHost(&v, 1, idx); }
  Attribute *Data() { return data_->data(); }
};

// base.h:79-100
class DataChunk {
 public:
  size_t count_;
  vector<Vector> data_;
  vector<AttributeType> types_;
  shared_ptr<SelectionVector> selection_vector_;

  explicit DataChunk(const vector<AttributeType> &types)
      : count_(0), types_(types), selection_vector_(std::make_shared<SelectionVector>(kBlockSize, false)) {
    for (auto &type : types) data_.emplace_back(type);
    Check(cc_sel_identity(selection_vector_->data(), kBlockSize, nullptr));
  }

  void Append(DataChunk &chunk, size_t num, size_t offset = 0) {  // base.cpp:15-27
    vector<int64_t *> dst, src;
    for (auto &v : data_) dst.push_back(v.Data());
    for (auto &v : chunk.data_) src.push_back(v.Data());
    Check(cc_chunk_append(dst.data(), count_, src.data(), chunk.selection_vector_->data(), num, offset, types_.size(), nullptr));
    count_ += num;
  }
  void AppendTuple(vector<Attribute> &tuple) {  // base.cpp:29-35
    for (size_t i = 0; i < types_.size(); ++i) data_[i].SetValue(count_, tuple[i]);
    ++count_;
  }
  void Slice(DataChunk &other, SelectionVector &selection_vector, size_t count) {  // base.cpp:37-47
    count_ = count;
    for (size_t c = 0; c < other.data_.size(); ++c) data_[c].Reference(other.data_[c]);
    Check(cc_sel_compose(selection_vector_->data(), other.selection_vector_->data(), selection_vector.data(), count, nullptr));
  }
  void SIMDSlice(DataChunk &other, SelectionVector &selection_vector, size_t count) { Slice(other, selection_vector, count); }
  void Reset() {  // base.h:96-99
    count_ = 0;
    Check(cc_sel_identity(selection_vector_->data(), kBlockSize, nullptr));
  }
};

// data_collection.h:15-33 -- row-major host rows like the reference; FetchChunk / AppendChunk do the
// row <-> column transposes on the device (cc_rows_to_columns / cc_columns_to_rows)
class DataCollection {
 public:
  explicit DataCollection(vector<AttributeType> &types) : types_(types), n_tuples_(0) {}

  void AppendTuple(vector<Attribute> &tuple) {
    collection_.insert(collection_.end(), tuple.begin(), tuple.end());
    ++n_tuples_;
  }
  void AppendChunk(DataChunk &chunk) {  // data_collection.cpp:10-21
    if (chunk.count_ == 0) return;
    size_t nc = types_.size();
    DeviceArray<Attribute> rows(chunk.count_ * nc, false);
    vector<const int64_t *> cols;
    for (auto &v : chunk.data_) cols.push_back(v.Data());
    Check(cc_columns_to_rows(cols.data(), chunk.selection_vector_->data(), chunk.count_, nc, rows.data(), nullptr));
    auto h = rows.ToHost();
    collection_.insert(collection_.end(), h.begin(), h.end());
    n_tuples_ += chunk.count_;
  }
  DataChunk FetchChunk(size_t start, size_t end) {  // data_collection.cpp:23-27
    DataChunk chunk(types_);
    size_t nc = types_.size(), n = end - start;
    if (n) {
      DeviceArray<Attribute> rows(n * nc, false);
      rows.FromHost(collection_.data() + start * nc, n * nc);
      vector<int64_t *> cols;
      for (auto &v : chunk.data_) cols.push_back(v.Data());
      Check(cc_rows_to_columns(rows.data(), n, nc, cols.data(), nullptr));
      Check(cc_stream_sync(nullptr));
    }
    chunk.count_ = n;
    return chunk;
  }
  inline size_t NumTuples() const { return n_tuples_; }
  void Print(size_t n_tuple) {  // data_collection.cpp:29-45
    n_tuple = std::min(n_tuple, n_tuples_);
    for (size_t i = 0; i < n_tuple; ++i) {
      for (size_t j = 0; j < types_.size(); ++j)
        if (types_[j] == AttributeType::INTEGER) std::cout << collection_[i * types_.size() + j] << ", ";
      std::cout << "\n";
    }
  }
  const vector<Attribute> &Rows() const { return collection_; }  // row-major, NumTuples() x #columns

 private:
  vector<AttributeType> types_;
  size_t n_tuples_;
  vector<Attribute> collection_;
};

// ---- chaining_ht.h / linear_probing_ht.h ------------------------------------------------------------
template <int KIND>
class BasicScanStructure {
 public:
  explicit BasicScanStructure(cc_scan *scan) : scan_(scan) {}
  BasicScanStructure(BasicScanStructure &&o) noexcept : scan_(o.scan_) { o.scan_ = nullptr; }
  BasicScanStructure(const BasicScanStructure &) = delete;
  ~BasicScanStructure() {
    if (scan_) cc_scan_destroy(scan_);
  }

  size_t Next(Vector &join_key, DataChunk &input, DataChunk &result) { return Step(0, join_key, input, result); }
  size_t InOneNext(Vector &join_key, DataChunk &input, DataChunk &result) { return Step(1, join_key, input, result); }
  // the AVX-512 twins compute the same results; compact_mode is unused in the reference too
  size_t SIMDNext(Vector &join_key, DataChunk &input, DataChunk &result, bool = true) { return Step(0, join_key, input, result); }
  size_t SIMDInOneNext(Vector &join_key, DataChunk &input, DataChunk &result, bool = false) { return Step(1, join_key, input, result); }
  inline bool HasNext() const { return cc_scan_has_next(scan_) != 0; }

 private:
  size_t Step(int in_one, Vector &join_key, DataChunk &input, DataChunk &result) {
    size_t n_in = input.data_.size();
    for (size_t c = 0; c < n_in; ++c) result.data_[c].Reference(input.data_[c]);  // Slice: base.cpp:40
    size_t count = 0;
    Check(cc_scan_next(scan_, in_one, join_key.Data(), input.selection_vector_->data(), result.selection_vector_->data(),
                       result.data_[n_in + 1].Data(), &count, nullptr));
    result.count_ = count;
    return count;
  }
  cc_scan *scan_;
};

template <int KIND>
class BasicHashTable {
 public:
  using Scan = BasicScanStructure<KIND>;
  BasicHashTable(size_t n_rhs_tuples, size_t chunk_factor) { Check(cc_ht_build_reference(&ht_, KIND, n_rhs_tuples, chunk_factor, nullptr)); }
  // keep_payload: keep the `payload = cnt + 10000000` column the reference generates and drops (chaining_ht.cpp:21-23,34)
  BasicHashTable(size_t n_rhs_tuples, size_t chunk_factor, bool keep_payload) {
    Check(keep_payload ? cc_ht_build_reference_payload(&ht_, KIND, n_rhs_tuples, chunk_factor, nullptr)
                       : cc_ht_build_reference(&ht_, KIND, n_rhs_tuples, chunk_factor, nullptr));
  }
  // explicit build tuples (host): keys plus payload columns
  BasicHashTable(const vector<Attribute> &keys, const vector<vector<Attribute>> &payload_cols) : BasicHashTable(keys) {
    if (payload_cols.empty()) return;
    DeviceArray<Attribute> k(keys.size() ? keys.size() : 1, false);
    k.FromHost(keys.data(), keys.size());
    vector<std::unique_ptr<DeviceArray<Attribute>>> cols;
    vector<const int64_t *> ptrs;
    for (auto &c : payload_cols) {
      cols.push_back(std::make_unique<DeviceArray<Attribute>>(keys.size() ? keys.size() : 1, false));
      cols.back()->FromHost(c.data(), std::min(c.size(), keys.size()));
      ptrs.push_back(cols.back()->data());
    }
    Check(cc_ht_attach_payload(ht_, k.data(), ptrs.data(), ptrs.size(), nullptr));
  }
  size_t PayloadColumns() const { return cc_ht_payload_cols(ht_); }
  // Whole-column probe with dense output (cc_probe_batch_payload): returns the result rows
  // (probe key, matched build key, payload columns...) column-major on the host.
  vector<vector<Attribute>> ProbeBatchPayload(const vector<Attribute> &probe_keys, size_t capacity) {
    const size_t ncol = PayloadColumns();
    DeviceArray<Attribute> k(probe_keys.size() ? probe_keys.size() : 1, false);
    k.FromHost(probe_keys.data(), probe_keys.size());
    vector<std::unique_ptr<DeviceArray<Attribute>>> out;
    vector<int64_t *> ptrs;
    for (size_t c = 0; c < 2 + ncol; ++c) {
      out.push_back(std::make_unique<DeviceArray<Attribute>>(capacity ? capacity : 1, false));
      ptrs.push_back(out.back()->data());
    }
    DeviceArray<uint64_t> res(sizeof(cc_probe_payload_result) / sizeof(uint64_t));
    Check(cc_probe_batch_payload(ht_, k.data(), probe_keys.size(), ptrs[0], ptrs[1], ptrs.data() + 2, ncol, nullptr, capacity,
                                 reinterpret_cast<cc_probe_payload_result *>(res.data()), nullptr));
    Check(cc_stream_sync(nullptr));
    auto r = res.ToHost();
    const size_t m = std::min<size_t>(r[0], capacity);
    vector<vector<Attribute>> cols;
    for (auto &o : out) cols.push_back(o->ToHost(m));
    return cols;
  }
  // explicit build keys (host) -- the reference has no external build-input API
  explicit BasicHashTable(const vector<Attribute> &keys) {
    DeviceArray<Attribute> d(keys.size() ? keys.size() : 1, false);
    d.FromHost(keys.data(), keys.size());
    Check(cc_ht_build(&ht_, KIND, d.data(), keys.size(), CC_BUILD_ORDERED, nullptr));
  }
  BasicHashTable(const BasicHashTable &) = delete;
  ~BasicHashTable() {
    if (ht_) cc_ht_destroy(ht_);
  }
  Scan Probe(Vector &join_key, size_t count, SelectionVector &sel_vec) {
    cc_scan *s = nullptr;
    Check(cc_probe_chunk(ht_, join_key.Data(), count, sel_vec.data(), kBlockSize, &s, nullptr));
    return Scan(s);
  }
  Scan SIMDProbe(Vector &join_key, size_t count, SelectionVector &sel_vec) { return Probe(join_key, count, sel_vec); }
  cc_ht *Handle() const { return ht_; }

 private:
  cc_ht *ht_ = nullptr;
};

using HashTable = BasicHashTable<CC_HT_CHAIN>;         // chaining_ht.h:86
using ScanStructure = BasicScanStructure<CC_HT_CHAIN>; // chaining_ht.h:29
using LPHashTable = BasicHashTable<CC_HT_LP>;          // linear_probing_ht.h:56
using LPScanStructure = BasicScanStructure<CC_HT_LP>;  // linear_probing_ht.h:24

// ---- compactor.h --------------------------------------------------------------------------------------
// One class, three aliases: the threshold decides (setting.h:17-29 picks the alias at compile time).
class NaiveCompactor {
 public:
  explicit NaiveCompactor(vector<AttributeType> &types) : types_(types) { Check(cc_compactor_create(&c_, types.size(), kBlockSize, kBlockSize)); }
  NaiveCompactor(const NaiveCompactor &) = delete;
  ~NaiveCompactor() {
    if (c_) cc_compactor_destroy(c_);
  }
  void SetThreshold(size_t t) { Check(cc_compactor_set_threshold(c_, t)); }  // main.cpp:141
  size_t GetThreshold() const { return cc_compactor_get_threshold(c_); }     // main.cpp:166

  void Compact(unique_ptr<DataChunk> &chunk) {  // compactor.cpp:5-41
    size_t nc = types_.size(), count = chunk->count_;
    vector<int64_t *> cols, out(nc);
    for (auto &v : chunk->data_) cols.push_back(v.Data());
    const uint32_t *out_sel = nullptr;
    Check(cc_compactor_compact(c_, cols.data(), chunk->selection_vector_->data(), &count, out.data(), &out_sel, nullptr));
    if (out_sel == chunk->selection_vector_->data()) {  // pass-through or buffered (count == 0)
      if (count == 0) chunk->Reset();
      return;
    }
    chunk = Wrap(out, out_sel, count);
  }
  inline void Flush(unique_ptr<DataChunk> &chunk) {  // compactor.h:23
    size_t nc = types_.size(), count = 0;
    vector<int64_t *> out(nc);
    const uint32_t *out_sel = nullptr;
    Check(cc_compactor_flush(c_, out.data(), &out_sel, &count, nullptr));
    chunk = Wrap(out, out_sel, count);
  }

 private:
  // The reference hands the full cache chunk to the caller by swapping unique_ptrs (compactor.cpp:33-34) and the
  // caller then reuses it as its scratch `result` for the following Next() calls.  The C ABI lends out library-owned
  // storage instead (valid until the next Compact/Flush), so the facade copies the dense rows into a chunk of its own.
  unique_ptr<DataChunk> Wrap(const vector<int64_t *> &cols, const uint32_t *sel, size_t count) {
    (void) sel;  // dense chunk: identity selection, set up by the DataChunk constructor
    auto ch = std::make_unique<DataChunk>(types_);
    for (size_t j = 0; j < types_.size(); ++j)
      if (count) Check(cc_memcpy_d2d(ch->data_[j].Data(), cols[j], count * sizeof(Attribute), nullptr));
    ch->count_ = count;
    return ch;
  }
  vector<AttributeType> types_;
  cc_compactor *c_ = nullptr;
};
using BinaryCompactor = NaiveCompactor;   // setting.h:21 (absent from the reference)
using DynamicCompactor = NaiveCompactor;  // setting.h:24 (absent from the reference)

// ---- negative_feedback.hpp:165-260 ----------------------------------------------------------------------
class CompactTuner {
 public:
  static CompactTuner &Get() {
    static CompactTuner instance;
    return instance;
  }
  inline void Initialize(size_t address, const vector<size_t> &arms = {0, 32, 64, 128, 256, 384, 512, 768, 1024}) {
    Check(cc_tuner_initialize(t_, address, arms.data(), arms.size()));
  }
  inline size_t SelectArm(idx_t id) {
    size_t v = 0;
    Check(cc_tuner_select_arm(t_, id, &v));
    return v;
  }
  inline void UpdateArm(idx_t id, size_t arm, double reward) { Check(cc_tuner_update_arm(t_, id, arm, reward)); }
  inline void Reset(bool enable_log = false) { Check(cc_tuner_reset(t_, enable_log, nullptr)); }
  inline int64_t GetId(size_t address) { return cc_tuner_get_id(t_, address); }
  inline size_t GetBanditSize() { return cc_tuner_bandit_size(t_); }

 private:
  CompactTuner() { Check(cc_tuner_create(&t_)); }
  ~CompactTuner() { cc_tuner_destroy(t_); }
  cc_tuner *t_ = nullptr;
};


// ---- partitioned multi-GPU join (new functionality; cc_pjoin_* of the C ABI) -------------------------------
// LocalComm: the control plane for ONE NODE without any library -- the ranks are processes forked from one parent AFTER the
// communicator was created (and BEFORE any of them touches CUDA), talking through an anonymous shared mapping: a
// sense-reversing barrier and a staging area for all-gathers.  Hand `Comm()` to cc_pjoin_create / PartitionedJoin.
}  // namespace simd_compaction
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <cstring>
namespace simd_compaction {

class LocalComm {
 public:
  static constexpr size_t kSlotBytes = 256;
  explicit LocalComm(int world) : world_(world) {
    const size_t bytes = sizeof(Shared) + (size_t) world * kSlotBytes;
    void *m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (m == MAP_FAILED) throw std::runtime_error("LocalComm: mmap failed");
    sh_ = new (m) Shared();
    slots_ = reinterpret_cast<unsigned char *>(m) + sizeof(Shared);
    bytes_ = bytes;
  }
  ~LocalComm() { munmap(sh_, bytes_); }
  // forks world - 1 children; returns this process's rank (the parent is rank 0)
  int Fork() {
    for (int r = 1; r < world_; ++r) {
      pid_t pid = fork();
      if (pid < 0) throw std::runtime_error("LocalComm: fork failed");
      if (pid == 0) {
        rank_ = r;
        children_.clear();
        return r;
      }
      children_.push_back(pid);
    }
    rank_ = 0;
    return 0;
  }
  // rank 0: waits for the children; returns true if all of them exited with status 0
  bool Join() {
    bool ok = true;
    for (pid_t pid : children_) {
      int st = 0;
      if (waitpid(pid, &st, 0) < 0 || !WIFEXITED(st) || WEXITSTATUS(st) != 0) ok = false;
    }
    children_.clear();
    return ok;
  }
  void Barrier() {
    const unsigned sense = local_sense_ ^= 1u;
    if (sh_->arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == (unsigned) world_) {
      sh_->arrived.store(0, std::memory_order_relaxed);
      sh_->sense.store(sense, std::memory_order_release);
    } else {
      // a rank that failed never arrives: give up instead of waiting for it forever
      for (unsigned long spins = 0; sh_->sense.load(std::memory_order_acquire) != sense; ++spins) {
        if (sh_->failed.load(std::memory_order_acquire)) throw std::runtime_error("LocalComm: a peer rank failed");
        if (spins > 2400000ul) throw std::runtime_error("LocalComm: barrier timed out (120 s)");
        usleep(50);
      }
    }
  }
  void Fail() { sh_->failed.store(1, std::memory_order_release); }  // call before leaving on an error: releases the waiting peers
  void AllGather(const void *send, void *recv, size_t bytes) {
    if (bytes > kSlotBytes) throw std::runtime_error("LocalComm: all-gather record too large");
    std::memcpy(slots_ + (size_t) rank_ * kSlotBytes, send, bytes);
    Barrier();
    for (int r = 0; r < world_; ++r) std::memcpy(static_cast<unsigned char *>(recv) + (size_t) r * bytes, slots_ + (size_t) r * kSlotBytes, bytes);
    Barrier();
  }
  cc_comm Comm() {
    cc_comm c;
    c.rank = rank_;
    c.world = world_;
    c.user = this;
    c.allgather = [](void *u, const void *s, void *r, size_t b) -> int {
      try {
        static_cast<LocalComm *>(u)->AllGather(s, r, b);
        return 0;
      } catch (...) {
        return 1;
      }
    };
    c.barrier = [](void *u) -> int {
      try {
        static_cast<LocalComm *>(u)->Barrier();
        return 0;
      } catch (...) {
        return 1;
      }
    };
    return c;
  }
  int Rank() const { return rank_; }
  int World() const { return world_; }

 private:
  struct Shared {
    std::atomic<unsigned> arrived{0};
    std::atomic<unsigned> sense{0};
    std::atomic<unsigned> failed{0};
  };
  Shared *sh_ = nullptr;
  unsigned char *slots_ = nullptr;
  size_t bytes_ = 0;
  int world_ = 1, rank_ = 0;
  unsigned local_sense_ = 0;
  vector<pid_t> children_;
};

// The partitioned join as a class: build once (collective), probe many times (collective, one call per rank and step).
class PartitionedJoin {
 public:
  PartitionedJoin(const cc_comm &comm, int kind, const Attribute *d_build_keys, size_t n_build_local, size_t max_probe_rows, int n_sub = 4) {
    Check(cc_pjoin_create(&h_, &comm, kind, d_build_keys, n_build_local, max_probe_rows, n_sub, nullptr));
  }
  PartitionedJoin(const PartitionedJoin &) = delete;
  PartitionedJoin &operator=(const PartitionedJoin &) = delete;
  ~PartitionedJoin() {
    if (h_) cc_pjoin_destroy(h_);
  }
  // enqueues partition + exchange + probe of this rank's keys; nothing is synchronised
  void Probe(const Attribute *d_keys, size_t n, Attribute *d_out_key, Attribute *d_out_payload, size_t out_capacity, cc_probe_result *d_result) {
    Check(cc_pjoin_probe(h_, d_keys, n, d_out_key, d_out_payload, out_capacity, d_result, nullptr));
  }
  // the two halves, for software pipelining across steps: ProbeBegin(t + 1) may be called before ProbeEnd(t)
  void ProbeBegin(const Attribute *d_keys, size_t n) { Check(cc_pjoin_probe_begin(h_, d_keys, n, nullptr)); }
  void ProbeEnd(Attribute *d_out_key, Attribute *d_out_payload, size_t out_capacity, cc_probe_result *d_result) {
    Check(cc_pjoin_probe_end(h_, d_out_key, d_out_payload, out_capacity, d_result, nullptr));
  }
  cc_ht_info TableInfo() const {
    const cc_ht *t = nullptr;
    Check(cc_pjoin_table(h_, &t));
    cc_ht_info i;
    Check(cc_ht_get_info(t, &i));
    return i;
  }

 private:
  cc_pjoin *h_ = nullptr;
};

}  // namespace simd_compaction
