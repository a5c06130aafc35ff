// pjoin_main.cpp -- the hash-partitioned multi-GPU join (BASELINE config 5) driven from C++ through the C ABI alone:
// no Python, no NCCL, no MPI.  One process per GPU, forked from this binary; the control plane is LocalComm (fork + shared
// memory, host/simd_compaction.hpp), the data plane is libccb200 (cc_pjoin_*: partition kernels, copy-engine block copies into
// CUDA-IPC-mapped peer memory over NVLink, device-side ready / consumed flags).
//
// The build side is the reference's own key column (chaining_ht.cpp:15-26: n_build rows, chunk_factor copies per key), the
// probe side the counter generator of SURVEY 8d (murmurhash64(seed + i) & (n_build * hit - 1)); both start range-partitioned
// over the ranks.  Every step each rank calls cc_pjoin_probe on its share; results stay sharded.  Checked on every run: the
// match count and both column checksums against a host recomputation (every probe key k matches chunk_factor build rows iff
// k is one of the generated keys), and the owner property of every result row.  --dump writes each rank's result rows for a
// sorted-tuple comparison against the oracle (tests/test_gpu_multirank.py).
//   pjoin_main --gpus 8 [--log2-build 20] [--log2-probe 22] [--table lp|chain] [--chunk-factor 1] [--hit 1] [--steps 3]
//              [--sub-batches 4] [--pipeline 0|1] [--dump prefix]
#include <chrono>
#include <cstdlib>
#include <cstring>

#include "simd_compaction.hpp"

using namespace simd_compaction;

static uint64_t mm64(uint64_t x) {  // hash_functions.h:8-16
  x ^= x >> 32;
  x *= 0xd6e8feb86659fd93ULL;
  x ^= x >> 32;
  x *= 0xd6e8feb86659fd93ULL;
  x ^= x >> 32;
  return x;
}

int main(int argc, char **argv) {
  int gpus = 1, log2_build = 20, log2_probe = 22, steps = 3, n_sub = 4, pipeline = 0;
  size_t cf = 1, hit = 1;
  std::string table = "lp", dump;
  for (int i = 1; i + 1 < argc; i += 2) {
    std::string a = argv[i], v = argv[i + 1];
    if (a == "--gpus") gpus = std::stoi(v);
    else if (a == "--log2-build") log2_build = std::stoi(v);
    else if (a == "--log2-probe") log2_probe = std::stoi(v);
    else if (a == "--table") table = v;
    else if (a == "--chunk-factor") cf = std::stoull(v);
    else if (a == "--hit") hit = std::stoull(v);
    else if (a == "--steps") steps = std::stoi(v);
    else if (a == "--sub-batches") n_sub = std::stoi(v);
    else if (a == "--dump") dump = v;
    else if (a == "--pipeline") pipeline = std::stoi(v);
  }
  const size_t n_build = (size_t) 1 << log2_build, n_probe = (size_t) 1 << log2_probe;
  const int kind = table == "lp" ? CC_HT_LP : CC_HT_CHAIN;
  LocalComm comm(gpus);
  const int rank = comm.Fork();  // BEFORE the first CUDA call: a forked child must not inherit a CUDA context
  int rc = 0;
  try {
    Check(cc_device_init(rank));
    const size_t nb = n_build / gpus, np = n_probe / gpus;
    DeviceArray<Attribute> build(nb ? nb : 1, false), probe(np ? np : 1, false);
    Check(cc_gen_build_keys_range(build.data(), rank * nb, nb, n_build, cf, nullptr));
    const uint64_t seed = 2, mask = n_build * hit - 1;
    Check(cc_gen_keys_counter(probe.data(), np, seed, rank * np, mask, nullptr));
    Check(cc_stream_sync(nullptr));
    auto t_build = std::chrono::steady_clock::now();
    cc_comm c = comm.Comm();
    PartitionedJoin join(c, kind, build.data(), nb, np, n_sub);
    Check(cc_stream_sync(nullptr));
    const double build_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_build).count();
    const size_t cap = (2 * np + (1 << 16)) * cf;
    DeviceArray<Attribute> out_key(cap, false), out_payload(cap, false);
    DeviceArray<uint64_t> res(sizeof(cc_probe_result) / sizeof(uint64_t));
    double best = 1e30;
    auto *result = reinterpret_cast<cc_probe_result *>(res.data());
    if (pipeline) join.ProbeBegin(probe.data(), np);  // prologue: batch 0 is on its way
    for (int s = 0; s < steps + 1; ++s) {             // one warm-up step
      comm.Barrier();
      auto t0 = std::chrono::steady_clock::now();
      if (pipeline) {  // a step = the exchange of batch t + 1 is enqueued, then batch t (which had a whole step to land) is probed
        join.ProbeBegin(probe.data(), np);
        join.ProbeEnd(out_key.data(), out_payload.data(), cap, result);
      } else {
        join.Probe(probe.data(), np, out_key.data(), out_payload.data(), cap, result);
      }
      Check(cc_stream_sync(nullptr));
      comm.Barrier();  // a step ends when its slowest rank is done
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (s > 0 && dt < best) best = dt;
    }
    if (pipeline) {  // epilogue: the batch still in flight (same keys, same result)
      join.ProbeEnd(out_key.data(), out_payload.data(), cap, result);
      Check(cc_stream_sync(nullptr));
    }
    auto r = res.ToHost();
    // ---- host recomputation of this rank's SENT keys: count and checksums of the whole join are sums over the probe keys
    const size_t num_unique = n_build / cf + (n_build % cf != 0), step = n_build / num_unique;
    uint64_t want[3] = {0, 0, 0};  // matches, key sum, payload sum contributed by the keys this rank SENT
    for (size_t i = 0; i < np; ++i) {
      const uint64_t k = mm64(seed + rank * np + i) & mask;
      if (k % step == 0 && k / step < num_unique) {
        // rows of key k in the build column: the reference emits `cf` copies of every key but the last may be cut at n_build
        const size_t first_row = (k / step) * cf, copies = std::min(cf, n_build - first_row);
        want[0] += copies;
        want[1] += k * copies;
        want[2] += k * copies;
      }
    }
    // ---- owner property of the rows this rank HOLDS
    // (all rows when they are dumped, else the first 4 Mi rows: a full-size run holds 2^30 rows per rank)
    const size_t m = std::min<size_t>(std::min<size_t>(r[0], cap), dump.empty() ? (size_t) 1 << 22 : SIZE_MAX);
    auto hk = out_key.ToHost(m), hp = out_payload.ToHost(m);
    int log2p = 0;
    while ((1 << log2p) < gpus) ++log2p;
    uint64_t owner_ok = 1;
    for (size_t i = 0; i < m; ++i) {
      const uint64_t owner = log2p ? mm64((uint64_t) hk[i]) >> (64 - log2p) : 0;
      if (owner != (uint64_t) rank || hk[i] != hp[i]) owner_ok = 0;
    }
    if (!dump.empty()) {
      std::string path = dump + "." + std::to_string(rank) + ".bin";
      FILE *f = fopen(path.c_str(), "wb");
      if (!f) throw std::runtime_error("cannot write " + path);
      for (size_t i = 0; i < m; ++i) {
        fwrite(&hk[i], 8, 1, f);
        fwrite(&hp[i], 8, 1, f);
      }
      fclose(f);
    }
    // ---- reduce over the ranks through the communicator
    uint64_t mine[8] = {r[0], r[1], r[2], r[3], want[0], want[1], want[2], owner_ok};
    std::vector<uint64_t> all(8 * (size_t) gpus);
    comm.AllGather(mine, all.data(), sizeof(mine));
    double times[1] = {best};
    std::vector<double> all_t(gpus);
    comm.AllGather(times, all_t.data(), sizeof(times));
    if (rank == 0) {
      uint64_t got[3] = {0, 0, 0}, exp[3] = {0, 0, 0}, overflow = 0, owned = 1;
      for (int q = 0; q < gpus; ++q) {
        for (int t = 0; t < 3; ++t) got[t] += all[8 * q + t], exp[t] += all[8 * q + 4 + t];
        overflow |= all[8 * q + 3];
        owned &= all[8 * q + 7];
      }
      double worst = 0;
      for (double t : all_t) worst = std::max(worst, t);
      const bool ok = got[0] == exp[0] && got[1] == exp[1] && got[2] == exp[2] && overflow == 0 && owned == 1;
      cc_ht_info info = join.TableInfo();
      printf("{\"gpus\": %d, \"table\": \"%s\", \"n_build\": %zu, \"n_probe\": %zu, \"chunk_factor\": %zu, \"hit\": %zu, \"sub_batches\": %d, "
             "\"pipelined\": %d, \"n_matches\": %llu, \"key_sum\": %llu, \"payload_sum\": %llu, \"expected_matches\": %llu, \"overflow\": %llu, \"owner_property\": %s, "
             "\"checks_ok\": %s, \"ms_per_step\": %.3f, \"probe_tuples_per_sec\": %.4g, \"build_seconds\": %.3f, \"local_table_slots\": %zu}\n",
             gpus, table.c_str(), n_build, n_probe, cf, hit, n_sub, pipeline, (unsigned long long) got[0], (unsigned long long) got[1],
             (unsigned long long) got[2], (unsigned long long) exp[0], (unsigned long long) overflow, owned ? "true" : "false", ok ? "true" : "false",
             worst * 1e3, n_probe / worst, build_s, (size_t) info.n_slots);
      if (!ok) rc = 3;
    }
  } catch (const std::exception &e) {
    fprintf(stderr, "pjoin_main[rank %d]: %s\n", rank, e.what());
    comm.Fail();
    rc = 2;
  }
  if (rank != 0) _exit(rc);
  if (!comm.Join() && rc == 0) rc = 4;
  return rc;
}
