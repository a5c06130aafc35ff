// CPU check of LocalComm (host/simd_compaction.hpp): the fork + shared-memory control plane of the C++ partitioned-join driver.
// Forks `world` ranks, runs all-gathers of different record sizes and barriers through the cc_comm callbacks the library would
// call, checks every rank sees every rank's record; a second run lets one rank fail and checks that the others are released.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "simd_compaction.hpp"

using namespace simd_compaction;

int main(int argc, char **argv) {
  const int world = argc > 1 ? atoi(argv[1]) : 4;
  const bool inject_failure = argc > 2 && atoi(argv[2]) != 0;
  LocalComm comm(world);
  const int rank = comm.Fork();
  cc_comm c = comm.Comm();
  int rc = 0;
  try {
    if (c.rank != rank || c.world != world) throw std::runtime_error("cc_comm rank / world wrong");
    for (size_t bytes : {(size_t) 8, (size_t) 64, (size_t) 24}) {
      unsigned char mine[64], all[64 * 16];
      for (size_t i = 0; i < bytes; ++i) mine[i] = (unsigned char) (rank * 37 + i);
      if (c.allgather(c.user, mine, all, bytes) != 0) throw std::runtime_error("allgather callback failed");
      for (int r = 0; r < world; ++r)
        for (size_t i = 0; i < bytes; ++i)
          if (all[r * bytes + i] != (unsigned char) (r * 37 + i)) throw std::runtime_error("allgather delivered wrong bytes");
      if (c.barrier(c.user) != 0) throw std::runtime_error("barrier callback failed");
    }
    if (inject_failure) {
      if (rank == 1) throw std::runtime_error("injected failure");
      // the others enter a barrier the failed rank never reaches: they must be released with an error, not hang
      if (c.barrier(c.user) == 0) throw std::runtime_error("barrier succeeded although a rank failed");
      rc = 7;  // expected outcome of the surviving ranks
    }
  } catch (const std::exception &e) {
    if (!(inject_failure && rank == 1)) fprintf(stderr, "localcomm_check[rank %d]: %s\n", rank, e.what());
    comm.Fail();
    rc = 2;
  }
  if (rank != 0) _exit(rc);
  const bool children_ok = comm.Join();
  if (inject_failure) return (rc == 7 && !children_ok) ? 0 : 1;  // rank 0 survived with the expected error, some child exited non-zero
  return (rc == 0 && children_ok) ? 0 : 1;
}
