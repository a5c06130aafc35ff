// Host-side check of the addressing logic the partitioned join relies on (compiled with nvcc, runs WITHOUT a GPU):
// PartFn (plain owner / slice functions and the fused owner x slice function) and SegIn::region (walk order -> arena region).
// Prints one line per probe: "<key> <owner P=8> <slice S=64 of a 2^20-slot table> <fused id> <plain owner id> <plain slice id>"
// for tests/test_partition_host.py to compare against an independent numpy computation, then the region mapping table.
#include <cstdio>

#include "partition.cuh"

using namespace ccb;

int main() {
  const int log2p = 3, log2_slots = 20, log2s = 6;
  const uint64_t mask = ((uint64_t) 1 << log2_slots) - 1;
  const PartFn fused = PartFn::owner_and_slice(log2p, mask, log2_slots, log2s);
  const PartFn owner = PartFn::high_bits(log2p);
  const PartFn slice = PartFn::slot_bits(mask, log2_slots, log2s);
  const PartFn one_rank = PartFn::owner_and_slice(0, mask, log2_slots, log2s);  // a single rank: slice only
  if (fused.parts() != 512 || owner.parts() != 8 || slice.parts() != 64 || one_rank.parts() != 64) return 2;
  uint64_t x = 88172645463325252ULL;
  for (int i = 0; i < 2000; ++i) {
    x = x * 6364136223846793005ULL + 1442695040888963407ULL;
    const uint64_t key = i < 8 ? (uint64_t) i : (i < 16 ? ~(uint64_t) i : x);
    printf("%llu %u %u %u %u\n", (unsigned long long) key, fused(key), owner(key), slice(key), one_rank(key));
  }
  // arena walk: pieces = 3, senders = 4, slices allocated = 5 -> inner = 12, outer_stride = 5
  SegIn seg;
  seg.inner = 12;
  seg.outer_stride = 5;
  printf("regions");
  for (uint32_t p = 0; p < 60; ++p) printf(" %u", seg.region(p));
  printf("\n");
  SegIn plain;
  if (plain.region(7) != 7) return 3;
  return 0;
}
