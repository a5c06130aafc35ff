"""Multi-GPU hash-partitioned join (new functionality) -- host layer.

One process per GPU (torch.distributed, NCCL over NVLink 5 / NVSwitch).  An equi-join
shards naturally: rows with equal keys must meet, nothing else is shared.  So both
sides are hash-partitioned with  p = murmurhash64(key) >> (64 - log2 P)  (high hash
bits, independent of the low bits that address the owner's table), exchanged ONCE with
a variable-size all-to-all, and every GPU then builds / probes its own table with the
single-GPU kernels.  Results stay sharded; only counts and checksums are reduced.

  * partition kernels : csrc/partition.cu (cc_partition_count / cc_partition_scatter)
  * exchange          : all_to_all_single on the segment buffers (this file) -- the
                        only collective on the data path
  * small build sides : broadcast (all_gather) instead, zero probe-side traffic; this
                        is also the right plan for join CHAINS on different key columns

The exchange helpers are device-agnostic (they are exercised on CPU tensors with the
gloo backend in tests/test_distributed_cpu.py); partitioning and probing are CUDA only.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def log2_exact(p: int) -> int:
    l = p.bit_length() - 1
    if p <= 0 or (1 << l) != p:
        raise ValueError(f"world size {p} must be a power of two for hash partitioning")
    return l


def exchange_counts(send_counts: torch.Tensor, group=None) -> torch.Tensor:
    """send_counts[p] = rows this rank sends to rank p  ->  recv_counts[p] = rows rank p sends here."""
    recv = torch.empty_like(send_counts)
    dist.all_to_all_single(recv, send_counts, group=group)
    return recv


def exchange_rows(send: torch.Tensor, send_counts: List[int], recv_counts: List[int], out: Optional[torch.Tensor] = None,
                  group=None) -> torch.Tensor:
    """Variable-size all-to-all of a partition-grouped 1-D buffer (segments in rank order)."""
    total = int(sum(recv_counts))
    if out is None:
        out = torch.empty(max(total, 1), dtype=send.dtype, device=send.device)
    recv = out[:total]
    dist.all_to_all_single(recv, send[: int(sum(send_counts))], output_split_sizes=[int(c) for c in recv_counts],
                           input_split_sizes=[int(c) for c in send_counts], group=group)
    return recv


def choose_plan(n_build_total: int, n_probe_total: int, world: int) -> str:
    """SURVEY 8e rule: broadcast the build side when replicating it costs much less traffic than
    moving the probe side (n_build * 16 B * P  <<  n_probe * 8 B), else hash-partition both sides."""
    return "broadcast" if n_build_total * 16 * world * 4 < n_probe_total * 8 else "partition"


class PartitionedJoin:
    """Build once, probe many times.  `pkg` is the product package (passed in to avoid a circular import)."""

    def __init__(self, pkg, kind: int, local_build_keys: torch.Tensor, group=None, plan: str = "partition"):
        self.pkg = pkg
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.log2p = log2_exact(self.world)
        self.plan = plan
        T = pkg.LPHashTable if kind == pkg.CC_HT_LP else pkg.HashTable
        if plan == "broadcast":
            n_local = torch.tensor([local_build_keys.numel()], dtype=torch.int64, device=local_build_keys.device)
            sizes = [torch.empty_like(n_local) for _ in range(self.world)]
            dist.all_gather(sizes, n_local, group=group)
            sizes = [int(s.item()) for s in sizes]
            parts = [torch.empty(max(s, 1), dtype=torch.int64, device=local_build_keys.device)[:s] for s in sizes]
            # all_gather needs equal sizes: pad to the maximum
            mx = max(sizes + [1])
            padded = torch.zeros(mx, dtype=torch.int64, device=local_build_keys.device)
            padded[: local_build_keys.numel()] = local_build_keys
            gathered = [torch.empty_like(padded) for _ in range(self.world)]
            dist.all_gather(gathered, padded, group=group)
            keys = torch.cat([g[:s] for g, s in zip(gathered, sizes)])
            del parts
        else:
            keys = self.shuffle(local_build_keys)
        self.n_build_local = keys.numel()
        self.table = T(keys=keys)

    def shuffle(self, keys: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Hash-partition `keys` and exchange: returns the rows this rank owns."""
        part, counts, _ = self.pkg.partition_keys(keys, self.log2p)
        send_counts = torch.from_numpy(counts).to(keys.device)
        recv_counts = exchange_counts(send_counts, self.group)
        return exchange_rows(part, counts.tolist(), recv_counts.cpu().tolist(), out=out, group=self.group)

    def probe(self, local_probe_keys: torch.Tensor, **kw) -> dict:
        """One probe pass: (partition + all-to-all unless broadcast plan) + local batch probe."""
        keys = local_probe_keys if self.plan == "broadcast" else self.shuffle(local_probe_keys)
        kw.setdefault("capacity", max(1, keys.numel()))
        return self.table.probe_batch(keys, **kw)


def reduce_result(n_matches: int, key_sum: int, payload_sum: int, device, group=None) -> Tuple[int, int, int]:
    """Sum of the per-rank counts / wrapping checksums (all-reduce of three int64)."""
    def wrap(v):
        v &= (1 << 64) - 1
        return v - (1 << 64) if v >= (1 << 63) else v
    t = torch.tensor([wrap(n_matches), wrap(key_sum), wrap(payload_sum)], dtype=torch.int64, device=device)
    dist.all_reduce(t, group=group)
    return tuple(int(x) & ((1 << 64) - 1) for x in t.cpu().tolist())
