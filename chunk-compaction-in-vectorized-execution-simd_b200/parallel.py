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


def table_slots(lp: bool, n_total: int, n_local: int, world: int) -> int:
    """Slot / bucket count of one rank's table in the partitioned join: the reference's sizing rule on the GLOBAL key count
    (LP: pow2 >= 4 n, linear_probing_ht.cpp:5-6; chain: pow2 >= 2 n, chaining_ht.cpp:5-6) divided by the (power-of-two) number
    of ranks -- doubled while a skewed partition would push an LP table beyond half full."""
    per_key = 4 if lp else 2
    total = 1
    while total < per_key * n_total:
        total <<= 1
    slots = max(1, total // world)
    while lp and slots < 2 * n_local:
        slots <<= 1
    return slots


class PeerExchange:
    """Fused scatter + exchange over peer memory (NVLink 5 / NVSwitch).

    Every rank owns `n_buffers` receive buffers (plain cudaMalloc via cc_malloc) and maps the buffers
    of all other ranks into its address space with CUDA IPC.  A shuffle is then
        histogram -> all-gather of the P counts (tiny) -> ONE scatter kernel that stores every row
        straight into its owner's receive buffer -> stream-ordered all-reduce as the barrier
    i.e. the partition kernel IS the all-to-all: no staging copy, no second pass over the data, and
    the NVLink transfer overlaps the kernel's own reads tile by tile.  Buffers alternate between
    consecutive shuffles so that a fast rank can already scatter step i+1 while a slow one still
    probes step i."""

    def __init__(self, pkg, capacity_rows: int, group=None, n_buffers: int = 3, peer_blocks: int = 0):
        import ctypes as C

        self.pkg, self.group = pkg, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.log2p = log2_exact(self.world)
        self.capacity = int(capacity_rows)
        self.n_buffers = n_buffers
        pkg._lib.check(pkg.lib().cc_partition_set_peer_blocks(int(peer_blocks)))
        self.step = 0
        lib = pkg.lib()
        self.local, self.peers, self._opened = [], [], []
        for _ in range(n_buffers):
            ptr = C.c_void_p()
            pkg._lib.check(lib.cc_malloc(C.byref(ptr), self.capacity * 8))
            handle = (C.c_ubyte * 64)()
            pkg._lib.check(lib.cc_ipc_export(ptr, handle))
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            ptrs = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    ptrs.append(ptr.value)
                else:
                    hb = (C.c_ubyte * 64).from_buffer_copy(h)
                    q = C.c_void_p()
                    pkg._lib.check(lib.cc_ipc_open(hb, C.byref(q)))
                    ptrs.append(q.value)
                    self._opened.append(q.value)
            self.local.append(ptr.value)
            self.peers.append((C.c_void_p * self.world)(*ptrs))
        dev = torch.device("cuda", torch.cuda.current_device())
        self._counts = torch.zeros(self.world, dtype=torch.int64, device=dev)
        self._matrix = torch.zeros(self.world * self.world, dtype=torch.int64, device=dev)
        self._cursors = torch.zeros(self.world, dtype=torch.int64, device=dev)
        self._token = torch.zeros(1, dtype=torch.int32, device=dev)
        dist.barrier(group=group)

    def shuffle(self, keys: torch.Tensor, reader_done: Optional[torch.cuda.Event] = None) -> torch.Tensor:
        """Hash-partition `keys` by owner and deliver them: returns this rank's rows (a view of the current
        receive buffer; it is overwritten by the n_buffers-th shuffle from now).
        reader_done: event after which this rank no longer reads the buffer that the NEXT shuffle will
        fill on the peers' behalf -- the barrier waits for it, so no peer can overwrite rows still in use."""
        pkg, lib = self.pkg, self.pkg.lib()
        stream = torch.cuda.current_stream().cuda_stream
        n = keys.numel()
        b = self.step % self.n_buffers
        self.step += 1
        pkg._lib.check(lib.cc_partition_count(keys.data_ptr() if n else None, n, self.log2p, self._counts.data_ptr(), stream))
        dist.all_gather_into_tensor(self._matrix, self._counts, group=self.group)
        m = self._matrix.view(self.world, self.world)  # m[sender][owner]
        base = m[: self.rank].sum(dim=0).contiguous()  # rows of earlier senders in each owner's buffer
        # EVERY rank holds the full P x P matrix, so every rank checks EVERY receive column: if any owner's buffer would
        # overflow, all ranks raise together -- nobody stores past the end of a peer's buffer over NVLink and nobody is
        # left alone in the barrier below
        recv_all = m.sum(dim=0).cpu()
        n_recv = int(recv_all[self.rank])
        if int(recv_all.max()) > self.capacity:
            worst = int(recv_all.argmax())
            raise RuntimeError(f"receive buffer of rank {worst} too small: {int(recv_all[worst])} rows > capacity {self.capacity}")
        pkg._lib.check(lib.cc_partition_scatter_peers(keys.data_ptr() if n else None, n, self.log2p, base.data_ptr(),
                                                      self._cursors.data_ptr(), self.peers[b], stream))
        if reader_done is not None:
            torch.cuda.current_stream().wait_event(reader_done)
        dist.all_reduce(self._token, group=self.group)  # stream-ordered barrier: every rank's stores have landed
        return pkg._wrap_ptr(self.local[b], max(n_recv, 1), torch.int64)[:n_recv]

    def close(self) -> None:
        lib = self.pkg.lib()
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for q in self._opened:
            lib.cc_ipc_close(q)
        for p in self.local:
            lib.cc_free(p)
        self._opened, self.local = [], []


class CopyExchange:
    """Exchange by copy engine: SM-free, host-sync-free, overlapped with the probe (NVLink 5 / NVSwitch).

    A shuffle is split into a START and a FINISH half:
      start  : ONE single-pass partition kernel (cc_partition_single) groups the keys by owner into fixed regions of a local
               send buffer -- no histogram pass, the region fill counts stay on the device.  Then, on a COPY stream, region p is
               copied as one block into slot [this rank] of owner p's receive buffer (CUDA IPC mapping, cudaMemcpyAsync: the
               copy engines move it over NVLink without occupying a single SM), and the counts are all-gathered.
      finish : wait for the own copies, then a stream-ordered all-reduce as the barrier ("every rank's copies have landed").
    The receive buffer is a SEGMENTED column (one segment per sender, counts on the device) that cc_probe_batch_segmented
    consumes as is -- the host never learns a count, so nothing ever blocks the launch queue.  PartitionedJoin.probe_pipelined
    starts sub-batch b + 1 before it finishes sub-batch b: the copies of b + 1 run underneath the probe of b, and per key the
    SMs only do the owner partition, the slice partition and the probe.  (The fused peer-scatter kernel of PeerExchange moves
    the same bytes but needs every SM while it runs, so it cannot overlap a probe that also wants every SM.)
    Cost: the regions are copied with their slack (3 % + two tiles more NVLink bytes; a hash partition of n rows into P regions
    fills each to n / P +- sqrt(n / P), so even 1 % is dozens of standard deviations).  Heavily skewed keys overrun a region: this is
    detected on the device and reported by check_overflow() -- use exchange="p2p" or "nccl" for such inputs."""

    TILE = 4096  # region capacities are multiples of the partition kernel's tile

    def __init__(self, pkg, max_rows: int, group=None, n_buffers: int = 3, copy_streams: int = 4):
        import ctypes as C

        self.pkg, self.group = pkg, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.log2p = log2_exact(self.world)
        self.n_buffers = n_buffers
        self.step = 0
        self.max_rows = int(max_rows)
        per = (self.max_rows + self.world - 1) // self.world
        self.cap = ((per + per // 32 + 2 * self.TILE + self.TILE - 1) // self.TILE) * self.TILE  # rows per (sender, owner) region
        self.rows = self.cap * self.world
        lib = pkg.lib()
        dev = torch.device("cuda", torch.cuda.current_device())
        self.local, self.peers, self._opened = [], [], []
        for _ in range(n_buffers):
            ptr = C.c_void_p()
            pkg._lib.check(lib.cc_malloc(C.byref(ptr), self.rows * 8))
            handle = (C.c_ubyte * 64)()
            pkg._lib.check(lib.cc_ipc_export(ptr, handle))
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            ptrs = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    ptrs.append(ptr.value)
                else:
                    hb = (C.c_ubyte * 64).from_buffer_copy(h)
                    q = C.c_void_p()
                    pkg._lib.check(lib.cc_ipc_open(hb, C.byref(q)))
                    ptrs.append(q.value)
                    self._opened.append(q.value)
            self.local.append(ptr.value)
            self.peers.append(ptrs)
        self.send = [torch.empty(self.rows, dtype=torch.int64, device=dev) for _ in range(n_buffers)]
        self.counts = [torch.zeros(self.world, dtype=torch.int64, device=dev) for _ in range(n_buffers)]
        self.matrix = [torch.zeros(self.world * self.world, dtype=torch.int64, device=dev) for _ in range(n_buffers)]
        self.mine = [torch.zeros(self.world, dtype=torch.int64, device=dev) for _ in range(n_buffers)]
        self.overflow = torch.zeros(n_buffers, dtype=torch.int32, device=dev)
        self._token = torch.zeros(1, dtype=torch.int32, device=dev)
        # The block copies of one shuffle go out on several copy streams: one stream drives one copy engine at a time, and a
        # single engine does not fill an NVLink 5 port (round 1: 450 GB/s effective with one stream).  Destinations are dealt
        # round-robin, so that concurrent copies of a rank always target different peers.
        self._copy_streams = [torch.cuda.Stream() for _ in range(max(1, min(int(copy_streams), self.world - 1)))]
        self._copied = [None] * n_buffers
        dist.barrier(group=group)

    def start(self, keys: torch.Tensor) -> int:
        """First half of shuffle k (returned): owner partition on the current stream, block copies on the copy stream."""
        pkg, lib = self.pkg, self.pkg.lib()
        n = keys.numel()
        if n > self.max_rows:
            raise ValueError(f"{n} keys exceed the {self.max_rows} rows per shuffle this exchange was sized for")
        k = self.step
        self.step += 1
        b = k % self.n_buffers
        main = torch.cuda.current_stream()
        # the rows this rank keeps go straight into slot [rank] of its own receive buffer (same region offset rank * cap)
        pkg.partition_single(keys, self.log2p, self.cap, out=self.send[b], counts=self.counts[b], overflow=self.overflow[b:b + 1],
                             self_part=self.rank if self.world <= 16 else -1, self_out_ptr=self.local[b])
        parted = torch.cuda.Event()
        parted.record(main)
        dist.all_gather_into_tensor(self.matrix[b], self.counts[b], group=self.group)  # matrix[sender * P + owner]
        self.mine[b].copy_(self.matrix[b].view(self.world, self.world)[:, self.rank])  # rows every sender delivers to this rank
        for cs in self._copy_streams:
            cs.wait_event(parted)
        src = self.send[b].data_ptr()
        block = self.cap * 8
        for j, i in enumerate(range(0 if self.world > 16 else 1, self.world)):
            p = (self.rank + i) % self.world  # stagger the destinations so that the ranks do not all hit the same peer at once
            cs = self._copy_streams[j % len(self._copy_streams)]
            pkg._lib.check(lib.cc_memcpy_d2d(self.peers[b][p] + self.rank * block, src + p * block, block, cs.cuda_stream))
        done = []
        for cs in self._copy_streams:
            e = torch.cuda.Event()
            e.record(cs)
            done.append(e)
        self._copied[b] = done
        return k

    def finish(self, k: int):
        """Second half of shuffle k: returns (segmented receive column, segment capacity, device counts[P])."""
        b = k % self.n_buffers
        for e in self._copied[b]:
            torch.cuda.current_stream().wait_event(e)
        dist.all_reduce(self._token, group=self.group)  # stream-ordered barrier: every rank's copies have landed
        return self.pkg._wrap_ptr(self.local[b], self.rows, torch.int64), self.cap, self.mine[b]

    def check_overflow(self) -> None:
        """Raises ON EVERY RANK if any shuffle of any rank since the last check overran a region (synchronises; collective).
        The per-buffer flags are sticky (cc_partition_single only ORs into them), so an overrun in ANY shuffle since the
        last check is seen, not just in the last n_buffers ones."""
        bad_t = (self.overflow != 0).sum().to(torch.int32).reshape(1)
        dist.all_reduce(bad_t, op=dist.ReduceOp.MAX, group=self.group)
        bad = int(bad_t.item())
        self.overflow.zero_()
        if bad:
            raise RuntimeError("copy-engine exchange: a partition region overran (heavily skewed keys); use exchange='p2p' or 'nccl'")

    def close(self) -> None:
        lib = self.pkg.lib()
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for q in self._opened:
            lib.cc_ipc_close(q)
        for p in self.local:
            lib.cc_free(p)
        self._opened, self.local = [], []


class PartitionedJoin:
    """Build once, probe many times.  `pkg` is the product package (passed in to avoid a circular import)."""

    def __init__(self, pkg, kind: int, local_build_keys: torch.Tensor, group=None, plan: str = "partition",
                 exchange: str = "nccl", capacity_rows: int = 0, peer_blocks: int = 0, ce_probe: str = "auto", copy_streams: int = 4):
        """plan: "partition" (hash-partition both sides) or "broadcast" (replicate the build side).
        exchange: "nccl" (scatter locally, then all_to_all_single), "p2p" (PeerExchange: the scatter kernel
        writes into the owners' buffers over NVLink) or "ce" (CopyExchange: single-pass partition + copy-engine block
        copies, overlapped with the probe by probe_pipelined; the build side then travels by all_to_all_single);
        capacity_rows sizes the p2p receive buffers / the largest "ce" sub-batch.
        ce_probe: how the "ce" pipeline probes -- "stream": ONE incremental probe per call (the table is streamed from HBM
        once; best while the NVLink copies hide under the partition kernels, i.e. few GPUs), "batch": every sub-batch is
        probed as soon as it has landed (the table is streamed once per sub-batch, but the probes run underneath the copy
        chain, which is what bounds a step on many GPUs), "auto": stream up to 2 GPUs, batch beyond."""
        if ce_probe not in ("auto", "stream", "batch"):
            raise ValueError(f"ce_probe must be auto, stream or batch, not {ce_probe!r}")
        self.pkg = pkg
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.log2p = log2_exact(self.world)
        self.plan = plan
        self.ce_probe = ce_probe if ce_probe != "auto" else ("stream" if self.world <= 2 else "batch")
        import time

        def lap(name, t0):  # wall-clock phases of the (untimed) build, reported by bench.py as build_phases
            torch.cuda.synchronize()
            self.build_phases[name] = time.perf_counter() - t0
            return time.perf_counter()

        self.build_phases = {}
        t0 = time.perf_counter()
        self.peer = PeerExchange(pkg, capacity_rows, group, peer_blocks=peer_blocks) if (exchange == "p2p" and plan == "partition") else None
        self.copier = CopyExchange(pkg, capacity_rows, group, copy_streams=copy_streams) if (exchange == "ce" and plan == "partition") else None
        t0 = lap("exchange_setup_s", t0)
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
            n_slots = 0  # every rank holds the whole build side: the reference's sizing rule applies as is
        else:
            keys = self.shuffle(local_build_keys)
            # Size the local table from the GLOBAL key count: the reference's rule (pow2 >= 4n for LP, linear_probing_ht.cpp:5-6;
            # pow2 >= 2n for chains, chaining_ht.cpp:5-6) applied to the whole build side, divided by the number of ranks.
            # Applied to the LOCAL count it doubles the table whenever a hash partition lands a hair above n_total / P
            # (round 1: 134 224 497 keys > 2^27 by 0.005 % -> 2^30 slots = 8 GiB at load factor 0.125 on every rank).
            n_total = torch.tensor([local_build_keys.numel()], dtype=torch.int64, device=local_build_keys.device)
            dist.all_reduce(n_total, group=group)
            n_slots = table_slots(kind == pkg.CC_HT_LP, int(n_total.item()), keys.numel(), self.world)
        t0 = lap("build_exchange_s", t0)
        self.n_build_local = keys.numel()
        self.table = T(keys=keys, n_slots=n_slots)
        lap("table_build_s", t0)

    def shuffle(self, keys: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Hash-partition `keys` and exchange: returns the rows this rank owns."""
        if self.peer is not None:
            return self.peer.shuffle(keys)
        part, counts, _ = self.pkg.partition_keys(keys, self.log2p)
        send_counts = torch.from_numpy(counts).to(keys.device)
        recv_counts = exchange_counts(send_counts, self.group)
        return exchange_rows(part, counts.tolist(), recv_counts.cpu().tolist(), out=out, group=self.group)

    def probe_pipelined(self, local_probe_keys: torch.Tensor, n_sub: int, out_key: torch.Tensor, out_payload: torch.Tensor,
                        results: torch.Tensor) -> None:
        """One probe pass in n_sub sub-batches: the shuffle of sub-batch b+1 (NVLink-bound) runs on a side
        stream while sub-batch b is probed on the current stream.  results: int64[n_sub, 4] device tensor
        (one cc_probe_result per sub-batch); outputs are written into n_sub equal slices of out_key/out_payload.
        Buffer safety: with 3 receive buffers a peer may only fill buffer k % 3 again at shuffle k + 3, i.e. after
        it passed barrier k + 2, which this rank enters only once its probe of sub-batch k is complete."""
        if self.copier is not None:
            return self._probe_pipelined_ce(local_probe_keys, n_sub, out_key, out_payload, results)
        assert self.peer is not None, "pipelined probing needs the peer-memory or the copy-engine exchange"
        main = torch.cuda.current_stream()
        if not hasattr(self, "_xstream"):
            self._xstream = torch.cuda.Stream()
            self._probe_done = {}
        xs = self._xstream
        xs.wait_stream(main)
        cap = out_key.numel() // n_sub
        for b, chunk in enumerate(local_probe_keys.chunk(n_sub)):
            k = self.peer.step  # global shuffle index
            with torch.cuda.stream(xs):
                recv = self.peer.shuffle(chunk, reader_done=self._probe_done.pop(k - 2, None))
                ready = torch.cuda.Event()
                ready.record(xs)
            main.wait_event(ready)
            self.table.probe_batch(recv, capacity=cap, out_key=out_key[b * cap:(b + 1) * cap], out_payload=out_payload[b * cap:(b + 1) * cap],
                                   result=results[b], sync=False)
            done = torch.cuda.Event()
            done.record(main)
            self._probe_done[k] = done

    def _probe_pipelined_ce(self, local_probe_keys, n_sub, out_key, out_payload, results, ready=None, landed=None) -> None:
        """Copy-engine variant: everything that needs SMs is enqueued on the CURRENT stream in the order
        P(0) P(1) B(0) S(0) P(2) B(1) S(1) ... B(n-1) S(n-1) PROBE  (P = owner partition of a sub-batch, B = barrier, S = scatter
        of the received sub-batch into the table-slice regions of ONE incremental probe, cc_probe_stream_*) while the block
        copies C(b) run on the copy stream underneath S(b - 1) / P(b + 1).  The table is streamed once per call, not once per
        sub-batch.  Only results[0] is written (one dense output over all of out_key / out_payload).
        (ce_probe == "batch" instead probes every sub-batch on arrival, see __init__.)
        Buffer safety with 3 rotating buffers: P(k + 3) is enqueued behind B(k + 1), which completes only when every rank has
        entered it, i.e. after its S(k)."""
        import os
        cx = self.copier
        chunks = list(local_probe_keys.chunk(n_sub))
        main = torch.cuda.current_stream()

        def start(b):  # ready[b] (optional): event after which chunk b holds its keys (the H2D copy of probe_host)
            if ready is not None:
                main.wait_event(ready[b])
            return cx.start(chunks[b])

        if self.ce_probe == "batch":
            # P(0) P(1) B(0) L(0) P(2) B(1) L(1) ...  with L = slice partition + probe of ONE received sub-batch: the block
            # copies C(b + 1) run underneath L(b).  results[b] / slice b of the output columns belong to sub-batch b.
            cap = out_key.numel() // n_sub
            pending = [start(0)]
            for b in range(len(chunks)):
                if b + 1 < len(chunks):
                    pending.append(start(b + 1))
                recv, seg_cap, counts = cx.finish(pending[b])
                self.table.probe_batch_segmented(recv, self.world, seg_cap, counts, capacity=cap, out_key=out_key[b * cap:(b + 1) * cap],
                                                 out_payload=out_payload[b * cap:(b + 1) * cap], result=results[b], sync=False)
                if landed is not None:
                    landed(b)  # sub-batch b's rows and counter are final once the work enqueued so far has run
            return
        trace = [] if os.environ.get("CCB_CE_TRACE") else None  # evidence switch: CUDA-event timeline of one call (synchronises)

        def mark(name):
            if trace is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                trace.append((name, e))

        mark("begin")
        n_local = local_probe_keys.numel()
        stream = self.table.probe_stream(n_local + n_local // 16 + (1 << 16), capacity=out_key.numel(), out_key=out_key, out_payload=out_payload,
                                         result=results[0])
        pending = [start(0)]
        mark("P0")
        for b in range(len(chunks)):
            if b + 1 < len(chunks):
                pending.append(start(b + 1))
                mark(f"P{b + 1}")
            recv, seg_cap, counts = cx.finish(pending[b])
            mark(f"B{b}")
            stream.add_segmented(recv, self.world, seg_cap, counts)
            mark(f"S{b}")
        stream.finish(sync=False)
        mark("PROBE")
        if trace is not None and self.rank == 0:
            torch.cuda.synchronize()
            print("ce timeline (ms since begin): " + "  ".join(f"{n}={trace[0][1].elapsed_time(e):.2f}" for n, e in trace[1:]), flush=True)

    def probe_host(self, h_keys: torch.Tensor, h_out_key: torch.Tensor, h_out_payload: torch.Tensor, n_sub: int = 4) -> int:
        """End to end with HOST buffers (pinned int64 tensors): this rank's probe keys travel host -> device, through partition +
        exchange + probe, and the result rows this rank ends up owning travel back.  Returns their number (rows [0, n) of
        h_out_key / h_out_payload).  Everything overlaps: the H2D copies run chunk by chunk on their own stream ahead of the
        owner partitions, and with the copy-engine exchange in "batch" mode the rows of sub-batch b go back to the host (on a
        third stream) as soon as its 32-byte result record has landed in pinned memory, while sub-batch b + 1 is still being
        exchanged and probed -- the host never waits for more than one small record per sub-batch."""
        n = h_keys.numel()
        dev = torch.device("cuda", torch.cuda.current_device())
        ws = getattr(self, "_host_ws", None)
        cap = n + n // 8 + (1 << 16)
        cap -= cap % n_sub
        if ws is None or ws["n"] < n or ws["n_sub"] != n_sub:
            ws = {"n": n, "n_sub": n_sub, "dk": torch.empty(n, dtype=torch.int64, device=dev),
                  "ok": torch.empty(cap, dtype=torch.int64, device=dev), "op": torch.empty(cap, dtype=torch.int64, device=dev),
                  "res": torch.zeros((n_sub, 4), dtype=torch.int64, device=dev),
                  "hres": torch.zeros((n_sub, 4), dtype=torch.int64).pin_memory(),
                  "s_in": torch.cuda.Stream(), "s_out": torch.cuda.Stream(), "cap": cap}
            self._host_ws = ws
        dk, ok, op, res, hres, cap = ws["dk"][:n], ws["ok"], ws["op"], ws["res"], ws["hres"], ws["cap"]
        main = torch.cuda.current_stream()
        s_in, s_out = ws["s_in"], ws["s_out"]
        s_in.wait_stream(main)
        ready = []
        with torch.cuda.stream(s_in):
            for dchunk, hchunk in zip(dk.chunk(n_sub), h_keys.chunk(n_sub)):
                dchunk.copy_(hchunk, non_blocking=True)
                e = torch.cuda.Event()
                e.record(s_in)
                ready.append(e)
        hcap = min(h_out_key.numel(), h_out_payload.numel())
        res.zero_()
        batch = self.copier is not None and self.ce_probe == "batch" and self.plan == "partition"
        if batch:
            done = []

            def landed(b):
                hres[b].copy_(res[b], non_blocking=True)
                e = torch.cuda.Event()
                e.record(main)
                done.append(e)

            self._probe_pipelined_ce(dk, n_sub, ok, op, res, ready=ready, landed=landed)
            capb, off = cap // n_sub, 0
            for b, e in enumerate(done):
                e.synchronize()  # the 32-byte record of sub-batch b is in pinned memory
                m = min(int(hres[b, 0]), capb, hcap - off)
                if m > 0:
                    s_out.wait_event(e)
                    with torch.cuda.stream(s_out):
                        h_out_key[off:off + m].copy_(ok[b * capb:b * capb + m], non_blocking=True)
                        h_out_payload[off:off + m].copy_(op[b * capb:b * capb + m], non_blocking=True)
                off += max(m, 0)
            s_out.synchronize()
            return off
        # one dense run of rows: stream-mode copy-engine pipeline, the fused peer scatter, NCCL all-to-all or the broadcast plan
        if self.copier is not None and self.plan == "partition":
            self._probe_pipelined_ce(dk, n_sub, ok, op, res, ready=ready)
        else:
            main.wait_event(ready[-1])
            if self.peer is not None and self.plan == "partition":
                self.probe_pipelined(dk, 1, ok, op, res[:1])
            else:
                self.probe(dk, capacity=cap, out_key=ok, out_payload=op, result=res[0], sync=False)
        hres.copy_(res, non_blocking=True)
        main.synchronize()
        m = min(int(hres[0, 0]), cap, hcap)
        if m > 0:
            h_out_key[:m].copy_(ok[:m], non_blocking=True)
            h_out_payload[:m].copy_(op[:m], non_blocking=True)
            main.synchronize()
        return max(m, 0)

    def probe(self, local_probe_keys: torch.Tensor, **kw) -> dict:
        """One probe pass: (partition + all-to-all unless broadcast plan) + local batch probe."""
        keys = local_probe_keys if self.plan == "broadcast" else self.shuffle(local_probe_keys)
        kw.setdefault("capacity", max(1, keys.numel()))
        return self.table.probe_batch(keys, **kw)


def make_comm(pkg, group=None):
    """cc_comm (include/cc_api.h) on top of torch.distributed: the control plane of the C-ABI partitioned join.  Returns the
    ctypes struct; it keeps its callbacks alive.  Works with NCCL (staging through device tensors) and gloo."""
    import ctypes as C

    L = pkg._lib
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    on_gpu = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")

    def allgather(user, send, recv, nbytes):
        try:
            mine = torch.frombuffer(bytearray(C.string_at(send, nbytes)), dtype=torch.uint8).to(dev)
            out = torch.empty(world * nbytes, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(out, mine, group=group)
            C.memmove(recv, out.cpu().numpy().tobytes(), world * nbytes)
            return 0
        except Exception:  # noqa: BLE001 -- reported to the C side as a failed callback
            return 1

    def barrier(user):
        try:
            dist.barrier(group=group)
            return 0
        except Exception:  # noqa: BLE001
            return 1

    comm = L.Comm()
    comm.rank, comm.world = rank, world
    comm._ag, comm._ba = L.ALLGATHER_FN(allgather), L.BARRIER_FN(barrier)  # keep the trampolines alive
    comm.allgather, comm.barrier, comm.user = comm._ag, comm._ba, None
    return comm


class CPartitionedJoin:
    """The partitioned join entirely behind the C ABI (cc_pjoin_*, csrc/pjoin.cu): partition kernels, copy-engine block copies
    into IPC-mapped peer memory and device-side ready / consumed flags -- no collective on the data path; torch.distributed is
    only the control plane (IPC handles and sizes at create, a barrier at destroy).  A C or C++ host does the same with its own
    cc_comm (host/simd_compaction.hpp: LocalComm, host/pjoin_main.cpp)."""

    def __init__(self, pkg, kind: int, local_build_keys: torch.Tensor, max_probe_rows: int, n_sub: int = 4, group=None):
        import ctypes as C

        self.pkg, self.group = pkg, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.log2p = log2_exact(self.world)
        self.n_sub = n_sub
        self._comm = make_comm(pkg, group)
        h = C.c_void_p()
        n = local_build_keys.numel()
        pkg._lib.check(pkg.lib().cc_pjoin_create(C.byref(h), C.byref(self._comm), kind, local_build_keys.data_ptr() if n else None, n,
                                                  int(max_probe_rows), int(n_sub), torch.cuda.current_stream().cuda_stream))
        self._h = h

    def table_info(self):
        import ctypes as C

        t = C.c_void_p()
        self.pkg._lib.check(self.pkg.lib().cc_pjoin_table(self._h, C.byref(t)))
        i = self.pkg._lib.HtInfo()
        self.pkg._lib.check(self.pkg.lib().cc_ht_get_info(t, C.byref(i)))
        return i

    def probe(self, local_probe_keys: torch.Tensor, out_key: Optional[torch.Tensor], out_payload: Optional[torch.Tensor], result: torch.Tensor) -> None:
        """Enqueues partition + exchange + probe of this rank's keys on the current stream (collective; nothing is synchronised).
        result: int64[4] device tensor (cc_probe_result); the rows this rank owns land densely in out_key / out_payload."""
        n = local_probe_keys.numel()
        cap = out_key.numel() if out_key is not None else (out_payload.numel() if out_payload is not None else 0)
        self.pkg._lib.check(self.pkg.lib().cc_pjoin_probe(self._h, local_probe_keys.data_ptr() if n else None, n,
                                                           out_key.data_ptr() if out_key is not None else None,
                                                           out_payload.data_ptr() if out_payload is not None else None, cap,
                                                           result.data_ptr(), torch.cuda.current_stream().cuda_stream))

    def probe_begin(self, local_probe_keys: torch.Tensor) -> None:
        """First half of a probe (cc_pjoin_probe_begin): partition + copies of this batch are enqueued.  May be called for batch
        t + 1 before probe_end of batch t: the exchange then runs underneath the probe of the batch before it."""
        n = local_probe_keys.numel()
        self.pkg._lib.check(self.pkg.lib().cc_pjoin_probe_begin(self._h, local_probe_keys.data_ptr() if n else None, n,
                                                                 torch.cuda.current_stream().cuda_stream))

    def probe_end(self, out_key: Optional[torch.Tensor], out_payload: Optional[torch.Tensor], result: torch.Tensor) -> None:
        """Second half (cc_pjoin_probe_end): waits for the oldest batch begun and probes it."""
        cap = out_key.numel() if out_key is not None else (out_payload.numel() if out_payload is not None else 0)
        self.pkg._lib.check(self.pkg.lib().cc_pjoin_probe_end(self._h, out_key.data_ptr() if out_key is not None else None,
                                                               out_payload.data_ptr() if out_payload is not None else None, cap,
                                                               result.data_ptr(), torch.cuda.current_stream().cuda_stream))

    def probe_host(self, h_keys: torch.Tensor, h_out_key: torch.Tensor, h_out_payload: torch.Tensor, n_chunks: int = 4) -> int:
        """End to end with HOST buffers (pinned int64 tensors; collective): this rank's probe keys travel host -> device in n_chunks
        chunks, every chunk is one batch of the C-ABI join (begin / end, pipelined two deep: the exchange of chunk c + 1 runs under
        the probe of chunk c), and the rows of chunk c go back to the host -- on their own stream, as soon as its 32-byte result
        record has landed in pinned memory -- while the later chunks are still being copied in, exchanged and probed.  Returns the
        number of rows this rank owns (rows [0, n) of h_out_key / h_out_payload)."""
        n = h_keys.numel()
        dev = torch.device("cuda", torch.cuda.current_device())
        per = -(-n // n_chunks) if n else 0
        capc = per + per // 8 + (1 << 16)  # rows a rank can end up owning from one chunk (hash partition: per +- a fraction of a percent)
        ws = getattr(self, "_host_ws", None)
        if ws is None or ws["per"] < per or ws["n_chunks"] != n_chunks:
            ws = {"per": per, "n_chunks": n_chunks, "dk": torch.empty(max(n, 1), dtype=torch.int64, device=dev),
                  "ok": [torch.empty(capc, dtype=torch.int64, device=dev) for _ in range(n_chunks)],
                  "op": [torch.empty(capc, dtype=torch.int64, device=dev) for _ in range(n_chunks)],
                  "res": torch.zeros((n_chunks, 4), dtype=torch.int64, device=dev),
                  "hres": torch.zeros((n_chunks, 4), dtype=torch.int64).pin_memory(),
                  "s_in": torch.cuda.Stream(), "s_out": torch.cuda.Stream()}
            self._host_ws = ws
        dk, res, hres, s_in, s_out = ws["dk"], ws["res"], ws["hres"], ws["s_in"], ws["s_out"]
        main = torch.cuda.current_stream()
        s_in.wait_stream(main)
        bounds = [(min(n, c * per), min(n, (c + 1) * per)) for c in range(n_chunks)]
        ready = []
        with torch.cuda.stream(s_in):
            for lo, hi in bounds:
                if hi > lo:
                    dk[lo:hi].copy_(h_keys[lo:hi], non_blocking=True)
                e = torch.cuda.Event()
                e.record(s_in)
                ready.append(e)
        done = []

        def begin(c):
            main.wait_event(ready[c])
            lo, hi = bounds[c]
            self.probe_begin(dk[lo:hi])

        def end(c):
            self.probe_end(ws["ok"][c], ws["op"][c], res[c])
            hres[c].copy_(res[c], non_blocking=True)
            e = torch.cuda.Event()
            e.record(main)
            done.append(e)

        begin(0)
        for c in range(n_chunks):
            if c + 1 < n_chunks:
                begin(c + 1)
            end(c)
        hcap = min(h_out_key.numel(), h_out_payload.numel())
        off = 0
        for c, e in enumerate(done):
            e.synchronize()  # the 32-byte record of chunk c is in pinned memory
            m = min(int(hres[c, 0]), ws["ok"][c].numel(), hcap - off)
            if m > 0:
                s_out.wait_event(e)
                with torch.cuda.stream(s_out):
                    h_out_key[off:off + m].copy_(ws["ok"][c][:m], non_blocking=True)
                    h_out_payload[off:off + m].copy_(ws["op"][c][:m], non_blocking=True)
            off += max(m, 0)
        s_out.synchronize()
        return off

    def close(self) -> None:
        if getattr(self, "_h", None):
            h, self._h = self._h, None
            self.pkg._lib.check(self.pkg.lib().cc_pjoin_destroy(h))


def reduce_result(n_matches: int, key_sum: int, payload_sum: int, device, group=None) -> Tuple[int, int, int]:
    """Sum of the per-rank counts / wrapping checksums (all-reduce of three int64)."""
    def wrap(v):
        v &= (1 << 64) - 1
        return v - (1 << 64) if v >= (1 << 63) else v
    t = torch.tensor([wrap(n_matches), wrap(key_sum), wrap(payload_sum)], dtype=torch.int64, device=device)
    dist.all_reduce(t, group=group)
    return tuple(int(x) & ((1 << 64) - 1) for x in t.cpu().tolist())
