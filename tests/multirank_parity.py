#!/usr/bin/env python
"""Multi-rank GPU parity gate for the hash-partitioned join ("1/2/4/8-GPU partitioned == 1-GPU == CPU").

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29611 \
        tests/multirank_parity.py [--report profiles/r2_multirank_parity_pP.txt]

One process per GPU (NCCL).  For BOTH table kinds (linear probing: linear_probing_ht.cpp:62-115 -- lanes walk past a match,
duplicates sit in separate slots; separate chaining: chaining_ht.cpp:60-136), chunk_factor in {1, 4, 8}, hit in {1, 2}, every
exchange (copy-engine "stream" and "batch", fused peer-memory scatter, NCCL all-to-all) and both plans (partition, broadcast):
  * the build and probe sides start range-partitioned over the ranks, like bench.py's C5 share;
  * the sharded result rows (probe key, matched build key) of all ranks are gathered on rank 0 and compared AS SORTED TUPLES
    with the oracle's pipeline over the undivided inputs (tests/oracle_lib.py: O.pipeline, the pinned restatement of the
    reference's scalar Probe + Next path), together with the match count and both wrapping column sums;
  * every rank checks the OWNER PROPERTY: each result row it holds, and each key its local table holds, hashes to this rank
    (murmurhash64(key) >> (64 - log2 P)); broadcast plan: every rank's table holds the whole build side instead.
Plus the two failure paths: skewed keys overrunning a copy-engine region in an early sub-batch (n_sub > n_buffers) must raise
on EVERY rank at check_overflow(), and an undersized peer-exchange buffer must raise on every rank before anyone scatters.

The oracle is only the checker here (test infrastructure); the product path is libccb200.so on every rank.
Also runnable through pytest on a multi-GPU box: tests/test_gpu_multirank.py launches it with torchrun.
"""
from __future__ import annotations

import argparse
import importlib
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
PKG_NAME = "chunk-compaction-in-vectorized-execution-simd_b200"
U64 = (1 << 64) - 1


def gather_rows(cols, rank, world, dev):
    """all ranks -> rank 0: variable-length int64 columns (padded all_gather); returns a list of per-rank arrays on rank 0"""
    n = torch.tensor([cols[0].numel()], dtype=torch.int64, device=dev)
    sizes = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(x.item()) for x in sizes]
    mx = max(sizes + [1])
    out = []
    for c in cols:
        pad = torch.zeros(mx, dtype=torch.int64, device=dev)
        pad[: c.numel()] = c
        got = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(got, pad)
        out.append([g[:s].cpu().numpy() for g, s in zip(got, sizes)] if rank == 0 else None)
    return out


def all_ok(flag: bool, dev) -> bool:
    t = torch.tensor([1 if flag else 0], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--report", default=None, help="write the PASS / FAIL lines to this file (rank 0)")
    ap.add_argument("--log2-build", type=int, default=18, help="build keys of the whole join")
    ap.add_argument("--log2-probe", type=int, default=20, help="probe keys of the whole join")
    ap.add_argument("--quick", action="store_true", help="cf in {1, 4} and hit 2 only")
    ap.add_argument("--partitioned-probe", action="store_true",
                    help="force the partitioned probe strategy with 1 MiB table slices, so that these small tables take the large-table path "
                         "(slice partition behind the exchange, incremental probe regions) instead of the direct probe")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pkg = importlib.import_module(PKG_NAME)
    pkg.init(local)
    par = importlib.import_module(PKG_NAME + ".parallel")
    import oracle_lib as O

    if args.partitioned_probe:
        pkg.set_probe_strategy(2, 1 << 20)
        os.environ["CCB_PJ_SLICE_BYTES"] = str(64 << 10)  # the C-ABI join then groups by (owner, 64 KiB table slice) on the sender

    sys.stdout.flush()
    dist.init_process_group("nccl", device_id=dev)
    log2p = par.log2_exact(world)
    lines, failures = [], []

    def say(msg):
        if rank == 0:
            print(msg, flush=True)
            lines.append(msg)

    n_build, n_probe = 1 << args.log2_build, 1 << args.log2_probe
    nb_local, np_local = n_build // world, n_probe // world
    say(f"# multi-rank parity, P = {world} ranks ({torch.cuda.get_device_name(local)}), whole join: {n_build} build keys, {n_probe} probe keys; "
        f"probe strategy {'partitioned (forced, 1 MiB slices)' if args.partitioned_probe else 'auto (direct at this size)'}; "
        f"git {os.popen('git -C ' + ROOT + ' rev-parse --short HEAD 2>/dev/null').read().strip() or 'n/a'}")
    t_start = time.time()

    def owner_of(t: torch.Tensor) -> torch.Tensor:
        return (pkg.murmurhash64(t) >> (64 - log2p)) & (world - 1) if world > 1 else torch.zeros_like(t)

    cfs = (1, 4) if args.quick else (1, 4, 8)
    hits = (2,) if args.quick else (1, 2)
    # "cabi": the whole partitioned join behind the C ABI (cc_pjoin_*: device-side flags instead of collectives on the data path)
    modes = [("partition", "ce", "stream", 4), ("partition", "ce", "batch", 3), ("partition", "p2p", None, 2), ("partition", "nccl", None, 1),
             ("partition", "cabi", None, 5), ("broadcast", "nccl", None, 1)]
    for kind, kname in ((pkg.CC_HT_LP, "lp"), (pkg.CC_HT_CHAIN, "chain")):
        OT = O.OracleLP if kind == pkg.CC_HT_LP else O.OracleChain
        for cf in cfs:
            build_all = O.build_keys(n_build, cf)  # the reference's own generator (chaining_ht.cpp:15-26)
            oracle_tab = OT(build_all) if rank == 0 else None
            my_build = torch.from_numpy(build_all[rank * nb_local:(rank + 1) * nb_local].copy()).to(dev)
            for plan, exchange, ce_probe, n_sub in modes:
                cap_rows = -(-np_local // n_sub) if exchange == "ce" else np_local * 4 + nb_local * 4 + (1 << 16)
                if exchange == "cabi":
                    cjoin = par.CPartitionedJoin(pkg, kind, my_build, np_local, n_sub=n_sub)
                    info = cjoin.table_info()
                    tot = torch.tensor([int(info.n_keys)], dtype=torch.int64, device=dev)
                    dist.all_reduce(tot)
                    own_b = int(tot.item()) == n_build  # (the owner property of the table keys follows from that of the probe rows below)
                    join = None
                else:
                    join = par.PartitionedJoin(pkg, kind, my_build, plan=plan, exchange=exchange, capacity_rows=cap_rows,
                                               ce_probe=ce_probe or "auto")
                    # build side: every key in this rank's table hashes here (partition plan) / the table holds everything (broadcast)
                    exp = join.table.export()
                    tkeys = exp[exp != -1] if kind == pkg.CC_HT_LP else exp[2]
                if exchange == "cabi":
                    pass
                elif plan == "partition":
                    own_b = bool((owner_of(torch.from_numpy(np.ascontiguousarray(tkeys)).to(dev)) == rank).all().item()) if tkeys.size else True
                    tot = torch.tensor([tkeys.size], dtype=torch.int64, device=dev)
                    dist.all_reduce(tot)
                    own_b = own_b and int(tot.item()) == n_build
                else:
                    own_b = tkeys.size == n_build and np.array_equal(np.sort(tkeys), np.sort(build_all))
                for hit in hits:
                    probe_all = O.gen_keys_counter(n_probe, 2 + cf + 10 * hit, n_build * hit - 1)
                    my_probe = torch.from_numpy(probe_all[rank * np_local:(rank + 1) * np_local].copy()).to(dev)
                    cap = (2 * np_local + (1 << 16)) * max(1, cf) // 1
                    cap -= cap % n_sub
                    ok = torch.full((cap,), -7, dtype=torch.int64, device=dev)
                    op = torch.full((cap,), -7, dtype=torch.int64, device=dev)
                    res = torch.zeros((n_sub, 4), dtype=torch.int64, device=dev)
                    tag = f"{kname:5s} cf={cf} hit={hit} plan={plan:9s} exchange={exchange + ('/' + ce_probe if ce_probe else ''):9s} n_sub={n_sub}"
                    try:
                        for rep in range(2):  # twice: the second pass runs on rotated buffers
                            res.zero_()
                            if exchange == "cabi":
                                cjoin.probe(my_probe, ok, op, res[0])
                            elif plan == "partition" and exchange in ("ce", "p2p"):
                                join.probe_pipelined(my_probe, n_sub, ok, op, res)
                            else:
                                join.probe(my_probe, capacity=cap, out_key=ok, out_payload=op, result=res[0], sync=False)
                            torch.cuda.synchronize()
                        if join is not None and join.copier is not None:
                            join.copier.check_overflow()
                        rr = res.cpu().numpy().view(np.uint64)
                        host_ok = True
                        if exchange == "cabi" or (exchange == "ce" and plan == "partition"):
                            # the host-buffer path of the same join (pinned keys in, the rows this rank owns out): same rows
                            hk = my_probe.cpu().pin_memory()
                            hok = torch.empty(cap, dtype=torch.int64).pin_memory()
                            hop = torch.empty(cap, dtype=torch.int64).pin_memory()
                            m_host = cjoin.probe_host(hk, hok, hop) if exchange == "cabi" else join.probe_host(hk, hok, hop, n_sub=n_sub)
                            m_dev = int(rr[:, 0].sum())
                            host_ok = m_host == m_dev and int(hok[:m_host].sum()) & U64 == int(rr[:, 1].sum()) & U64 and bool((hok[:m_host] == hop[:m_host]).all())
                        dense = n_sub == 1 or (exchange == "ce" and ce_probe == "stream") or exchange == "cabi"
                        capb = cap // n_sub
                        if dense:
                            m = int(rr[0, 0])
                            gk, gp = ok[:m], op[:m]
                        else:
                            gk = torch.cat([ok[b * capb:b * capb + int(rr[b, 0])] for b in range(n_sub)])
                            gp = torch.cat([op[b * capb:b * capb + int(rr[b, 0])] for b in range(n_sub)])
                        overflow = int(rr[:, 3].sum())
                        own_p = True if plan == "broadcast" else (bool((owner_of(gk.contiguous()) == rank).all().item()) if gk.numel() else True)
                        sums = rr.sum(axis=0, dtype=np.uint64)
                        n_red, ks_red, ps_red = par.reduce_result(int(sums[0]), int(sums[1]), int(sums[2]), dev)
                        cols = gather_rows([gk.contiguous(), gp.contiguous()], rank, world, dev)
                        verdict = ""
                        if rank == 0:
                            want = O.pipeline([oracle_tab], probe_all.reshape(-1, 1), 2048, collect=True)
                            wt = want["tuples"][:, [0, 2]] if want["n_tuples"] else np.empty((0, 2), dtype=np.int64)
                            got = np.stack([np.concatenate(cols[0]), np.concatenate(cols[1])], axis=1)
                            same_n = got.shape[0] == want["n_tuples"] == n_red
                            same_sums = (ks_red, ps_red) == (want["colsum"][0], want["colsum"][2])
                            same_rows = same_n and np.array_equal(O.sort_rows(got), O.sort_rows(wt))
                            good = same_n and same_sums and same_rows and overflow == 0
                            verdict = (f"rows {got.shape[0]} (oracle {want['n_tuples']})  count {'ok' if same_n else 'MISMATCH'}  checksums "
                                       f"{'ok' if same_sums else 'MISMATCH'}  sorted tuples {'ok' if same_rows else 'MISMATCH'}")
                        else:
                            good = overflow == 0
                        good = all_ok(good and own_p and own_b and host_ok, dev)
                        say(f"{'PASS' if good else 'FAIL'}  {tag}  {verdict}  owner property (probe rows, table keys) on every rank "
                            f"{'ok' if good or (own_p and own_b) else 'VIOLATED'}")
                        if not good:
                            failures.append(tag)
                    except Exception as e:  # noqa: BLE001 -- a crash in one combination is a failure of that combination
                        say(f"FAIL  {tag}  {type(e).__name__}: {e}")
                        failures.append(tag)
                        raise
                if join is None:
                    cjoin.close()
                    del cjoin
                    continue
                if join.copier is not None:
                    join.copier.close()
                if join.peer is not None:
                    join.peer.close()
                del join

    # ---- failure paths ------------------------------------------------------------------------------------------------
    if world > 1:
        # (1) skewed keys: every key of sub-batch 0 (of 5, more than the 3 rotating buffers) is the same -> one region overruns in
        #     shuffle 0; shuffles 1..4 are well-behaved and reuse the flags.  check_overflow() must still raise, on EVERY rank.
        my_build = torch.arange(rank * nb_local, (rank + 1) * nb_local, dtype=torch.int64, device=dev)
        n_sub = 5
        per = -(-np_local // n_sub)
        join = par.PartitionedJoin(pkg, pkg.CC_HT_LP, my_build, plan="partition", exchange="ce", capacity_rows=per, ce_probe="batch")
        keys = pkg.gen_keys_counter(np_local, 77, n_build - 1, first=rank * np_local).clone()
        if rank == 0:
            keys[:per] = 42  # only rank 0 is skewed: the others must learn about it through the collective check
        cap = (np_local * 2 // n_sub) * n_sub
        ok, op = torch.empty(cap, dtype=torch.int64, device=dev), torch.empty(cap, dtype=torch.int64, device=dev)
        res = torch.zeros((n_sub, 4), dtype=torch.int64, device=dev)
        join.probe_pipelined(keys, n_sub, ok, op, res)
        torch.cuda.synchronize()
        raised = False
        try:
            join.copier.check_overflow()
        except RuntimeError:
            raised = True
        good = all_ok(raised, dev)
        say(f"{'PASS' if good else 'FAIL'}  copy-engine exchange: region overrun in sub-batch 0 of {n_sub} reported on every rank by check_overflow()")
        if not good:
            failures.append("ce overflow")
        # a clean pass afterwards: the flags were reset by the check
        keys = pkg.gen_keys_counter(np_local, 78, n_build - 1, first=rank * np_local)
        join.probe_pipelined(keys, n_sub, ok, op, res)
        torch.cuda.synchronize()
        clean = True
        try:
            join.copier.check_overflow()
        except RuntimeError:
            clean = False
        rr = res.cpu().numpy().view(np.uint64).sum(axis=0, dtype=np.uint64)
        n_red, ks_red, _ = par.reduce_result(int(rr[0]), int(rr[1]), int(rr[2]), dev)
        ksum = torch.tensor([int(keys.sum().item())], dtype=torch.int64, device=dev)
        dist.all_reduce(ksum)
        good = all_ok(clean and n_red == n_probe and ks_red == (int(ksum.item()) & U64), dev)
        say(f"{'PASS' if good else 'FAIL'}  copy-engine exchange: clean pass after the reported overrun (count {n_red}, checksum {'ok' if good else 'MISMATCH'})")
        if not good:
            failures.append("ce after overflow")
        join.copier.close()
        del join
        # (2) peer exchange with receive buffers too small for ONE owner (rank 1): every rank must raise before the scatter
        join = par.PartitionedJoin(pkg, pkg.CC_HT_LP, my_build, plan="partition", exchange="p2p", capacity_rows=nb_local * 2 + 4096)
        skew = torch.full((nb_local * 4,), 0, dtype=torch.int64, device=dev)
        target = 1 % world
        base = torch.arange(1, 1 << 16, dtype=torch.int64, device=dev)
        cand = base[owner_of(base) == target]
        skew[:] = cand[0]  # all rows of every rank go to rank `target`: 4 * nb_local * P rows > its capacity
        raised = False
        try:
            join.peer.shuffle(skew)
        except RuntimeError:
            raised = True
        good = all_ok(raised, dev)
        say(f"{'PASS' if good else 'FAIL'}  peer exchange: an overflowing receive buffer of rank {target} raises on every rank before any store")
        if not good:
            failures.append("p2p capacity")
        join.peer.close()
        del join

    say(f"# {'ALL GREEN' if not failures else 'FAILURES: ' + '; '.join(failures)}  ({time.time() - t_start:.1f} s)")
    if rank == 0 and args.report:
        os.makedirs(os.path.dirname(os.path.abspath(args.report)), exist_ok=True)
        with open(args.report, "w") as f:
            f.write("\n".join(lines) + "\n")
    dist.barrier()
    dist.destroy_process_group()
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
