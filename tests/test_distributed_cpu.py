"""world_size-2 gloo tests (CPU) of the multi-GPU host layer: partition -> count exchange ->
variable-size all-to-all -> local join.  The CUDA partition/probe kernels need a GPU, so the
per-rank partitioning and local joins here use the ORACLE (test infrastructure) -- what is under
test is parallel.py's exchange logic and the partition-id contract (high hash bits)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG_NAME, ROOT


def _worker(rank: int, world: int, port: int, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib

    import oracle_lib as O

    par = importlib.import_module(PKG_NAME + ".parallel")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        log2p = par.log2_exact(world)
        n_build, n_probe = 5000, 40000
        # range-partitioned inputs, as in bench.py: every rank owns a slice of the build keys (cf = 2)
        build_all = O.build_keys(n_build * world, 2)
        probe_all = O.gen_keys_counter(n_probe * world, 2, (1 << 14) - 1)
        my_build = build_all[rank * n_build:(rank + 1) * n_build]
        my_probe = probe_all[rank * n_probe:(rank + 1) * n_probe]

        def shuffle(keys):
            pid = (O.murmurhash64(keys.view(np.uint64)) >> np.uint64(64 - log2p)).astype(np.int64)
            order = np.argsort(pid, kind="stable")
            counts = np.bincount(pid, minlength=world)
            send = torch.from_numpy(keys[order].copy())
            recv_counts = par.exchange_counts(torch.from_numpy(counts.astype(np.int64)))
            got = par.exchange_rows(send, counts.tolist(), recv_counts.tolist())
            return got.numpy().copy()

        mine_b = shuffle(my_build)
        mine_p = shuffle(my_probe)
        # ownership: every received key hashes to this rank
        for arr in (mine_b, mine_p):
            pid = O.murmurhash64(arr.view(np.uint64)) >> np.uint64(64 - log2p)
            assert np.all(pid == rank)
        local = O.pipeline([O.OracleLP(mine_b)], mine_p.reshape(-1, 1), 2048)
        n, ks, ps = par.reduce_result(local["n_tuples"], local["colsum"][0], local["colsum"][2], torch.device("cpu"))
        # single-process answer over the union
        want = O.pipeline([O.OracleLP(build_all)], probe_all.reshape(-1, 1), 2048)
        assert (n, ks, ps) == (want["n_tuples"], want["colsum"][0], want["colsum"][2])
        # nothing lost or duplicated in the exchange
        tot = torch.tensor([mine_b.size, mine_p.size], dtype=torch.int64)
        dist.all_reduce(tot)
        assert tot.tolist() == [n_build * world, n_probe * world]
        assert par.choose_plan(2_000_000, 20_000_000 * 8, world) in ("broadcast", "partition")
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_join_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, "ok") for r in range(world)], results


def test_plan_rule_and_log2():
    import importlib

    par = importlib.import_module(PKG_NAME + ".parallel")
    assert par.log2_exact(8) == 3 and par.log2_exact(1) == 0
    with pytest.raises(ValueError):
        par.log2_exact(6)
    # a 2M-key build against 8B probe keys is broadcast; 1B build against 8B probe is partitioned (C5)
    assert par.choose_plan(2_000_000, 8 << 30, 8) == "broadcast"
    assert par.choose_plan(1 << 30, 8 << 30, 8) == "partition"


def _e2e_worker(rank: int, world: int, port: int, q):
    """bench.distributed_e2e under gloo: the step is a CPU stand-in (modulo partition + the real exchange helpers + an
    identity 'probe'), what is under test is the leg's own accounting -- H2D / D2H slicing, counters, cross-rank checks."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    import importlib

    import bench

    par = importlib.import_module(PKG_NAME + ".parallel")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev = torch.device("cpu")
        ne, n_sub = 1 << 12, 4
        cap = (ne * 2 // n_sub) * n_sub
        capb = cap // n_sub
        out_key = torch.zeros(cap, dtype=torch.int64)
        out_payload = torch.zeros(cap, dtype=torch.int64)
        result = torch.zeros((n_sub, 4), dtype=torch.int64)

        def make_step(dense, corrupt=False):
            def step(dk):
                result.zero_()
                off = 0
                for b, chunk in enumerate(np.array_split(dk.numpy(), n_sub)):
                    pid = chunk % world
                    order = np.argsort(pid, kind="stable")
                    counts = np.bincount(pid, minlength=world).astype(np.int64)
                    recv_counts = par.exchange_counts(torch.from_numpy(counts))
                    got = par.exchange_rows(torch.from_numpy(chunk[order].copy()), counts.tolist(), recv_counts.tolist())
                    m = got.numel()
                    at = off if dense else b * capb
                    out_key[at:at + m] = got
                    out_payload[at:at + m] = got + (1 if corrupt and b == 1 else 0)
                    if dense:
                        off += m
                        result[0, 0] += m
                    else:
                        result[b, 0] = m
            return step

        gen = lambda n, first: torch.arange(first, first + n, dtype=torch.int64)
        for dense in (True, False):
            e2e, note = bench.distributed_e2e(dist, dev, ne=ne, world=world, rank=rank, n_sub=n_sub, dense=dense, cap=cap, step=make_step(dense),
                                              result=result, out_key=out_key, out_payload=out_payload, gen_keys=gen, iters=2)
            assert note is None and e2e is not None, note
            assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] == ne * 8 * world and e2e["d2h_bytes_per_step"] == (16 * ne + 32 * n_sub) * world
        e2e, note = bench.distributed_e2e(dist, dev, ne=ne, world=world, rank=rank, n_sub=n_sub, dense=False, cap=cap, step=make_step(False, corrupt=True),
                                          result=result, out_key=out_key, out_payload=out_payload, gen_keys=gen, iters=1)
        assert e2e is None and "check failed" in note

        def broken(dk):
            raise RuntimeError("boom")

        e2e, note = bench.distributed_e2e(dist, dev, ne=ne, world=world, rank=rank, n_sub=n_sub, dense=True, cap=cap, step=broken,
                                          result=result, out_key=out_key, out_payload=out_payload, gen_keys=gen, iters=1)
        assert e2e is None and "boom" in note
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


def test_bench_distributed_e2e_leg_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 77
    procs = [ctx.Process(target=_e2e_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, "ok") for r in range(world)], results


def _comm_worker(rank: int, world: int, port: int, q):
    """parallel.make_comm: the cc_comm control plane (include/cc_api.h) over torch.distributed, called the way libccb200 calls it
    (through the C function pointers), under gloo"""
    import ctypes as C

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    import importlib

    pkg = importlib.import_module(PKG_NAME)
    par = importlib.import_module(PKG_NAME + ".parallel")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = par.make_comm(pkg)
        assert (comm.rank, comm.world) == (rank, world)
        for nbytes in (8, 64, 24):
            send = (C.c_ubyte * nbytes)(*[(rank * 37 + i) % 256 for i in range(nbytes)])
            recv = (C.c_ubyte * (nbytes * world))()
            assert comm.allgather(None, C.cast(send, C.c_void_p), C.cast(recv, C.c_void_p), nbytes) == 0
            assert list(recv) == [(r * 37 + i) % 256 for r in range(world) for i in range(nbytes)]
        assert comm.barrier(None) == 0
        assert par.table_slots(True, 1 << 30, 134224497, 8) == 1 << 29 and par.table_slots(False, 1 << 30, 134224497, 8) == 1 << 28
        assert par.table_slots(True, 1000, 900, 2) == 2048  # a skewed partition doubles the LP table instead of overfilling it
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


def test_c_abi_comm_callbacks_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 131
    procs = [ctx.Process(target=_comm_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, "ok") for r in range(world)], results
