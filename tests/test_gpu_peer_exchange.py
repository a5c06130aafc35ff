"""GPU test of the fused scatter+exchange path (PeerExchange, probe_pipelined) in a single-process
NCCL group (world_size 1: the peer buffer is this GPU's own, every code path except the IPC mapping runs)."""
import importlib
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist

from conftest import PKG_NAME

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pg():
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    if not dist.is_initialized():
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    yield
    dist.destroy_process_group()


@pytest.mark.parametrize("n_sub,peer_blocks", [(1, 0), (2, 0), (4, 16), (3, 48)])
def test_pipelined_peer_exchange_matches_direct(ccb, pg, n_sub, peer_blocks):
    par = importlib.import_module(PKG_NAME + ".parallel")
    n_build, n_probe = 1 << 22, 3 * (1 << 22)
    build = torch.arange(n_build, dtype=torch.int64, device="cuda")
    join = par.PartitionedJoin(ccb, ccb.CC_HT_LP, build, plan="partition", exchange="p2p", capacity_rows=n_probe + 4096,
                               peer_blocks=peer_blocks)
    assert join.n_build_local == n_build
    keys = ccb.gen_keys_counter(n_probe, 7, 2 * n_build - 1)  # hit rate 1/2
    hits = keys[keys < n_build]
    want_n, want_sum = hits.numel(), int(hits.sum().item()) & ((1 << 64) - 1)
    cap = n_probe - n_probe % n_sub
    ok = torch.empty(cap, dtype=torch.int64, device="cuda")
    op = torch.empty(cap, dtype=torch.int64, device="cuda")
    res = torch.zeros((n_sub, 4), dtype=torch.int64, device="cuda")
    for rep in range(4):  # several passes: exercises the buffer rotation
        join.probe_pipelined(keys, n_sub, ok, op, res)
        torch.cuda.synchronize()
        r = res.cpu().numpy().view(np.uint64).sum(axis=0, dtype=np.uint64)
        assert int(r[0]) == want_n and int(r[1]) == want_sum and int(r[2]) == want_sum, (rep, r)
    join.peer.close()


@pytest.mark.parametrize("n_sub", [1, 2, 4, 5])
def test_pipelined_copy_exchange_matches_direct(ccb, pg, n_sub):
    """CopyExchange (single-pass partition + block copies + segmented probe) in a one-rank group: every code path except
    the IPC mapping runs; several passes exercise the rotation of the three send / receive buffers."""
    par = importlib.import_module(PKG_NAME + ".parallel")
    n_build, n_probe = 1 << 22, 3 * (1 << 22) + 777
    build = torch.arange(n_build, dtype=torch.int64, device="cuda")
    join = par.PartitionedJoin(ccb, ccb.CC_HT_LP, build, plan="partition", exchange="ce", capacity_rows=-(-n_probe // n_sub))
    keys = ccb.gen_keys_counter(n_probe, 7, 2 * n_build - 1)  # hit rate 1/2
    hits = keys[keys < n_build]
    want_n, want_sum = hits.numel(), int(hits.sum().item()) & ((1 << 64) - 1)
    cap = n_probe
    ok = torch.empty(cap, dtype=torch.int64, device="cuda")
    op = torch.empty(cap, dtype=torch.int64, device="cuda")
    res = torch.zeros((n_sub, 4), dtype=torch.int64, device="cuda")
    for rep in range(4):
        join.probe_pipelined(keys, n_sub, ok, op, res)
        torch.cuda.synchronize()
        r = res.cpu().numpy().view(np.uint64).sum(axis=0, dtype=np.uint64)
        assert int(r[0]) == want_n and int(r[1]) == want_sum and int(r[2]) == want_sum and int(r[3]) == 0, (rep, r)
    join.copier.check_overflow()
    # the materialised rows are exactly the matching probe keys (one dense output over all sub-batches)
    assert torch.equal(torch.sort(ok[:want_n])[0], torch.sort(hits)[0]) and torch.equal(torch.sort(op[:want_n])[0], torch.sort(hits)[0])
    join.copier.close()


@pytest.mark.parametrize("n_sub", [1, 3, 4])
def test_pipelined_copy_exchange_batch_mode(ccb, pg, n_sub):
    """ce_probe="batch" (the default beyond 2 GPUs): every sub-batch is probed as it lands; its rows go to slice b of the
    output columns and its counters to results[b]."""
    par = importlib.import_module(PKG_NAME + ".parallel")
    n_build, n_probe = 1 << 22, 3 * (1 << 22) + 777
    build = torch.arange(n_build, dtype=torch.int64, device="cuda")
    join = par.PartitionedJoin(ccb, ccb.CC_HT_LP, build, plan="partition", exchange="ce", capacity_rows=-(-n_probe // n_sub),
                               ce_probe="batch")
    assert join.ce_probe == "batch"
    keys = ccb.gen_keys_counter(n_probe, 7, 2 * n_build - 1)  # hit rate 1/2
    hits = keys[keys < n_build]
    want_n, want_sum = hits.numel(), int(hits.sum().item()) & ((1 << 64) - 1)
    cap = n_probe
    capb = cap // n_sub
    ok = torch.empty(cap, dtype=torch.int64, device="cuda")
    op = torch.empty(cap, dtype=torch.int64, device="cuda")
    res = torch.zeros((n_sub, 4), dtype=torch.int64, device="cuda")
    for rep in range(4):
        join.probe_pipelined(keys, n_sub, ok, op, res)
        torch.cuda.synchronize()
        rr = res.cpu().numpy().view(np.uint64)
        r = rr.sum(axis=0, dtype=np.uint64)
        assert int(r[0]) == want_n and int(r[1]) == want_sum and int(r[2]) == want_sum and int(r[3]) == 0, (rep, r)
    join.copier.check_overflow()
    got_k = torch.cat([ok[b * capb: b * capb + int(rr[b, 0])] for b in range(n_sub)])
    got_p = torch.cat([op[b * capb: b * capb + int(rr[b, 0])] for b in range(n_sub)])
    assert torch.equal(torch.sort(got_k)[0], torch.sort(hits)[0]) and torch.equal(torch.sort(got_p)[0], torch.sort(hits)[0])
    join.copier.close()
    with pytest.raises(ValueError):
        par.PartitionedJoin(ccb, ccb.CC_HT_LP, build, plan="broadcast", ce_probe="sometimes")


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("strategy", [1, 2])
def test_probe_stream_equals_one_batch(ccb, kind, strategy):
    """cc_probe_stream_*: dense and segmented pieces added one by one == cc_probe_batch over their concatenation (small-table
    path: every piece probed at once; large-table path: pieces scattered into the slice regions, one probe at the end)."""
    n_build = 1 << 21
    T = ccb.LPHashTable if kind == 0 else ccb.HashTable
    tab = T(n_build, 1)
    seg_cap, fills = 4096 * 16, [4096 * 16, 0, 777, 4096 * 3 + 1]
    seg_col = torch.full((len(fills) * seg_cap,), 1, dtype=torch.int64, device="cuda")
    pieces = [ccb.gen_keys_counter(n, 50 + i, 2 * n_build - 1) for i, n in enumerate([300000, 1, 123456])]
    seg_parts = []
    for s, f in enumerate(fills):
        k = ccb.gen_keys_counter(max(f, 1), 90 + s, 2 * n_build - 1)[:f]
        seg_col[s * seg_cap: s * seg_cap + f] = k
        seg_parts.append(k)
    dense = torch.cat(pieces + seg_parts)
    counts = torch.tensor(fills, dtype=torch.int64, device="cuda")
    ccb.set_probe_strategy(strategy, 4 << 20)
    try:
        want = tab.probe_batch(dense)
        cap = dense.numel()
        ok = torch.empty(cap, dtype=torch.int64, device="cuda")
        op = torch.empty(cap, dtype=torch.int64, device="cuda")
        st = tab.probe_stream(dense.numel(), capacity=cap, out_key=ok, out_payload=op)
        st.add(pieces[0])
        st.add_segmented(seg_col, len(fills), seg_cap, counts)
        st.add(pieces[1])
        st.add(pieces[2])
        st.add(pieces[2][:0])
        got = st.finish()
    finally:
        ccb.set_probe_strategy(0, 32 << 20)
    for f in ("n_matches", "key_sum", "payload_sum", "overflow"):
        assert got[f] == want[f], (f, got[f], want[f])
    n = got["n_matches"]
    assert torch.equal(torch.sort(ok[:n])[0], torch.sort(want["out_key"][:n])[0]) and torch.equal(torch.sort(op[:n])[0], torch.sort(ok[:n])[0])


def test_probe_stream_reports_region_overrun(ccb):
    """Heavily skewed keys overrun a slice region of the incremental probe: reported as overflow bit 1, never silently wrong."""
    tab = ccb.LPHashTable(1 << 21, 1)
    same = torch.full((1 << 20,), 42, dtype=torch.int64, device="cuda")
    ccb.set_probe_strategy(2, 4 << 20)
    try:
        st = tab.probe_stream(same.numel(), capacity=same.numel(), out_key=torch.empty_like(same), out_payload=torch.empty_like(same))
        st.add(same)
        got = st.finish()
    finally:
        ccb.set_probe_strategy(0, 32 << 20)
    assert got["overflow"] & 2


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("strategy", [1, 2, 3])
def test_probe_batch_segmented_equals_dense(ccb, kind, strategy):
    """cc_probe_batch_segmented over ragged segments (empty, one row, full to capacity) == cc_probe_batch over the
    concatenation, for the direct probe and both partitioned strategies, both table kinds."""
    n_build = 1 << 21
    T = ccb.LPHashTable if kind == 0 else ccb.HashTable
    tab = T(n_build, 1)
    seg_cap, fills = 4096 * 40, [4096 * 40, 0, 1, 4096 * 7 + 5, 123457, 4096 * 40 - 1, 4096, 99999]
    col = torch.full((len(fills) * seg_cap,), 1, dtype=torch.int64, device="cuda")  # slack rows hold a MATCHING key: reading one changes the count
    parts = []
    for s, f in enumerate(fills):
        k = ccb.gen_keys_counter(max(f, 1), 100 + s, 2 * n_build - 1)[:f]
        col[s * seg_cap: s * seg_cap + f] = k
        parts.append(k)
    dense = torch.cat(parts)
    counts = torch.tensor(fills, dtype=torch.int64, device="cuda")
    ccb.set_probe_strategy(strategy, 4 << 20)
    try:
        want = tab.probe_batch(dense)
        cap = dense.numel()
        ok = torch.empty(cap, dtype=torch.int64, device="cuda")
        op = torch.empty(cap, dtype=torch.int64, device="cuda")
        got = tab.probe_batch_segmented(col, len(fills), seg_cap, counts, capacity=cap, out_key=ok, out_payload=op)
    finally:
        ccb.set_probe_strategy(0, 32 << 20)
    for f in ("n_matches", "key_sum", "payload_sum", "overflow"):
        assert got[f] == want[f], (f, got[f], want[f])
    n = got["n_matches"]
    assert torch.equal(torch.sort(ok[:n])[0], torch.sort(want["out_key"][:n])[0]) and torch.equal(torch.sort(op[:n])[0], torch.sort(ok[:n])[0])


def test_partition_single_regions(ccb):
    """cc_partition_single: fixed regions, device-side counts, overflow flag on skewed keys."""
    n, log2p = (1 << 20) + 4321, 3
    P = 1 << log2p
    keys = ccb.gen_keys_counter(n, 5, (1 << 40) - 1)
    cap = ((n // P) * 9 // 8 + 8192 + 4095) // 4096 * 4096
    other = torch.full((P * cap,), -9, dtype=torch.int64, device="cuda")
    out, counts, flag = ccb.partition_single(keys, log2p, cap, self_part=5, self_out_ptr=other.data_ptr())  # partition 5 is redirected
    torch.cuda.synchronize()
    assert int(flag.item()) == 0 and int(counts.sum().item()) == n
    c5 = int(counts[5].item())
    assert int((other != -9).sum().item()) == c5 and bool((other[5 * cap: 5 * cap + c5] != -9).all())
    out[5 * cap: 5 * cap + c5] = other[5 * cap: 5 * cap + c5]
    h = ccb.murmurhash64(keys)
    pid = (h.view(torch.int64) >> (64 - log2p)) & (P - 1)
    for p in range(P):
        c = int(counts[p].item())
        assert torch.equal(torch.sort(out[p * cap: p * cap + c])[0], torch.sort(keys[pid == p])[0])
    same = torch.full((n,), 42, dtype=torch.int64, device="cuda")  # every key in one partition: the region must overrun
    _, _, flag = ccb.partition_single(same, log2p, cap)
    assert int(flag.item()) != 0
    # the flag is sticky: a later, well-behaved call on the same flag must not erase the report (one flag watches over
    # all shuffles of a step in the copy-engine exchange)
    _, counts2, flag = ccb.partition_single(keys, log2p, cap, overflow=flag)
    assert int(flag.item()) != 0 and int(counts2.sum().item()) == n


@pytest.mark.parametrize("log2p", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("n", [4096 * 37, 4096 * 600 + 17, 1000])
def test_partition_single_fanouts(ccb, log2p, n):
    """Few partitions (<= 16: the multi-GPU owner partition) take the ballot-ranked path of the scatter kernel for full tiles,
    more take the shared-memory-atomic path; either way every region must hold exactly its partition's keys."""
    P = 1 << log2p
    for keys in (ccb.gen_keys_counter(n, 11, (1 << 44) - 1), torch.arange(n, dtype=torch.int64, device="cuda")):
        cap = ((n // P) * 5 // 4 + 8192 + 4095) // 4096 * 4096
        out, counts, flag = ccb.partition_single(keys, log2p, cap)
        torch.cuda.synchronize()
        assert int(flag.item()) == 0 and int(counts.sum().item()) == n
        pid = (ccb.murmurhash64(keys).view(torch.int64) >> (64 - log2p)) & (P - 1)
        for p in range(P):
            c = int(counts[p].item())
            assert c == int((pid == p).sum().item())
            assert torch.equal(torch.sort(out[p * cap: p * cap + c])[0], torch.sort(keys[pid == p])[0])
