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
