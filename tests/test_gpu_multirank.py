"""Launches the multi-rank parity gate (tests/multirank_parity.py: sorted result tuples of the partitioned join on P GPUs ==
the oracle's pipeline over the undivided inputs, owner property on every rank, both failure paths) with torchrun on every
power-of-two rank count the box offers.  Skipped on a single-GPU box; profiles/r2_multirank_parity_p*.txt hold the logs of the
round's multi-GPU runs."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multirank_parity_gate(ccb, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, the box has {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29650 + world), os.path.join(ROOT, "tests", "multirank_parity.py"), "--quick"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    assert p.returncode == 0, p.stdout[-4000:] + p.stderr[-4000:]
    assert "ALL GREEN" in p.stdout and "FAIL" not in p.stdout, p.stdout[-4000:]


PJOIN = os.path.join(ROOT, "chunk-compaction-in-vectorized-execution-simd_b200", "host", "pjoin_main")


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("table,cf,hit", [("lp", 1, 2), ("chain", 4, 1), ("lp", 8, 1)])
def test_cpp_pjoin_driver_matches_oracle(ccb, tmp_path, world, table, cf, hit):
    """host/pjoin_main.cpp: the partitioned join run by a C++ host through the C ABI alone (cc_pjoin_*, fork + shared-memory
    control plane, no Python / NCCL on the path).  The sharded result rows it dumps, taken together, must equal the oracle's
    pipeline over the undivided inputs as sorted tuples; the driver's own checks (count, checksums, owner property) must hold.
    world = 1 runs on any box (every code path except the IPC mapping)."""
    import json

    import numpy as np

    import oracle_lib as O

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, the box has {torch.cuda.device_count()}")
    assert os.path.exists(PJOIN), "build it with make -C <pkg>/csrc driver"
    lb, lp = 16, 19
    prefix = str(tmp_path / "rows")
    # tables of 2^16 keys live in L2; CCB_PJ_SLICE_BYTES makes the join slice them anyway (fused owner x slice partition on the sender,
    # arenas probed slice by slice in two groups of pieces), once per table kind
    env = dict(os.environ)
    if cf != 8:
        env["CCB_PJ_SLICE_BYTES"] = str(32 << 10)
    else:
        env["CCB_PJ_SM_COPY_PCT"] = "25"  # a quarter of every block copy is moved by the SM copy kernel behind the partition pass
    out = subprocess.run([PJOIN, "--gpus", str(world), "--log2-build", str(lb), "--log2-probe", str(lp), "--table", table, "--chunk-factor", str(cf),
                          "--hit", str(hit), "--steps", "2", "--sub-batches", "5", "--pipeline", "1" if table == "chain" else "0", "--dump", prefix],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["checks_ok"] and r["owner_property"] and r["overflow"] == 0 and r["n_matches"] == r["expected_matches"]
    n_build, n_probe = 1 << lb, 1 << lp
    probe = O.gen_keys_counter(n_probe, 2, n_build * hit - 1)  # the driver's generator: murmurhash64(2 + i) & mask, range-partitioned over the ranks
    bk = O.build_keys(n_build, cf)
    want = O.pipeline([(O.OracleLP if table == "lp" else O.OracleChain)(bk)], probe.reshape(-1, 1), 2048, collect=True)
    got = np.concatenate([np.fromfile(f"{prefix}.{q}.bin", dtype=np.int64).reshape(-1, 2) for q in range(world)])
    assert got.shape[0] == want["n_tuples"] == r["n_matches"]
    assert np.array_equal(O.sort_rows(got), O.sort_rows(want["tuples"][:, [0, 2]]))
