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
