"""CPU tests of bench.py's contract: the reference arm prints ONE JSON line with the agreed keys (rank 0 only), and the
product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=600, env=e)


def test_reference_arm_prints_one_json_line():
    p = run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-log2-build", "16", "--cpu-log2-probe", "18"])
    assert p.returncode == 0, p.stderr
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "probe_tuples_per_sec" and d["unit"] == "tuples/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("C4")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "2^16" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    p = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_arm_fails_loudly_without_a_gpu():
    p = run(["--steps", "1", "--warmup", "3", "--no-cpu-baseline", "--no-e2e"])
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
