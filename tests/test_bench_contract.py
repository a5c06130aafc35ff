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
    # the reference arm names the sample it REALLY timed (table and probe sizes), not the GPU arm's workload string
    assert d["config"]["workload"].startswith("C4 sample timed on the host CPU") and "2^16 keys" in d["config"]["workload"] and "2^18 probe keys" in d["config"]["workload"]
    assert d["config"]["reference_of"].startswith("C4: LP hash join, 2^28 build keys")
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


def test_cpu_table_size_respects_host_memory():
    """cpu_table_log2: the largest LP table `procs` private copies of the reference's LPHashTable fit in half of the available host
    memory (its constructor peaks at ~112 B per build key), never more than requested"""
    import bench

    avail = next(int(ln.split()[1]) * 1024 for ln in open("/proc/meminfo") if ln.startswith("MemAvailable"))
    for procs in (1, 16, 64):
        k = bench.cpu_table_log2(26, procs)
        assert 16 <= k <= 26
        assert k == 16 or procs * (112 << k) <= avail // 2
        assert k == 26 or procs * (112 << (k + 1)) > avail // 2
    assert bench.cpu_table_log2(18, 1) == 18
