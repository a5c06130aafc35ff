"""Generates the golden fixtures in this directory from the REAL reference classes.

Run in the build container (needs /root/reference and g++):
    make -C oracle ref && python tests/golden/make_golden.py
Every fixture is produced by oracle/_ref/ref_driver, which links the reference's own
base.cpp / chaining_ht.cpp / linear_probing_ht.cpp / compactor.cpp (see oracle/Makefile).
Inputs that the reference cannot generate itself (explicit key files) are produced here
with numpy's PCG64 at fixed seeds and stored alongside the outputs.
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DRV = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


def run(*args):
    out = subprocess.check_output([DRV] + [str(a) for a in args])
    return json.loads(out.decode().strip().splitlines()[-1])


def main():
    golden = {"pipeline_main": [], "pipeline_explicit": [], "nextdump": []}
    # 1. main.cpp pipelines (mt19937(2) LHS, chaining table), no compaction and fixed full compaction
    for (J, cf, lhs, rhs, block) in [(1, 1, 100000, 1000, 256), (2, 1, 200000, 20000, 256), (2, 2, 10000, 1000, 256),
                                     (3, 5, 300000, 20000, 256), (4, 8, 300000, 20000, 256), (4, 1, 300000, 20000, 256),
                                     (4, 5, 1000000, 100000, 256), (3, 5, 50000, 5000, 2048), (2, 3, 20000, 2000, 2048)]:
        for compact in (0, 1):
            r = run("main", J, cf, lhs, rhs, block, compact)
            r.pop("seconds")
            golden["pipeline_main"].append(dict(J=J, cf=cf, lhs=lhs, rhs=rhs, block=block, compact=compact, **r))
    # 2. explicit LHS (numpy PCG64) through both table kinds, with dumped result tuples
    rng = np.random.Generator(np.random.PCG64(20240607))
    for name, (rows, J, cf, rhs, block) in {"e1": (3000, 2, 2, 500, 256), "e2": (5000, 3, 4, 800, 2048), "e3": (700, 1, 1, 64, 256)}.items():
        lhs = rng.integers(0, rhs + rhs // 4, size=(rows, J), dtype=np.int64)
        lhs_path = os.path.join(HERE, f"{name}_lhs.bin")
        lhs.tofile(lhs_path)
        for kind in (0, 1):
            dump = os.path.join(HERE, f"{name}_k{kind}_tuples.bin")
            r = run("pipe", lhs_path, rows, J, cf, rhs, block, kind, 0, dump)
            r.pop("seconds")
            t = np.fromfile(dump, dtype=np.int64).reshape(-1, 3 * J)
            t = t[np.lexsort(t.T[::-1])]  # canonical order: sorted tuples
            t.tofile(dump)
            golden["pipeline_explicit"].append(dict(name=name, rows=rows, J=J, cf=cf, rhs=rhs, block=block, kind=kind, **r))
    # 3. per-Next records of the chunk-granular protocol (both kinds, Next and InOneNext)
    for name, (n, cf, block, nkeys, hit) in {"n1": (128, 1, 256, 1024, 1), "n2": (1024, 4, 2048, 6000, 2), "n3": (300, 3, 256, 1500, 1),
                                             "n4": (1000, 8, 512, 3000, 4)}.items():
        keys = rng.integers(0, max(1, n * hit), size=nkeys, dtype=np.int64)
        kpath = os.path.join(HERE, f"{name}_keys.bin")
        keys.tofile(kpath)
        for kind in (0, 1):
            for inone in (0, 1):
                out = os.path.join(HERE, f"{name}_k{kind}_i{inone}_next.bin")
                r = run("nextdump", kind, n, cf, block, kpath, nkeys, inone, out)
                golden["nextdump"].append(dict(name=name, n=n, cf=cf, block=block, nkeys=nkeys, kind=kind, inone=inone, **r))
    # 4. the reference CompactTuner / MultiArmedBandit under a deterministic reward stream
    golden["bandit"] = dict(steps=3000, **run("bandit", 3000))
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(golden, f, indent=1)
    print("wrote", len(golden["pipeline_main"]), len(golden["pipeline_explicit"]), len(golden["nextdump"]))


if __name__ == "__main__":
    sys.exit(main())
