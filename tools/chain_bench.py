"""C3 measurement (evidence tool): chain of J hash joins, fused kernel, thresholds sweep.
usage: python tools/chain_bench.py [J] [cf] [lhs] [rhs] [kind]"""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
pkg = importlib.import_module("chunk-compaction-in-vectorized-execution-simd_b200")
pkg.init(0)
J = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cf = int(sys.argv[2]) if len(sys.argv) > 2 else 5
lhs_n = int(sys.argv[3]) if len(sys.argv) > 3 else 20_000_000
rhs = int(sys.argv[4]) if len(sys.argv) > 4 else 2_000_000
kind = sys.argv[5] if len(sys.argv) > 5 else "chain"
T = pkg.HashTable if kind == "chain" else pkg.LPHashTable
lhs = O.gen_lhs_main(lhs_n, J, rhs)  # main.cpp:41-55 generator
cols = [torch.from_numpy(np.ascontiguousarray(lhs[:, j])).cuda() for j in range(J)]
tables = [T(rhs, cf) for _ in range(J)]  # J separate tables, like main.cpp:62-63
want = O.multiplicity_oracle([O.build_keys(rhs, cf)] * J, lhs)
print(f"# J={J} cf={cf} lhs={lhs_n} rhs={rhs} {kind}: oracle n_tuples={want['n_tuples']} probe_tuples={want['probe_tuples']}")
res = torch.zeros(512 // 8 * 2, dtype=torch.int64, device="cuda")
for name, thr in [("full (512)", None), ("256", [256] * J), ("64", [64] * J), ("none (0)", [0] * J)]:
    best = 1e9
    for _ in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = pkg.chain_execute(tables, cols, thresholds=thr, sync=False)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    d = pkg.parse_chain_result(r["result_tensor"], J)
    assert (d["n_tuples"], d["digest"], d["colsum"]) == (want["n_tuples"], want["digest"], want["colsum"]), name
    dens = [d["level_lanes"][l] / max(1, d["level_steps"][l]) for l in range(J)]
    print(f"threshold {name:10s}: {best:8.3f} ms  {d['probe_tuples'] / best / 1e6:7.2f} G probe tuples/s  {lhs_n / best / 1e6:6.2f} G LHS rows/s  "
          f"HBM {lhs_n * 8 * J / best / 1e6:6.1f} GB/s  steps/level {d['level_steps']}  lanes/step {[round(x) for x in dens]}  device_ns {d['device_ns']}")
