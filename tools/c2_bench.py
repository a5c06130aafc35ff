"""C2 measurement (evidence tool): chaining / LP table with n=2M build keys, chunk_factor 1..8 (fanout cf, hit rate 1/cf),
20M probe keys from the main.cpp generator, dense compacted key+payload output."""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
pkg = importlib.import_module("chunk-compaction-in-vectorized-execution-simd_b200")
pkg.init(0)
n, nprobe = 2_000_000, 20_000_000
lhs = O.gen_lhs_main(nprobe, 1, n)[:, 0].copy()
keys = torch.from_numpy(lhs).cuda()
cap = nprobe * 2
ok = torch.empty(cap, dtype=torch.int64, device="cuda"); op = torch.empty(cap, dtype=torch.int64, device="cuda")
res = torch.zeros(4, dtype=torch.int64, device="cuda")
for kind, T in (("chain", pkg.HashTable), ("lp", pkg.LPHashTable)):
    for cf in (1, 2, 4, 8):
        tab = T(n, cf)
        want = O.multiplicity_oracle([O.build_keys(n, cf)], lhs.reshape(-1, 1))
        best = 1e9
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); tab.probe_batch(keys, capacity=cap, out_key=ok, out_payload=op, result=res, sync=False); b.record()
            torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
        r = res.cpu().numpy().view(np.uint64)
        assert int(r[0]) == want["n_tuples"] and int(r[1]) == want["colsum"][0] and int(r[2]) == want["colsum"][2], (kind, cf)
        m = int(r[0])
        print(f"{kind:5s} cf={cf}: {best:7.3f} ms  {nprobe / best / 1e6:6.1f} G probe tuples/s  matches {m}  HBM (8 B in + 16 B x m out) {(8 * nprobe + 16 * m) / best / 1e6:7.1f} GB/s", flush=True)
