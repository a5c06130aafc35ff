"""C1 measurement (evidence tool): single LP join, dense key+payload output.
 (i)  the reference's `simd_bench --scale 3`: 1024 build keys (4096 slots), 2^27 glibc-rand probe keys, hit 1 and 2
      (#tuples must equal the reference's 134217728 / 67114250);
 (ii) main.cpp-sized: 2 000 000 build keys (64 MiB table), 20 000 000 mt19937 probe keys."""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
pkg = importlib.import_module("chunk-compaction-in-vectorized-execution-simd_b200")
pkg.init(0)


def timed(tab, keys, cap):
    ok = torch.empty(cap, dtype=torch.int64, device="cuda"); op = torch.empty(cap, dtype=torch.int64, device="cuda")
    res = torch.zeros(4, dtype=torch.int64, device="cuda")
    best = 1e9
    for _ in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); tab.probe_batch(keys, capacity=cap, out_key=ok, out_payload=op, result=res, sync=False); b.record()
        torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best, int(res.cpu().numpy().view(np.uint64)[0])


for hit, want in ((1, 134217728), (2, 67114250)):
    n = 1024
    keys = torch.from_numpy(O.gen_keys_rand(1 << 27, n * hit - 1)).cuda()
    ms, m = timed(pkg.LPHashTable(n, 1), keys, 1 << 27)
    assert m == want, (m, want)
    print(f"C1(i)  scale 3, hit {hit}: {ms:7.3f} ms  {keys.numel() / ms / 1e6:6.1f} G probe tuples/s  #tuples {m} (reference: {want})  "
          f"HBM (8 B in + 16 B x m out) {(8 * keys.numel() + 16 * m) / ms / 1e6:7.1f} GB/s", flush=True)
n, nprobe = 2_000_000, 20_000_000
lhs = O.gen_lhs_main(nprobe, 1, n)[:, 0].copy()
want = O.multiplicity_oracle([O.build_keys(n, 1)], lhs.reshape(-1, 1))["n_tuples"]
ms, m = timed(pkg.LPHashTable(n, 1), torch.from_numpy(lhs).cuda(), nprobe)
assert m == want
print(f"C1(ii) 2M build keys, 20M probe keys: {ms:7.3f} ms  {nprobe / ms / 1e6:6.1f} G probe tuples/s  #tuples {m}  "
      f"HBM {(8 * nprobe + 16 * m) / ms / 1e6:7.1f} GB/s", flush=True)
