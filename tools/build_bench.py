"""Build-side evidence tool (north_star: "DRAM efficiency for the streaming build"): times the LP and chain table builds
over 2^log2 keys (the reference's generator, cf = 1, and shuffled arbitrary keys) -- what replaces linear_probing_ht.cpp:28-36
and chaining_ht.cpp:29-35.  Run it under ncu for the per-kernel DRAM figures (tools/gpu_call_r2a.sh).
usage: python tools/build_bench.py [log2_keys=28] [kind=both|lp|chain]"""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("chunk-compaction-in-vectorized-execution-simd_b200")
pkg.init(0)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 28
kind = sys.argv[2] if len(sys.argv) > 2 else "both"
n = 1 << lg
seq = pkg.gen_build_keys(n, 1)
# arbitrary-order keys (what a rank of the partitioned join receives): a bijective scramble of 0..n-1
shuf = (pkg.murmurhash64(seq) & ((1 << 62) - 1))
for name, T in (("lp", pkg.LPHashTable), ("chain", pkg.HashTable)):
    if kind not in ("both", name):
        continue
    for label, keys in (("sequential keys (reference generator, cf=1)", seq), ("scrambled keys", shuf)):
        ts = []
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            tab = T(keys=keys)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
            info = tab.info()
            tab.destroy()
        t = min(ts)
        print(f"{name:5s} build 2^{lg} {label:44s}: {t * 1e3:8.2f} ms  {n / t / 1e9:6.2f} G keys/s  table {info.bytes / 2**30:.2f} GiB  "
              f"({(8 * n + info.bytes) / t / 1e9:7.1f} GB/s of keys read + table written once)", flush=True)
