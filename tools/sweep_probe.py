"""Tuning sweep for the C4 probe (evidence tool): strategy x cache mode x slice size.
usage: python tools/sweep_probe.py [log2_build] [log2_probe]"""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("chunk-compaction-in-vectorized-execution-simd_b200")
pkg.init(0)
lb = int(sys.argv[1]) if len(sys.argv) > 1 else 28
lp = int(sys.argv[2]) if len(sys.argv) > 2 else 30
kind = sys.argv[3] if len(sys.argv) > 3 else "lp"
n, npb = 1 << lb, 1 << lp
T = pkg.LPHashTable if kind == "lp" else pkg.HashTable
tab = T(n, 1)
keys = pkg.gen_keys_counter(npb, 2, n - 1)
ok = torch.empty(npb, dtype=torch.int64, device="cuda")
op = torch.empty(npb, dtype=torch.int64, device="cuda")
res = torch.zeros(4, dtype=torch.int64, device="cuda")

def run(reps=3):
    best = 1e9
    for _ in range(reps + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        tab.probe_batch(keys, capacity=npb, out_key=ok, out_payload=op, result=res, sync=False)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    assert int(res[0].item()) == npb
    return best

only = os.environ.get("SWEEP_ONLY")  # "slice_mb:mode" -> run just that partitioned config (for ncu)
if only:
    sm, mode = (int(x) for x in only.split(":"))
    pkg.set_probe_strategy(2, sm << 20)
    pkg._lib.check(pkg.lib().cc_probe_set_cache_mode(mode & 4, mode))
    print(f"partitioned mode={mode} slice={sm} MiB {run(1):.2f} ms")
    os._exit(0)
print(f"# {kind} table 2^{lb} keys, 2^{lp} probe keys; ms and G tuples/s")
for mode in (0, 2):
    pkg.set_probe_strategy(1)
    pkg.set_probe_cache_mode(mode, 3)
    ms = run()
    print(f"direct       mode={mode}              {ms:8.2f} ms  {npb / ms / 1e6:7.1f} G/s", flush=True)
for slice_mb in (4, 8, 16, 32, 64):
    for mode in (0, 2):
        pkg.set_probe_strategy(2, slice_mb << 20)
        pkg.set_probe_cache_mode(0, mode)
        ms = run()
        print(f"partitioned  mode={mode} slice={slice_mb:3d} MiB {ms:8.2f} ms  {npb / ms / 1e6:7.1f} G/s", flush=True)
print("# TMA on/off A/B (mode bit 2 disables the TMA key ring)")
for mode in (2,):
    pkg.set_probe_strategy(2, 16 << 20)
    pkg._lib.check(pkg.lib().cc_probe_set_cache_mode(mode & 4, mode))
    print(f"partitioned  mode={mode} slice= 16 MiB {run():8.2f} ms", flush=True)
