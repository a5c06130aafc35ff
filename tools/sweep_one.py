"""Slice-size sweep of the partitioned probe at full C4 size (evidence tool)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("chunk-compaction-in-vectorized-execution-simd_b200")
pkg.init(0)
lb, lp = 28, 31
n, npb = 1 << lb, 1 << lp
tab = pkg.LPHashTable(n, 1)
keys = pkg.gen_keys_counter(npb, 2, n - 1)
ok = torch.empty(npb, dtype=torch.int64, device="cuda")
op = torch.empty(npb, dtype=torch.int64, device="cuda")
res = torch.zeros(4, dtype=torch.int64, device="cuda")
pkg.set_probe_profiling(True)
import itertools
for strat, slice_mb in ((2, 32),):
    pkg.set_probe_strategy(strat, slice_mb << 20)
    best = None
    for _ in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        tab.probe_batch(keys, capacity=npb, out_key=ok, out_payload=op, result=res, sync=False)
        b.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(b)
        ph = pkg.probe_last_phase_ms()
        if best is None or t < best[0]:
            best = (t, ph)
    assert int(res[0].item()) == npb
    print(f"strategy {strat} ({'single-pass' if strat == 2 else 'two-pass'}) slice {slice_mb:4d} MiB (P={max(2, (8 << 30) // (slice_mb << 20))}): total {best[0]:6.2f} ms  count {best[1][0]:5.2f}  scatter {best[1][1]:5.2f}  probe {best[1][2]:5.2f}", flush=True)
