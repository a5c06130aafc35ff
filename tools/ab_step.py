"""A/B tool: per-kernel times of the C4 step (scatter + probe) for the library CCB_LIB_PATH selects and the env knobs set.
usage: [CCB_LIB_PATH=...] [CCB_...=...] python tools/ab_step.py [tag] [log2_build=28] [log2_probe=31]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("chunk-compaction-in-vectorized-execution-simd_b200")
pkg.init(0)
tag = sys.argv[1] if len(sys.argv) > 1 else "default"
lb = int(sys.argv[2]) if len(sys.argv) > 2 else 28
lp = int(sys.argv[3]) if len(sys.argv) > 3 else 31
n, npb = 1 << lb, 1 << lp
tab = pkg.LPHashTable(n, 1)
keys = pkg.gen_keys_counter(npb, 2, n - 1)
want = int(keys.sum().item()) & ((1 << 64) - 1)
ok = torch.empty(npb, dtype=torch.int64, device="cuda")
op = torch.empty(npb, dtype=torch.int64, device="cuda")
res = torch.zeros(4, dtype=torch.int64, device="cuda")
pkg.set_probe_profiling(True)
rows = []
for _ in range(6):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    tab.probe_batch(keys, capacity=npb, out_key=ok, out_payload=op, result=res, sync=False)
    b.record()
    torch.cuda.synchronize()
    rows.append((a.elapsed_time(b), pkg.probe_last_phase_ms()))
r = res.cpu().numpy()
nocheck = os.environ.get("AB_NOCHECK") == "1"
assert nocheck or int(r[0]) == npb and (int(r[1]) & ((1 << 64) - 1)) == want and int(r[3]) == 0, r
rows = sorted(rows[2:])
t, ph = rows[len(rows) // 2]
print(f"{tag:40s} total {t:6.2f} ms  scatter {ph[1]:5.2f}  probe {ph[2]:5.2f}   ({npb / t / 1e6:5.1f} G tuples/s)", flush=True)
