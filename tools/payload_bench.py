#!/usr/bin/env python
"""Payload-carrying probe at the C4 shape: LP table over 2^LB build keys with NC payload columns,
2^LP counter-generated probe keys (hit = 1), dense output (probe key, build key, payloads).  Prints one JSON line per
variant with CUDA-event times and the algorithmic bytes:  python tools/payload_bench.py [LB=28] [LP=31] [NC=1]"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("chunk-compaction-in-vectorized-execution-simd_b200")


def timed(fn, warm=2, reps=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / reps


def main():
    lb = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    lp = int(sys.argv[2]) if len(sys.argv) > 2 else 31
    nc = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    pkg.init(0)
    peak = 6547.5
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    n, npr = 1 << lb, 1 << lp
    bk = torch.arange(n, dtype=torch.int64, device="cuda")
    pay = [bk * (3 + c) + 1 + c for c in range(nc)]
    tab = pkg.LPHashTable(keys=bk, payload=pay)
    del pay
    keys = pkg.gen_keys_counter(npr, 2, n - 1)
    want_sum = int(keys.sum().item())
    out_key = torch.empty(npr, dtype=torch.int64, device="cuda")
    out_build = torch.empty(npr, dtype=torch.int64, device="cuda")
    out_cols = [torch.empty(npr, dtype=torch.int64, device="cuda") for _ in range(nc)]
    import ctypes as C
    L = pkg._lib
    lib = pkg.lib()
    res = torch.zeros(8, dtype=torch.int64, device="cuda")
    arr = (C.c_void_p * nc)(*[c.data_ptr() for c in out_cols])
    st = torch.cuda.current_stream().cuda_stream

    def probe_pay():
        L.check(lib.cc_probe_batch_payload(tab._h, keys.data_ptr(), npr, out_key.data_ptr(), out_build.data_ptr(), arr, nc, None, npr, res.data_ptr(), st))

    def probe_key_only():
        tab.probe_batch(keys, capacity=npr, out_key=out_key, out_payload=out_build, result=res[:4], sync=False)

    info = tab.info()
    for name, fn, bytes_per in (("key-only (lean kernel, [k, k])", probe_key_only, 8 + 16),
                                (f"payload x{nc} (generic kernel, [k, k, p...])", probe_pay, 8 + 16 + 8 * nc)):
        ms = timed(fn)
        r = res.cpu().numpy()
        M = (1 << 64) - 1
        assert int(r[0]) == npr and (int(r[1]) & M) == (want_sum & M), (name, r)
        if fn is probe_pay:
            for c in range(nc):
                assert (int(r[4 + c]) & M) == (((3 + c) * want_sum + (1 + c) * npr) & M)
            assert torch.equal(out_cols[0][:1 << 20], out_key[:1 << 20] * 3 + 1)
        table_bytes = int(info.n_slots) * 8 * (1 + (nc if fn is probe_pay else 0))
        algo = bytes_per * npr + 16 * npr + table_bytes  # + partition pass (8 read + 8 written), table streamed once
        print(json.dumps({"variant": name, "log2_build": lb, "log2_probe": lp, "ms": ms, "G_tuples_per_s": npr / ms / 1e6,
                          "algorithmic_GB": algo / 1e9, "achieved_GBps": algo / ms / 1e6, "frac_of_measured_peak": algo / ms / 1e6 / peak,
                          "table_bytes": table_bytes}))


if __name__ == "__main__":
    main()
