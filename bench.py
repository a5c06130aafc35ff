#!/usr/bin/env python
"""bench.py -- probe tuples/sec of the B200 hash-join probe + compaction path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workloads (BASELINE.json `configs`):
  N = 1 : C4 -- LP table over 2^28 build keys (8 GiB, far beyond L2), one step = one probe of
          2^31 counter-generated keys (hit = 1) with dense (compacted) key+payload output.
  N > 1 : C5 share per GPU (weak scaling) -- 2^27 build keys and 2^30 probe keys per rank,
          both sides hash-partitioned and exchanged with an NCCL all-to-all; one step =
          partition + exchange + local probe of every rank's probe keys.
One JSON line on stdout (rank 0).  `--impl reference` times the reference's own CPU code
(oracle/_ref/ref_driver, else the oracle port) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
PKG_NAME = "chunk-compaction-in-vectorized-execution-simd_b200"

ALGO_BYTES_PER_TUPLE = 59.0  # 8 (key) + 32 * 1.1 (table sectors) + 16 * m (key + payload out), m = 1
METRIC = "probe_tuples_per_sec"
UNIT = "tuples/s"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arms
def have_avx512() -> bool:
    try:
        flags = open("/proc/cpuinfo").read()
        return all(x in flags for x in ("avx512f", "avx512dq", "avx512vl", "avx512bw"))
    except OSError:
        return False


def cpu_table_log2(requested: int, procs: int) -> int:
    """Largest table (log2 of its build keys, at most `requested`) that `procs` private copies of the reference's LPHashTable fit
    in host memory: its constructor holds about 112 B per build key at its peak (a vector<vector<int64>> of all build
    tuples next to the 32 B / key slot array, linear_probing_ht.cpp:7,14-25), and at most half of MemAvailable is used."""
    try:
        avail = next(int(ln.split()[1]) * 1024 for ln in open("/proc/meminfo") if ln.startswith("MemAvailable"))
    except Exception:
        avail = 32 << 30
    k = requested
    while k > 16 and procs * (112 << k) > avail // 2:
        k -= 1
    return k


def cpu_sample_desc(log2_build: int, log2_probe: int) -> str:
    return (f"LP table 2^{log2_build} keys ({32 << log2_build >> 20} MiB of slots per worker, DRAM-resident on the host), 2^{log2_probe} probe keys per step "
            f"(counter generator seed 2, hit=1), 2048-row chunks")


def cpu_sample(log2_build: int, log2_probe: int, procs: int, variants=(0,), include_single=True, reps: int = 1):
    """Times the reference's CPU probe (LP, 2048-row chunks: Probe + while(HasNext) Next, simd_micro_bench.cpp:83-116) on a bounded
    sample of the C4 workload: same key generators, a 2^log2_build-key table and 2^log2_probe probe keys per repetition.
    Every worker process builds its own table ONCE (a table shared copy-on-write by forked workers probes 2x slower per key
    here, so the private copy is the fair arm) and then runs `reps` timed repetitions.  Returns per-variant rates and, in
    "rep_rates", the tuples/s of every repetition of the fastest all-core variant."""
    import oracle_lib as O

    n, nk = 1 << log2_build, 1 << log2_probe
    keys = O.gen_keys_counter(nk, 2, n - 1)
    drv = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    out = {"sample": cpu_sample_desc(log2_build, log2_probe), "cores": procs, "runs": {}, "rep_rates": []}
    if os.path.exists(drv) and have_avx512():
        out["kind"] = "reference"
        with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
            keys.tofile(f)
            path = f.name
        try:
            names = {0: "scalar Probe+Next", 1: "AVX-512 SIMDProbe+SIMDNext", 2: "scalar InOneNext", 3: "AVX-512 SIMDInOneNext"}
            best = 0.0
            for v in variants:
                for p in (sorted({1, procs}) if include_single else [procs]):
                    nk_run = nk if p > 1 else min(nk, 1 << 24)  # the one-process figure is a per-core rate: a smaller slice does
                    r = json.loads(subprocess.check_output([drv, "micro", "0", str(v), str(n), "1", "2048", path, str(nk_run), str(p), "0",
                                                            str(reps if p > 1 else 1)], timeout=1500).decode().strip().splitlines()[-1])
                    assert r["n_tuples"] == nk_run, r
                    rates = [nk_run / t for t in r["rep_seconds"]]
                    out["runs"][f"{names[v]} x{p}"] = statistics.mean(rates)
                    if p == procs and statistics.mean(rates) > best:
                        best, out["rep_rates"] = statistics.mean(rates), rates
        finally:
            os.unlink(path)
    else:
        out["kind"] = "port"
        out["cores"] = 1
        tab = O.OracleLP(O.build_keys(n, 1))
        for _ in range(reps):
            t0 = time.perf_counter()
            cnt, _ = O.microbench(tab, keys, 2048)
            out["rep_rates"].append(nk / (time.perf_counter() - t0))
            assert cnt == nk
        out["runs"]["oracle port scalar Probe+Next x1"] = statistics.mean(out["rep_rates"])
    out["value"] = max(out["runs"].values())
    out["unit"] = UNIT
    return out


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path on all host cores.  ONE driver invocation per
    variant builds the tables once and runs warmup + steps timed repetitions; a step = one repetition = 2^cpu_log2_probe probe
    keys of the C4 generator against a 2^K-key LP table, K = the largest table every worker can hold privately (<= 2^26:
    the reference builds its 2^28-key table through ~30 GiB of per-tuple heap vectors per worker).  `config.workload` names
    exactly what was timed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    procs = os.cpu_count() or 1
    reps = args.warmup + args.steps
    k = cpu_table_log2(args.cpu_log2_build, procs)
    t0 = time.perf_counter()
    cb = cpu_sample(k, args.cpu_log2_probe, procs, variants=(0, 1), include_single=False, reps=reps)
    wall = time.perf_counter() - t0
    rates = cb["rep_rates"][args.warmup:] or cb["rep_rates"]
    value = statistics.mean(rates)
    nk = 1 << args.cpu_log2_probe
    cfg = {"workload": f"C4 sample timed on the host CPU: LP hash join probe, {cb['sample']}; the GPU arm's C4 is the same generators at 2^{args.log2_build} build / "
                       f"2^{args.log2_probe} probe keys", "table": "linear_probing", "chunk": 2048,
           "reference_of": workload_config(args)["workload"]}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(rates), "warmup": len(cb["rep_rates"]) - len(rates),
            "ms_per_step": 1e3 * statistics.mean(nk / r for r in rates), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cb["cores"], "kind": cb["kind"], "sample": cb["sample"], "runs": cb["runs"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0, "wall_s": wall}
    print(json.dumps(line))
    return 0


EXCHANGES = {"cabi": "everything behind the C ABI (cc_pjoin_*): single-pass owner partition + copy-engine block copies into IPC-mapped peer memory over NVLink, "
                     "device-side ready / consumed flags, no collective on the data path",
             "ce": "probe side: single-pass owner partition + copy-engine block copies over NVLink underneath the probe of the previous sub-batch",
             "p2p": "scatter kernel stores into peer memory over NVLink", "nccl": "NCCL all-to-all"}


def workload_config(args) -> dict:
    if args.gpus == 1:
        return {"workload": f"C4: LP hash join, 2^{args.log2_build} build keys (cf=1, {8 * 4 << args.log2_build >> 30} GiB table), "
                            f"2^{args.log2_probe} probe keys/step (counter generator seed 2, hit=1), dense key+payload output",
                "table": "linear_probing", "chunk": 1024,
                "chunk_note": "1024-row tiles are the GPU work unit of one CTA iteration (a design choice, DESIGN.md section 3); the reference's 2048-tuple "
                              "chunk (simd_micro_bench --scale 3) is the CPU arms' block size and the chunk protocol's default kBlockSize",
                "l2_policy": "inputs (16 GiB keys, 8 GiB table) far exceed the 126 MB L2; no flush needed"}
    return {"workload": f"C5 share: hash-partitioned LP join, per GPU 2^{args.log2_build} build keys and 2^{args.log2_probe} probe keys/step, "
                        f"exchange of both sides ({EXCHANGES[args.exchange]}), "
                        f"dense key+payload output (results stay sharded)",
            "table": "linear_probing", "parallelism": f"hash-partition x{args.gpus}", "exchange": args.exchange, "sub_batches": args.sub_batches,
            "ce_probe": (args.ce_probe if args.ce_probe != "auto" else ("stream" if args.gpus <= 2 else "batch")) if args.exchange == "ce" else None,
            "copy_streams": args.copy_streams if args.exchange == "ce" else (4 if args.exchange == "cabi" else None),
            "pipelined_across_steps": (not args.no_pipeline) if args.exchange == "cabi" else False,
            "l2_policy": "inputs far exceed L2; no flush needed"}


def distributed_e2e(dist, dev, *, ne, world, rank, n_sub, dense, cap, step, result, out_key, out_payload, gen_keys, iters, host_step=None):
    """End-to-end leg at N > 1 (every rank runs it): the rank's `ne` probe keys come from pinned host memory, go to the device,
    through `step(device_keys)` (partition + exchange + probe, writing result[n_sub, 4] and the output columns), and the rows the
    rank ends up owning go back to pinned host memory.  dense: the step writes ONE run of rows from row 0 (counter in
    result[0]); otherwise sub-batch b owns slice [b * cap // n_sub, ...) of the output columns and the counter result[b].
    Returns (e2e dict or None, note or None).  The host buffers are allocated first and the ranks agree on whether all of
    them succeeded, so that no rank enters the collectives of the timed loop alone; a deterministic failure or a failed
    check is reported instead of costing the whole JSON line.  (Device-agnostic: tests/test_distributed_cpu.py runs it
    under gloo on CPU tensors.)"""
    import numpy as np
    import torch

    on_gpu = dev.type == "cuda"
    sync = torch.cuda.synchronize if on_gpu else (lambda: None)
    result.zero_()
    hcap = ne + ne // 8 + (1 << 16)  # rows this rank can end up owning (hash partition: ne +- a fraction of a percent)
    hk = hok = hop = dk = None
    try:
        hk = torch.empty(ne, dtype=torch.int64, pin_memory=on_gpu)
        hok = torch.empty(hcap, dtype=torch.int64, pin_memory=on_gpu)
        hop = torch.empty(hcap, dtype=torch.int64, pin_memory=on_gpu)
        dk = torch.empty(ne, dtype=torch.int64, device=dev)
        ready = 1
    except Exception:
        ready = 0
    flag = torch.tensor([ready], dtype=torch.int64, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        return None, "pinned host buffers could not be allocated on every rank"
    try:
        hk.copy_(gen_keys(ne, 12345 + rank * ne))
        in_sum = int(hk.sum().item())
        capb = cap // n_sub

        def e2e_step():
            if host_step is not None:  # the product's own host-buffer call (PartitionedJoin.probe_host): H2D, exchange, probe and D2H overlapped
                return host_step(hk, hok, hop)
            dk.copy_(hk, non_blocking=True)
            step(dk)
            counts = [int(c) for c in result.cpu().numpy().view(np.uint64)[:, 0]]  # the host must learn the row counts: part of the cost
            off = 0
            for b, m in enumerate(counts):
                m = min(m, cap if dense else capb, hcap - off)
                if m:
                    hok[off:off + m].copy_(out_key[b * capb:b * capb + m], non_blocking=True)
                    hop[off:off + m].copy_(out_payload[b * capb:b * capb + m], non_blocking=True)
                off += m
            sync()
            return off

        e2e_step()
        ts, rows = [], 0
        for _ in range(iters):
            dist.barrier()
            sync()
            t0 = time.perf_counter()
            rows = e2e_step()
            ts.append(time.perf_counter() - t0)
        t = torch.tensor(ts, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # every step ends when its slowest rank is done
        chk = torch.tensor([rows, int(hok[:rows].sum().item()), in_sum, int((hok[:rows] != hop[:rows]).sum().item())], dtype=torch.int64, device=dev)
        dist.all_reduce(chk)
        e2e_s = float(t.mean().item())
        if not (int(chk[0]) == ne * world and int(chk[1]) == int(chk[2]) and int(chk[3]) == 0):
            return None, f"end-to-end check failed (rows, key sum out, key sum in, key != payload): {chk.tolist()}"
        return {"value": ne * world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": ne * 8 * world, "d2h_bytes_per_step": (16 * ne + 32 * n_sub) * world,
                "sample": f"2^{ne.bit_length() - 1} probe keys per GPU and call through the join's probe_host: pinned host keys -> H2D -> partition + "
                          f"exchange + probe -> D2H of the rows each rank owns into pinned host memory, the three legs overlapped chunk by chunk",
                "ms_per_step": 1e3 * e2e_s}, None
    except Exception as e:  # noqa: BLE001 -- reported in the line
        return None, f"end-to-end leg failed: {type(e).__name__}: {e}"


# ----------------------------------------------------------------------------- GPU arm
def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-build", type=int, default=None, help="build keys per GPU (default 28 at N=1, 27 at N>1)")
    ap.add_argument("--log2-probe", type=int, default=None, help="probe keys per GPU per step (default 31 at N=1, 30 at N>1)")
    ap.add_argument("--e2e-log2-probe", type=int, default=28, help="probe keys of the host-buffer end-to-end sample")
    ap.add_argument("--cpu-log2-build", type=int, default=26, help="CPU arms: build keys of the LP table (capped by host memory, see cpu_table_log2)")
    ap.add_argument("--cpu-log2-probe", type=int, default=27, help="CPU arms: probe keys per step (SURVEY 8d: a 2^27-key slice of the C4 probe side)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sub-batches", type=int, default=None,
                    help="N>1: the exchange of sub-batch b+1 overlaps the probe of sub-batch b (default 4 with --exchange ce, else 1)")
    ap.add_argument("--peer-blocks", type=int, default=0, help="N>1 with p2p: CTA cap of the NVLink-bound peer scatter (0 = all SMs)")
    ap.add_argument("--ce-probe", default="auto", choices=["auto", "stream", "batch"],
                    help="N>1 with --exchange ce: one incremental probe per step (stream), a probe per landed sub-batch (batch), "
                         "or stream up to 2 GPUs and batch beyond (auto)")
    ap.add_argument("--copy-streams", type=int, default=4, help="N>1 with --exchange ce: copy streams the block copies of one shuffle are dealt over")
    ap.add_argument("--no-chain", action="store_true", help="N=1: skip the C3 join-chain sub-record")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="N>1 with --exchange cabi: one cc_pjoin_probe per step instead of begin(t + 1) + end(t) (the exchange of the next batch under the probe of this one)")
    ap.add_argument("--exchange", default="cabi", choices=["cabi", "ce", "p2p", "nccl"],
                    help="N>1: the C-ABI join (cc_pjoin_*: fused owner x slice partition on the sender, copy engines, device-side flags; default), "
                         "the round-1 copy-engine pipeline on torch.distributed, fused peer-memory scatter kernel, or NCCL all-to-all")
    args = ap.parse_args()
    if args.impl == "ours":
        args.warmup = max(args.warmup, 3)  # timing rule: at least 3 warm-up steps
    if args.log2_build is None:
        args.log2_build = 28 if args.gpus == 1 else 27
    if args.log2_probe is None:
        args.log2_probe = 31 if args.gpus == 1 else 30
    if args.sub_batches is None:
        args.sub_batches = 4 if (args.gpus > 1 and args.exchange in ("ce", "cabi")) else 1
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    pkg = importlib.import_module(PKG_NAME)
    pkg.init(local_rank)
    saved_stdout = None
    if distributed:
        import torch.distributed as dist

        # stdout carries exactly ONE JSON line: libraries that print there (NCCL's version banner) go to stderr until then
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_build, n_probe = 1 << args.log2_build, 1 << args.log2_probe
    peak, peak_src = measured_peak()
    dev = torch.device("cuda", local_rank)

    # ---- build side (untimed, like the reference: main.cpp:58-68 is outside the timer)
    t0 = time.perf_counter()
    if not distributed:
        table = pkg.LPHashTable(n_build, 1)
        join = None
        key_space = n_build
    else:
        par = importlib.import_module(PKG_NAME + ".parallel")
        key_space = n_build * world
        local_build = torch.arange(rank * n_build, (rank + 1) * n_build, dtype=torch.int64, device=dev)  # keys 0..N*nb-1, cf=1
        cap_rows = -(-n_probe // args.sub_batches) if args.exchange == "ce" else int(n_probe * 1.05) + (1 << 20)
        if args.exchange == "cabi":
            join = par.CPartitionedJoin(pkg, pkg.CC_HT_LP, local_build, n_probe, n_sub=args.sub_batches)
            table = None
        else:
            join = par.PartitionedJoin(pkg, pkg.CC_HT_LP, local_build, plan="partition", exchange=args.exchange,
                                       capacity_rows=cap_rows, peer_blocks=args.peer_blocks, ce_probe=args.ce_probe, copy_streams=args.copy_streams)
            table = join.table
        del local_build
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    info = table.info() if table is not None else join.table_info()

    # ---- probe side resident in HBM
    keys = pkg.gen_keys_counter(n_probe, 2, key_space - 1, first=rank * n_probe)
    cap = n_probe if not distributed else int(n_probe * 1.05) + (1 << 20)
    cap -= cap % max(1, args.sub_batches)
    out_key = torch.empty(cap, dtype=torch.int64, device=dev)
    out_payload = torch.empty(cap, dtype=torch.int64, device=dev)
    n_sub = args.sub_batches if (distributed and args.exchange in ("p2p", "ce")) else 1  # result records: "cabi" accumulates into ONE
    result = torch.zeros((n_sub, 4), dtype=torch.int64, device=dev)
    recv_buf = torch.empty(cap, dtype=torch.int64, device=dev) if (distributed and args.exchange == "nccl") else None
    expected_sum = int(keys.sum().item()) & ((1 << 64) - 1)

    in_flight = []  # cabi pipeline: a batch whose exchange is under way

    def drain():
        while in_flight:
            in_flight.pop()
            join.probe_end(out_key, out_payload, result[0])

    def step(k=None):
        k = keys if k is None else k
        if not distributed:
            return table.probe_batch(k, capacity=cap, out_key=out_key, out_payload=out_payload, result=result[0], sync=False)
        if args.exchange == "cabi":
            if k is not keys or args.no_pipeline:  # (the end-to-end leg passes its own keys: one self-contained call)
                drain()
                return join.probe(k, out_key, out_payload, result[0])
            # software pipelining across steps: a step enqueues the exchange of batch t + 1 and then probes batch t, which had the
            # whole previous step to land; every step still does the full work of one batch (one partition + exchange, one probe)
            if not in_flight:
                join.probe_begin(k)
                in_flight.append(1)
            join.probe_begin(k)
            return join.probe_end(out_key, out_payload, result[0])
        if args.exchange in ("p2p", "ce"):
            return join.probe_pipelined(k, n_sub, out_key, out_payload, result)
        shuffled = join.shuffle(k, out=recv_buf)
        return table.probe_batch(shuffled, capacity=cap, out_key=out_key, out_payload=out_payload, result=result[0], sync=False)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        step()
        if w == 0 and distributed:
            # owner property, checked on the first step of every multi-GPU run: every result row this rank holds must hash to this
            # rank (murmurhash64(key) >> (64 - log2 P)) -- a key delivered to the wrong owner cannot hide behind a checksum
            torch.cuda.synchronize()
            rr = result.cpu().numpy().view(np.uint64)
            dense = n_sub == 1 or (args.exchange == "ce" and join.ce_probe == "stream")
            capb = cap // n_sub
            owned_ok = True
            for b in range(1 if dense else n_sub):
                m = min(int(rr[b, 0]), cap if dense else capb)
                rows = out_key[b * capb:b * capb + m]
                owner = (pkg.murmurhash64(rows.contiguous()) >> (64 - join.log2p)) & (world - 1)
                owned_ok = owned_ok and bool((owner == rank).all().item()) and bool((out_payload[b * capb:b * capb + m] == rows).all().item())
            t = torch.tensor([int(owned_ok)], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            assert int(t.item()) == 1, "owner property violated: a result row sits on a rank its key does not hash to"
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = pkg.launch_count()
    phase_ms = []
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_all0, t_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_all0.record()
    for a, b in ev:
        a.record()
        step()
        b.record()
    t_all1.record()
    if distributed and args.exchange == "cabi":
        drain()  # the batch still in flight (its result equals every other step's: same keys)
    barrier()
    launches = pkg.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if not distributed:  # per-kernel times: extra steps AFTER the timed region, CUDA events inside the library
        pkg.set_probe_profiling(True)
        for _ in range(3):
            step()
            phase_ms.append(pkg.probe_last_phase_ms())
        pkg.set_probe_profiling(False)
    total_ms = t_all0.elapsed_time(t_all1)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    if distributed:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())

    # ---- correctness properties at full size (hit = 1: every probe matches exactly once)
    r = result.cpu().numpy().view(np.uint64).sum(axis=0, dtype=np.uint64)  # wrapping sums over the sub-batches
    n_matches, key_sum, payload_sum, overflow = int(r[0]), int(r[1]), int(r[2]), int(r[3])
    if distributed:
        n_matches, key_sum, payload_sum = par.reduce_result(n_matches, key_sum, payload_sum, dev)
        t = torch.tensor([expected_sum - (1 << 64) if expected_sum >= (1 << 63) else expected_sum], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        expected_sum = int(t.item()) & ((1 << 64) - 1)
    if distributed and getattr(join, "copier", None) is not None:
        join.copier.check_overflow()
    assert overflow == 0, "output capacity overflow"
    assert n_matches == n_probe * world, (n_matches, n_probe * world)
    assert key_sum == expected_sum and payload_sum == expected_sum, "checksum mismatch"

    ms_per_step = total_ms / args.steps
    value = n_probe * world / (ms_per_step * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
            "data": "synthetic", "config": workload_config(args), "gpu_launches": int(launches), "clocks": clocks,
            "build_seconds": build_s, "table": {"n_keys": int(info.n_keys), "n_slots": int(info.n_slots), "bytes": int(info.bytes)},
            "checks": {"n_matches": n_matches, "key_sum_ok": True, "owner_property": True if distributed else None}}
    if distributed and hasattr(join, "build_phases"):
        line["build_phases"] = {k: round(v, 3) for k, v in join.build_phases.items()}

    if rank == 0 or not distributed:
        # roofline of the dominant kernel (probe_unique_kernel): algorithmic bytes per launch / event time
        if not distributed:
            kernel_ms = statistics.mean(step_ms)
            achieved = ALGO_BYTES_PER_TUPLE * n_probe / (kernel_ms * 1e-3) / 1e9
            ph = [statistics.mean(p[i] for p in phase_ms) for i in range(3)] if phase_ms else [0, 0, 0]
            tb = int(info.n_slots) * 8  # the LP slot array (the occupancy bitmap beside it is not read by the batch probe)
            kernels = [  # live CUDA-event time of every kernel of the step with its own algorithmic bytes
                {"kernel": "partition_count_kernel", "ms": ph[0], "algorithmic_bytes": 8 * n_probe},
                {"kernel": "partition_scatter_kernel", "ms": ph[1], "algorithmic_bytes": 16 * n_probe},
                {"kernel": "probe_unique_kernel<LP,hints,key+payload> (+ the gated no-op launches of the two-pass fallback)", "ms": ph[2],
                 "algorithmic_bytes": 24 * n_probe + tb},
            ]
            if ph[0] < 0.01:  # single-pass partition: there is no histogram pass
                kernels = kernels[1:]
            for k in kernels:
                k["achieved_GBps"] = k["algorithmic_bytes"] / (k["ms"] * 1e-3) / 1e9 if k["ms"] > 0 else None
                k["frac"] = k["achieved_GBps"] / peak if k["achieved_GBps"] else None
            hw_bytes = 16 * n_probe + 24 * n_probe + tb  # what the step really moves: scatter (8 in + 8 out) + probe (8 key + 16 out + table once)
            line["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                "hw_frac": hw_bytes / (kernel_ms * 1e-3) / 1e9 / peak, "hw_bytes_per_step": hw_bytes,
                                "hw_note": "hw_frac = bytes the partitioned step REALLY moves (44 B per tuple of pure streaming + the table once) / step time / peak: "
                                           "the hardware's share; frac = the SURVEY 8d model bytes (59 B per tuple) over the same time",
                                "traffic": load_traffic(), "kernel": "whole step = partition_scatter_kernel + probe_unique_kernel (dominant)",
                                "kernel_ms": kernel_ms, "algorithmic_bytes_per_tuple": ALGO_BYTES_PER_TUPLE, "peak_source": peak_src,
                                "note": "achieved = 59 B x probe tuples / step time (SURVEY 8d C4 model, assumes one 32 B table sector per probe); "
                                        "the partitioned strategy streams the table once instead, see kernels[] for per-kernel figures",
                                "kernels": kernels}
        else:
            line["roofline"] = {"bound": "hbm", "achieved": ALGO_BYTES_PER_TUPLE * n_probe / (ms_per_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                "frac": ALGO_BYTES_PER_TUPLE * n_probe / (ms_per_step * 1e-3) / 1e9 / peak, "traffic": None,
                                "note": "per-GPU, whole step (partition + all-to-all + probe); NVLink moves 8 B x (N-1)/N per key each way",
                                "peak_source": peak_src,
                                # the exchange against the NVLink roofline: bytes each GPU sends (= receives) per step
                                "nvlink": {"bytes_per_gpu_per_step_each_way": 8 * n_probe * (world - 1) // world,
                                           "achieved_GBps_each_way": 8 * n_probe * (world - 1) / world / (ms_per_step * 1e-3) / 1e9,
                                           "peak_GBps_each_way": 900.0,
                                           "frac": 8 * n_probe * (world - 1) / world / (ms_per_step * 1e-3) / 1e9 / 900.0,
                                           "note": "nominal NVLink 5 rate per direction; measured on this pool: copy engines 660 GB/s on an idle GPU, "
                                                   "280-375 GB/s while the partition / probe kernels run (profiles/r2_nvlink_bench_n2.txt, DESIGN.md section 6)"}}

    # ---- end to end through the host-buffer C-ABI call (H2D + probe + D2H inside the timed region)
    if not distributed:
        del out_key, out_payload, keys
        torch.cuda.empty_cache()
        if not args.no_chain:
            try:
                line["chain"] = chain_record(pkg, torch, peak)
            except Exception as e:  # noqa: BLE001 -- a sub-record never costs the headline line
                line["chain"] = {"error": f"{type(e).__name__}: {e}"}
            try:
                line["small_configs"] = small_config_records(pkg, torch, peak)
            except Exception as e:  # noqa: BLE001
                line["small_configs"] = {"error": f"{type(e).__name__}: {e}"}
    if not args.no_e2e and not distributed:
        ne = 1 << min(args.e2e_log2_probe, args.log2_probe)
        hk = torch.empty(ne, dtype=torch.int64).pin_memory()
        hk.copy_(pkg.gen_keys_counter(ne, 2, key_space - 1, first=12345).cpu())
        hok = torch.empty(ne, dtype=torch.int64).pin_memory()
        hop = torch.empty(ne, dtype=torch.int64).pin_memory()
        hk_np, hok_np, hop_np = hk.numpy(), hok.numpy(), hop.numpy()
        table.probe_batch_host(hk_np, hok_np, hop_np)
        ts = []
        for _ in range(max(3, args.steps)):
            t0 = time.perf_counter()
            re = table.probe_batch_host(hk_np, hok_np, hop_np)
            ts.append(time.perf_counter() - t0)
        assert re["n_matches"] == ne and re["overflow"] == 0
        assert np.array_equal(np.sort(hop_np[:1 << 16]), np.sort(hok_np[:1 << 16]))
        e2e_s = statistics.mean(ts)
        line["e2e"] = {"value": ne / e2e_s, "unit": UNIT, "h2d_bytes_per_step": ne * 8, "d2h_bytes_per_step": ne * 16,
                       "sample": f"2^{ne.bit_length() - 1} probe keys per call through cc_probe_batch_host (pinned host buffers, same table)",
                       "ms_per_step": 1e3 * e2e_s}
    elif distributed and args.no_e2e:
        line["e2e"] = None
    elif distributed:
        def cabi_host_step(hk, hok, hop):
            drain()  # no batch of the device-resident loop may be in flight
            return join.probe_host(hk, hok, hop)

        e2e, note = distributed_e2e(dist, dev, ne=1 << min(args.e2e_log2_probe - 1, args.log2_probe), world=world, rank=rank, n_sub=n_sub,
                                    dense=n_sub == 1 or (args.exchange == "ce" and join.ce_probe == "stream"), cap=cap, step=step,
                                    result=result, out_key=out_key, out_payload=out_payload,
                                    gen_keys=lambda n, first: pkg.gen_keys_counter(n, 2, key_space - 1, first=first),
                                    iters=max(3, args.steps),
                                    host_step=cabi_host_step if args.exchange == "cabi" else (lambda hk, hok, hop: join.probe_host(hk, hok, hop, n_sub=n_sub)))
        line["e2e"] = e2e
        if note:
            line["e2e_note"] = note

    if rank == 0 and not args.no_cpu_baseline and not distributed:
        try:
            procs = os.cpu_count() or 1
            cb = cpu_sample(cpu_table_log2(args.cpu_log2_build, procs), args.cpu_log2_probe, procs, variants=(0, 1), reps=2)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "runs")}
        except Exception as e:  # the baseline is informative; never lose the GPU number over it
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if distributed:
        os.dup2(2, 1)  # teardown chatter, if any, also goes to stderr
        dist.barrier()
        dist.destroy_process_group()
    return 0


def chain_record(pkg, torch, peak, reps=5):
    """N = 1 sub-record for the metric's "join chain" (BASELINE config 3 / SURVEY 8d C3): the reference's own example
    (main.cpp:36: --join-num 4 --lhs-size 20000000 --rhs-size 2000000) at chunk_factor 5 and 20, through cc_chain_execute (the
    fused kernel, ExecutePipeline + FlushPipelineCache of main.cpp:119-191) with full compaction, without compaction, and with the
    thresholds chosen by the negative-feedback bandits (cc_chain_execute_tuned, main.cpp:137-167).  Results are counted and
    checksummed, not materialised (flag_collect_tuples=false, setting.h:31).  Roofline: 32 B per LHS row (4 key columns), HBM."""
    J, lhs_n, rhs = 4, 20_000_000, 2_000_000
    g = torch.Generator(device="cuda")
    g.manual_seed(2)
    cols = [torch.randint(0, rhs + 1, (lhs_n,), generator=g, device="cuda", dtype=torch.int64) for _ in range(J)]  # main.cpp:42-43: uniform in [0, rhs]
    rec = {"workload": f"C3: chain of {J} separate-chaining hash joins, {lhs_n} LHS rows x {rhs} build keys per table (main.cpp:36), LHS keys uniform in "
                       f"[0, rhs] (torch.randint, seed 2), fused kernel, in-kernel compaction, results counted + checksummed",
           "algorithmic_bytes_per_lhs_row": 8 * J, "runs": []}
    for cf in (5, 20):
        tables = [pkg.HashTable(rhs, cf) for _ in range(J)]
        num_unique = -(-rhs // cf)
        step = rhs // num_unique
        hit = [((c % step) == 0) & ((c // step) < num_unique) for c in cols]  # build keys are i * step, i < num_unique, each cf times (chaining_ht.cpp:15-26)
        level_in, alive = [], torch.ones(lhs_n, dtype=torch.bool, device="cuda")
        for l in range(J):
            level_in.append(int(alive.sum().item()) * cf ** l)
            alive = alive & hit[l]
        want_tuples, want_probe = int(alive.sum().item()) * cf ** J, sum(level_in)
        del hit, alive
        policies = [("full compaction", None), ("no compaction", [0] * J)]
        for name, thr in policies:
            ts = []
            for _ in range(reps + 1):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                r = pkg.chain_execute(tables, cols, thresholds=thr, sync=False)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            d = pkg.parse_chain_result(r["result_tensor"], J)
            assert d["n_tuples"] == want_tuples and d["probe_tuples"] == want_probe and d["level_in"] == level_in, (name, cf, d, want_tuples, level_in)
            ms = statistics.mean(ts[1:])
            rec["runs"].append({"chunk_factor": cf, "policy": name, "ms": ms, "probe_tuples_per_sec": want_probe / ms * 1e3, "lhs_rows_per_sec": lhs_n / ms * 1e3,
                                "result_tuples": want_tuples, "probe_tuples": want_probe,
                                "roofline": {"bound": "hbm", "achieved": 8 * J * lhs_n / ms / 1e6, "peak": peak, "unit": "GB/s", "frac": 8 * J * lhs_n / ms / 1e6 / peak}})
        # negative-feedback policy: one bandit per join picks the threshold of every batch; timed in steady state (the bandits keep
        # their state over `reps` passes over the LHS table, the last pass is the one reported)
        # Batches must be large: a batch is one kernel, and a pipeline instance needs many chunks for its load to even out (one LHS
        # row that matches at every level fans out to chunk_factor^4 result rows, all handled by the instance that pulled it).
        tuner = pkg.CompactTuner()
        for l in range(J):
            tuner.Initialize(0x1000 + l)
        batch, passes = 1 << 21, 12  # 10 batches per pass; 120 pulls per bandit (its forced round-robin warm-up is 36)
        for i in range(passes):
            t0 = time.perf_counter()
            d = pkg.chain_execute_tuned(tables, cols, tuner, batch)
            torch.cuda.synchronize()
            wall_ms = (time.perf_counter() - t0) * 1e3
        assert d["n_tuples"] == want_tuples and d["probe_tuples"] == want_probe, ("tuned", cf, d["n_tuples"], want_tuples)
        dev_ms = d["device_ns"] / 1e6
        rec["runs"].append({"chunk_factor": cf, "policy": f"negative-feedback bandits (batches of {batch} LHS rows, pass {passes} of {passes} over the LHS table)", "ms": wall_ms,
                            "device_ms": dev_ms,
                            "probe_tuples_per_sec": want_probe / wall_ms * 1e3, "lhs_rows_per_sec": lhs_n / wall_ms * 1e3, "result_tuples": want_tuples,
                            "probe_tuples": want_probe,
                            "roofline": {"bound": "hbm", "achieved": 8 * J * lhs_n / wall_ms / 1e6, "peak": peak, "unit": "GB/s", "frac": 8 * J * lhs_n / wall_ms / 1e6 / peak}})
        del tables
    return rec


def small_config_records(pkg, torch, peak, reps=5):
    """N = 1 sub-records for BASELINE configs 1 and 2, through cc_probe_batch with dense key + payload output:
    C1 = single LP join, the reference's `simd_bench --scale 3` table (1024 build keys, 2048-tuple chunks on the CPU side) and a
    main.cpp-sized one (2 M build keys), 2^27 probe keys; C2 = separate-chaining table, 2 M build keys, chunk_factor (fanout)
    1 / 2 / 4 / 8, 20 M probe keys, every sparse match vector compacted into dense output rows.  Probe keys: counter generator
    (murmurhash64(seed + i) & mask).  Checked: every probe key k matches chunk_factor build rows iff k is a generated build key.
    Roofline: 8 B read + 16 B written per result row (the tables live in L2)."""
    recs = []

    def run(name, T, n_build, cf, n_probe, mask):
        tab = T(n_build, cf)
        keys = pkg.gen_keys_counter(n_probe, 5, mask)
        num_unique = -(-n_build // cf)
        step = n_build // num_unique
        hit = ((keys % step) == 0) & ((keys // step) < num_unique)
        copies = torch.clamp(n_build - (keys // step) * cf, max=cf)  # the last key's copies may be cut at n_build rows
        want = int((copies * hit).sum().item())
        want_sum = int((keys * copies * hit).sum().item()) & ((1 << 64) - 1)
        del hit, copies
        cap = want + 1024
        ok = torch.empty(cap, dtype=torch.int64, device="cuda")
        op = torch.empty(cap, dtype=torch.int64, device="cuda")
        res = torch.zeros(4, dtype=torch.int64, device="cuda")
        ts = []
        for _ in range(reps + 1):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            tab.probe_batch(keys, capacity=cap, out_key=ok, out_payload=op, result=res, sync=False)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        r = [int(x) & ((1 << 64) - 1) for x in res.cpu().tolist()]
        assert r[0] == want and r[1] == want_sum and r[2] == want_sum and r[3] == 0, (name, r, want, want_sum)
        ms = statistics.mean(ts[1:])
        gbs = (8 * n_probe + 16 * want) / ms / 1e6
        recs.append({"config": name, "ms": ms, "probe_tuples_per_sec": n_probe / ms * 1e3, "result_rows": want,
                     "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak}})

    run("C1(i): LP, 1024 build keys (simd_bench --scale 3), 2^27 probe keys, hit 1", pkg.LPHashTable, 1024, 1, 1 << 27, 1023)
    run("C1(i): LP, 1024 build keys (simd_bench --scale 3), 2^27 probe keys, hit 2", pkg.LPHashTable, 1024, 1, 1 << 27, 2047)
    run("C1(ii): LP, 2^21 build keys, 2^27 probe keys, hit 1", pkg.LPHashTable, 1 << 21, 1, 1 << 27, (1 << 21) - 1)
    for cf in (1, 2, 4, 8):
        run(f"C2: chaining, 2^21 build keys, fanout {cf} (hit rate 1/{cf}), 20 M probe keys, dense output", pkg.HashTable, 1 << 21, cf, 20_000_000, (1 << 21) - 1)
    return recs


def load_traffic():
    """dram bytes per launch of the probe kernel from the committed ncu capture (profiles/), if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "probe_batch_traffic.json")) as f:
            d = json.load(f)
        return d.get("dram_bytes_per_launch")
    except Exception:
        return None


if __name__ == "__main__":
    sys.exit(main())
