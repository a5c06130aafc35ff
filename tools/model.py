#!/usr/bin/env python
"""Back-of-the-envelope models behind DESIGN.md sections 4 and 6, kept as code so that the numbers can be re-derived.

  python tools/model.py            # single-GPU L1 / HBM bounds of the C4 step + the multi-GPU step timeline for N = 2, 4, 8

Inputs are MEASURED constants of this pool's B200 (profiles/): copy bandwidth, the L2-resident gather rate, per-kernel
times of the C4 step.  The only fitted number is the effective copy-engine bandwidth over NVLink with every GPU sending
at once (`NVLINK_GBPS`), chosen so that the N = 8 "stream" timeline reproduces the measured 31.5 ms.
"""
SM, CLK = 148, 1.92e9          # SMs, SM clock under this load (bench `clocks`)
HBM = 6547.5e9                 # measured copy bandwidth, B/s (MEASURED_PEAKS.json)
GATHER = 288e9                 # measured L2-resident 8-byte gathers per second (profiles/r1_gather_bench_b200.txt)
SCATTER_MS_PER_2_31 = 9.5      # partition_scatter_kernel, 256 partitions (profiles/r1_ncu_summary.md)
OWNER_MS_PER_2_28 = 1.12       # same kernel, 8 partitions, 2^28 keys (profiles/r1_owner_partition_ballot_vs_atomics.txt)
PROBE_MS_PER_2_31 = 17.75      # probe_unique_kernel incl. one pass over an 8 GiB table
NVLINK_GBPS = 450e9            # fitted, see above (nominal 900 GB/s per direction)


def single_gpu():
    n = 2 ** 31
    print("== C4 step, one GPU (2^31 probe keys, 8 GiB table)")
    hbm_bytes = {"scatter": 16 * n, "probe": 24 * n + 8 * 2 ** 30}
    for k, b in hbm_bytes.items():
        print(f"  {k:8s} HBM bound      {b / HBM * 1e3:6.2f} ms for {b / 1e9:6.1f} GB")
    # L1 wavefronts per 1024-key tile of the probe kernel: one per gathered line, ~13 % second visits (one sector each),
    # keys 8 B x 1024 coalesced (64), parked matches written + read through shared memory (~128), stores 16 B x 1024 (128)
    per_tile = 1024 * 1.13 + 64 + 128 + 128
    tiles_per_sm = n / 1024 / SM
    print(f"  probe    L1 bound       {per_tile * tiles_per_sm / CLK * 1e3:6.2f} ms ({per_tile:.0f} wavefronts per 1024-key tile, 1 per cycle and SM)")
    print(f"  probe    gather-rate bound {n * 1.13 / GATHER * 1e3:5.2f} ms (measured gather ceiling {GATHER / 1e9:.0f} G/s)")
    print(f"  measured: scatter {SCATTER_MS_PER_2_31} ms + probe {PROBE_MS_PER_2_31} ms = {SCATTER_MS_PER_2_31 + PROBE_MS_PER_2_31:.2f} ms")


def multi_gpu(world, mode, n_sub=4, log2_probe=30, slack=1.03):
    """Event timeline of PartitionedJoin._probe_pipelined_ce: one SM stream, one copy stream (parallel.py)."""
    n = 2 ** log2_probe
    sub = n // n_sub
    t_owner = OWNER_MS_PER_2_28 * sub / 2 ** 28
    t_slice = SCATTER_MS_PER_2_31 * sub / 2 ** 31
    table_gb = 8 * 4 * 2 ** 27 / 1e9                      # per-GPU table of the C5 share: 2^27 keys, 2^29 slots
    probe_stream_ms = PROBE_MS_PER_2_31 * (24 * n + table_gb * 1e9) / (24 * 2 ** 31 + 8 * 2 ** 30)
    probe_sub_ms = PROBE_MS_PER_2_31 * (24 * sub + table_gb * 1e9) / (24 * 2 ** 31 + 8 * 2 ** 30)
    copy_ms = (world - 1) / world * sub * 8 * slack / NVLINK_GBPS * 1e3
    sm = 0.0          # SM stream clock
    ce = 0.0          # copy stream clock
    done_p, done_c = [], []
    # order on the SM stream: P0 P1 B0 L0 P2 B1 L1 ... ; copy b starts when P(b) and copy b-1 are done
    def P(b):
        nonlocal sm, ce
        sm += t_owner
        done_p.append(sm)
        ce = max(ce, sm) + copy_ms
        done_c.append(ce)
    P(0)
    for b in range(n_sub):
        if b + 1 < n_sub:
            P(b + 1)
        sm = max(sm, done_c[b])                           # barrier B(b): every rank's copies have landed
        sm += t_slice + (probe_sub_ms if mode == "batch" else 0.0)
    if mode == "stream":
        sm += probe_stream_ms
    return sm, copy_ms * n_sub, t_owner * n_sub + t_slice * n_sub + (probe_stream_ms if mode == "stream" else probe_sub_ms * n_sub)


if __name__ == "__main__":
    single_gpu()
    print("== C5 share per GPU (2^27 build keys, 2^30 probe keys per step), copy-engine exchange, 4 sub-batches")
    measured = {(2, "stream"): 21.3, (4, "stream"): 26.4, (4, "batch"): 25.3, (8, "stream"): 31.5}
    for world in (2, 4, 8):
        for mode in ("stream", "batch"):
            for slack in (1.125, 1.03):
                t, nv, smw = multi_gpu(world, mode, slack=slack)
                m = measured.get((world, mode)) if slack == 1.125 else None
                print(f"  N={world} {mode:6s} slack {slack:5.3f}: step {t:5.1f} ms  (NVLink chain {nv:5.1f} ms, SM work {smw:5.1f} ms)"
                      + (f"   measured {m} ms" if m else ""))
