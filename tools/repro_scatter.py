"""Stress / diagnose cc_partition_scatter_peers (evidence tool): repeats the call and reports any mismatch in detail."""
import ctypes as C, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
ccb = importlib.import_module("chunk-compaction-in-vectorized-execution-simd_b200")
ccb.init(0)
log2p, n = 3, (3 << 20) + 12345
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.Generator(np.random.PCG64(n + log2p))
keys = rng.integers(-(1 << 62), 1 << 62, size=n, dtype=np.int64)
P = 1 << log2p
dkeys = torch.from_numpy(keys).cuda()
h = ccb.murmurhash64(dkeys)
pid = ((h.view(torch.int64) >> (64 - log2p)) & (P - 1)).cpu().numpy()
counts = np.bincount(pid, minlength=P)
base = np.arange(P, dtype=np.int64) * 7 + 3
bufs = [torch.full((int(counts[p] + base[p]) + 16,), -7, dtype=torch.int64, device="cuda") for p in range(P)]
want = [torch.sort(dkeys[torch.from_numpy(pid == p).cuda()])[0] for p in range(P)]
dbase = torch.from_numpy(base).cuda()
cursors = torch.zeros(P, dtype=torch.int64, device="cuda")
ptrs = (C.c_void_p * P)(*[b.data_ptr() for b in bufs])
bad = 0
for rep in range(reps):
    for blocks in (0, 8):
        ccb._lib.check(ccb.lib().cc_partition_set_peer_blocks(blocks))
        for b in bufs:
            b.fill_(-7)
        ccb._lib.check(ccb.lib().cc_partition_scatter_peers(dkeys.data_ptr(), n, log2p, dbase.data_ptr(), cursors.data_ptr(), ptrs,
                                                            torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        for p in range(P):
            got = bufs[p][base[p]: base[p] + counts[p]]
            gs = torch.sort(got)[0]
            if not torch.equal(gs, want[p]):
                bad += 1
                nm7 = int((got == -7).sum())
                wset = want[p]
                # keys present in got but not wanted / wanted but missing
                gi = gs.cpu().numpy(); wi = wset.cpu().numpy()
                extra = np.setdiff1d(gi, wi); missing = np.setdiff1d(wi, gi)
                pos = np.nonzero(np.isin(keys, missing))[0]
                print(f"rep {rep} blocks {blocks} part {p}: stale(-7)={nm7} extra={len(extra)} missing={len(missing)} "
                      f"missing input rows (first 8)={pos[:8]} tiles={np.unique(pos // 4096)[:8]} row%4096={(pos % 4096)[:8]}", flush=True)
                if len(extra):
                    epos = np.nonzero(np.isin(keys, extra))[0]
                    print("   extra values:", extra[:4], "their input rows:", epos[:8], "tiles", np.unique(epos // 4096)[:8])
print("mismatching (rep, blocks, part) triples:", bad, "of", reps * 2 * P)
