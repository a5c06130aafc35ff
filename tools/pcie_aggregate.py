"""Host-side ceiling of the multi-GPU end-to-end leg (evidence tool): every rank copies 1 GiB pinned host -> device and 1 GiB device
-> pinned host at the same time, all ranks together; prints per-rank and aggregate GB/s.  Run under torchrun."""
import os, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 27
h_in, h_out = torch.empty(n, dtype=torch.int64).pin_memory(), torch.empty(n, dtype=torch.int64).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.int64, device="cuda"), torch.ones(n, dtype=torch.int64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = []
for mode in ("h2d", "d2h", "both"):
    best = 1e9
    for _ in range(4):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(); dist.barrier()
        best = min(best, time.perf_counter() - t0)
    gb = (2 if mode == "both" else 1) * n * 8 / 1e9
    res.append((mode, gb / best))
if rank == 0:
    print(f"# {world} ranks, 1 GiB per direction and rank, pinned host memory, all ranks at once (time incl. the barrier)")
    for mode, r in res:
        print(f"{mode:5s}: {r:6.1f} GB/s per rank, {r * world:7.1f} GB/s aggregate")
dist.destroy_process_group()
