"""NVLink exchange ceilings (evidence tool, run under torchrun): every rank sends one block to every peer at the same time --
(a) copy engines (cudaMemcpyAsync into IPC-mapped peer memory) on 1 / 4 / 8 streams, GPU otherwise idle;
(b) the same while a memory-bound kernel (the owner partition of 2^28 keys, back to back) keeps the SMs busy;
(c) SM-driven copies (cc_peer_copy_sm) with 8 .. 296 CTAs.
Prints GB/s sent per GPU (each GPU receives as much)."""
import ctypes as C, importlib, os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
pkg = importlib.import_module("chunk-compaction-in-vectorized-execution-simd_b200")
pkg.init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib, L = pkg.lib(), pkg._lib
block = 128 << 20  # bytes per (sender, receiver) pair
ptr = C.c_void_p()
L.check(lib.cc_malloc(C.byref(ptr), block * world))
handle = (C.c_ubyte * 64)()
L.check(lib.cc_ipc_export(ptr, handle))
handles = [None] * world
dist.all_gather_object(handles, bytes(handle))
peers = []
for r, h in enumerate(handles):
    if r == rank:
        peers.append(ptr.value)
    else:
        q = C.c_void_p()
        L.check(lib.cc_ipc_open((C.c_ubyte * 64).from_buffer_copy(h), C.byref(q)))
        peers.append(q.value)
src = torch.ones(block * world // 8, dtype=torch.int64, device="cuda")
streams = [torch.cuda.Stream() for _ in range(8)]
keys = pkg.gen_keys_counter(1 << 28, 3, (1 << 40) - 1)
dests = [(rank + i) % world for i in range(1, world)]

def all_to_all(mode, n_streams=4, blocks=0):
    for j, p in enumerate(dests):
        s = streams[j % n_streams]
        if mode == "ce":
            L.check(lib.cc_memcpy_d2d(peers[p] + rank * block, src.data_ptr() + p * block, block, s.cuda_stream))
        else:
            L.check(lib.cc_peer_copy_sm(peers[p] + rank * block, src.data_ptr() + p * block, block, max(1, blocks // len(dests)), s.cuda_stream))

def measure(tag, fn, busy=False, reps=5):
    best = 1e9
    for _ in range(reps):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        if busy:
            for _ in range(6):
                pkg.partition_single(keys, 3, ((1 << 25) * 9 // 8 + 8192 + 4095) // 4096 * 4096)
        fn()
        for s in streams: s.synchronize()
        t_copy = time.perf_counter() - t0
        torch.cuda.synchronize(); dist.barrier()
        best = min(best, t_copy)
    if rank == 0:
        print(f"{tag:58s}: {len(dests) * block / best / 1e9:7.1f} GB/s sent per GPU ({best * 1e3:6.2f} ms)", flush=True)

if rank == 0:
    print(f"# {world} GPUs, {block >> 20} MiB per (sender, receiver) pair, all pairs at once")
for ns in (1, 4, 8):
    measure(f"copy engines, {ns} stream(s), idle GPU", lambda ns=ns: all_to_all("ce", ns))
measure("copy engines, 4 streams, partition kernels running", lambda: all_to_all("ce", 4), busy=True)
for g in (8, 16, 32, 64, 148, 296):
    measure(f"SM copy kernels, {g} CTAs in total, idle GPU", lambda g=g: all_to_all("sm", 8, g))
measure("SM copy kernels, 32 CTAs, partition kernels running", lambda: all_to_all("sm", 8, 32), busy=True)
torch.cuda.synchronize(); dist.barrier()
dist.destroy_process_group()
