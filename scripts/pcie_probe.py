"""Development aid: aggregate pinned-memory copy bandwidth with all ranks copying at once (torchrun, one rank per GPU)."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(kind, reps=6):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        if kind in ("d2h", "both"):
            with torch.cuda.stream(s1):
                host.copy_(dev, non_blocking=True)
        if kind in ("h2d", "both"):
            with torch.cuda.stream(s2):
                dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t], device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    gb = reps * n / 1e9 * (2 if kind == "both" else 1)
    if rank == 0:
        print(f"{kind:5s}: {gb / dt.item():7.1f} GB/s per GPU, {world * gb / dt.item():7.1f} GB/s aggregate over {world} GPUs", flush=True)


for k in ("d2h", "h2d", "both"):
    run(k, 2)
    run(k)
if rank == 0:
    print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
    os.system("numactl -H 2>/dev/null | head -5; nvidia-smi topo -m 2>/dev/null | head -14")
if world > 1:
    dist.destroy_process_group()
