"""Development aid: aggregate pinned-memory copy bandwidth of the box (torchrun, one rank per GPU).
Device-to-host and host-to-device with 1, 2, 4, ... ranks copying at once, and device-to-host by a kernel that
stores straight into mapped pinned host memory instead of the copy engines."""
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


def run(kind, active, reps=4):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    if rank < active:
        for _ in range(reps):
            if kind in ("d2h", "both"):
                with torch.cuda.stream(s1):
                    host.copy_(dev, non_blocking=True)
            if kind in ("h2d", "both"):
                with torch.cuda.stream(s2):
                    dev.copy_(host, non_blocking=True)
            if kind == "d2h_2streams":          # the same bytes as two halves on two streams
                with torch.cuda.stream(s1):
                    host[: n // 2].copy_(dev[: n // 2], non_blocking=True)
                with torch.cuda.stream(s2):
                    host[n // 2:].copy_(dev[n // 2:], non_blocking=True)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t], device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    gb = reps * n / 1e9 * (2 if kind == "both" else 1)
    if rank == 0:
        print(f"{kind:13s} {active} GPU(s) active: {gb / dt.item():7.1f} GB/s per GPU, {active * gb / dt.item():7.1f} GB/s aggregate", flush=True)


k = 1
while k <= world:
    for kind in ("d2h", "h2d"):
        run(kind, k, 1)
        run(kind, k)
    k *= 2
run("d2h_2streams", world)
run("both", world)
if rank == 0:
    print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
    os.system("lscpu | grep -E 'Model name|Socket|NUMA node|Core|Thread' ; numactl -H 2>/dev/null | head -8; nvidia-smi topo -m 2>/dev/null | head -12; "
              "nvidia-smi --query-gpu=index,pcie.link.gen.current,pcie.link.width.current --format=csv 2>/dev/null | head -10; free -g | head -2")
if world > 1:
    dist.destroy_process_group()
