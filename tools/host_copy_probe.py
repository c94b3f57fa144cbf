"""Pinned-memory host <-> device copy ceiling of the box with all ranks copying at once (the floor of any
end-to-end number that starts and ends in host memory).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/host_copy_probe.py [GB per rank]
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

gb = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(gb * 1e9 / 8)
h = torch.empty(n, dtype=torch.float64).pin_memory()
h2 = torch.empty(n, dtype=torch.float64).pin_memory()
d = torch.empty(n, dtype=torch.float64, device="cuda")
d2 = torch.ones(n, dtype=torch.float64, device="cuda")
side = torch.cuda.Stream()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps=3):
    best = float("inf")
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
    return best


def both():
    with torch.cuda.stream(side):
        h2.copy_(d2, non_blocking=True)
    d.copy_(h, non_blocking=True)
    side.synchronize()


t_h2d = timed(lambda: d.copy_(h, non_blocking=True))
t_d2h = timed(lambda: h2.copy_(d2, non_blocking=True))
t_both = timed(both)
if rank == 0:
    nb = n * 8 / 1e9
    print(json.dumps({"ranks": world, "GB_per_rank": nb,
                      "h2d_GBps_per_gpu": nb / t_h2d, "h2d_GBps_aggregate": nb * world / t_h2d,
                      "d2h_GBps_per_gpu": nb / t_d2h, "d2h_GBps_aggregate": nb * world / t_d2h,
                      "both_directions_GBps_per_gpu_each_way": nb / t_both,
                      "both_directions_GBps_aggregate_each_way": nb * world / t_both,
                      "numa_nodes": len([p for p in os.listdir("/sys/devices/system/node") if p.startswith("node")])
                      if os.path.isdir("/sys/devices/system/node") else None}))
if world > 1:
    dist.destroy_process_group()
