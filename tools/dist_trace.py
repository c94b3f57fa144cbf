"""Device timeline of one multi-GPU Iman-Conover call (developer tool): CUDA events on the compute
stream around every wait / stage of DistributedImanConover.run, printed for rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29512 tools/dist_trace.py [rows_per_rank] [d]
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from probabilit_b200.distributed import DistributedImanConover  # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    X = torch.randn((d, n), generator=g, device="cuda", dtype=torch.float64).T
    Y = torch.empty_strided(X.shape, X.stride(), dtype=X.dtype, device=X.device)
    A = np.random.default_rng(0).normal(size=(2 * d, d))
    Ct = 0.9 * np.corrcoef(A, rowvar=False) + 0.1 * np.eye(d)
    runner = DistributedImanConover(n, d, Ct, dist)
    for _ in range(2):
        runner.run(X, Y)
    torch.cuda.synchronize()
    dist.barrier()
    runner.trace = []
    runner.run(X, Y)
    torch.cuda.synchronize()
    tr = runner.trace
    rows = [(tr[i][0], round(tr[i - 1][1].elapsed_time(tr[i][1]), 3)) for i in range(1, len(tr))]
    total = tr[0][1].elapsed_time(tr[-1][1])
    allr = [None] * world
    dist.all_gather_object(allr, {"rank": rank, "total_ms": total, "segments": rows})
    if rank == 0:
        print(json.dumps(allr[0], indent=1))
        print(json.dumps([{"rank": a["rank"], "total_ms": a["total_ms"],
                           "waits_ms": sum(v for k, v in a["segments"] if k.startswith("wait"))} for a in allr]))
    runner.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
