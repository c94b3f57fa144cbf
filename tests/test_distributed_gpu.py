"""Row-sharded Iman-Conover on >= 2 GPUs (NCCL): bit-exact against the single-process oracle."""
import os
import socket

import numpy as np
import pytest

from conftest import random_target

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n_local, k, result_dir):
    import torch
    import torch.distributed as dist

    from probabilit_b200.distributed import DistributedImanConover

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        rng = np.random.default_rng(3)
        n_total = n_local * world
        X = rng.normal(size=(n_total, k))
        X[:, 1] = rng.poisson(3.0, n_total)
        C = random_target(rng, k)
        Xl = X[rank * n_local:(rank + 1) * n_local]
        Xt = torch.from_numpy(np.ascontiguousarray(Xl.T)).cuda().T
        Yt = torch.empty_strided(Xt.shape, Xt.stride(), dtype=Xt.dtype, device=Xt.device)
        runner = DistributedImanConover(n_local, k, C, dist)
        runner.run(Xt, Yt)
        torch.cuda.synchronize()
        np.save(os.path.join(result_dir, f"y{rank}.npy"), Yt.cpu().numpy())
        if rank == 0:
            np.save(os.path.join(result_dir, "x.npy"), X)
            np.save(os.path.join(result_dir, "c.npy"), C)
        runner.close()
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n_local,k", [(40_000, 5), (300_000, 3)])
def test_two_gpus_bit_exact(tmp_path, n_local, k):
    import torch
    import torch.multiprocessing as mp

    from oracle import iman_conover as oic

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    mp.spawn(_worker, args=(world, _free_port(), n_local, k, str(tmp_path)), nprocs=world, join=True)
    X, C = np.load(tmp_path / "x.npy"), np.load(tmp_path / "c.npy")
    got = np.vstack([np.load(tmp_path / f"y{r}.npy") for r in range(world)])
    np.testing.assert_array_equal(got, oic.iman_conover(X, C))
