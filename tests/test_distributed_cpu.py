"""The multi-GPU choreography of probabilit_b200.distributed on CPU: world_size 2 and 3 over gloo,
with a NumPy stand-in (built from the oracle) for the CUDA stage calls.  Checks that the exchanges
(rows<->columns all-to-alls, Gram all-reduce) assemble exactly the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import random_target
from oracle import iman_conover as oic
from probabilit_b200.distributed import DistributedImanConover, column_blocks


class NumpyStages:
    """Test stand-in for CudaStages (same attributes / methods), NumPy on CPU tensors."""

    def __init__(self, n_local, n_total, k, kc, P):
        self.torch = torch
        self.n_local, self.n_total, self.k, self.kc, self.P = n_local, n_total, k, kc, P
        f64 = torch.float64
        self.x_cols = torch.zeros((kc, n_total), dtype=f64)
        self.y_cols = self.x_cols
        self.scores_cols = torch.zeros((kc, n_total), dtype=f64)
        self.scores_rows = torch.zeros((k, n_local), dtype=f64)
        self.gram = torch.zeros(k * k, dtype=f64)
        self.colsum = torch.zeros(k, dtype=f64)
        self.sortedX = None
        self._status = 0

    def begin(self):
        self._status = 0

    def rank_scores(self, ci=0, nci=None):
        from scipy.special import ndtri
        nci = self.kc - ci if nci is None else nci
        x = self.x_cols.numpy()
        if self.sortedX is None:
            self.sortedX = np.empty_like(x)
        for c in range(ci, ci + nci):
            self.sortedX[c] = np.sort(x[c])
            r, _ = oic.average_ranks(x[c])
            self.scores_cols[c] = torch.from_numpy(ndtri(r / (self.n_total + 1)))

    def gram_partial(self):
        s = self.scores_rows.numpy()
        self.gram.copy_(torch.from_numpy((s @ s.T).ravel()))
        self.colsum.copy_(torch.from_numpy(s.sum(axis=1)))

    def solve_and_transform(self):
        n, k = self.n_total, self.k
        G = self.gram.numpy().reshape(k, k)
        cs = self.colsum.numpy()
        c = (G - np.outer(cs, cs) / n) / (n - 1)
        sd = np.sqrt(np.diag(c))
        R = np.clip(c / sd[:, None] / sd[None, :], -1, 1)
        try:
            _, T = oic.transform_matrix(R, self.P)
        except np.linalg.LinAlgError:
            self._status = 1
            return
        s = self.scores_rows.numpy()
        self.scores_rows.copy_(torch.from_numpy((s.T @ T).T.copy()))

    def rank_gather(self, ci=0, nci=None):
        nci = self.kc - ci if nci is None else nci
        corr = self.scores_cols.numpy()
        for c in range(ci, ci + nci):
            self.y_cols[c] = torch.from_numpy(self.sortedX[c][oic.midpoint_index(corr[c])])

    def status(self):
        return self._status


def _worker(rank, world, port, n_local, k, seed, ties, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(seed)
        n_total = n_local * world
        X = rng.normal(size=(n_total, k))
        if ties:
            X[:, 0] = rng.poisson(2.0, n_total)
        C = random_target(rng, k)
        Xl = np.asfortranarray(X[rank * n_local:(rank + 1) * n_local])
        Xt = torch.from_numpy(np.ascontiguousarray(Xl.T)).T  # (n_local, k) column-major
        Yt = torch.empty_strided(Xt.shape, Xt.stride(), dtype=Xt.dtype)
        kc = column_blocks(k, world)[rank]
        stages = NumpyStages(n_local, n_total, k, kc[1] - kc[0], np.linalg.cholesky(C))
        DistributedImanConover(n_local, k, C, dist, stages=stages).run(Xt, Yt)
        np.save(os.path.join(result_dir, f"y{rank}.npy"), Yt.numpy())
        if rank == 0:
            np.save(os.path.join(result_dir, "want.npy"), oic.iman_conover(X, C))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,n_local,k,ties", [(2, 500, 5, False), (2, 301, 4, True), (3, 200, 2, False)])
def test_rows_sharded_equals_single_process(tmp_path, world, n_local, k, ties):
    mp.spawn(_worker, args=(world, _free_port(), n_local, k, 7, ties, str(tmp_path)), nprocs=world, join=True)
    want = np.load(tmp_path / "want.npy")
    got = np.vstack([np.load(tmp_path / f"y{r}.npy") for r in range(world)])
    np.testing.assert_array_equal(got, want)


def test_column_blocks():
    assert column_blocks(16, 8) == [(2 * g, 2 * g + 2) for g in range(8)]
    assert column_blocks(3, 2) == [(0, 2), (2, 3)]
    assert column_blocks(2, 3) == [(0, 1), (1, 2), (2, 2)]
    b = column_blocks(1024, 8)
    assert b[0] == (0, 128) and b[-1] == (896, 1024)
