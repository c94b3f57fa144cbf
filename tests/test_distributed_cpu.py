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

    def __init__(self, n_local, n_total, k, kc, P, buffers=None):
        self.torch = torch
        self.n_local, self.n_total, self.k, self.kc, self.P = n_local, n_total, k, kc, P
        f64 = torch.float64
        if buffers is None:
            buffers = {"x": np.zeros((kc, n_total)), "scols": np.zeros((kc, n_total)), "srows": np.zeros((k, n_local))}
        self.x_cols = torch.from_numpy(buffers["x"])
        self.y_cols = self.x_cols
        self.scores_cols = torch.from_numpy(buffers["scols"])
        self.scores_rows = torch.from_numpy(buffers["srows"])
        self.gram = torch.zeros(k * k, dtype=f64)
        self.colsum = torch.zeros(k, dtype=f64)
        self.sortedX = None
        self._status = 0

    def begin(self):
        self._status = 0

    def _chunks(self, on_chunk, first_chunk):
        if on_chunk is not None:  # the row-chunk hook of the CUDA stages: every chunk once, first_chunk first
            world = self.n_total // self.n_local
            for i in range(world):
                on_chunk((first_chunk + i) % world)

    def rank_scores(self, ci=0, nci=None, on_chunk=None, first_chunk=0):
        from scipy.special import ndtri
        nci = self.kc - ci if nci is None else nci
        x = self.x_cols.numpy()
        if self.sortedX is None:
            self.sortedX = np.empty_like(x)
        for c in range(ci, ci + nci):
            self.sortedX[c] = np.sort(x[c])
            r, _ = oic.average_ranks(x[c])
            self.scores_cols[c] = torch.from_numpy(ndtri(r / (self.n_total + 1)))
        if nci > 0:
            self._chunks(on_chunk, first_chunk)

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

    def rank_gather(self, ci=0, nci=None, on_chunk=None, first_chunk=0):
        nci = self.kc - ci if nci is None else nci
        corr = self.scores_cols.numpy()
        for c in range(ci, ci + nci):
            self.y_cols[c] = torch.from_numpy(self.sortedX[c][oic.midpoint_index(corr[c])])
        if nci > 0:
            self._chunks(on_chunk, first_chunk)

    def status(self):
        return self._status


class SharedMemoryTransport:
    """CPU stand-in for CudaPeerTransport: every rank's "x" / "scols" / "srows" buffers live in POSIX
    shared memory that the peers map (CUDA IPC's role), copies are NumPy slice assignments done at once
    (the copy engines' role), the barrier is gloo's.  Same interface, same call sequence."""

    NAMES = ("x", "scols", "srows")

    def __init__(self, dist_, tag, shapes):
        from multiprocessing import shared_memory
        self.dist, self.rank, self.world = dist_, dist_.get_rank(), dist_.get_world_size()
        self._own, self._peers, self.arr = [], [], {}
        for name in self.NAMES:
            size = max(8, int(np.prod(shapes[self.rank][name])) * 8)
            self._own.append(shared_memory.SharedMemory(create=True, size=size, name=f"{tag}_{self.rank}_{name}"))
        dist_.barrier()
        for name in self.NAMES:
            self.arr[name] = []
            for g in range(self.world):
                shm = shared_memory.SharedMemory(name=f"{tag}_{g}_{name}")
                self._peers.append(shm)
                self.arr[name].append(np.ndarray(shapes[g][name], dtype=np.float64, buffer=shm.buf))
        for name in self.NAMES:
            self.arr[name][self.rank][...] = 0.0
        dist_.barrier()

    def local(self):
        return {name: self.arr[name][self.rank] for name in self.NAMES}

    def _view(self, loc, count):
        rank, what, off = loc
        flat = what.numpy().reshape(-1) if rank is None else self.arr[what][rank].reshape(-1)
        return flat[off:off + count]

    def _copies(self, copies):
        for dst, src, count in copies:
            self._view(dst, count)[:] = self._view(src, count)

    def compute_event(self):
        return None

    def push(self, copies, after, barrier=True, also_after=()):
        self._copies(copies)
        if barrier:
            self.dist.barrier()

    def pull(self, copies, after):
        self.dist.barrier()
        self._copies(copies)

    def wait(self, ev):
        pass

    def close(self):
        self.dist.barrier()
        self.arr = {}
        for shm in self._peers:
            shm.close()
        self.dist.barrier()
        for shm in self._own:
            shm.close()
            shm.unlink()


def _peer_worker(rank, world, port, n_local, k, seed, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(seed)
        n_total = n_local * world
        X = rng.normal(size=(n_total, k))
        X[:, k - 1] = rng.poisson(2.0, n_total)
        C = random_target(rng, k)
        Xt = torch.from_numpy(np.ascontiguousarray(X[rank * n_local:(rank + 1) * n_local].T)).T
        Yt = torch.empty_strided(Xt.shape, Xt.stride(), dtype=Xt.dtype)
        blocks = column_blocks(k, world)
        shapes = [{"x": (b - a, n_total), "scols": (b - a, n_total), "srows": (k, n_local)} for a, b in blocks]
        tp = SharedMemoryTransport(dist, f"pblt{port}", shapes)
        kc = blocks[rank][1] - blocks[rank][0]
        stages = NumpyStages(n_local, n_total, k, kc, np.linalg.cholesky(C), buffers=tp.local())
        runner = DistributedImanConover(n_local, k, C, dist, stages=stages, transport=tp)
        for _ in range(2):  # twice: the second call finds every buffer in its end-of-call state
            runner.run(Xt, Yt)
        np.save(os.path.join(result_dir, f"y{rank}.npy"), Yt.numpy())
        if rank == 0:
            np.save(os.path.join(result_dir, "want.npy"), oic.iman_conover(X, C))
        stages.x_cols = stages.y_cols = stages.scores_cols = stages.scores_rows = None
        tp.close()
    finally:
        dist.destroy_process_group()


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


class FailingStages(NumpyStages):
    """A stage call fails (CUDA / internal error) on ONE rank only."""

    def __init__(self, *args, fail=False, **kw):
        super().__init__(*args, **kw)
        self.fail = fail
        self.failure = None

    def rank_gather(self, *args, **kw):
        super().rank_gather(*args, **kw)
        if self.fail:
            self.failure = (5, "injected failure")  # what CudaStages._note records

    def status(self):
        return 5 if self.failure else super().status()


def _failing_worker(rank, world, port, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from probabilit_b200._lib import PblError

        rng = np.random.default_rng(3)
        n_local, k = 120, 4
        X = rng.normal(size=(n_local * world, k))
        C = random_target(rng, k)
        Xt = torch.from_numpy(np.ascontiguousarray(X[rank * n_local:(rank + 1) * n_local].T)).T
        Yt = torch.empty_strided(Xt.shape, Xt.stride(), dtype=Xt.dtype)
        a, b = column_blocks(k, world)[rank]
        stages = FailingStages(n_local, n_local * world, k, b - a, np.linalg.cholesky(C), fail=(rank == 1))
        outcome = "no exception"
        try:
            DistributedImanConover(n_local, k, C, dist, stages=stages).run(Xt, Yt)
        except PblError as e:
            outcome = f"PblError: {e}"
        with open(os.path.join(result_dir, f"outcome{rank}.txt"), "w") as f:
            f.write(outcome)
    finally:
        dist.destroy_process_group()


def test_a_failure_on_one_rank_is_raised_by_all_ranks(tmp_path):
    """A CUDA / internal failure of a stage on one rank must not leave the others blocked in a collective:
    it is folded into the agreed status and every rank raises PblError after the collective."""
    world = 3
    mp.spawn(_failing_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outcomes = [(tmp_path / f"outcome{r}.txt").read_text() for r in range(world)]
    assert all(o.startswith("PblError") for o in outcomes), outcomes
    assert "injected failure" in outcomes[1] and "peer rank failed" in outcomes[0]


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


@pytest.mark.parametrize("world,n_local,k", [(2, 400, 5), (3, 150, 7), (3, 200, 2), (4, 90, 16), (8, 40, 16)])
def test_peer_copy_choreography_equals_single_process(tmp_path, world, n_local, k):
    """The peer-copy pipeline (pushes into the owners' buffers, row-chunk hook, barriers) with shared
    memory standing in for CUDA IPC: exactly the single-process result, also with uneven column blocks
    and with ranks that own no column."""
    mp.spawn(_peer_worker, args=(world, _free_port(), n_local, k, 11, str(tmp_path)), nprocs=world, join=True)
    want = np.load(tmp_path / "want.npy")
    got = np.vstack([np.load(tmp_path / f"y{r}.npy") for r in range(world)])
    np.testing.assert_array_equal(got, want)


def test_column_blocks():
    assert column_blocks(16, 8) == [(2 * g, 2 * g + 2) for g in range(8)]
    assert column_blocks(3, 2) == [(0, 2), (2, 3)]
    assert column_blocks(2, 3) == [(0, 1), (1, 2), (2, 2)]
    b = column_blocks(1024, 8)
    assert b[0] == (0, 128) and b[-1] == (896, 1024)
