"""Iman-Conover over the GPUs of one box: rows sharded, columns partitioned for the sorts.

The reference is single-process; this is the B200-native scale-out of
``ImanConover.__call__`` (reference src/probabilit/correlation.py:368-425), one process per GPU,
``torch.distributed`` (NCCL over NVLink) for the plumbing:

    rank r holds rows [r*n_local, (r+1)*n_local) of X, all K columns           (row shard)
    1. all-to-all  rows -> columns : rank g receives its columns C_g at full length N = G*n_local
    2. rank_scores on C_g          : sort, tie-run average ranks, van der Waerden scores (:394-395)
    3. all-to-all  columns -> rows : every rank gets the scores of its rows, all K columns
       (1-3 and 6-8 are pipelined per local column: the exchange of column r+1 and the return of
       column r-1 overlap the sort of column r)
    4. local Gram + column sums, NCCL all-reduce of the (K*K + K) doubles   (np.corrcoef, :398)
    5. solve (replicated, K x K) and transform of the local rows            (:398-414)
    6. all-to-all  rows -> columns of the correlated scores
    7. rank_gather on C_g          : sort, midpoint index, gather of the sorted marginal (:419-423)
    8. all-to-all  columns -> rows : Y comes back row sharded like X

The per-column sorts never communicate (columns partition naturally); the only reduction is the
K x K Gram.  The choreography is independent of the compute backend: ``CudaStages`` drives the C ABI
(the product path); the CPU test-suite plugs a NumPy stand-in to exercise the exchanges on gloo.
"""
import ctypes as C

import numpy as np

from . import _lib
from .correlation import _IcPlan, _raise_for_status


def column_blocks(k, world):
    """Contiguous blocks of columns per rank (sizes differ by at most one)."""
    sizes = [k // world + (1 if g < k % world else 0) for g in range(world)]
    starts = np.concatenate([[0], np.cumsum(sizes)]).astype(int)
    return [(int(starts[g]), int(starts[g + 1])) for g in range(world)]


class _DevBuf:
    """Zero-copy torch view of a device buffer owned by the C library."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 3, "strides": None}


class CudaStages:
    """Compute backend: the C-ABI stage calls on two plans (column shard for the sorts at full
    length, row shard for Gram / solve / transform).  All buffers are [columns][rows] tensors."""

    def __init__(self, n_local, n_total, k, kc, P, device):
        import torch

        self.torch = torch
        self.lib = _lib.require_gpu()
        self.n_local, self.n_total, self.k, self.kc = n_local, n_total, k, kc
        dev = torch.device("cuda", device)
        self.row_plan = _IcPlan(n_local, k, device, rows_only=True)
        self.row_plan.set_target(P)
        self.scores_rows = torch.as_tensor(_DevBuf(self.row_plan.buffer(0)[0], (k, n_local)), device=dev)
        self.gram = torch.as_tensor(_DevBuf(self.row_plan.buffer(2)[0], (k * k,)), device=dev)
        self.colsum = torch.as_tensor(_DevBuf(self.row_plan.buffer(3)[0], (k,)), device=dev)
        self.sort_plan = None
        if kc > 0:
            self.sort_plan = _IcPlan(n_total, kc, device)
            self.sort_plan.set_target(np.eye(kc))
            self.scores_cols = torch.as_tensor(_DevBuf(self.sort_plan.buffer(0)[0], (kc, n_total)), device=dev)
        else:
            self.scores_cols = torch.empty((0, n_total), dtype=torch.float64, device=dev)
        self.x_cols = torch.empty((kc, n_total), dtype=torch.float64, device=dev)
        self.y_cols = self.x_cols  # X's columns are dead once ranked: reuse for Y's columns

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def begin(self):
        for p in (self.row_plan, self.sort_plan):
            if p is not None:
                _lib.check(self.lib.pbl_ic_stage_begin(p.handle, self._stream()))

    def rank_scores(self, ci=0, nci=None):  # x_cols -> scores_cols (+ sortedX kept inside the sort plan)
        nci = self.kc - ci if nci is None else nci
        if nci > 0:
            _lib.check(self.lib.pbl_ic_stage_rank_scores(
                self.sort_plan.handle, self.x_cols.data_ptr(), 1, self.n_total, ci, nci, self._stream()))

    def gram_partial(self):  # scores_rows -> gram, colsum (local partial sums)
        _lib.check(self.lib.pbl_ic_stage_gram(self.row_plan.handle, self._stream()))

    def solve_and_transform(self):  # gram/colsum (global) -> T ; scores_rows <- scores_rows @ T
        _lib.check(self.lib.pbl_ic_stage_solve(self.row_plan.handle, self.n_total, self._stream()))
        _lib.check(self.lib.pbl_ic_stage_transform(self.row_plan.handle, self._stream()))

    def rank_gather(self, ci=0, nci=None):  # scores_cols (correlated) -> y_cols
        nci = self.kc - ci if nci is None else nci
        if nci > 0:
            _lib.check(self.lib.pbl_ic_stage_rank_gather(
                self.sort_plan.handle, self.y_cols.data_ptr(), 1, self.n_total, ci, nci, self._stream()))

    def status(self):
        st = 0
        for p in (self.row_plan, self.sort_plan):
            if p is not None:
                st = max(st, _lib.check(self.lib.pbl_ic_stage_status(p.handle, self._stream())))
        return st

    def close(self):
        for p in (self.row_plan, self.sort_plan):
            if p is not None:
                p.close()


class DistributedImanConover:
    """``ImanConover().set_target(C)(X)`` for a row-sharded X (every rank: n_local rows, K columns).

    ``X_local`` / ``Y_local`` are (n_local, K) tensors stored column-major (stride (1, n_local)).
    """

    def __init__(self, n_local, k, correlation_matrix, dist, stages=None, device=None):
        self.dist = dist
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        self.n_local, self.k = int(n_local), int(k)
        self.n_total = self.n_local * self.world
        if self.n_total <= self.k:
            raise ValueError(f"The matrix X must have rows > columns. Got shape: {(self.n_total, self.k)}")
        self.C = np.asarray(correlation_matrix, dtype=np.float64)
        self.P = np.linalg.cholesky(self.C)
        self.blocks = column_blocks(self.k, self.world)
        self.c0, self.c1 = self.blocks[self.rank]
        self.kc = self.c1 - self.c0
        if stages is None:
            import torch
            device = torch.cuda.current_device() if device is None else device
            stages = CudaStages(self.n_local, self.n_total, self.k, self.kc, self.P, device)
        self.st = stages

    # ------------------------------------------------------------------ exchanges
    # A transpose is done in ROUNDS: in round r every rank g exchanges what belongs to its r-th local
    # column, so that the sort of local column r can start as soon as its own round has landed while
    # the later rounds are still in flight (NCCL runs them on its own stream), and the scores of
    # column r travel back while column r+1 is being sorted.  With kc columns per rank the exposed
    # communication is ~1/kc of a bulk transpose.
    def _round(self, r, rows_buf, cols_buf, to_cols):
        """Issue round r of a transpose (asynchronously); returns the requests to wait on.
        to_cols: rows_buf [K][n_local] -> cols_buf [kc][n_total]; else the reverse."""
        dist, nl = self.dist, self.n_local
        ops = []
        mine = r < self.kc
        for g, (a, b) in enumerate(self.blocks):
            theirs = r < b - a
            if g == self.rank:
                if mine:
                    if to_cols:
                        cols_buf[r, g * nl:(g + 1) * nl].copy_(rows_buf[self.c0 + r])
                    else:
                        rows_buf[self.c0 + r].copy_(cols_buf[r, g * nl:(g + 1) * nl])
                continue
            if to_cols:
                if mine:
                    ops.append(dist.P2POp(dist.irecv, cols_buf[r, g * nl:(g + 1) * nl], g))
                if theirs:
                    ops.append(dist.P2POp(dist.isend, rows_buf[a + r], g))
            else:
                if theirs:
                    ops.append(dist.P2POp(dist.irecv, rows_buf[a + r], g))
                if mine:
                    ops.append(dist.P2POp(dist.isend, cols_buf[r, g * nl:(g + 1) * nl], g))
        return dist.batch_isend_irecv(ops) if ops else []

    @staticmethod
    def _wait(reqs):
        for r in reqs:
            r.wait()

    @property
    def rounds(self):
        return max(b - a for a, b in self.blocks)

    def rows_to_cols(self, rows_buf, cols_buf):
        """Bulk form (all rounds, then wait)."""
        for r in range(self.rounds):
            self._wait(self._round(r, rows_buf, cols_buf, True))

    def cols_to_rows(self, cols_buf, rows_buf):
        for r in range(self.rounds):
            self._wait(self._round(r, cols_buf=cols_buf, rows_buf=rows_buf, to_cols=False))

    # ------------------------------------------------------------------ the transform
    def run(self, X_local, Y_local):
        """Y_local <- Iman-Conover(X) restricted to this rank's rows.  Raises like the reference."""
        st = self.st
        dist = self.dist
        Xc = X_local.T  # [K][n_local] view of the column-major storage
        Yc = Y_local.T
        assert Xc.is_contiguous() and Yc.is_contiguous(), "X_local / Y_local must be column-major"
        R = self.rounds
        for _attempt in range(2):
            st.begin()
            # 1-3: X rows -> columns, rank + score each column, scores columns -> rows, pipelined
            pending = self._round(0, Xc, st.x_cols, True)
            back = []
            for r in range(R):
                self._wait(pending)
                pending = self._round(r + 1, Xc, st.x_cols, True) if r + 1 < R else []
                if r < self.kc:
                    st.rank_scores(r, 1)
                back.append(self._round(r, st.scores_rows, st.scores_cols, False))
            for reqs in back:
                self._wait(reqs)
            st.gram_partial()                                # 4
            dist.all_reduce(st.gram)
            dist.all_reduce(st.colsum)
            st.solve_and_transform()                         # 5
            # 6-8: correlated scores rows -> columns, rank + gather, Y columns -> rows, pipelined
            pending = self._round(0, st.scores_rows, st.scores_cols, True)
            back = []
            for r in range(R):
                self._wait(pending)
                pending = self._round(r + 1, st.scores_rows, st.scores_cols, True) if r + 1 < R else []
                if r < self.kc:
                    st.rank_gather(r, 1)
                back.append(self._round(r, Yc, st.y_cols, False))
            for reqs in back:
                self._wait(reqs)
            status = self._agree(st.status())
            if status != 6:  # PBL_RETRY: some rank switched to the exact 64-bit sort; run again
                break
        _raise_for_status(status)
        return Y_local

    def _agree(self, status):
        """Every rank must take the same branch: reduce the status with MAX."""
        torch_mod = getattr(self.st, "torch", None)
        if torch_mod is None:
            import torch as torch_mod
        dev = "cuda" if torch_mod.cuda.is_available() and self.dist.get_backend() == "nccl" else "cpu"
        t = torch_mod.tensor([int(status)], dtype=torch_mod.int32, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return int(t.item())

    def close(self):
        if hasattr(self.st, "close"):
            self.st.close()
