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

Two transports move the transposes:

* **peer copies** (``CudaPeerTransport``, the product path on a box of B200s): every rank maps its peers'
  column / row buffers with CUDA IPC and the COPY ENGINES write (or, for Y, read) the slices straight
  over NVLink (``pbl_peer_copy_many``) on a side stream -- no SM is taken from the sorts that run
  meanwhile, which NCCL's send/recv kernels do (measured: +10 % on every sort they overlap).  A
  one-element NCCL all-reduce on the same side stream is the cross-rank "round has landed" barrier.
* **send/recv** (``torch.distributed`` P2P ops; gloo on CPU, NCCL if ``PBL_DIST_EXCHANGE=sendrecv``).
"""
import ctypes as C
import os

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
        # allocated by the library so that the pointer is the base of a device allocation (CUDA IPC)
        self._x_ptr = C.c_void_p()
        _lib.check(self.lib.pbl_device_malloc(C.byref(self._x_ptr), max(kc, 1) * n_total * 8), "pbl_device_malloc")
        self.x_cols = torch.as_tensor(_DevBuf(self._x_ptr.value, (max(kc, 1), n_total)), device=dev)[:kc]
        self.y_cols = self.x_cols  # X's columns are dead once ranked: reuse for Y's columns

    def exportable(self):
        """Base device pointers peers may map: name -> pointer (None if this rank has no such buffer)."""
        return {"x": self._x_ptr.value if self.kc > 0 else None,
                "scols": self.sort_plan.buffer(0)[0] if self.sort_plan is not None else None,
                "srows": self.row_plan.buffer(0)[0]}

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def _note(self, status, what):
        """A CUDA / internal failure (status 4 / 5) of a stage call is REMEMBERED, not raised: raising here
        would take this rank out of the choreography while its peers block in the next collective.  It
        comes back through status(), is agreed across the ranks and is raised by all of them together."""
        if status in (_lib.STATUS_CUDA, _lib.STATUS_INTERNAL) and self.failure is None:
            self.failure = (int(status), f"{what} failed: {_lib.last_error()}")
        return status

    failure = None

    def begin(self):
        self.failure = None
        for p in (self.row_plan, self.sort_plan):
            if p is not None:
                self._note(self.lib.pbl_ic_stage_begin(p.handle, self._stream()), "pbl_ic_stage_begin")

    class _ChunkHook:
        """Scope of the row-chunk hook (pbl_ic_plan_set_chunk_hook): on_chunk(g) is called as soon as the
        rows [g * n_local, (g+1) * n_local) of the stage's output have been enqueued, rank `first` first."""

        def __init__(self, stages, on_chunk, first):
            self.st, self.on_chunk, self.first, self.error = stages, on_chunk, first, None

        def __enter__(self):
            if self.on_chunk is not None:
                def trampoline(_col, g, _user):
                    try:
                        self.on_chunk(int(g))
                    except BaseException as e:  # must not propagate through the C frame
                        self.error = self.error or e
                self.cb = _lib.CHUNK_FN(trampoline)
                _lib.check(self.st.lib.pbl_ic_plan_set_chunk_hook(
                    self.st.sort_plan.handle, self.st.n_local, self.first, self.cb, None))
            return self

        def __exit__(self, *exc):
            if self.on_chunk is not None:
                self.st.lib.pbl_ic_plan_set_chunk_hook(self.st.sort_plan.handle, 0, 0, None, None)
                if self.error is not None and exc[0] is None:
                    raise self.error
            return False

    def rank_scores(self, ci=0, nci=None, on_chunk=None, first_chunk=0):
        """x_cols -> scores_cols (+ sortedX kept inside the sort plan)"""
        nci = self.kc - ci if nci is None else nci
        if nci > 0:
            with self._ChunkHook(self, on_chunk, first_chunk):
                self._note(self.lib.pbl_ic_stage_rank_scores(
                    self.sort_plan.handle, self.x_cols.data_ptr(), 1, self.n_total, ci, nci, self._stream()),
                    "pbl_ic_stage_rank_scores")

    def gram_partial(self):  # scores_rows -> gram, colsum (local partial sums)
        self._note(self.lib.pbl_ic_stage_gram(self.row_plan.handle, self._stream()), "pbl_ic_stage_gram")

    def solve_and_transform(self):  # gram/colsum (global) -> T ; scores_rows <- scores_rows @ T
        self._note(self.lib.pbl_ic_stage_solve(self.row_plan.handle, self.n_total, self._stream()), "pbl_ic_stage_solve")
        self._note(self.lib.pbl_ic_stage_transform(self.row_plan.handle, self._stream()), "pbl_ic_stage_transform")

    def rank_gather(self, ci=0, nci=None, on_chunk=None, first_chunk=0):
        """scores_cols (correlated) -> y_cols"""
        nci = self.kc - ci if nci is None else nci
        if nci > 0:
            with self._ChunkHook(self, on_chunk, first_chunk):
                self._note(self.lib.pbl_ic_stage_rank_gather(
                    self.sort_plan.handle, self.y_cols.data_ptr(), 1, self.n_total, ci, nci, self._stream()),
                    "pbl_ic_stage_rank_gather")

    def status(self):
        st = 0
        for p in (self.row_plan, self.sort_plan):
            if p is not None:
                st = max(st, self._note(self.lib.pbl_ic_stage_status(p.handle, self._stream()), "pbl_ic_stage_status"))
        if self.failure is not None:
            st = max(st, self.failure[0])
        return st

    def close(self):
        for p in (self.row_plan, self.sort_plan):
            if p is not None:
                p.close()
        if self._x_ptr is not None and self._x_ptr.value:
            self.x_cols = self.y_cols = None
            self.lib.pbl_device_free(self._x_ptr)
            self._x_ptr = None


class CudaPeerTransport:
    """Peer copies over NVLink: CUDA IPC mappings of every rank's "x" (X / Y columns), "scols" (scores of
    the owned columns) and "srows" (scores of the owned rows) buffers, copy engines for the data, a
    one-element all-reduce as barrier, all on one side stream.  A location is (rank, name, element
    offset) or (None, tensor, element offset) for the caller's own tensors."""

    def __init__(self, stages, dist):
        self.torch = torch = stages.torch
        self.lib = stages.lib
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        mine = {}
        for name, ptr in stages.exportable().items():
            if ptr is not None:
                h = (C.c_ubyte * 64)()
                _lib.check(self.lib.pbl_ipc_export(C.c_void_p(ptr), h), "pbl_ipc_export")
                mine[name] = bytes(h)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine)
        self.ptr = {}
        self._opened = []
        for name, local in stages.exportable().items():
            table = []
            for g in range(self.world):
                if g == self.rank:
                    table.append(local)
                elif name in everyone[g]:
                    p = C.c_void_p()
                    h = (C.c_ubyte * 64).from_buffer_copy(everyone[g][name])
                    _lib.check(self.lib.pbl_ipc_open(h, C.byref(p)), "pbl_ipc_open")
                    self._opened.append(p.value)
                    table.append(p.value)
                else:
                    table.append(None)
            self.ptr[name] = table
        self.side = torch.cuda.Stream()
        self.token = torch.zeros(1, device="cuda")
        self.barrier()

    def _addr(self, loc):
        rank, what, off = loc
        base = what.data_ptr() if rank is None else self.ptr[what][rank]
        return base + 8 * off

    def barrier(self):
        self.dist.all_reduce(self.token)

    def compute_event(self):
        ev = self.torch.cuda.Event()
        ev.record()
        return ev

    def _copies(self, copies):
        # local <- local moves go through a copy KERNEL on the side stream (a copy engine moves local HBM
        # at well under 1 TB/s, a kernel at 2.5 TB/s); everything that crosses NVLink uses the engines
        def is_local(c):
            return c[0][0] in (None, self.rank) and c[1][0] in (None, self.rank)

        for dst, src, cnt in (c for c in copies if is_local(c)):
            d = self.torch.as_tensor(_DevBuf(self._addr(dst), (cnt,)), device="cuda")
            d.copy_(self.torch.as_tensor(_DevBuf(self._addr(src), (cnt,)), device="cuda"))
        copies = [c for c in copies if not is_local(c)]
        n = len(copies)
        if n:
            dst = (C.c_void_p * n)(*[self._addr(c[0]) for c in copies])
            src = (C.c_void_p * n)(*[self._addr(c[1]) for c in copies])
            nb = (C.c_uint64 * n)(*[8 * c[2] for c in copies])
            _lib.check(self.lib.pbl_peer_copy_many(n, dst, src, nb, C.c_void_p(self.side.cuda_stream)),
                       "pbl_peer_copy_many")

    def push(self, copies, after, barrier=True, also_after=()):
        """After `after` (a compute-stream event) and the events in `also_after`: copy, then (optionally)
        barrier.  Returns the event that says "every rank's copies of this call have landed"."""
        torch = self.torch
        with torch.cuda.stream(self.side):
            self.side.wait_event(after)
            for ev in also_after:
                self.side.wait_event(ev)
            self._copies(copies)
            if barrier:
                self.barrier()
            ev = torch.cuda.Event()
            ev.record()
        return ev

    def pull(self, copies, after):
        """After every rank has reached `after`: copy (reads of the peers' buffers)."""
        torch = self.torch
        with torch.cuda.stream(self.side):
            self.side.wait_event(after)
            self.barrier()
            self._copies(copies)
            ev = torch.cuda.Event()
            ev.record()
        return ev

    def wait(self, ev):
        self.torch.cuda.current_stream().wait_event(ev)

    def close(self):
        self.torch.cuda.synchronize()
        self.dist.barrier()
        for p in self._opened:
            self.lib.pbl_ipc_close(C.c_void_p(p))
        self._opened = []
        self.dist.barrier()  # nobody frees a buffer a peer still has mapped


class DistributedImanConover:
    """``ImanConover().set_target(C)(X)`` for a row-sharded X (every rank: n_local rows, K columns).

    ``X_local`` / ``Y_local`` are (n_local, K) tensors stored column-major (stride (1, n_local)).
    """

    def __init__(self, n_local, k, correlation_matrix, dist, stages=None, device=None, transport=None):
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
        if transport is None and isinstance(stages, CudaStages) and self.world > 1 \
                and os.environ.get("PBL_DIST_EXCHANGE", "peer") == "peer":
            transport = CudaPeerTransport(stages, dist)
        self.tp = transport  # None: torch.distributed send/recv rounds
        self._x_ready = self._y_done = None
        self._host = None  # device staging + copy stream of run_host
        self.trace = None  # developer aid (tools/dist_trace.py): list of (label, CUDA event) when enabled

    def _mark(self, label):
        if self.trace is not None:
            ev = self.st.torch.cuda.Event(enable_timing=True)
            ev.record()
            self.trace.append((label, ev))

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

    def _reduce_gram(self):
        """Global Gram and column sums from the ranks' partials: gathered, then summed IN RANK ORDER by
        every rank (a fixed reduction order, so the result does not depend on the
        collective's algorithm, tree or chunking and is identical on every rank and from run to run).
        k*k + k doubles per rank: latency-bound either way."""
        st, dist = self.st, self.dist
        torch_mod = st.torch
        part = torch_mod.cat([st.gram, st.colsum])
        parts = [torch_mod.empty_like(part) for _ in range(self.world)]
        dist.all_gather(parts, part)
        total = parts[0].clone()
        for g in range(1, self.world):
            total += parts[g]
        st.gram.copy_(total[:st.gram.numel()])
        st.colsum.copy_(total[st.gram.numel():])

    def _finish(self, status):
        """Agreed status -> the same outcome on every rank."""
        if status in (_lib.STATUS_CUDA, _lib.STATUS_INTERNAL):
            mine = getattr(self.st, "failure", None)
            raise _lib.PblError(mine[1] if mine else "a peer rank failed inside the multi-GPU Iman-Conover call")
        _raise_for_status(status)

    # ------------------------------------------------------------------ the transform
    def run(self, X_local, Y_local):
        """Y_local <- Iman-Conover(X) restricted to this rank's rows.  Raises like the reference."""
        st = self.st
        dist = self.dist
        Xc = X_local.T  # [K][n_local] view of the column-major storage
        Yc = Y_local.T
        assert Xc.is_contiguous() and Yc.is_contiguous(), "X_local / Y_local must be column-major"
        R = self.rounds
        if self.tp is not None:
            return self._run_peer(Xc, Yc, Y_local)
        for _attempt in range(2):
            st.begin()
            self._mark("begin")
            # 1-3: X rows -> columns, rank + score each column, scores columns -> rows, pipelined
            pending = self._round(0, Xc, st.x_cols, True)
            back = []
            for r in range(R):
                self._wait(pending)
                self._mark(f"wait x{r}")
                pending = self._round(r + 1, Xc, st.x_cols, True) if r + 1 < R else []
                if r < self.kc:
                    st.rank_scores(r, 1)
                    self._mark(f"rank_scores {r}")
                back.append(self._round(r, st.scores_rows, st.scores_cols, False))
            for reqs in back:
                self._wait(reqs)
            self._mark("wait scores back")
            st.gram_partial()                                # 4
            self._reduce_gram()
            st.solve_and_transform()                         # 5
            self._mark("gram+allreduce+transform")
            # 6-8: correlated scores rows -> columns, rank + gather, Y columns -> rows, pipelined
            pending = self._round(0, st.scores_rows, st.scores_cols, True)
            back = []
            for r in range(R):
                self._wait(pending)
                self._mark(f"wait s{r}")
                pending = self._round(r + 1, st.scores_rows, st.scores_cols, True) if r + 1 < R else []
                if r < self.kc:
                    st.rank_gather(r, 1)
                    self._mark(f"rank_gather {r}")
                back.append(self._round(r, Yc, st.y_cols, False))
            for reqs in back:
                self._wait(reqs)
            self._mark("wait y back")
            status = self._agree(st.status())
            self._mark("status")
            if status != 6:  # PBL_RETRY: some rank switched to the exact 64-bit sort; run again
                break
        self._finish(status)
        return Y_local

    # ------------------------------------------------------------------ host buffers
    def run_host(self, X_host, Y_host):
        """The same call with this rank's row shard in HOST memory (column-major (n_local, K) NumPy arrays,
        ideally page-locked): the host->device copy of the columns of round r+1 overlaps the exchange and the
        sorts of round r, and the columns of Y go back to the host round by round while the later rounds are
        still being ranked -- the pipelining the single-GPU host entry point does (pbl_iman_conover_host_f64),
        applied per exchange round.  Needs the peer-copy transport."""
        if self.tp is None:
            raise NotImplementedError("run_host needs the peer-copy transport (PBL_DIST_EXCHANGE=peer)")
        torch_mod = self.st.torch
        lib = self.st.lib
        nl, k = self.n_local, self.k
        for a in (X_host, Y_host):
            if a.shape != (nl, k) or a.dtype != np.float64 or not (a.flags.f_contiguous or k == 1):
                raise ValueError("run_host: (n_local, K) float64 column-major arrays expected")
        if self._host is None:
            dev = torch_mod.device("cuda", torch_mod.cuda.current_device())
            self._host = {"x": torch_mod.empty((k, nl), dtype=torch_mod.float64, device=dev),
                          "y": torch_mod.empty((k, nl), dtype=torch_mod.float64, device=dev),
                          "copy": torch_mod.cuda.Stream()}
        h = self._host
        copy = h["copy"]
        cs = C.c_void_p(copy.cuda_stream)
        me = self.rank
        ring = [(me + 1 + i) % self.world for i in range(self.world)]
        cols_of_round = [[self.blocks[g][0] + r for g in ring if r < self.blocks[g][1] - self.blocks[g][0]]
                         for r in range(self.rounds)]
        # all H2D copies are queued at once, in the order the rounds need them; the copy stream first waits for
        # whatever the caller queued on the compute stream (e.g. the previous call's last reads of the staging)
        copy.wait_event(self.tp.compute_event())
        ev_x = {}
        for cols in cols_of_round:
            for c in cols:
                _lib.check(lib.pbl_memcpy_h2d(C.c_void_p(h["x"].data_ptr() + c * nl * 8),
                                              C.c_void_p(X_host.ctypes.data + c * nl * 8), nl * 8, cs), "pbl_memcpy_h2d")
                ev = torch_mod.cuda.Event()
                ev.record(copy)
                ev_x[c] = ev

        def y_done(r, ev):
            copy.wait_event(ev)
            for c in cols_of_round[r]:
                _lib.check(lib.pbl_memcpy_d2h(C.c_void_p(Y_host.ctypes.data + c * nl * 8),
                                              C.c_void_p(h["y"].data_ptr() + c * nl * 8), nl * 8, cs), "pbl_memcpy_d2h")

        self._x_ready = lambda r: [ev_x[c] for c in cols_of_round[r]]
        self._y_done = y_done
        try:
            self._run_peer(h["x"], h["y"], None)
        finally:
            self._x_ready = self._y_done = None
            copy.synchronize()
        return Y_host

    # ------------------------------------------------------------------ the transform, peer copies
    def _run_peer(self, Xc, Yc, Y_local):
        """Same pipeline with the peer-copy transport.  Buffer hazards: a rank's buffers are only written by
        peers (a) after a barrier that follows the owner's last read of them, and every call ends with an
        all-reduce (the status agreement), so the next call's first pushes find all buffers free."""
        st, tp, dist = self.st, self.tp, self.dist
        nl, nt, me = self.n_local, self.n_total, self.rank
        R = self.rounds
        blocks = self.blocks

        # peers in staggered order (me+1, me+2, ..., me): at any moment every rank writes to (reads from)
        # a different peer, so no receiver's NVLink ingress is shared by several senders
        ring = [(me + 1 + i) % self.world for i in range(self.world)]

        def x_to_cols(r, src):      # every rank -> owner g: row slice of local column a_g + r
            return [((g, "x" if src is Xc else "scols", r * nt + me * nl), (None, src, (blocks[g][0] + r) * nl), nl)
                    for g in ring if r < blocks[g][1] - blocks[g][0]]

        x_ready = self._x_ready or (lambda r: ())   # host staging (run_host): events of round r's source columns
        y_done = self._y_done or (lambda r, ev: None)  # ... and "round r's columns of Y are complete"
        for _attempt in range(2):
            st.begin()
            self._mark("begin")
            # 1-3: X rows -> columns (push), rank + score each column, scores columns -> rows (push)
            # (the way back is sent row range by row range while the scatter that ends the stage is still
            # delivering the rest: chunk g = the rows of rank g, the ring successor first)
            landed = tp.push(x_to_cols(0, Xc), tp.compute_event(), also_after=x_ready(0))
            for r in range(R):
                tp.wait(landed)
                self._mark(f"wait x{r}")
                if r + 1 < R:
                    landed = tp.push(x_to_cols(r + 1, Xc), tp.compute_event(), also_after=x_ready(r + 1))
                if r < self.kc:
                    st.rank_scores(r, 1, first_chunk=ring[0], on_chunk=lambda g, r=r: tp.push(
                        [((g, "srows", (self.c0 + r) * nl), (me, "scols", r * nt + g * nl), nl)],
                        tp.compute_event(), barrier=False))
                    self._mark(f"rank_scores {r}")
            # the side stream is ordered: one barrier after the last chunk covers every round
            tp.wait(tp.push([], tp.compute_event(), barrier=True))
            self._mark("wait scores back")
            st.gram_partial()                                # 4
            self._reduce_gram()
            st.solve_and_transform()                         # 5
            self._mark("gram+allreduce+transform")
            # 6-8: correlated scores rows -> columns (push), rank + gather, Y columns -> rows (pull)
            landed = tp.push(x_to_cols(0, st.scores_rows), tp.compute_event())
            pulls = []
            for r in range(R):
                tp.wait(landed)
                self._mark(f"wait s{r}")
                if r + 1 < R:
                    landed = tp.push(x_to_cols(r + 1, st.scores_rows), tp.compute_event())
                # Y travels like the scores did: pushed, chunk by chunk, into the owners' row buffer (whose
                # column a_g + r is dead: it was sent off in round r of step 6), then copied locally into Y
                # (NVLink reads are slower than writes: 350-400 vs 530 GB/s per GPU on this box)
                if r < self.kc:
                    st.rank_gather(r, 1, first_chunk=ring[0], on_chunk=lambda g, r=r: tp.push(
                        [((g, "srows", (self.c0 + r) * nl), (me, "x", r * nt + g * nl), nl)],
                        tp.compute_event(), barrier=False))
                    self._mark(f"rank_gather {r}")
                pulls.append(tp.pull([((None, Yc, (blocks[g][0] + r) * nl), (me, "srows", (blocks[g][0] + r) * nl), nl)
                                      for g in ring if r < blocks[g][1] - blocks[g][0]], tp.compute_event()))
                y_done(r, pulls[-1])
            tp.wait(pulls[-1])
            self._mark("wait y back")
            status = self._agree(st.status())
            self._mark("status")
            if status != 6:
                break
        self._finish(status)
        return Y_local

    def _agree(self, status):
        """Every rank must take the same branch: reduce the status with MAX."""
        torch_mod = getattr(self.st, "torch", None)
        if torch_mod is None:
            import torch as torch_mod
        dev = "cuda" if torch_mod.cuda.is_available() and self.dist.get_backend() == "nccl" else "cpu"
        # a failure (4 / 5) on any rank outranks a retry request (6) of another
        failed = status in (_lib.STATUS_CUDA, _lib.STATUS_INTERNAL)
        t = torch_mod.tensor([int(status) + (10 if failed else 0)], dtype=torch_mod.int32, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        agreed = int(t.item())
        return agreed - 10 if agreed >= 10 else agreed

    def close(self):
        self._host = None
        if self.tp is not None:
            self.tp.close()
            self.tp = None
        if hasattr(self.st, "close"):
            self.st.close()
