"""Correlators with the reference's interface, executing on the B200.

Mirrors src/probabilit/correlation.py of the reference:

* ``Correlator.set_target(C) -> self`` (:162-179) and ``_validate_X`` (:181-202) stay host NumPy
  (k x k work; same checks, same exception types and messages);
* ``ImanConover.__call__(X)`` (:368-425) runs the whole transform through the C ABI
  (``pbl_ic_plan_*`` in include/probabilit_b200.h) -- there is no CPU path in this class.

``X`` may be a NumPy array (host: copied to the device and back, like any NumPy caller of the
reference would expect) or a CUDA ``torch.Tensor`` (device resident: no host traffic; the result
is a tensor on the same device).  The class can be passed as ``correlator=`` to the modeling
graph's ``sample`` exactly like the reference's own (modeling.py:505-507, :577-581).
"""
import abc
import ctypes as C

import numpy as np

from . import _lib


class CorrelatorError(Exception):
    pass


def nearest_correlation_matrix(matrix, *, weights=None, eps=1e-6, verbose=False):
    """Correlation matrix nearest to ``matrix`` in the Frobenius norm, with smallest eigenvalue
    >= 10 * eps / K -- the problem the reference hands to cvxpy/SCS (correlation.py:59-150).

    Not on the hot path (one K x K problem per ``sample`` call; SURVEY.md section 8 keeps it on the
    host): solved here with Higham's alternating projections + Dykstra's correction (N. Higham,
    "Computing the nearest correlation matrix", 2002) in NumPy, so that ``.correlate()`` graphs
    work without cvxpy.  A matrix that is already feasible is returned unchanged (SCS returns it
    to ~1e-6).  With elementwise ``weights`` H the objective is ``||H o (X - G)||_F`` (equation (3) of
    Qi & Sun, the reference's docstring), which has no closed-form projection: it is solved by ADMM on
    the splitting X = Z (X: weighted least squares with a unit diagonal, elementwise; Z: projection on
    the eigenvalue floor), reproducing the reference's doctest values and MATLAB's ``nearcorr`` example
    (reference tests/test_correlation.py:37-78) to 1e-4."""
    if not isinstance(matrix, np.ndarray):
        raise TypeError("Input argument `matrix` must be np.ndarray.")
    if not (matrix.ndim == 2 and matrix.shape[0] == matrix.shape[1]):
        raise ValueError("Input argument `matrix` must be square.")
    if weights is not None:
        if not isinstance(weights, np.ndarray):
            raise TypeError("Input argument `weights` must be np.ndarray.")
        if weights.shape != matrix.shape:
            raise ValueError("Argument `weights` must have same shape as `matrix`.")
    K = matrix.shape[0]
    floor = 10.0 * eps / K
    G = np.array(matrix, dtype=float)
    if weights is not None and not np.allclose(weights, weights.flat[0]):
        return _nearest_correlation_matrix_weighted(G, np.array(weights, dtype=float), floor, verbose)
    sym = 0.5 * (G + G.T)
    if np.allclose(G, G.T) and np.allclose(np.diag(G), 1.0) and np.linalg.eigvalsh(sym).min() >= floor:
        return G
    Y, dS = sym.copy(), np.zeros_like(sym)
    for it in range(10000):
        R = Y - dS
        w, V = np.linalg.eigh(R)
        X = (V * np.maximum(w, floor)) @ V.T  # projection on {X : X - floor*I >= 0}
        dS = X - R
        Y_new = X.copy()
        np.fill_diagonal(Y_new, 1.0)  # projection on the unit diagonal
        done = np.linalg.norm(Y_new - Y, "fro") <= 1e-3 * eps * max(1.0, np.linalg.norm(Y_new, "fro"))
        Y = Y_new
        if verbose:
            print(f"nearest_correlation_matrix: iteration {it}, min eig {np.linalg.eigvalsh(Y).min():.3e}")
        if done:
            break
    # finish on the PSD side with an exact unit diagonal: shrink towards the identity if needed
    lam = np.linalg.eigvalsh(Y).min()
    if lam < floor:
        t = (floor - lam) / (1.0 - lam)
        Y = (1.0 - t) * Y + t * np.eye(K)
    return 0.5 * (Y + Y.T)


def _nearest_correlation_matrix_weighted(G, H, floor, verbose=False):
    """argmin ||H o (X - G)||_F  s.t.  diag(X) = 1,  X - floor*I >= 0   (reference correlation.py:124-137).

    ADMM on  f(X) + g(Z),  X = Z:   f = 0.5 ||H o (X - G)||^2 + indicator(diag X = 1)  (separable: the X-step
    is elementwise),  g = indicator(Z - floor*I >= 0)  (the Z-step clamps eigenvalues).  The penalty is
    re-balanced from the primal / dual residuals.  K x K, once per ``sample`` call: host NumPy."""
    K = G.shape[0]
    G = 0.5 * (G + G.T)
    H2 = (0.5 * (H + H.T)) ** 2
    rho = max(float(np.mean(H2)), 1e-3)
    Z = G.copy()
    np.fill_diagonal(Z, 1.0)
    U = np.zeros_like(G)
    eye = np.eye(K, dtype=bool)
    for it in range(50000):
        X = (H2 * G + rho * (Z - U)) / (H2 + rho)
        X[eye] = 1.0
        w, V = np.linalg.eigh(X + U)
        Z_new = (V * np.maximum(w, floor)) @ V.T
        r = np.linalg.norm(X - Z_new, "fro")           # primal residual
        s_ = rho * np.linalg.norm(Z_new - Z, "fro")    # dual residual
        Z = Z_new
        U = U + X - Z
        if verbose and it % 100 == 0:
            print(f"nearest_correlation_matrix (weighted): iteration {it}, residuals {r:.3e} {s_:.3e}, rho {rho:.3g}")
        if r < 1e-11 * K and s_ < 1e-11 * K:
            break
        if it % 50 == 49:  # residual balancing (scaled dual variable is rescaled with rho)
            if r > 10.0 * s_:
                rho, U = rho * 2.0, U / 2.0
            elif s_ > 10.0 * r:
                rho, U = rho / 2.0, U * 2.0
    Y = 0.5 * (Z + Z.T)
    np.fill_diagonal(Y, 1.0)
    lam = np.linalg.eigvalsh(Y).min()
    if lam < floor:  # finish on the feasible side with an exact unit diagonal
        t = (floor - lam) / (1.0 - lam)
        Y = (1.0 - t) * Y + t * np.eye(K)
    return Y


def _is_positive_definite(X):
    try:
        np.linalg.cholesky(X)
        return True
    except np.linalg.LinAlgError:
        return False


class Correlator(abc.ABC):
    def set_target(self, correlation_matrix):
        """Set target correlation matrix (reference correlation.py:162-179)."""
        if not isinstance(correlation_matrix, np.ndarray):
            raise TypeError("Input argument `correlation_matrix` must be NumPy array.")
        if not correlation_matrix.ndim == 2:
            raise ValueError("Correlation matrix must be square.")
        if not correlation_matrix.shape[0] == correlation_matrix.shape[1]:
            raise ValueError("Correlation matrix must be square.")
        if not np.allclose(np.diag(correlation_matrix), 1.0):
            raise ValueError("Correlation matrix must have 1.0 on diagonal.")
        if not np.allclose(correlation_matrix.T, correlation_matrix):
            raise ValueError("Correlation matrix must be symmetric.")
        if not _is_positive_definite(correlation_matrix):
            raise ValueError("Correlation matrix must be positive definite.")

        self.C = correlation_matrix.copy()
        self.P = np.linalg.cholesky(self.C)
        return self

    def _validate_X(self, X, check_rows_cols=True):
        """Validate array X of shape (observations, variables) (reference :181-202)."""
        if not (hasattr(self, "C") and hasattr(self, "P")):
            raise CorrelatorError("User must call `set_target` first.")
        if not (isinstance(X, np.ndarray) or _is_cuda_tensor(X)):
            raise TypeError("Input argument `X` must be NumPy array.")
        if not X.ndim == 2:
            raise ValueError("Correlation matrix must be square.")
        N, K = X.shape
        if self.P.shape[0] != K:
            msg = f"Shape of `X` ({tuple(X.shape)}) does not match shape of "
            msg += f"correlation matrix ({self.P.shape})"
            raise ValueError(msg)
        if check_rows_cols and N <= K:
            msg = f"The matrix X must have rows > columns. Got shape: {tuple(X.shape)}"
            raise ValueError(msg)
        return N, K


def _is_cuda_tensor(X):
    return type(X).__module__.startswith("torch") and hasattr(X, "is_cuda") and X.is_cuda


def _strides_elems(shape, strides_bytes, itemsize):
    return tuple(s // itemsize for s in strides_bytes)


class _IcPlan:
    """Owns one pbl_ic_plan (device workspace for an (n, k) problem on one device)."""

    def __init__(self, n, k, device, col_batch=0, rows_only=False):
        self.lib = _lib.require_gpu()
        self.n, self.k, self.device = n, k, device
        h = C.c_void_p()
        with _lib.device_guard(device):
            st = _lib.check(
                self.lib.pbl_ic_plan_create_ex(n, k, col_batch, 1 if rows_only else 0, C.byref(h)),
                "pbl_ic_plan_create")
        if st != _lib.STATUS_OK:
            raise ValueError(_lib.last_error())
        self.handle = h
        self._P = None

    def set_target(self, P):
        P = np.ascontiguousarray(P, dtype=np.float64)
        if self._P is not None and np.array_equal(P, self._P):
            return
        _lib.check(self.lib.pbl_ic_plan_set_target(self.handle, P.ctypes.data), "pbl_ic_plan_set_target")
        self._P = P.copy()

    def buffer(self, what):
        ptr, nbytes = C.c_void_p(), C.c_uint64()
        _lib.check(self.lib.pbl_ic_plan_buffer(self.handle, what, C.byref(ptr), C.byref(nbytes)))
        return ptr.value, nbytes.value

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.pbl_ic_plan_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_NOT_PD_MSG = (
    "Rank data correlation not positive definite."
    "There are perfect correlations in the ranked data."
    "Supply more data (rows in X) or sample differently."
)


def _raise_for_status(st):
    """Status of the device pipeline -> the exception the reference raises at that point."""
    if st == _lib.STATUS_OK:
        return
    if st == _lib.STATUS_NOT_PD:
        raise ValueError(_NOT_PD_MSG)  # correlation.py:399-403
    if st == _lib.STATUS_NON_FINITE:
        raise ValueError("array must not contain infs or NaNs")  # scipy check_finite at :409
    raise ValueError(_lib.last_error())


class _PlanCorrelator(Correlator):
    """Shared plumbing of the correlators that run on a pbl_ic_plan: workspace reuse, host and
    device entry points, status -> exception mapping."""

    _entry = None        # C symbol that runs the transform on a plan
    _entry_host = None   # C symbol with host buffers and pipelined copies (optional)
    _rows_only = False   # plan without sort workspace

    def __init__(self, device=None, col_batch=0):
        self.device = device
        self.col_batch = col_batch
        self._plan = None
        self._dev_bufs = None

    def set_target(self, correlation_matrix):
        super().set_target(correlation_matrix)
        return self

    # ------------------------------------------------------------------ plan management
    def _get_plan(self, n, k, device):
        p = self._plan
        if p is None or (p.n, p.k, p.device) != (n, k, device):
            if p is not None:
                p.close()
            self._free_dev_bufs()
            p = self._plan = _IcPlan(n, k, device, self.col_batch, rows_only=self._rows_only)
        p.set_target(self.P)
        return p

    def _free_dev_bufs(self):
        if self._dev_bufs is not None:
            lib = _lib.load()
            for ptr in self._dev_bufs[:2]:
                lib.pbl_device_free(C.c_void_p(ptr))
            self._dev_bufs = None

    def close(self):
        """Release the device workspace."""
        self._free_dev_bufs()
        if self._plan is not None:
            self._plan.close()
            self._plan = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _raise(self, st):
        _raise_for_status(st)

    def _run(self, plan, x_ptr, xrs, xcs, y_ptr, yrs, ycs, stream):
        fn = getattr(plan.lib, self._entry)
        with _lib.device_guard(plan.device):  # the plan's device, whatever the caller's current one is
            st = _lib.check(fn(plan.handle, C.c_void_p(x_ptr), xrs, xcs, C.c_void_p(y_ptr), yrs, ycs, stream),
                            self._entry)
        self._raise(st)

    # ------------------------------------------------------------------ the transform
    def __call__(self, X, *, out=None):
        """Transform an input matrix X of shape (N, K); same contract as the reference's
        ``__call__`` (correlation.py:248-285 / :368-425).

        ``out`` (extension, optional): a float64 array with X's shape and memory order to
        receive the result (e.g. page-locked memory, so that the device-to-host copy runs at
        PCIe speed); by default a fresh ``np.empty_like(X)`` is returned like the reference."""
        N, K = self._validate_X(X)
        if _is_cuda_tensor(X):
            return self._call_device(X, N, K)
        _lib.require_gpu()
        with _lib.device_guard(0 if self.device is None else int(self.device)):
            return self._call_host(X, N, K, out)

    def _call_host(self, X, N, K, out=None):
        lib = _lib.require_gpu()
        device = 0 if self.device is None else int(self.device)
        Xd = X
        if Xd.dtype != np.float64 or not (Xd.flags.f_contiguous or Xd.flags.c_contiguous):
            Xd = np.asfortranarray(X, dtype=np.float64)
        if out is not None:
            if not (isinstance(out, np.ndarray) and out.dtype == np.float64 and out.shape == Xd.shape
                    and out.strides == Xd.strides):
                raise ValueError("`out` must be a float64 array with the shape and strides of X")
            result = out
        else:
            result = np.empty_like(Xd)  # correlation.py:418 (keeps the memory order of X)
        plan = self._get_plan(N, K, device)
        nbytes = Xd.nbytes
        if self._dev_bufs is None or self._dev_bufs[2] != nbytes:
            self._free_dev_bufs()
            dX, dY = C.c_void_p(), C.c_void_p()
            _lib.check(lib.pbl_device_malloc(C.byref(dX), nbytes), "pbl_device_malloc")
            _lib.check(lib.pbl_device_malloc(C.byref(dY), nbytes), "pbl_device_malloc")
            self._dev_bufs = (dX.value, dY.value, nbytes)
        dX, dY, _ = self._dev_bufs
        rs, cs = _strides_elems(Xd.shape, Xd.strides, 8)
        yrs, ycs = _strides_elems(result.shape, result.strides, 8)
        if self._entry_host is not None:
            fn = getattr(lib, self._entry_host)
            st = _lib.check(fn(plan.handle, Xd.ctypes.data, rs, cs, result.ctypes.data, yrs, ycs,
                               C.c_void_p(dX), C.c_void_p(dY), None), self._entry_host)
            self._raise(st)
        else:
            _lib.check(lib.pbl_memcpy_h2d(dX, Xd.ctypes.data, nbytes, None), "pbl_memcpy_h2d")
            self._run(plan, dX, rs, cs, dY, yrs, ycs, None)
            _lib.check(lib.pbl_memcpy_d2h(result.ctypes.data, dY, nbytes, None), "pbl_memcpy_d2h")
            _lib.check(lib.pbl_stream_synchronize(None), "pbl_stream_synchronize")
        if result.dtype != X.dtype and self._keeps_dtype:
            result = result.astype(X.dtype)  # np.empty_like(X) keeps X's dtype in the reference
        return result

    def _call_device(self, X, N, K):
        import torch

        _lib.require_gpu()
        if X.dtype != torch.float64:
            raise TypeError("device-resident X must be float64")
        device = X.device.index if X.device.index is not None else torch.cuda.current_device()
        plan = self._get_plan(N, K, device)
        Y = torch.empty_strided(X.shape, X.stride(), dtype=X.dtype, device=X.device)
        stream = torch.cuda.current_stream(X.device).cuda_stream
        self._run(plan, X.data_ptr(), X.stride(0), X.stride(1), Y.data_ptr(), Y.stride(0), Y.stride(1),
                  C.c_void_p(stream))
        return Y

    def correlate_device(self, X):
        """(n, k) ``DeviceColumns`` -> new ``DeviceColumns`` (no host traffic): the graph path's
        ``correlator_instance(samples_input)`` (reference modeling.py:580-581) on device columns."""
        from ._device import DeviceColumns

        if not (hasattr(self, "C") and hasattr(self, "P")):
            raise CorrelatorError("User must call `set_target` first.")
        N, K = X.n, X.k
        if self.P.shape[0] != K:
            raise ValueError(f"Shape of `X` ({(N, K)}) does not match shape of correlation matrix ({self.P.shape})")
        if N <= K:
            raise ValueError(f"The matrix X must have rows > columns. Got shape: {(N, K)}")
        device = 0 if self.device is None else int(self.device)
        with _lib.device_guard(device):
            plan = self._get_plan(N, K, device)
            Y = DeviceColumns(N, K)
            self._run(plan, X.ptr, 1, N, Y.ptr, 1, N, None)
        return Y


class ImanConover(_PlanCorrelator):
    """Iman-Conover transform on the GPU (reference correlation.py:288-425).

    >>> transform = ImanConover().set_target(np.array([[1, 0.7], [0.7, 1]]))   # doctest: +SKIP
    >>> X_transformed = transform(X)                                           # doctest: +SKIP
    """

    _entry = "pbl_ic_plan_run"
    _entry_host = "pbl_ic_plan_run_host"
    _keeps_dtype = True


class Cholesky(_PlanCorrelator):
    """Cholesky transform on the GPU (reference correlation.py:205-285): standardise, remove the
    sample correlation with its Cholesky factor, impose the target's, restore mean and scale.
    Does not preserve the marginals.

    >>> transform = Cholesky().set_target(np.array([[1, 0.7], [0.7, 1]]))      # doctest: +SKIP
    >>> X_transformed = transform(X)                                           # doctest: +SKIP
    """

    _entry = "pbl_cholesky_plan_run"
    _rows_only = True
    _keeps_dtype = False  # `mean + X_n @ ...` is float64 whatever X was

    def _raise(self, st):
        if st == _lib.STATUS_NOT_PD:
            raise np.linalg.LinAlgError("Matrix is not positive definite")  # np.linalg.cholesky(cov), :271
        _raise_for_status(st)


def corrcoef(X, *, spearman=False, device=None):
    """``np.corrcoef(X, rowvar=False)`` or (``spearman=True``) the Spearman rank-correlation matrix of
    an (N, K) matrix, computed on the GPU: X may be a NumPy array, a CUDA ``torch.Tensor`` or
    ``DeviceColumns`` (no host traffic).  Lets the Pearson / Spearman acceptance checks of the
    correlators (reference tests/test_permutation_correlator.py) run at N = 1e8."""
    from ._device import DeviceColumns, as_device_columns

    lib = _lib.require_gpu()
    if _is_cuda_tensor(X):
        N, K = X.shape
        ptr, rs, cs, keep = X.data_ptr(), X.stride(0), X.stride(1), X
        device = X.device.index if device is None else device
    else:
        cols, keep = as_device_columns(X)
        N, K, ptr, rs, cs = cols.n, cols.k, cols.ptr, 1, cols.n
    device = 0 if device is None else int(device)
    stream = None
    if _is_cuda_tensor(X):  # ordered after whatever produced X on the caller's current stream
        import torch

        stream = C.c_void_p(torch.cuda.current_stream(X.device).cuda_stream)
    plan = _IcPlan(N, K, device, rows_only=not spearman)
    try:
        out = np.empty((K, K))
        with _lib.device_guard(device):
            st = _lib.check(lib.pbl_corrcoef_f64(plan.handle, C.c_void_p(ptr), rs, cs, 1 if spearman else 0,
                                                 out.ctypes.data, stream), "pbl_corrcoef_f64")
        _raise_for_status(st)
    finally:
        plan.close()
    del keep
    return out


class SwapIndexGenerator:
    """Disjoint swap index pairs drawn from a NumPy generator exactly like the reference's
    (correlation.py:428-470): consumes one ``rng.permutation(n)`` 2*size entries at a time and
    re-draws when it runs out.  Host control plane: this stream *is* the reference's stream, which
    is what makes the device hill climb reproduce the reference's accept/reject sequence."""

    def __init__(self, rng, n):
        assert n >= 2
        self.rng = rng
        self.indices = np.arange(n)
        self.permutation = self.rng.permutation(self.indices)

    def snapshot(self):
        """State to come back to with restore(): the unconsumed part of the current permutation (a view: draws
        re-slice it, a redraw replaces it, nothing writes into it) and the generator's state."""
        import copy

        return self.permutation, copy.deepcopy(self.rng.bit_generator.state)

    def restore(self, snap):
        self.permutation, state = snap
        self.rng.bit_generator.state = state

    def __call__(self, size):
        assert size >= 1
        size = min(size, len(self.indices) // 2)
        chosen, self.permutation = self.permutation[: 2 * size], self.permutation[2 * size:]
        if len(chosen) < 2 * size:
            self.permutation = self.rng.permutation(self.indices)
            return self(size)
        return chosen[:size], chosen[size:]


class PermutationCorrelator(Correlator):
    """Randomised hill climbing on row swaps within each column (reference correlation.py:473-703),
    executed as one persistent CUDA block per chunk of steps (csrc/permcorr.cu) with the incremental
    correlation update of the reference's ``CorrelationMatrix`` (:757-921).

    Same constructor and call signature as the reference.  The swap proposals are the reference's own
    NumPy stream (``SwapIndexGenerator``), so for the same ``seed`` the result equals the
    reference's entry for entry."""

    _CHUNK_STEPS = 1 << 16  # steps per launch (bounds the host-side index staging)

    def __init__(self, *, weights=None, iterations=1000, tol=0.01, correlation_type="pearson", seed=None,
                 verbose=False, device=None):
        if not (weights is None or np.all(weights > 0)):
            raise ValueError("`weights` must have positive entries.")
        if not (isinstance(iterations, int) and iterations >= 0):
            raise ValueError("`iterations` must be non-negative integer.")
        if not isinstance(tol, float) and tol > 0:
            raise ValueError("`tol` must be a positive float.")
        if not (seed is None or isinstance(seed, int)):
            raise TypeError("`seed` must be None or an integer")
        if not isinstance(verbose, bool):
            raise TypeError("`verbose` must be boolean")
        self.iters = iterations
        self.tol = tol
        self.rng = np.random.default_rng(seed)
        self.verbose = verbose
        self.correlation_type = correlation_type
        self.device = device

    def set_target(self, correlation_matrix, *, weights=None):
        super().set_target(correlation_matrix)
        weights = np.ones_like(self.C) if weights is None else weights
        self.weights = weights / np.sum(weights)
        self.triu_indices = np.triu_indices(self.C.shape[0], k=1)
        return self

    def _error(self, observed, target):
        """RMSE over the upper triangle of corr(X) - target (reference :597-601); K x K, host."""
        idx = self.triu_indices
        return float(np.sqrt(np.sum(self.weights[idx] * (observed[idx] - target[idx]) ** 2.0)))

    @staticmethod
    def subiters(n, i):
        """Swaps per step: long swap lists early, single swaps for the last half (reference :603-617)."""
        C = np.log2(n) + 1
        return int(np.ceil(C ** (1 - (2 * i / n))))

    def __call__(self, X):
        self._validate_X(X, check_rows_cols=False)
        if not (isinstance(X, np.ndarray) and X.ndim == 2):
            raise ValueError("`X` must be a 2D numpy array.")
        if self.correlation_type not in ("pearson", "spearman"):
            raise ValueError(f"`correlation_type` must be in ('pearson', 'spearman'), got {self.correlation_type}")
        from ._device import DeviceColumns

        lib = _lib.require_gpu()
        N, K = X.shape
        if N < 2:
            raise AssertionError("need at least two observations")
        if self.verbose:
            print(f"Running permutation correlator for {self.iters if self.iters else 'inf'} iterations.")
        device = 0 if self.device is None else int(self.device)
        spearman = self.correlation_type == "spearman"
        plan = _IcPlan(N, K, device, rows_only=not spearman)
        # X goes to the device in the layout it has (the begin stage reads any strides and makes the
        # column-major working copy itself); the result comes back in C order like the reference's X.copy()
        Xh = X if (X.dtype == np.float64 and (X.flags.c_contiguous or X.flags.f_contiguous)) \
            else np.ascontiguousarray(X, dtype=np.float64)
        xrs, xcs = _strides_elems(Xh.shape, Xh.strides, 8)
        raw = C.c_void_p()
        dY = DeviceColumns(N, K)
        try:
            _lib.check(lib.pbl_device_malloc(C.byref(raw), max(Xh.nbytes, 16)), "pbl_device_malloc")
            if Xh.nbytes:
                _lib.check(lib.pbl_memcpy_h2d(raw, Xh.ctypes.data, Xh.nbytes, None), "pbl_memcpy_h2d")
            target = np.ascontiguousarray(self.C, dtype=np.float64)
            weights = np.ascontiguousarray(self.weights, dtype=np.float64)
            st = _lib.check(lib.pbl_permcorr_begin(plan.handle, raw, xrs, xcs, C.c_void_p(dY.ptr),
                                                   1 if spearman else 0, target.ctypes.data, weights.ctypes.data,
                                                   None), "pbl_permcorr_begin")
            if st == _lib.STATUS_NOT_PD:
                raise ValueError("X has one or several constant columns")  # correlation.py:847-848
            _raise_for_status(st)
            swaps_gen = SwapIndexGenerator(rng=self.rng, n=N)
            n_sched = self.iters if self.iters else 10_000
            iters_per_chunk = max(1, self._CHUNK_STEPS // K)
            iteration, finished = 1, False
            while not finished and (not self.iters or iteration <= self.iters):
                last = iteration + iters_per_chunk - 1
                if self.iters:
                    last = min(last, self.iters)
                snapshot = swaps_gen.snapshot()  # to replay the exact consumption on early exit
                cols, offs, cnts, idx, counts_per_iter = [], [], [], [], []
                total = 0
                for it in range(iteration, last + 1):
                    num_swaps = self.subiters(n=n_sched, i=it)
                    counts_per_iter.append(num_swaps)
                    for k in range(K):
                        i, j = swaps_gen(num_swaps)
                        cols.append(k)
                        offs.append(total)
                        cnts.append(len(i))
                        idx.append(i)
                        idx.append(j)
                        total += len(i)
                n_steps = len(cols)
                cols_a = np.asarray(cols, dtype=np.int32)
                offs_a = np.asarray(offs, dtype=np.int32)
                cnts_a = np.asarray(cnts, dtype=np.int32)
                idx_a = np.ascontiguousarray(np.concatenate(idx), dtype=np.int64)
                errors = np.zeros(last - iteration + 3)
                done, conv, nerr = C.c_int64(), C.c_int32(), C.c_int64()
                st = _lib.check(lib.pbl_permcorr_steps(
                    plan.handle, C.c_void_p(dY.ptr), cols_a.ctypes.data, offs_a.ctypes.data, cnts_a.ctypes.data,
                    idx_a.ctypes.data, total, n_steps, float(self.tol), C.byref(done), C.byref(conv),
                    errors.ctypes.data, errors.size, C.byref(nerr), None), "pbl_permcorr_steps")
                _raise_for_status(st)
                if self.verbose:
                    self._print_progress(iteration, last, errors, int(nerr.value), counts_per_iter, bool(conv.value))
                if conv.value:
                    # leave self.rng where the reference would have left it: replay the consumed draws only
                    swaps_gen.restore(snapshot)
                    for t in range(int(done.value)):
                        swaps_gen(counts_per_iter[t // K])
                    self.rng = swaps_gen.rng
                    finished = True
                iteration = last + 1
            # corr_mat.X is X.copy(): C order (reference :831) -- transposed on the device, into the input's
            # staging buffer (dead since the begin stage), then one contiguous copy to the host
            result = np.empty((N, K), dtype=np.float64)
            _lib.check(lib.pbl_copy_strided_f64(C.c_void_p(dY.ptr), 1, N, raw, K, 1, N, K, None), "pbl_copy_strided_f64")
            if result.nbytes:
                _lib.check(lib.pbl_memcpy_d2h(result.ctypes.data, raw, result.nbytes, None), "pbl_memcpy_d2h")
            _lib.check(lib.pbl_stream_synchronize(None), "pbl_stream_synchronize")
        finally:
            if raw.value:
                lib.pbl_device_free(raw)
            dY.free()
            plan.close()
        return result if result.dtype == X.dtype else result.astype(X.dtype)

    def _print_progress(self, first, last, errors, nerr, counts, converged):
        # errors[0] = before the chunk, errors[t] = after the check of the t-th iteration of the chunk
        every = self.iters // 10 if self.iters else 1000
        for t, it in enumerate(range(first, first + max(nerr - 1, 0) + (0 if converged else 0))):
            if it > last:
                break
            if every and it % every == 0 and t < len(errors):
                print(f" Iter {it:>6}  Error: {errors[t]:.6f} Swaps: {counts[t]:>2}")
        if converged and nerr >= 1:
            it = first + nerr - 2
            print(f" Terminating at iteration {it} due to tolerance. Error: {errors[nerr - 1]:.6f}")
