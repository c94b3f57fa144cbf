"""Unit-cube generators with the ``scipy.stats.qmc`` call surface the reference uses
(``Engine(d=d, rng=seed_or_generator).random(n)``, reference src/probabilit/modeling.py:479-489;
``LatinHypercube(d=2, seed=42, scramble=True)``, README.md:113), generated on the B200.

* ``Sobol``  -- bit-exact with ``scipy.stats.qmc.Sobol`` (scrambled or not) for the same seed:
  the host draws the scrambling bits from the NumPy generator exactly as SciPy does; direction
  numbers, LMS scrambling and the points are computed by CUDA kernels.
* ``Halton`` -- bit-exact with ``scipy.stats.qmc.Halton`` for the same seed (digit permutations
  drawn on the host like SciPy, radical inverse on the device).
* ``LatinHypercube`` and ``PhiloxUniform`` -- GPU-native counter-based streams: statistically
  equivalent to SciPy/NumPy (exact stratification for LHS), not bit-equal.

``random(n)`` returns a NumPy (n, d) float64 array (column-major); ``random(n, device=True)``
returns a CUDA ``torch.Tensor`` with the same shape/strides and no host traffic.
There is no CPU fallback: without the CUDA library / a GPU the calls raise.
"""
import copy
import ctypes as C
import math
import numbers
import os
import warnings

import numpy as np

from . import _lib


# ------------------------------------------------------------------------------------------
# small device-memory helper over the C ABI (no torch needed for NumPy in / NumPy out)
# ------------------------------------------------------------------------------------------
class _DevMem:
    def __init__(self, nbytes):
        self.lib = _lib.require_gpu()
        self.nbytes = int(nbytes)
        self.ptr = C.c_void_p()
        _lib.check(self.lib.pbl_device_malloc(C.byref(self.ptr), max(self.nbytes, 16)), "pbl_device_malloc")

    @classmethod
    def from_host(cls, arr):
        arr = np.ascontiguousarray(arr)
        m = cls(arr.nbytes)
        if arr.nbytes:
            _lib.check(m.lib.pbl_memcpy_h2d(m.ptr, arr.ctypes.data, arr.nbytes, None), "pbl_memcpy_h2d")
            _lib.check(m.lib.pbl_stream_synchronize(None))
        return m

    def to_host(self, shape, dtype, order="C"):
        out = np.empty(shape, dtype=dtype, order=order)
        if out.nbytes:
            _lib.check(self.lib.pbl_memcpy_d2h(out.ctypes.data, self.ptr, out.nbytes, None), "pbl_memcpy_d2h")
            _lib.check(self.lib.pbl_stream_synchronize(None))
        return out

    def free(self):
        if self.ptr is not None and self.ptr.value:
            self.lib.pbl_device_free(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def check_random_state(seed=None):
    """scipy/stats/_qmc.py:61-84: int / None -> np.random.default_rng, Generator / RandomState as is."""
    if seed is None or isinstance(seed, (numbers.Integral, np.integer)):
        return np.random.default_rng(seed)
    if isinstance(seed, (np.random.RandomState, np.random.Generator)):
        return seed
    raise ValueError(f"{seed!r} cannot be used to seed a numpy.random.Generator instance")


def _rng_integers(gen, low, size, dtype):
    """scipy._lib._util.rng_integers for [0, low)."""
    if isinstance(gen, np.random.Generator):
        return gen.integers(low, size=size, dtype=dtype)
    return gen.randint(low, size=size, dtype=dtype)


def n_primes(d):
    primes, c = [], 2
    while len(primes) < d:
        if all(c % p for p in primes if p * p <= c):
            primes.append(c)
        c += 1
    return primes


class QMCEngine:
    """Common state of the engines (scipy/stats/_qmc.py:930-950)."""

    def __init__(self, d, *, rng=None, seed=None):
        if not np.issubdtype(type(d), np.integer) or d < 0:
            raise ValueError("d must be a non-negative integer value")
        if rng is None:
            rng = seed
        self.d = int(d)
        if isinstance(rng, np.random.Generator):
            bg = rng._bit_generator  # own a spawned child, like scipy's _rng_spawn
            self.rng = np.random.Generator(type(bg)(bg._seed_seq.spawn(1)[0]))
        else:
            self.rng = check_random_state(rng)
        self.rng_seed = copy.deepcopy(self.rng)
        self.num_generated = 0

    # -- output plumbing ------------------------------------------------------------------
    def _generate(self, n, launch, device):
        """launch(out_ptr, row_stride, col_stride, stream) writes the (n, d) block column-major."""
        n, d = int(n), self.d
        if device == "columns":  # torch-free device buffer (what the graph evaluator consumes)
            from ._device import DeviceColumns

            cols = DeviceColumns(n, d)
            if n and d:
                launch(C.c_void_p(cols.ptr), 1, n, None)
            return cols
        if device:
            import torch
            buf = torch.empty((d, n), dtype=torch.float64, device="cuda")
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            if n and d:
                launch(C.c_void_p(buf.data_ptr()), 1, n, stream)
            return buf.T
        mem = _DevMem(n * d * 8)
        try:
            if n and d:
                launch(mem.ptr, 1, n, None)
            return mem.to_host((n, d), np.float64, order="F")
        finally:
            mem.free()

    def random(self, n=1, *, device=False):
        sample = self._random(int(n), device)
        self.num_generated += int(n)
        return sample

    def fast_forward(self, n):
        self.num_generated += int(n)
        return self

    def reset(self):
        self.rng = copy.deepcopy(self.rng_seed)
        self.num_generated = 0
        return self


class Sobol(QMCEngine):
    """scipy.stats.qmc.Sobol (scipy/stats/_qmc.py:1761-1957), computed on the GPU, bit-exact."""

    MAXDIM = 21201

    def __init__(self, d, *, scramble=True, bits=None, rng=None, seed=None, optimization=None):
        super().__init__(d, rng=rng, seed=seed)
        if optimization is not None:
            raise NotImplementedError("optimization is host-side post-processing in scipy; not on the hot path")
        if d > self.MAXDIM:
            raise ValueError(f"Maximum supported dimensionality is {self.MAXDIM}.")
        self.bits = 30 if bits is None else int(bits)
        if not 1 <= self.bits <= 64:
            raise ValueError("Maximum supported 'bits' is 64")
        self.maxn = 2 ** self.bits
        self.scramble = scramble
        lib = _lib.require_gpu()
        import scipy.stats
        z = np.load(os.path.join(os.path.dirname(scipy.stats.__file__), "_sobol_direction_numbers.npz"))
        poly = np.ascontiguousarray(z["poly"][: max(self.d, 1)], dtype=np.int64)
        vinit = np.ascontiguousarray(z["vinit"][: max(self.d, 1)], dtype=np.int64)
        d_poly, d_vinit = _DevMem.from_host(poly), _DevMem.from_host(vinit)
        self._sv_dev = _DevMem(max(self.d, 1) * self.bits * 8)
        self._shift_dev = _DevMem.from_host(np.zeros(max(self.d, 1), dtype=np.uint64))
        if self.d:
            _lib.check(lib.pbl_sobol_direction_numbers(d_poly.ptr, d_vinit.ptr, vinit.shape[1], self.d,
                                                       self.bits, self._sv_dev.ptr, None))
            if scramble:
                dt = np.uint32 if self.bits <= 32 else np.uint64
                # the two draws scipy makes, in scipy's order (scipy/stats/_qmc.py:1812-1824)
                shift_bits = _rng_integers(self.rng, 2, (self.d, self.bits), dt)
                ltm_bits = _rng_integers(self.rng, 2, (self.d, self.bits, self.bits), dt)
                d_shift = _DevMem.from_host(shift_bits.astype(np.uint8))
                d_ltm = _DevMem.from_host(ltm_bits.astype(np.uint8))
                _lib.check(lib.pbl_sobol_scramble(d_ltm.ptr, d_shift.ptr, self.d, self.bits, self._sv_dev.ptr,
                                                  self._shift_dev.ptr, None))
                _lib.check(lib.pbl_stream_synchronize(None))
                d_shift.free()
                d_ltm.free()
        _lib.check(lib.pbl_stream_synchronize(None))
        d_poly.free()
        d_vinit.free()

    @property
    def _sv(self):
        return self._sv_dev.to_host((self.d, self.bits), np.uint64)

    @property
    def _shift(self):
        return self._shift_dev.to_host((self.d,), np.uint64)

    def _random(self, n, device):
        total = self.num_generated + n
        if total > self.maxn:
            raise ValueError(f"At most 2**{self.bits}={self.maxn} distinct points can be generated.")
        if self.num_generated == 0 and n > 0 and (n & (n - 1)) != 0:
            warnings.warn("The balance properties of Sobol' points require n to be a power of 2.", stacklevel=3)
        lib = _lib.load()
        skip = self.num_generated

        def launch(out, rs, cs, stream):
            _lib.check(lib.pbl_sobol_f64(self._sv_dev.ptr, self._shift_dev.ptr, self.d, self.bits, skip, n, out,
                                         rs, cs, stream), "pbl_sobol_f64")
        return self._generate(n, launch, device)

    def random_base2(self, m, *, device=False):
        n = 2 ** int(m)
        total = self.num_generated + n
        if total & (total - 1) != 0:
            raise ValueError("The balance properties of Sobol' points require n to be a power of 2.")
        return self.random(n, device=device)


class Halton(QMCEngine):
    """scipy.stats.qmc.Halton (scipy/stats/_qmc.py:1120-1283), computed on the GPU, bit-exact."""

    def __init__(self, d, *, scramble=True, rng=None, seed=None, optimization=None):
        super().__init__(d, rng=rng, seed=seed)
        if optimization is not None:
            raise NotImplementedError("optimization is host-side post-processing in scipy; not on the hot path")
        self.base = n_primes(self.d)
        self.scramble = scramble
        self._bases_dev = _DevMem.from_host(np.asarray(self.base or [2], dtype=np.int32))
        self._perms_dev = self._off_dev = self._cnt_dev = None
        if scramble and self.d:
            flat, off, cnt = [], [], []
            pos = 0
            for b in self.base:  # scipy/stats/_qmc.py:724-729, same draws in the same order
                count = math.ceil(54 / math.log2(b)) - 1
                perms = np.repeat(np.arange(b)[None], count, axis=0)
                for row in perms:
                    self.rng.shuffle(row)
                flat.append(perms.astype(np.int64).ravel())
                off.append(pos)
                cnt.append(count)
                pos += count * b
            self._perms_dev = _DevMem.from_host(np.concatenate(flat))
            self._off_dev = _DevMem.from_host(np.asarray(off, dtype=np.int64))
            self._cnt_dev = _DevMem.from_host(np.asarray(cnt, dtype=np.int32))

    def _random(self, n, device):
        lib = _lib.load()
        start = self.num_generated
        null = C.c_void_p()

        def launch(out, rs, cs, stream):
            _lib.check(lib.pbl_halton_f64(
                self._bases_dev.ptr, self._perms_dev.ptr if self._perms_dev else null,
                self._off_dev.ptr if self._off_dev else null, self._cnt_dev.ptr if self._cnt_dev else null,
                self.d, start, n, out, rs, cs, stream), "pbl_halton_f64")
        return self._generate(n, launch, device)


class LatinHypercube(QMCEngine):
    """scipy.stats.qmc.LatinHypercube, strength 1 (scipy/stats/_qmc.py:1546-1559):
    ``(perm - U) / n`` with a GPU-native counter-based permutation and Philox jitter."""

    def __init__(self, d, *, scramble=True, strength=1, optimization=None, rng=None, seed=None):
        super().__init__(d, rng=rng, seed=seed)
        if strength != 1 or optimization is not None:
            raise NotImplementedError("only strength=1 without optimization is on the hot path")
        self.scramble = scramble

    def _random(self, n, device):
        lib = _lib.load()
        key = int(_rng_integers(self.rng, 2 ** 62, None, np.int64))  # one draw per call

        def launch(out, rs, cs, stream):
            _lib.check(lib.pbl_lhs_f64(key, n, self.d, 1 if self.scramble else 0, out, rs, cs, stream),
                       "pbl_lhs_f64")
        return self._generate(n, launch, device)


class PhiloxUniform(QMCEngine):
    """Pseudo-random (n, d) uniforms: the ``random_state.random((size, d))`` draw of
    Node.sample(method=None) (reference modeling.py:485-486) as a Philox4x32-10 counter stream."""

    def __init__(self, d, *, rng=None, seed=None):
        super().__init__(d, rng=rng, seed=seed)
        self._key = int(_rng_integers(self.rng, 2 ** 62, None, np.int64))

    def _random(self, n, device):
        lib = _lib.load()
        if self.num_generated % 2:
            raise ValueError("PhiloxUniform blocks must start at an even row (counter = row pair)")
        row0 = self.num_generated

        def launch(out, rs, cs, stream):
            _lib.check(lib.pbl_uniform_f64(self._key, row0, n, self.d, out, rs, cs, stream), "pbl_uniform_f64")
        return self._generate(n, launch, device)
