"""Host-memory stand-ins for the device plumbing of probabilit_b200.modeling, so that the CPU
test-suite can run the graph *compiler* end to end: DeviceColumns backed by NumPy, and
``_run_program`` executed by the oracle's bytecode VM (oracle/graph_vm.py) instead of the CUDA
kernel.  Test infrastructure only -- the product has no such path."""
import numpy as np

from oracle import graph_vm

_REGISTRY = {}


class FakeColumns:
    def __init__(self, n, k):
        self.n, self.k = int(n), int(k)
        self.data = np.zeros((self.k, self.n))
        self.base = (len(_REGISTRY) + 1) << 20
        for j in range(self.k):
            _REGISTRY[self.base + j] = self.data[j]

    def column_ptr(self, j):
        return self.base + j

    @classmethod
    def from_host(cls, arr):
        arr = np.asarray(arr, dtype=np.float64)
        out = cls(arr.shape[0], arr.shape[1])
        out.data[:] = arr.T
        return out

    def to_host(self):
        return np.asfortranarray(self.data.T.copy())

    def column_to_host(self, j):
        return self.data[j].copy()


def fake_as_device_columns(q):
    if isinstance(q, FakeColumns):
        return q, q
    cols = FakeColumns.from_host(q)
    return cols, cols


def device_form(prog):
    """The library's device form of an ABI program: (program, op words with the internal flags, slot count)."""
    import ctypes as C

    from probabilit_b200 import _lib

    lib = _lib.load()
    ops = (C.c_int32 * len(prog))()
    out = (_lib.GraphInstr * len(prog))()
    ns = C.c_int32(0)
    assert lib.pbl_graph_debug_translate(prog, len(prog), out, ops, C.byref(ns)) == 0
    assert sorted(o & 0xFFF for o in ops) == sorted(i.op & 0xFFF for i in prog)  # a permutation of the same work
    return out, list(ops), ns.value


DEVICE_FORM = False  # True: run the VM on the library's device form of the program (accumulator / dead-store flags)


def fake_run_program(em, n, row0, inputs, outputs):
    prog, n_slots = em.assemble()
    dev_ops = None
    if DEVICE_FORM and len(prog):
        prog, dev_ops, n_slots = device_form(prog)
    return graph_vm.run(list(prog), n_slots, n, [_REGISTRY[p] for p in inputs], [_REGISTRY[p] for p in outputs],
                        dev_ops=dev_ops)


def install(monkeypatch):
    import probabilit_b200.modeling as m

    monkeypatch.setattr(m, "DeviceColumns", FakeColumns)
    monkeypatch.setattr(m, "as_device_columns", fake_as_device_columns)
    monkeypatch.setattr(m, "_run_program", fake_run_program)
