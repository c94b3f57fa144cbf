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


def fake_run_program(em, n, row0, inputs, outputs):
    prog, n_slots = em.assemble()
    return graph_vm.run(list(prog), n_slots, n, [_REGISTRY[p] for p in inputs], [_REGISTRY[p] for p in outputs])


def install(monkeypatch):
    import probabilit_b200.modeling as m

    monkeypatch.setattr(m, "DeviceColumns", FakeColumns)
    monkeypatch.setattr(m, "as_device_columns", fake_as_device_columns)
    monkeypatch.setattr(m, "_run_program", fake_run_program)
