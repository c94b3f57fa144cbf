"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU and exports
every symbol include/probabilit_b200.h declares (no compute calls), the ctypes table covers the
header, and the product never imports the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "probabilit_b200.h")).read()
    return sorted(set(re.findall(r"PBL_API\s+[\w\s\*]+?\b(pbl_\w+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    from probabilit_b200 import _lib

    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.pbl_version() >= 100
    assert isinstance(_lib.kernel_launches(), int)


def test_ctypes_table_matches_header():
    from probabilit_b200 import _lib

    assert sorted(_lib.SIGNATURES) == header_symbols()
    assert ctypes.sizeof(_lib.GraphInstr) == 56  # pbl_graph_instr


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "probabilit_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, re.M) or "oracle/" in text and f.endswith(".py"):
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders


def test_compute_calls_fail_loudly_without_a_gpu():
    """No CPU fallback: without a CUDA device the public API raises instead of computing."""
    import numpy as np

    from probabilit_b200 import ImanConover, _lib

    if _lib.load().pbl_device_count() > 0:
        pytest.skip("a GPU is visible")
    X = np.random.default_rng(0).normal(size=(100, 2))
    with pytest.raises(_lib.PblError):
        ImanConover().set_target(np.array([[1, 0.5], [0.5, 1]]))(X)
    import probabilit_b200.modeling as m

    with pytest.raises(_lib.PblError):
        m.Distribution("norm").sample(10, random_state=0)
