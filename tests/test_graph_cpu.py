"""CPU tests of the modeling-graph path: the oracle restatement is pinned to the unmodified
reference's golden vectors, and the host compiler of probabilit_b200.modeling is run end to end
against the same vectors with the CUDA kernel replaced by the oracle's bytecode VM."""
import os

import numpy as np
import pytest

import fake_device
import graph_recipes
from oracle import iman_conover as oic
from oracle import modeling as omod

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "graph_reference.npz"))


def ic_correlate(X, C):
    return oic.iman_conover(X, C)


class OracleImanConover:
    """A 'foreign' correlator following the reference protocol (modeling.py:577-581)."""

    def set_target(self, C):
        self.C = C
        return self

    def __call__(self, X):
        return oic.iman_conover(X, self.C)


@pytest.mark.parametrize("name", list(graph_recipes.RECIPES))
def test_oracle_matches_reference_golden(name):
    import probabilit_b200.modeling as m

    recipe, n = graph_recipes.RECIPES[name]
    sink, named = recipe(m)
    samples = omod.evaluate(sink, GOLDEN[f"{name}__quantiles"], correlate=ic_correlate)
    for label, node in named:
        want = GOLDEN[f"{name}__{label}"]
        got = samples[node]
        assert got.dtype == want.dtype, (label, got.dtype, want.dtype)
        np.testing.assert_array_equal(got, want, err_msg=f"{name}:{label}")


@pytest.mark.parametrize("name", list(graph_recipes.RECIPES))
def test_compiler_with_vm_matches_reference_golden(name, monkeypatch):
    import probabilit_b200.modeling as m

    fake_device.install(monkeypatch)
    recipe, n = graph_recipes.RECIPES[name]
    sink, named = recipe(m)
    result = sink.sample_from_quantiles(GOLDEN[f"{name}__quantiles"], correlator=OracleImanConover)
    for label, node in named:
        want = GOLDEN[f"{name}__{label}"]
        got = node.samples_
        assert got.dtype == want.dtype, (label, got.dtype, want.dtype)
        np.testing.assert_array_equal(got, want, err_msg=f"{name}:{label}")
    if type(sink).__name__ != "NoOp":
        np.testing.assert_array_equal(result, GOLDEN[f"{name}__{named[-1][0]}"])


def test_gc_strategy_and_constants(monkeypatch):
    import probabilit_b200.modeling as m

    fake_device.install(monkeypatch)
    a = m.Distribution("norm")
    inter = (a + a) ** 2 - a
    final = m.Exp(inter)
    q = np.random.default_rng(0).random((50, 1))
    out = final.sample_from_quantiles(q, gc_strategy=[])
    assert not hasattr(a, "samples_") and not hasattr(inter, "samples_")
    final.sample_from_quantiles(q, gc_strategy=[a])
    assert hasattr(a, "samples_") and not hasattr(inter, "samples_")
    np.testing.assert_array_equal(out, final.samples_)
    final.sample_from_quantiles(q)
    two = [n for n in final.nodes() if isinstance(n, m.Constant)][0]
    assert two.samples_.dtype == np.int64 and two.samples_.shape == (50,)


def test_errors(monkeypatch):
    import probabilit_b200.modeling as m

    fake_device.install(monkeypatch)
    x = m.Distribution("norm")
    with pytest.raises(ValueError, match="Sampling this node gave non-finite values"):
        m.Log(x - 100).sample_from_quantiles(np.full((4, 1), 0.5))
    with pytest.raises(ValueError, match="Sampling this node gave non-finite values"):
        x.sample_from_quantiles(np.array([[0.0], [0.5]]))  # norm.ppf(0) = -inf
    y = m.Distribution("norm", loc=x)
    with pytest.raises(ValueError, match="Cannot correlate variable"):
        (x + y).correlate(x, y, corr_mat=np.eye(2)).sample_from_quantiles(np.full((4, 2), 0.5))
    with pytest.raises(ValueError, match="is not an ancestor"):
        x.correlate(x, m.Distribution("norm"), corr_mat=np.eye(2))
    with pytest.raises(NotImplementedError):
        m.Distribution("cauchy").sample_from_quantiles(np.full((4, 1), 0.5))
    with pytest.raises(TypeError):
        (-(x > 0)).sample_from_quantiles(np.full((4, 1), 0.5))  # numpy: boolean negative


def test_distribution_constructors(monkeypatch):
    """probabilit_b200.distributions mirrors the reference's constructors (distributions.py doctests)."""
    import probabilit_b200
    import probabilit_b200.distributions as d
    import probabilit_b200.modeling as m

    fake_device.install(monkeypatch)
    assert repr(d.PERT(0, 6, 10)) == 'Distribution("beta", a=3.4, b=2.6, loc=0, scale=10)'
    assert d.pert_to_beta(0, 9, 10, gamma=6) == (6.4, 1.6, 0, 10)
    assert repr(d.Triangular(low=1, mode=5, high=9, low_perc=0, high_perc=1)) == \
        'Distribution("triang", loc=1, scale=8, c=0.5)'
    loc, scale, c = d.fit_triangular_distribution(3, 8, 10, low_perc=0.10, high_perc=0.90)
    assert abs(loc + 0.207) < 1e-2 and abs(scale - 12.53) < 1e-2 and abs(c - 0.65) < 1e-2
    with pytest.raises(ValueError):
        d.Triangular(5, 4, 9)
    # Lognormal(mean, std): moments of the lognormal itself (reference distributions.py:39-45)
    q = np.random.default_rng(0).random((20000, 1))
    s = d.Lognormal(mean=2, std=1).sample_from_quantiles(q)
    assert abs(s.mean() - 2.0) < 0.02 and abs(s.std() - 1.0) < 0.05
    u = d.Uniform(2, 5).sample_from_quantiles(q)
    assert u.min() >= 2 and u.max() < 5
    pert = d.PERT(0, 6, 10).sample_from_quantiles(q)  # beta inverse CDF
    assert 0 <= pert.min() and pert.max() <= 10 and abs(pert.mean() - (0 + 4 * 6 + 10) / 6) < 0.05
    tn = d.TruncatedNormal(loc=3, scale=0.5, low=2.5, high=4).sample_from_quantiles(q)
    assert tn.min() >= 2.5 and tn.max() <= 4
    assert probabilit_b200.Distribution is m.Distribution and probabilit_b200.PERT is d.PERT


@pytest.mark.parametrize("name", list(graph_recipes.RECIPES))
def test_device_form_of_the_program_preserves_the_results(name, monkeypatch):
    """The library translates the ABI bytecode into its device form before launching (csrc/graph.cu:
    operands taken from the accumulator, dead slot stores skipped).  The translation is host code: run the
    oracle VM on the device form and require the reference's vectors."""
    import probabilit_b200.modeling as m

    fake_device.install(monkeypatch)
    monkeypatch.setattr(fake_device, "DEVICE_FORM", True)
    recipe, n = graph_recipes.RECIPES[name]
    sink, named = recipe(m)
    sink.sample_from_quantiles(GOLDEN[f"{name}__quantiles"], correlator=OracleImanConover)
    for label, node in named:
        np.testing.assert_array_equal(node.samples_, GOLDEN[f"{name}__{label}"], err_msg=f"{name}:{label}")


def test_device_form_chains_the_mutual_fund_graph_through_the_accumulator(monkeypatch):
    """README example 3: ppf -> MUL -> ADD per year.  With only the sink retained every interest rate and every
    product is consumed by the next instruction: no slot traffic except the running total."""
    import ctypes as C

    import probabilit_b200.modeling as m
    from probabilit_b200 import _lib

    fake_device.install(monkeypatch)
    seen = {}

    def capture(em, n, row0, inputs, outputs):
        prog, n_slots = em.assemble()
        seen["prog"], seen["ops"], seen["n_slots"] = fake_device.device_form(prog)
        seen["abi_slots"] = n_slots
        return fake_device.fake_run_program(em, n, row0, inputs, outputs)

    monkeypatch.setattr(m, "_run_program", capture)
    sink, _ = graph_recipes.mutual_fund(m)
    q = np.random.default_rng(0).random((50, 20))
    sink.sample_from_quantiles(q, gc_strategy=[])
    ops = seen["ops"]
    names = [o & 0xFF for o in ops]
    assert names.count(16) == 20  # the norm ppfs
    acc = sum(1 for o in ops if o & 0x3000)
    no_write = sum(1 for o in ops if o & 0x4000)
    # 20 x (ppf -> MUL via the accumulator -> ADD via the accumulator); only the running total lives in a slot
    assert acc >= 40 and no_write >= 40, (acc, no_write, len(ops), [hex(o) for o in ops])
    assert seen["abi_slots"] >= 20 and seen["n_slots"] <= 2, (seen["abi_slots"], seen["n_slots"])


@pytest.mark.parametrize("seed", range(20))
def test_device_form_on_random_graphs(seed, monkeypatch):
    """Random modeling graphs (inverse CDFs with immediate and node-valued parameters, arithmetic, comparisons,
    shared sub-expressions, all nodes retained or only the sink): the VM run on the ABI program and on the
    library's device form of it (producers sunk, accumulator operands, slots re-allocated) must agree on every
    retained node, bit for bit."""
    import probabilit_b200.modeling as m

    rng = np.random.default_rng(1000 + seed)

    def build():
        r = np.random.default_rng(2000 + seed)  # the same graph both times
        nodes = []
        for _ in range(int(r.integers(2, 6))):
            kind = r.choice(["norm", "uniform", "expon", "triang"])
            if kind == "norm":
                nodes.append(m.Distribution("norm", loc=float(r.normal()), scale=float(r.uniform(0.5, 2))))
            elif kind == "uniform":
                nodes.append(m.Distribution("uniform", loc=float(r.normal()), scale=float(r.uniform(0.5, 2))))
            elif kind == "expon":
                nodes.append(m.Distribution("expon", scale=float(r.uniform(0.5, 2))))
            else:
                nodes.append(m.Distribution("triang", c=float(r.uniform(0.1, 0.9)), loc=0.0, scale=float(r.uniform(1, 3))))
        if r.random() < 0.5:  # a composite distribution: its scale is another node's value
            nodes.append(m.Distribution("norm", loc=0.0, scale=m.Abs(nodes[0]) + 0.5))
        for _ in range(int(r.integers(4, 14))):
            a = nodes[int(r.integers(len(nodes)))]
            b = nodes[int(r.integers(len(nodes)))] if r.random() < 0.7 else float(r.normal())
            op = r.choice(["add", "mul", "sub", "gt", "neg", "abs", "max", "square"])
            if op == "add":
                nodes.append(a + b)
            elif op == "mul":
                nodes.append(a * b)
            elif op == "sub":
                nodes.append(b - a)
            elif op == "gt":
                nodes.append((a > b) * 1.0)  # (booleans only feed float arithmetic: bool ** 2 etc. is int64 in NumPy)
            elif op == "neg":
                nodes.append(-a)
            elif op == "abs":
                nodes.append(m.Abs(a))
            elif op == "max":
                nodes.append(m.Max(a, b))
            else:
                nodes.append(a ** 2)
        sink = nodes[-1]
        for extra in nodes[-4:-1]:
            sink = sink + extra
        return sink

    fake_device.install(monkeypatch)
    results = []
    for device_form in (False, True):
        monkeypatch.setattr(fake_device, "DEVICE_FORM", device_form)
        sink = build()
        d = sink.num_distribution_nodes()
        q = np.random.default_rng(3000 + seed).random((48, d))
        gc = [] if seed % 2 else None
        sink.sample_from_quantiles(q, gc_strategy=gc)
        G = sink.to_graph()
        import networkx as nx

        order = list(nx.topological_sort(G))
        results.append([np.array(n.samples_) if getattr(n, "samples_", None) is not None else None
                        for n in sorted(order, key=lambda n: n._id)])
    assert len(results[0]) == len(results[1])
    for a, b in zip(*results):
        assert (a is None) == (b is None)
        if a is not None:
            np.testing.assert_array_equal(a, b)
