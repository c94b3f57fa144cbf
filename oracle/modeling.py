"""CPU restatement of the reference's graph evaluation.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``evaluate(sink, quantiles, correlate=...)`` follows ``Node.sample_from_quantiles``
(src/probabilit/modeling.py:495-614) on any duck-typed node graph that exposes the reference's
structure (``_id``, ``get_parents()``, ``is_leaf``, ``_correlations`` and, per class *name*,
``value`` / ``distr, args, kwargs`` / ``parents`` / ``parent``), i.e. both the unmodified reference's
nodes and probabilit_b200.modeling's.  It makes the same NumPy / SciPy calls as the reference:

* column assignment: initial sampling nodes by ``_id`` (:521-538), the remaining distributions in
  ``networkx.topological_sort`` order of the MultiDiGraph built from the depth-first edge list
  (:586-592, :663-683);
* ``Distribution._sample`` = ``getattr(scipy.stats, distr)(*args, **kwargs).ppf(q)`` (:795-812);
* ``Constant._sample`` = ``np.ones(size, dtype=type(value)) * value`` (:760-763);
* transforms = the NumPy callables of the op table (:962-1169);
* per-node finite check on numeric dtypes (:600-606).

Returns {node: samples} for every node (gc_strategy=None semantics).

Pinned against the unmodified reference: tests/golden/graph_reference.npz (made by
tests/golden/make_graph_golden.py by importing /root/reference in the build container) and the
README.md:21-76 golden values (tests/test_oracle_graph.py).
"""
import functools
import itertools
import operator

import networkx as nx
import numpy as np
from scipy import stats

NUMPY_OPS = {
    # variadic (functools.reduce)            modeling.py:962-984
    "Add": operator.add, "Multiply": operator.mul, "Max": np.maximum, "Min": np.minimum,
    "All": np.logical_and, "Any": np.logical_or,
    # binary                                  modeling.py:1012-1062, :1121
    "FloorDivide": np.floor_divide, "Mod": np.mod, "Divide": operator.truediv, "Power": operator.pow,
    "Subtract": operator.sub, "Equal": np.equal, "NotEqual": np.not_equal, "LessThan": operator.lt,
    "LessThanOrEqual": operator.le, "GreaterThan": operator.gt, "GreaterThanOrEqual": operator.ge,
    "IsClose": np.isclose, "Arctan2": np.arctan2,
    # unary                                   modeling.py:1075-1169
    "Negate": operator.neg, "Abs": operator.abs, "Log": np.log, "Exp": np.exp, "Floor": np.floor,
    "Ceil": np.ceil, "Sign": np.sign, "Sqrt": np.sqrt, "Square": np.square, "Log10": np.log10,
    "Sin": np.sin, "Cos": np.cos, "Tan": np.tan, "Arcsin": np.arcsin, "Arccos": np.arccos,
    "Arctan": np.arctan, "Sinh": np.sinh, "Cosh": np.cosh, "Tanh": np.tanh, "Arcsinh": np.arcsinh,
    "Arccosh": np.arccosh, "Arctanh": np.arctanh,
}
VARIADIC = {"Add", "Multiply", "Max", "Min", "All", "Any"}
UNARY = {"Negate", "Abs", "Log", "Exp", "Floor", "Ceil", "Sign", "Sqrt", "Square", "Log10", "Sin", "Cos", "Tan",
         "Arcsin", "Arccos", "Arctan", "Sinh", "Cosh", "Tanh", "Arcsinh", "Arccosh", "Arctanh"}


def kind(node):
    names = {c.__name__ for c in type(node).__mro__}
    if "Constant" in names:
        return "constant"
    if "AbstractDistribution" in names:
        return "distribution"
    return "transform"


def dfs_nodes(sink):
    """modeling.py:403-420"""
    stack = [sink]
    while stack:
        node = stack.pop()
        yield node
        stack.extend(node.get_parents())


def build_graph(sink):
    """modeling.py:663-683"""
    nodes = list(dfs_nodes(sink))
    if len(nodes) == 1:
        G = nx.MultiDiGraph()
        G.add_node(sink)
        return G
    return nx.MultiDiGraph([(p, n) for n in nodes for p in n.get_parents() if not n.is_leaf])


def is_initial_sampling_node(node):
    if kind(node) != "distribution":
        return False
    return not any(kind(a) == "distribution" for a in set(dfs_nodes(node)) - {node})


def sample_distribution(node, q, samples):
    name = type(node).__name__
    if name == "EmpiricalDistribution":  # modeling.py:841-842
        return np.quantile(a=node.data, q=q, **node.kwargs)
    if name == "CumulativeDistribution":  # modeling.py:880-882
        return np.interp(x=q, xp=node.q, fp=node.cumulatives)
    if name == "DiscreteDistribution":  # modeling.py:910-913
        idx = np.searchsorted(np.cumsum(node.probabilities), v=q, side="right")
        return node.values[idx]

    def unpack(arg):
        return samples[arg] if hasattr(arg, "get_parents") else arg

    args = tuple(unpack(a) for a in node.args)
    kwargs = {k: unpack(v) for k, v in node.kwargs.items()}
    return getattr(stats, node.distr)(*args, **kwargs).ppf(q)  # modeling.py:805-808


def sample_transform(node, samples):
    name = type(node).__name__
    parents = [samples[p] for p in node.get_parents()]
    if name == "Avg":
        return np.average(np.vstack(parents), axis=0)  # :986-990
    if name == "NoOp":
        return None
    op = NUMPY_OPS[name]
    if name in VARIADIC:
        return functools.reduce(op, parents)
    return op(*parents)


def build_corrmat(correlations):
    """utils.py:92-115"""
    k = 1 + max(max(idx) for idx, _ in correlations)
    C = np.eye(k, dtype=float)
    for idx, block in correlations:
        C[np.ix_(idx, idx)] = block
    return C


def evaluate(sink, quantiles, correlate=None):
    """-> {node: samples}.  ``correlate(X, C) -> X'`` induces the correlations (the reference uses
    ``correlator().set_target(nearest_correlation_matrix(C))(X)``, :571-581); required only when the
    graph carries ``.correlate(...)`` declarations."""
    G = build_graph(sink)
    assert nx.is_directed_acyclic_graph(G)
    members = set(dfs_nodes(sink))
    size, n_dim = quantiles.shape
    assert n_dim == sum(1 for m in members if kind(m) == "distribution")
    columns = iter(list(quantiles.T))
    samples = {}

    isns = sorted((m for m in members if is_initial_sampling_node(m)), key=lambda m: m._id)
    for node in isns:  # :527-538
        for anc in nx.topological_sort(G.subgraph(nx.ancestors(G, node))):
            samples[anc] = (np.ones(size, dtype=type(anc.value)) * anc.value if kind(anc) == "constant"
                            else sample_transform(anc, samples))
        samples[node] = sample_distribution(node, next(columns), samples)

    correlations = []
    for node in members:
        correlations.extend(node._correlations)
    for variables, _ in correlations:
        for v in variables:
            if v not in isns:
                raise ValueError(f"Cannot correlate variable: {v}")
    var_sets = [set(v) for v, _ in correlations]
    for s1, s2 in itertools.combinations(var_sets, 2):
        if len(s1 & s2) > 1:
            raise ValueError(f"Correlations specified more than once: {s1 & s2}")
    variables = sorted(functools.reduce(set.union, var_sets, set()), key=lambda m: m._id)
    index = {v: i for i, v in enumerate(variables)}
    if correlations:
        C = build_corrmat([(tuple(index[v] for v in vs), m) for vs, m in correlations])
        X = np.vstack([samples[v] for v in variables]).T
        Xc = correlate(X, C)
        for v, col in zip(variables, Xc.T):
            samples[v] = np.copy(col)

    for node in nx.topological_sort(G):  # :586-606
        if node not in samples:
            k = kind(node)
            if k == "constant":
                samples[node] = np.ones(size, dtype=type(node.value)) * node.value
            elif k == "distribution":
                samples[node] = sample_distribution(node, next(columns), samples)
            else:
                samples[node] = sample_transform(node, samples)
        s = samples[node]
        if s is not None and np.issubdtype(s.dtype, np.number) and not np.all(np.isfinite(s)):
            raise ValueError(f"Sampling this node gave non-finite values: {node}\n{s}")
    return samples
