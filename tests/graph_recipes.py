"""Modeling graphs used by the parity tests, written against a *namespace* so that the same
recipe builds the graph from the unmodified reference's classes (tests/golden/make_graph_golden.py,
build container only) and from probabilit_b200.modeling's (the tests).

Each recipe returns (sink, [(name, node), ...]); the named nodes are the ones whose samples are
stored in / compared with tests/golden/graph_reference.npz.
"""
import numpy as np


def height(ns):
    """README.md:21-29 (example 1)."""
    male = ns.Distribution("norm", loc=176, scale=7.1)
    female = ns.Distribution("norm", loc=162.5, scale=7.1)
    stat = male > female
    return stat, [("male", male), ("female", female), ("stat", stat)]


def birds(ns):
    """README.md:55-60 (example 2): composite poisson -> binom."""
    eggs = ns.Distribution("poisson", mu=3)
    survived = ns.Distribution("binom", n=eggs, p=0.4)
    return survived, [("eggs", eggs), ("survived", survived)]


def mutual_fund(ns):
    """README.md:68-76 (example 3)."""
    returns = 0
    interests = []
    for _ in range(20):
        interest = ns.Distribution("norm", loc=1.11, scale=0.15)
        interests.append(interest)
        returns = returns * interest + 1200
    return returns, [("interest0", interests[0]), ("interest19", interests[19]), ("returns", returns)]


def marginals(ns):
    """Every continuous inverse CDF on the device path, with non-default loc/scale."""
    t = ns.Distribution("triang", c=0.3, loc=-1, scale=4)
    g = ns.Distribution("gamma", a=2.5, scale=1.5)
    g1 = ns.Distribution("gamma", 1, loc=0.25)
    e = ns.Distribution("expon", scale=1 / 3)
    ln = ns.Distribution("lognorm", s=0.7, scale=2.0)
    u = ns.Distribution("uniform", loc=2, scale=3)
    total = ns.Add(t, g, g1, e, ln, u)
    return total, [("t", t), ("g", g), ("g1", g1), ("e", e), ("ln", ln), ("u", u), ("total", total)]


def discrete(ns):
    """Discrete inverse CDFs incl. large parameters and a bernoulli."""
    p_small = ns.Distribution("poisson", mu=0.7)
    p_big = ns.Distribution("poisson", mu=250.5)
    b = ns.Distribution("binom", n=40, p=0.3)
    b_big = ns.Distribution("binom", n=5000, p=0.6)
    be = ns.Distribution("bernoulli", p=0.25)
    shifted = ns.Distribution("poisson", 4.0, loc=10)
    total = ns.Add(p_small, p_big, b, b_big, be, shifted)
    return total, [("p_small", p_small), ("p_big", p_big), ("b", b), ("b_big", b_big), ("be", be),
                   ("shifted", shifted), ("total", total)]


def arithmetic(ns):
    """The transform op table (modeling.py:962-1169) on float and bool operands.
    IsClose is left out on purpose: in the reference ``op = np.isclose`` is a plain Python function
    stored as a class attribute, so ``self.op(a, b)`` binds the node as first argument and evaluates
    ``np.isclose(node, a, rtol=b)`` -- all True for finite input.  probabilit_b200 implements the
    intended ``np.isclose(a, b)`` (tested separately)."""
    a = ns.Distribution("norm", loc=0.5, scale=2)
    b = ns.Distribution("uniform", loc=0.1, scale=1.9)
    c = ns.Distribution("expon", scale=2.0)
    nodes = [
        ("a", a), ("b", b), ("c", c),
        ("sub", a - b), ("rsub", 3 - a), ("div", a / b), ("rdiv", 2.0 / b), ("pow", b ** a), ("pow2", a ** 2),
        ("rpow", 2 ** a), ("floordiv", a // b), ("mod", a % b), ("rmod", 7 % b), ("neg", -a), ("abs", abs(a)),
        ("max", ns.Max(a, b, c)), ("min", ns.Min(a, b, 0.3)), ("avg", ns.Avg(a, b, c)),
        ("log", ns.Log(b)), ("exp", ns.Exp(a)), ("floor", ns.Floor(a)), ("ceil", ns.Ceil(a)),
        ("sign", ns.Sign(a)), ("sqrt", ns.Sqrt(c)), ("square", ns.Square(a)), ("log10", ns.Log10(b)),
        ("sin", ns.Sin(a)), ("cos", ns.Cos(a)), ("tan", ns.Tan(b)), ("arcsin", ns.Arcsin(b / 2)),
        ("arccos", ns.Arccos(b / 2)), ("arctan", ns.Arctan(a)), ("arctan2", ns.Arctan2(a, b)),
        ("sinh", ns.Sinh(a)), ("cosh", ns.Cosh(a)), ("tanh", ns.Tanh(a)), ("arcsinh", ns.Arcsinh(a)),
        ("arccosh", ns.Arccosh(b + 1)), ("arctanh", ns.Arctanh(b / 2)),
        ("lt", a < b), ("le", a <= 0.5), ("gt", a > b), ("ge", a >= b), ("eq", ns.Equal(ns.Floor(a), ns.Floor(b))),
        ("ne", ns.NotEqual(ns.Floor(a), 0)),
        ("all", ns.All(a < b, c > 1)), ("any", ns.Any(a < b, c > 1)), ("bool_add", (a < b) + (c > 1)),
        ("bool_mul", (a < b) * (c > 1)), ("bool_float", (a < b) * c + 1),
        ("const_fold", a + (ns.Constant(2) ** 3 - 1) / 4),
    ]
    sink = ns.NoOp(*[n for _, n in nodes])
    return sink, nodes


def composite(ns):
    """Distribution parameters that are themselves nodes (modeling.py:797-803)."""
    mu = ns.Distribution("norm", loc=0, scale=1)
    sigma = ns.Distribution("uniform", loc=0.5, scale=1.0)
    x = ns.Distribution("norm", loc=mu * 2, scale=sigma)
    shape = ns.Distribution("uniform", loc=1, scale=4)
    g = ns.Distribution("gamma", a=shape, scale=ns.Abs(mu) + 0.1)
    mode = ns.Distribution("uniform", loc=0.1, scale=0.8)
    t = ns.Distribution("triang", c=mode, loc=x, scale=2)
    rate = ns.Distribution("gamma", a=3.0)
    counts = ns.Distribution("poisson", mu=rate * 5)
    hits = ns.Distribution("binom", n=counts, p=mode)
    result = x + g + t + hits
    return result, [("mu", mu), ("sigma", sigma), ("x", x), ("shape", shape), ("g", g), ("mode", mode), ("t", t),
                    ("rate", rate), ("counts", counts), ("hits", hits), ("result", result)]


def correlated(ns):
    """.correlate() on initial sampling nodes (modeling.py:540-583)."""
    a = ns.Distribution("uniform")
    b = ns.Distribution("expon")
    c = ns.Distribution("norm", loc=1, scale=2)
    d = ns.Distribution("norm", loc=c, scale=1)  # not an initial sampling node
    expr = (a + b).correlate(a, b, corr_mat=np.array([[1, 0.6], [0.6, 1]]))
    expr = (expr + c + d).correlate(b, c, corr_mat=np.array([[1, -0.4], [-0.4, 1]]))
    return expr, [("a", a), ("b", b), ("c", c), ("d", d), ("expr", expr)]


def tables(ns):
    """The table-lookup distributions (modeling.py:825-927) alone and inside arithmetic."""
    rng = np.random.default_rng(5)
    emp = ns.EmpiricalDistribution(rng.lognormal(size=257))
    dice = ns.EmpiricalDistribution([1, 2, 3, 4, 5, 6], method="closest_observation")
    low = ns.EmpiricalDistribution(rng.normal(size=50), method="lower")
    mid = ns.EmpiricalDistribution(rng.normal(size=51), method="midpoint")
    near = ns.EmpiricalDistribution(rng.normal(size=64), method="nearest")
    cum = ns.CumulativeDistribution([0, 0.2, 0.8, 1], [10, 15, 20, 25])
    disc = ns.DiscreteDistribution([10, 15, 20], probabilities=[0.2, 0.3, 0.5])
    discf = ns.DiscreteDistribution([0.5, -1.25, 3.0, 8.0], probabilities=[0.1, 0.2, 0.3, 0.4])
    cat = ns.DiscreteDistribution(["A", "B", "C", "D", "E", "F"])
    noisy = ns.Distribution("norm", loc=cum, scale=0.1)
    expr = emp + cum * 2 + disc / discf + noisy
    sink = ns.NoOp(expr, dice, cat, low, mid, near)
    return sink, [("emp", emp), ("dice", dice), ("low", low), ("mid", mid), ("near", near), ("cum", cum),
                  ("disc", disc), ("discf", discf), ("cat", cat), ("noisy", noisy), ("expr", expr)]


def four_param(ns):
    """beta (PERT) and truncnorm: the four-parameter inverse CDFs."""
    pert = ns.Distribution("beta", a=3.4, b=2.6, loc=0, scale=10)
    small = ns.Distribution("beta", a=0.6, b=0.8, loc=-1, scale=2)
    tn = ns.Distribution("truncnorm", a=-1.5, b=2.0, loc=3.0, scale=0.5)
    tail = ns.Distribution("truncnorm", a=2.5, b=6.0)
    total = pert + small + tn + tail
    return total, [("pert", pert), ("small", small), ("tn", tn), ("tail", tail), ("total", total)]


RECIPES = {
    "height": (height, 999), "birds": (birds, 2000), "mutual_fund": (mutual_fund, 999),
    "marginals": (marginals, 3000), "discrete": (discrete, 3000), "arithmetic": (arithmetic, 500),
    "composite": (composite, 2000), "correlated": (correlated, 1000), "tables": (tables, 4000),
    "four_param": (four_param, 3000),
}
