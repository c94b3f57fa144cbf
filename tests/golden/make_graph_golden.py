"""Golden vectors for the modeling graph from the UNMODIFIED reference (build container only).

    python tests/golden/make_graph_golden.py

Builds every recipe of tests/graph_recipes.py from the reference's own classes
(/root/reference/src/probabilit/modeling.py, with empty ``cvxpy`` / ``seaborn`` stubs as in
SURVEY.md section 8c), draws the quantiles with the reference's default generator
(``check_random_state(seed).random((n, d))``, modeling.py:485-486), runs the reference's
``sample_from_quantiles`` and stores quantiles + the named nodes' samples in
tests/golden/graph_reference.npz.  ``nearest_correlation_matrix`` needs cvxpy (absent), so it is
patched to the identity for the (already valid) matrices used here -- the work-around SURVEY.md
section 8c validated.  Also stores the README.md:21-76 example outputs via the public ``sample``.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def main():
    for name in ("cvxpy", "seaborn"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, "/root/reference/src")
    import probabilit.modeling as ref

    ref.nearest_correlation_matrix = lambda m, **kw: m.copy()
    import graph_recipes

    out = {}
    for name, (recipe, n) in graph_recipes.RECIPES.items():
        sink, named = recipe(ref)
        d = sink.num_distribution_nodes()
        q = np.random.RandomState(len(name) * 7 + n).random((n, d))
        sink.sample_from_quantiles(q)
        out[f"{name}__quantiles"] = q
        for label, node in named:
            out[f"{name}__{label}"] = np.asarray(node.samples_)
        print(name, "d =", d, "n =", n, "ok")

    # README examples through the public API (seeded pseudo-random draw)
    s, _ = graph_recipes.height(ref)
    out["readme__height"] = s.sample(999, random_state=0)
    s, _ = graph_recipes.birds(ref)
    out["readme__birds"] = s.sample(9, random_state=0)
    s, _ = graph_recipes.mutual_fund(ref)
    out["readme__mutual_fund"] = s.sample(999, random_state=42)
    print("README: height mean", out["readme__height"].mean(), "fund mean/std",
          out["readme__mutual_fund"].mean(), out["readme__mutual_fund"].std())
    np.savez_compressed(os.path.join(HERE, "graph_reference.npz"), **out)


if __name__ == "__main__":
    main()
