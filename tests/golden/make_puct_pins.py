"""Pins the pUCT formula against the reference's own scalar implementation.

The batched search arithmetic lives in mctx (absent), but the reference tree carries an independent scalar pUCT:
/root/reference/muax/frameworks/acme/tf/mcts/search.py:463-497 `puct(node, c_base, c_init)`:

    pb_c   = log((n + c_base + 1) / c_base) + c_init
    scores = value + pb_c * prior * sqrt(n) / (visits + 1)        (raw Q; mctx normalises Q first)
    argmax with illegal (zero-prior) actions masked

This script lifts `puct`, `argmax` and `check_numerics` out of that file with `ast` (nothing is copied into the repo),
runs them on random nodes and records both the score vector (captured at the `argmax` call) and the chosen action.
tests/test_oracle_golden.py::test_puct_pins then requires the oracle's policy-score term + the same raw Q to reproduce
them.  Run in the build container only (needs /root/reference):  python tests/golden/make_puct_pins.py
"""
import ast
import os
import types

import numpy as np

SRC = "/root/reference/muax/frameworks/acme/tf/mcts/search.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "puct_pins.npz")


def lift(names):
    tree = ast.parse(open(SRC).read())
    picked = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            node.returns = None
            for a in node.args.args:
                a.annotation = None
            picked.append(node)
    ns = {"np": np}
    exec(compile(ast.Module(body=picked, type_ignores=[]), SRC, "exec"), ns)
    return ns


def main():
    ns = lift({"puct", "argmax", "check_numerics"})
    captured = []
    ref_argmax = ns["argmax"]

    def spy(values):
        captured.append(np.array(values, dtype=np.float64))
        return ref_argmax(values)

    ns["argmax"] = spy
    rng = np.random.default_rng(11)
    cases = 512
    A = 6
    visits = np.zeros((cases, A), np.int32)
    priors = np.zeros((cases, A), np.float64)
    values = np.zeros((cases, A), np.float64)
    node_visits = np.zeros(cases, np.int32)
    scores = np.zeros((cases, A), np.float64)
    actions = np.zeros(cases, np.int32)
    for i in range(cases):
        visits[i] = rng.integers(0, 40, A)
        node_visits[i] = 1 + visits[i].sum()       # mctx invariant
        p = rng.dirichlet(np.full(A, 0.7))
        if i % 7 == 0:
            p[rng.integers(0, A)] = 0.0            # an illegal action: zero prior -> masked by the reference
            p /= p.sum()
        priors[i] = p
        values[i] = rng.normal(0, 1, A)
        node = types.SimpleNamespace(visit_count=int(node_visits[i]), children={
            a: types.SimpleNamespace(value=float(values[i, a]), prior=float(priors[i, a]), visit_count=int(visits[i, a]))
            for a in range(A)})
        actions[i] = ns["puct"](node)
        scores[i] = captured[-1]
    np.savez_compressed(OUT, node_visits=node_visits, visits=visits, priors=priors, values=values, scores=scores,
                        actions=actions, c_base=np.float64(19652), c_init=np.float64(1.25))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
