"""Pins the few hot-path functions whose arithmetic DOES live in the reference tree.

The search itself is mctx (absent), but the value/reward transform and the min-max normaliser are plain
functions in /root/reference/muax.  jax is not installable here, so this script lifts the function
bodies out of the reference files with `ast` (no copy is kept in this repo), executes them with NumPy
float32 standing in for `jax.numpy`, and stores input/output vectors:

    muax/utils.py:65-76    _scaling, _inv_scaling
    muax/utils.py:94-102   support_to_scalar
    muax/nn.py:37-44       min_max_normalize
    muax/frameworks/acme/tf/mcts/search.py:475  pb_c = log((n + c_base + 1) / c_base) + c_init

Run in the build container only (needs /root/reference):  python tests/golden/make_reference_pins.py
"""
import ast
import os

import numpy as np

REF = "/root/reference/muax"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_pins.npz")


def lift(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    picked = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            node.decorator_list = []  # drop @jax.jit
            picked.append(node)
    mod = ast.Module(body=picked, type_ignores=[])
    ns = {"jnp": np, "np": np}
    exec(compile(mod, path, "exec"), ns)
    return ns


def main():
    rng = np.random.default_rng(7)
    u = lift(os.path.join(REF, "utils.py"), {"_scaling", "_inv_scaling", "support_to_scalar"})
    n = lift(os.path.join(REF, "nn.py"), {"min_max_normalize"})
    x = np.concatenate([rng.uniform(-10, 10, 4000), rng.uniform(-0.01, 0.01, 1000), [0.0, 1.0, -1.0, 10.0, -10.0]]
                       ).astype(np.float32)
    inv = u["_inv_scaling"](x).astype(np.float32)
    S = 10
    logits = rng.standard_normal((256, 2 * S + 1)).astype(np.float32) * 3
    e = np.exp(logits - logits.max(-1, keepdims=True))
    probs = (e / e.sum(-1, keepdims=True)).astype(np.float32)
    # support_to_scalar uses jnp.sum over the last axis; evaluate it with a left-to-right float32 sum

    class _Seq:
        def __getattr__(self, k):
            return getattr(np, k)

        @staticmethod
        def sum(a, axis=-1):
            acc = np.zeros(a.shape[:-1], np.float32)
            for i in range(a.shape[-1]):
                acc = (acc + a[..., i].astype(np.float32)).astype(np.float32)
            return acc

    u["support_to_scalar"].__globals__["jnp"] = _Seq()
    sts = u["support_to_scalar"](probs, S).astype(np.float32)
    u["support_to_scalar"].__globals__["jnp"] = np
    s = rng.standard_normal((128, 8)).astype(np.float32)
    s[3] = 0.5           # degenerate row: scale < 1e-5 branch
    s[4, :] = s[4, 0] + np.arange(8, dtype=np.float32) * 1e-7
    mm = n["min_max_normalize"](s).astype(np.float32)
    visits = np.arange(0, 300, dtype=np.int64)
    pb_c = (np.log((visits + 19652 + 1) / 19652) + 1.25)  # float64, as acme/tf/mcts/search.py:475 computes it
    np.savez_compressed(OUT, inv_x=x, inv_y=inv, sts_probs=probs, sts_y=sts, mm_x=s, mm_y=mm,
                        pbc_visits=visits.astype(np.int32), pbc_y=pb_c)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
