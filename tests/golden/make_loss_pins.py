"""Pins the learner's loss against the REFERENCE'S OWN `default_loss_fn`.

`muax/frameworks/coax/loss.py:10-78` (the released loss; HEAD's muax/loss.py:9-88 is the same body minus the `/ L`)
and `scalar_to_support` / `_scaling` / `scale_gradient` (muax/frameworks/coax/utils.py) are executed in place — their
function bodies are lifted with `ast`, nothing is copied into this repo — with NumPy stand-ins for the handful of
jax / optax calls they make (`fori_loop` = a Python loop, `stop_gradient` = identity, `optax.softmax_cross_entropy`
= its documented formula, `jax.nn.one_hot`, `tree_leaves`).  The three networks are supplied by a NumPy float64
evaluation of the stock haiku MLPs (muax/nn.py:59-115).  Stored: parameters, a batch, the loss value.

Run in the build container only (needs /root/reference):  python tests/golden/make_loss_pins.py
"""
import ast
import os
import types

import numpy as np

REF = "/root/reference/muax/frameworks/coax"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "loss_pins.npz")


def lift(path, names, ns):
    tree = ast.parse(open(path).read())
    picked = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            node.decorator_list = []  # drop @jax.jit
            picked.append(node)
    exec(compile(ast.Module(body=picked, type_ignores=[]), path, "exec"), ns)
    return ns


def stand_ins():
    jax = types.SimpleNamespace()
    jax.lax = types.SimpleNamespace(stop_gradient=lambda x: x)

    def fori_loop(lo, hi, body, init):
        val = init
        for i in range(lo, hi):
            val = body(i, val)
        return val

    jax.lax.fori_loop = fori_loop
    jax.nn = types.SimpleNamespace(one_hot=lambda idx, n: np.eye(n)[np.asarray(idx)])

    def tree_leaves(t):
        if isinstance(t, dict):
            return [leaf for v in t.values() for leaf in tree_leaves(v)]
        return [t]

    jax.tree_util = types.SimpleNamespace(tree_leaves=tree_leaves)

    def softmax_cross_entropy(logits, labels):
        z = logits - logits.max(-1, keepdims=True)
        return -(labels * (z - np.log(np.exp(z).sum(-1, keepdims=True)))).sum(-1)

    optax = types.SimpleNamespace(softmax_cross_entropy=softmax_cross_entropy)
    return jax, optax


def elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))


def min_max(s):
    lo, hi = s.min(-1, keepdims=True), s.max(-1, keepdims=True)
    scale = hi - lo
    scale = np.where(scale < 1e-5, scale + 1e-5, scale)
    return (s - lo) / scale


def mlp(params, prefix, first, n_layers, x):
    for l in range(n_layers):
        name = f"{prefix}/linear" if first + l == 0 else f"{prefix}/linear_{first + l}"
        x = x @ params[name]["w"] + params[name]["b"]
        if l < n_layers - 1:
            x = elu(x)
    return x


def main():
    rng = np.random.default_rng(21)
    obs_dim, E, A, S, H, B, L = 4, 8, 2, 10, 16, 12, 5
    F = 2 * S + 1

    def lin(i, o):
        return {"w": rng.standard_normal((i, o)) / np.sqrt(i), "b": rng.standard_normal(o) * 0.1}

    params = {
        "representation": {"representation/linear": lin(obs_dim, E)},
        "prediction": {"prediction/linear": lin(E, H), "prediction/linear_1": lin(H, F),
                       "prediction/linear_2": lin(E, H), "prediction/linear_3": lin(H, A)},
        "dynamic": {"dynamic/linear": lin(E + A, H), "dynamic/linear_1": lin(H, E),
                    "dynamic/linear_2": lin(E + A, H), "dynamic/linear_3": lin(H, F)},
    }
    jax, optax = stand_ins()
    ns = {"jnp": np, "np": np, "jax": jax, "optax": optax}
    lift(os.path.join(REF, "utils.py"), {"_scaling", "scalar_to_support", "scale_gradient"}, ns)
    lift(os.path.join(REF, "loss.py"), {"default_loss_fn"}, ns)

    inst = types.SimpleNamespace(_support_size=S)
    inst._repr_apply = lambda p, obs: min_max(mlp(p, "representation", 0, 1, obs))                    # nn.py:67-70
    inst._pred_apply = lambda p, s: (mlp(p, "prediction", 0, 2, s), mlp(p, "prediction", 2, 2, s))    # nn.py:86-90

    def dy(p, s, a):                                                                                  # nn.py:105-115
        sa = np.concatenate([s, np.eye(A)[a]], axis=-1)
        return mlp(p, "dynamic", 2, 2, sa), min_max(mlp(p, "dynamic", 0, 2, sa))

    inst._dy_apply = dy
    batch = types.SimpleNamespace(obs=rng.standard_normal((B, L, obs_dim)), a=rng.integers(0, A, (B, L)),
                                  r=rng.standard_normal((B, L)), Rn=rng.standard_normal((B, L)) * 5,
                                  pi=rng.dirichlet(np.ones(A), (B, L)))
    keep = {k: np.array(v) for k, v in vars(batch).items()}  # the reference overwrites batch.r / batch.Rn
    loss = float(ns["default_loss_fn"](inst, params, batch))
    out = {f"batch_{k}": v for k, v in keep.items()}
    for g, tree in params.items():
        for m, leaves in tree.items():
            for k, v in leaves.items():
                out[f"param|{g}|{m}|{k}"] = v
    sts_x = np.concatenate([rng.uniform(-300, 300, 200), rng.uniform(-2, 2, 100), [0.0, 1.0, -1.0]])
    out.update(loss=loss, support_size=S, sts_x=sts_x, sts_y=ns["scalar_to_support"](sts_x, S))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, "loss", loss)


if __name__ == "__main__":
    main()
