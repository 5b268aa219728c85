"""Regenerates tests/golden/*.npz.

The reference (bwfbowen/muax + mctx + jax) cannot be imported in this image, so these vectors come from
the batched NumPy restatement (oracle/np_mctx.py, ExactMath back-end) and are accepted only when the
independently written scalar C restatement (oracle/mz_oracle.c) reproduces every array bit-for-bit.
"Parity unpinned": they pin the two restatements and the CUDA path to each other, not to mctx itself.
If a JAX+mctx environment is ever available, tools/dump_mctx_golden.py writes files of the same layout.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import c_oracle, np_mctx, threefry  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
TREE_FIELDS = ("node_visits", "parents", "action_from_parent", "children_index", "children_visits", "raw_values",
               "node_values", "children_prior_logits", "children_values", "children_rewards", "children_discounts",
               "embeddings", "sim_depth", "root_noise")
STACKS = ("repr", "pred_v", "pred_pi", "dyn_ns", "dyn_r")


def haiku_linear(rng, fan_in, fan_out, bias_scale=0.0):
    """hk.Linear default init: w ~ TruncatedNormal(stddev=1/sqrt(fan_in)) cut at 2 sigma, b = 0."""
    w = rng.standard_normal((fan_in, fan_out))
    while (np.abs(w) > 2).any():
        bad = np.abs(w) > 2
        w[bad] = rng.standard_normal(int(bad.sum()))
    b = rng.standard_normal(fan_out) * bias_scale
    return (w / np.sqrt(fan_in)).astype(np.float32), b.astype(np.float32)


def make_nets(rng, obs_dim, E, A, F, hidden=(16,), bias_scale=0.0):
    def mlp(i, o):
        dims = [i, *hidden, o]
        return [haiku_linear(rng, a, b, bias_scale) for a, b in zip(dims[:-1], dims[1:])]
    return dict(repr=[haiku_linear(rng, obs_dim, E, bias_scale)], pred_v=mlp(E, F), pred_pi=mlp(E, A),
                dyn_ns=mlp(E + A, E), dyn_r=mlp(E + A, F))


CASES = {
    # name: (obs_dim, E, A, S, hidden, B, seed, search kwargs, extras)
    "c1_muzero_seed0": (4, 8, 2, 10, (16,), 4, 0, dict(policy=0, num_simulations=50), dict(noise="dirichlet")),
    "c1_muzero_seed1": (4, 8, 2, 10, (16,), 4, 1, dict(policy=0, num_simulations=50), dict(noise="dirichlet")),
    "c1_muzero_seed42": (4, 8, 2, 10, (16,), 4, 42, dict(policy=0, num_simulations=50), dict(noise="dirichlet")),
    "c1_muzero_nofrac_seed0": (4, 8, 2, 10, (16,), 4, 0,
                               dict(policy=0, num_simulations=50, dirichlet_fraction=0.0, temperature=0.0),
                               dict(noise="dirichlet")),
    "c1_muzero_sampler_seed0": (4, 8, 2, 10, (16,), 4, 0, dict(policy=0, num_simulations=50), dict(noise=None)),
    "c1_muzero_partitionable_seed1": (4, 8, 2, 10, (16,), 4, 1, dict(policy=0, num_simulations=50, prng_mode=1),
                                      dict(noise="dirichlet")),
    "c1_gumbel_act_seed0": (4, 8, 2, 10, (16,), 4, 0, dict(policy=1, qtransform=0, num_simulations=50), dict(noise=None)),
    "c1_gumbel_mctx_seed42": (4, 8, 2, 10, (16,), 4, 42, dict(policy=1, qtransform=1, num_simulations=32),
                              dict(noise="gumbel")),
    "lunar_muzero_invalid_seed1": (8, 16, 4, 10, (16,), 3, 1,
                                   dict(policy=0, num_simulations=40, temperature=0.25, discount=0.997),
                                   dict(noise="dirichlet", invalid=True, bias_scale=0.1)),
    "lunar_gumbel_invalid_seed0": (8, 16, 4, 10, (16,), 3, 0,
                                   dict(policy=1, qtransform=1, num_simulations=32, max_considered=3),
                                   dict(noise=None, invalid=True, bias_scale=0.1)),
    "lunar_notebook_arch_seed0": (8, 10, 4, 20, (64, 64, 16), 2, 0,
                                  dict(policy=0, num_simulations=30, discount=0.999, repr_minmax=0, dyn_minmax=0),
                                  dict(noise="dirichlet", bias_scale=0.05)),
    "maxdepth_muzero_seed0": (4, 8, 3, 5, (16,), 3, 0, dict(policy=0, num_simulations=30, max_depth=4),
                              dict(noise="dirichlet", bias_scale=0.1)),
    "atari18_muzero_seed0": (12, 32, 18, 10, (32,), 2, 0, dict(policy=0, num_simulations=24, activation=1),
                             dict(noise=None, bias_scale=0.1, shard=(5, 2))),
    "atari18_gumbel_seed1": (12, 32, 18, 10, (32,), 2, 1, dict(policy=1, qtransform=1, num_simulations=24),
                             dict(noise=None, bias_scale=0.1, shard=(5, 3))),
}


def build_case(name):
    obs_dim, E, A, S, hidden, B, seed, kw, ex = CASES[name]
    rng = np.random.default_rng(seed)
    nets = make_nets(rng, obs_dim, E, A, 2 * S + 1, hidden, ex.get("bias_scale", 0.0))
    obs = rng.standard_normal((B, obs_dim)).astype(np.float32)
    key = threefry.PRNGKey(seed)
    invalid = None
    if ex.get("invalid"):
        invalid = (rng.random((B, A)) < 0.35).astype(np.uint8)
        invalid[:, int(rng.integers(A))] = 0
    noise = None
    if ex.get("noise") == "dirichlet":
        noise = rng.dirichlet([0.3] * A, size=B).astype(np.float32)
    elif ex.get("noise") == "gumbel":
        noise = rng.gumbel(size=(B, A)).astype(np.float32)
    kw = dict(kw, support_size=S)
    if "shard" in ex:  # rows [offset, offset+B) of a larger global batch: exercises global-index PRNG
        kw["global_batch"], kw["batch_offset"] = ex["shard"]
    return nets, obs, key, invalid, noise, kw


def main():
    for name in CASES:
        nets, obs, key, invalid, noise, kw = build_case(name)
        ref = np_mctx.act(nets, key, obs=obs, invalid=invalid, noise=noise, **kw)
        chk = c_oracle.search(nets, key, obs=obs, invalid=invalid, noise=noise, **kw)
        for f in ("action", "action_weights", "root_value") + TREE_FIELDS:
            if not np.array_equal(ref[f], chk[f]):
                raise SystemExit(f"{name}: NumPy and C restatements disagree on {f}")
        blob = {}
        for s in STACKS:
            for l, (w, b) in enumerate(nets[s]):
                blob[f"net.{s}.{l}.w"] = w
                blob[f"net.{s}.{l}.b"] = b
        blob["in.obs"] = obs
        blob["in.key"] = key
        if invalid is not None:
            blob["in.invalid"] = invalid
        if noise is not None:
            blob["in.noise"] = noise
        for k, v in kw.items():
            blob[f"cfg.{k}"] = np.asarray(v)
        for f in ("action", "action_weights", "root_value") + TREE_FIELDS:
            blob[f"out.{f}"] = ref[f]
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **blob)
        print(f"{name}: ok  mean depth {ref['sim_depth'].mean():.2f}  {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
