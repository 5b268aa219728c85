#!/usr/bin/env python
"""Writes tests/golden/mctx_<case>.npz: the REAL mctx (+ jax.random) run on the inputs of the committed golden cases.

The committed goldens (tests/golden/<case>.npz) come from this repo's own restatements of mctx (oracle/np_mctx.py,
oracle/mz_oracle.c) because neither jax nor mctx can be installed in the build image: parity is "unpinned".  Run this
script anywhere `import jax, mctx` works (CPU is enough; `pip install jax mctx`, optionally the reference muax):

    JAX_PLATFORMS=cpu python tools/dump_mctx_golden.py            # all cases
    JAX_PLATFORMS=cpu python tools/dump_mctx_golden.py c1_muzero_seed0

and commit the mctx_*.npz files it writes next to the goldens.  tests/test_mctx_golden.py (skipped while no such file
exists) then holds both CPU restatements and the CUDA engines to them: integer tree state exact, floats within 1e-5
(BASELINE.json north_star).  The file layout is the goldens' (net.* / cfg.* / in.* / out.*) minus `out.sim_depth`
(mctx does not report per-simulation path lengths).

What runs: the reference call chain muax/model.py:222-282 (`_plan` -> `_root_inference` / `_recurrent_inference`) with
the declarative MLP stacks of the golden file evaluated in jax.numpy, then mctx:
  * Gumbel policy, noise drawn by mctx            -> mctx.gumbel_muzero_policy (muax/policy.py:33-47);
  * MuZero policy / Gumbel with INJECTED noise    -> the body of mctx.muzero_policy / gumbel_muzero_policy re-assembled
    from mctx's own public pieces exactly as the reference's in-tree mirror does
    (muax/frameworks/acme/jax/diffusion_muzero/policy.py:62-139), with the injected Dirichlet / Gumbel array in place
    of the sampled one (jax.random.dirichlet is not reproducible off-XLA, so parity cases inject it);
  * cases whose noise came from this repo's own Dirichlet sampler (in.noise absent, MuZero policy) are skipped.
"""
import functools
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def main():
    import jax
    import jax.numpy as jnp
    import mctx
    from mctx._src import action_selection, qtransforms, search, seq_halving

    from helpers import STACKS, load_golden

    try:  # the reference's own transforms when the package is importable ...
        from muax.nn import min_max_normalize
        from muax.utils import support_to_scalar
    except Exception:  # ... else the formulas of muax/nn.py:37-44 and muax/utils.py:70-102
        def min_max_normalize(s):
            s_min, s_max = s.min(axis=1, keepdims=True), s.max(axis=1, keepdims=True)
            scale = s_max - s_min
            scale = jnp.where(scale < 1e-5, scale + 1e-5, scale)
            return (s - s_min) / scale

        def support_to_scalar(probs, support_size):
            x = jnp.sum(jnp.arange(-support_size, support_size + 1) * probs, axis=-1)
            eps = 0.001
            return jnp.sign(x) * (((jnp.sqrt(1 + 4 * eps * (jnp.abs(x) + 1 + eps)) - 1) / (2 * eps)) ** 2 - 1)

    names = sys.argv[1:] or sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                                   if not p.endswith("_pins.npz") and not os.path.basename(p).startswith("mctx_"))
    for name in names:
        nets, inp, cfg, _ = load_golden(name)
        policy, qt = cfg.get("policy", 0), cfg.get("qtransform", 0)
        if policy == 0 and inp["noise"] is None and cfg.get("dirichlet_fraction", 0.25) != 0.0:
            print(f"{name}: skipped (its Dirichlet noise came from this repo's own sampler)")
            continue
        jax.config.update("jax_threefry_partitionable", bool(cfg.get("prng_mode", 0)))
        S, gamma = cfg.get("support_size", 10), cfg.get("discount", 0.99)
        act = jax.nn.elu if cfg.get("activation", 0) == 0 else jax.nn.relu
        A = nets["pred_pi"][-1][0].shape[1]
        W = {k: [(jnp.asarray(w), jnp.asarray(b)) for w, b in v] for k, v in nets.items()}

        def mlp(layers, x):
            for i, (w, b) in enumerate(layers):
                x = x @ w + b
                if i + 1 < len(layers):
                    x = act(x)
            return x

        def root_inference(obs):  # muax/model.py:251-263
            s = mlp(W["repr"], obs)
            if cfg.get("repr_minmax", 1):
                s = min_max_normalize(s)
            v = support_to_scalar(jax.nn.softmax(mlp(W["pred_v"], s)), S).flatten()
            return mctx.RootFnOutput(prior_logits=mlp(W["pred_pi"], s), value=v, embedding=s)

        def recurrent_fn(params, rng_key, action, embedding):  # muax/model.py:265-282, muax/nn.py:93-115
            sa = jnp.concatenate([embedding, jax.nn.one_hot(action, A)], axis=1)
            ns = mlp(W["dyn_ns"], sa)
            if cfg.get("dyn_minmax", 1):
                ns = min_max_normalize(ns)
            r = support_to_scalar(jax.nn.softmax(mlp(W["dyn_r"], sa)), S).flatten()
            v = support_to_scalar(jax.nn.softmax(mlp(W["pred_v"], ns)), S).flatten()
            out = mctx.RecurrentFnOutput(reward=r, discount=jnp.ones_like(r) * gamma,
                                         prior_logits=mlp(W["pred_pi"], ns), value=v)
            return out, ns

        qtransform = (qtransforms.qtransform_by_parent_and_siblings, qtransforms.qtransform_completed_by_mix_value)[qt]
        key = jnp.asarray(inp["key"], jnp.uint32)
        root = root_inference(jnp.asarray(inp["obs"]))
        raw_value = np.asarray(root.value)
        invalid = None if inp["invalid"] is None else jnp.asarray(inp["invalid"])
        NS, max_depth = cfg["num_simulations"], cfg.get("max_depth") or None
        tiny = jnp.finfo(jnp.float32).tiny

        def mask(logits):  # mctx._src.policies._mask_invalid_actions
            if invalid is None:
                return logits
            logits = logits - jnp.max(logits, axis=-1, keepdims=True)
            return jnp.where(invalid, jnp.finfo(logits.dtype).min, logits)

        if policy == 0:
            frac = cfg.get("dirichlet_fraction", 0.25)
            rng_key, _, search_key = jax.random.split(key, 3)
            probs = jax.nn.softmax(root.prior_logits)
            noise = jnp.asarray(inp["noise"]) if inp["noise"] is not None else jnp.zeros_like(probs)
            noisy = jnp.log(jnp.maximum((1 - frac) * probs + frac * noise, tiny))
            root = root.replace(prior_logits=mask(noisy))
            interior = functools.partial(action_selection.muzero_action_selection, pb_c_base=cfg.get("pb_c_base", 19652),
                                         pb_c_init=cfg.get("pb_c_init", 1.25), qtransform=qtransform)
            tree = search.search(params=(), rng_key=search_key, root=root, recurrent_fn=recurrent_fn,
                                 root_action_selection_fn=functools.partial(interior, depth=0),
                                 interior_action_selection_fn=interior, num_simulations=NS, max_depth=max_depth,
                                 invalid_actions=invalid)
            weights = tree.summary().visit_probs
            logits = jnp.log(jnp.maximum(weights, tiny))
            logits = (logits - jnp.max(logits, axis=-1, keepdims=True)) / jnp.maximum(tiny, cfg.get("temperature", 1.0))
            action = jax.random.categorical(rng_key, logits)
            root_noise = np.asarray(noise)
        elif inp["noise"] is None:
            out = mctx.gumbel_muzero_policy(params=(), rng_key=key, root=root, recurrent_fn=recurrent_fn,
                                            num_simulations=NS, invalid_actions=invalid, max_depth=max_depth,
                                            qtransform=qtransform,
                                            max_num_considered_actions=cfg.get("max_considered", 16),
                                            gumbel_scale=cfg.get("gumbel_scale", 1.0))
            tree, action, weights = out.search_tree, out.action, out.action_weights
            root_noise = np.asarray(tree.extra_data.root_gumbel)
        else:  # gumbel_muzero_policy's body with the injected root Gumbel
            root = root.replace(prior_logits=mask(root.prior_logits))
            rng_key, _ = jax.random.split(key)
            gumbel = jnp.asarray(inp["noise"])
            mc = cfg.get("max_considered", 16)
            tree = search.search(
                params=(), rng_key=rng_key, root=root, recurrent_fn=recurrent_fn,
                root_action_selection_fn=functools.partial(action_selection.gumbel_muzero_root_action_selection,
                                                           num_simulations=NS, max_num_considered_actions=mc,
                                                           qtransform=qtransform),
                interior_action_selection_fn=functools.partial(
                    action_selection.gumbel_muzero_interior_action_selection, qtransform=qtransform),
                num_simulations=NS, max_depth=max_depth, invalid_actions=invalid,
                extra_data=action_selection.GumbelMuZeroExtraData(root_gumbel=gumbel))
            summary = tree.summary()
            considered_visit = jnp.max(summary.visit_counts, axis=-1, keepdims=True)
            completed_q = jax.vmap(qtransform, in_axes=[0, None])(tree, tree.ROOT_INDEX)
            to_argmax = seq_halving.score_considered(considered_visit, gumbel, root.prior_logits, completed_q,
                                                     summary.visit_counts)
            action = action_selection.masked_argmax(to_argmax, invalid)
            weights = jax.nn.softmax(mask(root.prior_logits + completed_q))
            root_noise = np.asarray(gumbel)

        out = {"out.action": np.asarray(action, np.int32), "out.action_weights": np.asarray(weights, np.float32),
               "out.root_value": raw_value.astype(np.float32), "out.root_noise": root_noise.astype(np.float32)}
        for f in ("node_visits", "parents", "action_from_parent", "children_index", "children_visits", "raw_values",
                  "node_values", "children_prior_logits", "children_values", "children_rewards", "children_discounts",
                  "embeddings"):
            out["out." + f] = np.asarray(getattr(tree, f))
        z = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        z = {k: v for k, v in z.items() if not k.startswith("out.")}
        z.update(out)
        z["meta.versions"] = np.array(f"jax {jax.__version__}, mctx {getattr(mctx, '__version__', '?')}")
        path = os.path.join(GOLDEN, f"mctx_{name}.npz")
        np.savez_compressed(path, **z)
        print(f"{name}: wrote {path}; mean selected-path depth is not reported by mctx")
    _ = STACKS


if __name__ == "__main__":
    main()
