"""Shared helpers for the test-suite: golden loading and haiku-style synthetic nets."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STACKS = ("repr", "pred_v", "pred_pi", "dyn_ns", "dyn_r")
OUT_FIELDS = ("action", "action_weights", "root_value", "node_visits", "parents", "action_from_parent",
              "children_index", "children_visits", "raw_values", "node_values", "children_prior_logits",
              "children_values", "children_rewards", "children_discounts", "embeddings")
INT_FIELDS = ("action", "node_visits", "parents", "action_from_parent", "children_index", "children_visits")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not p.endswith("_pins.npz") and not os.path.basename(p).startswith("mctx_"))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    nets = {s: [] for s in STACKS}
    for s in STACKS:
        l = 0
        while f"net.{s}.{l}.w" in z:
            nets[s].append((z[f"net.{s}.{l}.w"], z[f"net.{s}.{l}.b"]))
            l += 1
    cfg = {k[4:]: z[k].item() for k in z.files if k.startswith("cfg.")}
    inputs = dict(obs=z["in.obs"], key=z["in.key"], invalid=z["in.invalid"] if "in.invalid" in z else None,
                  noise=z["in.noise"] if "in.noise" in z else None)
    out = {k[4:]: z[k] for k in z.files if k.startswith("out.")}
    return nets, inputs, cfg, out


def haiku_linear(rng, fan_in, fan_out, bias_scale=0.0):
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


def assert_same_search(got, want, fields=OUT_FIELDS, float_tol=0.0):
    """Integer tree state must be bit-exact; float fields exact (float_tol=0) or within float_tol."""
    for f in fields:
        if f not in got or f not in want:
            continue
        g, w = np.asarray(got[f]), np.asarray(want[f])
        assert g.shape == w.shape, (f, g.shape, w.shape)
        if f in INT_FIELDS or float_tol == 0.0:
            assert np.array_equal(g, w), f"{f}: {np.sum(g != w)} of {g.size} entries differ"
        else:
            np.testing.assert_allclose(g, w, rtol=0, atol=float_tol, err_msg=f)


def check_tree_invariants(t, num_simulations):
    """SURVEY.md Appendix A.8 — size-independent properties of a finished search (default max_depth)."""
    nv, cv, ci = t["node_visits"], t["children_visits"], t["children_index"]
    B, N, A = ci.shape
    assert (nv[:, 0] == num_simulations + 1).all()
    assert (cv[:, 0].sum(-1) == num_simulations).all()
    assert (nv == 1 + cv.sum(-1)).all()
    for b in range(B):
        idx = ci[b][ci[b] >= 0]
        assert sorted(idx.tolist()) == list(range(1, num_simulations + 1))
    p, a = np.nonzero(ci[0] >= 0)
    for b in range(B):
        pp, aa = np.nonzero(ci[b] >= 0)
        child = ci[b, pp, aa]
        assert (t["parents"][b, child] == pp).all()
        assert (t["action_from_parent"][b, child] == aa).all()
        assert np.array_equal(t["children_values"][b, pp, aa], t["node_values"][b, child])
