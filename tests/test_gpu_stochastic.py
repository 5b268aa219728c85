"""StochasticMuZeroPolicy (muax/policy.py:50-67 -> mctx.stochastic_muzero_policy): the CUDA tree kernels (callback
mode, A' = A + C pseudo-actions, decision / chance node selection by depth parity) against the NumPy restatement
(oracle/np_mctx.py::stochastic_muzero_policy) on the same keys — every integer tree field, the action and the visit
distribution bit for bit.  The two recurrent functions are lookup tables driven by small-integer embeddings, so the
NumPy and the torch evaluation agree exactly and the comparison isolates the search."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

T = 97  # table rows


def _tables(rng, A, C):
    return dict(cl=rng.standard_normal((T, C)).astype(np.float32), av=rng.standard_normal(T).astype(np.float32),
                al=rng.standard_normal((T, A)).astype(np.float32), v=rng.standard_normal(T).astype(np.float32),
                r=rng.standard_normal(T).astype(np.float32))


def _np_fns(tab, discount):
    def dec(action, state):
        idx = (state[:, 0].astype(np.int64) * 7 + action.astype(np.int64) * 3 + 1) % T
        after = np.stack([idx.astype(np.float32), (idx * 5 % T).astype(np.float32)], 1)
        return tab["cl"][idx], tab["av"][idx], after

    def ch(outcome, after):
        idx = (after[:, 0].astype(np.int64) * 11 + outcome.astype(np.int64) * 5 + 2) % T
        st = np.stack([idx.astype(np.float32), (idx * 3 % T).astype(np.float32), (idx * 2 % T).astype(np.float32)], 1)
        return tab["al"][idx], tab["v"][idx], tab["r"][idx], np.full(idx.shape, discount, np.float32), st
    return dec, ch


def _torch_fns(tab, discount, dev):
    from muax_b200.policy import ChanceRecurrentFnOutput, DecisionRecurrentFnOutput
    tt = {k: torch.from_numpy(v).to(dev) for k, v in tab.items()}

    def dec(params, rng_key, action, state):
        idx = (state[:, 0].long() * 7 + action.long() * 3 + 1) % T
        after = torch.stack([idx.float(), (idx * 5 % T).float()], 1)
        return DecisionRecurrentFnOutput(tt["cl"][idx], tt["av"][idx]), after

    def ch(params, rng_key, outcome, after):
        idx = (after[:, 0].long() * 11 + outcome.long() * 5 + 2) % T
        st = torch.stack([idx.float(), (idx * 3 % T).float(), (idx * 2 % T).float()], 1)
        disc = torch.full(idx.shape, discount, dtype=torch.float32, device=dev)
        return ChanceRecurrentFnOutput(tt["al"][idx], tt["v"][idx], tt["r"][idx], disc), st
    return dec, ch


@pytest.mark.parametrize("B,A,C,NS,max_depth,temperature,with_invalid", [
    (64, 3, 4, 24, None, 1.0, False),
    (37, 2, 6, 40, None, 0.25, True),
    (16, 5, 3, 30, 5, 1.0, True),      # max_depth: re-expansion, odd cap ends a walk on a chance node
    (8, 4, 2, 0, None, 1.0, False),    # no simulations: uniform weights over the decision actions
])
def test_stochastic_muzero_matches_the_numpy_restatement(B, A, C, NS, max_depth, temperature, with_invalid):
    from muax_b200.policy import RootFnOutput, StochasticMuZeroPolicy
    from oracle import np_mctx
    rng = np.random.default_rng(100 * A + C)
    tab = _tables(rng, A, C)
    discount = 0.97
    root = (rng.standard_normal((B, A)).astype(np.float32), rng.standard_normal(B).astype(np.float32),
            rng.integers(0, T, (B, 3)).astype(np.float32))
    noise = rng.dirichlet([0.3] * A, B).astype(np.float32)
    invalid = None
    if with_invalid:
        invalid = (rng.random((B, A)) < 0.3).astype(np.uint8)
        invalid[:, 0] = 0
        invalid[0, :] = 1  # one row with every action invalid
    key = np.array([0, 11], np.uint32)

    dec_np, ch_np = _np_fns(tab, discount)
    model = np_mctx.StochasticModel(np_mctx.ExactMath(), dec_np, ch_np, A, C, 3, 2)
    want = np_mctx.stochastic_muzero_policy(model, key, root, NS, invalid_actions=invalid, max_depth=max_depth,
                                            temperature=temperature, dirichlet_noise=noise)

    dev = torch.device("cuda")
    dec_t, ch_t = _torch_fns(tab, discount, dev)
    policy = StochasticMuZeroPolicy()
    out = policy(None, key, RootFnOutput(*root), decision_recurrent_fn=dec_t, chance_recurrent_fn=ch_t,
                 num_simulations=NS, max_depth=max_depth, temperature=temperature, invalid_actions=invalid, noise=noise)
    torch.cuda.synchronize()
    assert np.array_equal(out.action.cpu().numpy(), want["action"])
    assert np.array_equal(out.action_weights.cpu().numpy(), want["action_weights"])
    if NS > 0:
        tree = {k: v.cpu().numpy() for k, v in out.search_tree.tree().items()}
        wt = want["tree"]
        for f in ("node_visits", "parents", "action_from_parent", "children_index", "children_visits"):
            assert np.array_equal(tree[f], getattr(wt, f)), f
        for f in ("node_values", "raw_values", "children_values", "children_rewards", "children_discounts",
                  "children_prior_logits", "embeddings"):
            assert np.array_equal(tree[f], getattr(wt, f)), f
        assert np.array_equal(tree["sim_depth"], want["sim_depth"])
        # structure: decision and chance nodes alternate, chance nodes only ever expand chance slots
        ci = tree["children_index"]
        is_dec = tree["embeddings"][:, :, -1] != 0
        visited = tree["node_visits"] > 0
        live = np.ones(B, bool) if invalid is None else ~invalid.all(-1)  # an all-invalid root falls through to a chance slot
        assert not (((ci[:, :, A:] >= 0).any(-1) & is_dec & visited)[live]).any()
        assert not ((ci[:, :, :A] >= 0).any(-1) & ~is_dec & visited).any()


def test_stochastic_policy_needs_both_functions():
    from muax_b200.policy import RootFnOutput, StochasticMuZeroPolicy
    with pytest.raises(ValueError):
        StochasticMuZeroPolicy()(None, np.array([0, 0], np.uint32),
                                 RootFnOutput(np.zeros((1, 2), np.float32), np.zeros(1, np.float32),
                                              np.zeros((1, 1), np.float32)))
