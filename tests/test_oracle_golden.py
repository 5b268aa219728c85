"""The two CPU restatements against the committed golden vectors, plus structural invariants and the
sensitivity of results to ulp-level math differences (libm vs include/mz_math.h)."""
import numpy as np
import pytest

from helpers import assert_same_search, check_tree_invariants, golden_names, load_golden, make_nets
from oracle import np_mctx, threefry


@pytest.mark.parametrize("name", golden_names())
def test_c_restatement_matches_golden(name, c_oracle):
    nets, inp, cfg, want = load_golden(name)
    got = c_oracle.search(nets, inp["key"], obs=inp["obs"], invalid=inp["invalid"], noise=inp["noise"], **cfg)
    assert_same_search(got, want)
    assert np.array_equal(got["sim_depth"], want["sim_depth"])
    if not cfg.get("max_depth"):
        check_tree_invariants(got, cfg["num_simulations"])


@pytest.mark.parametrize("name", ["c1_muzero_seed0", "c1_gumbel_mctx_seed42", "lunar_muzero_invalid_seed1",
                                  "atari18_muzero_seed0"])
def test_numpy_restatement_matches_golden(name):
    nets, inp, cfg, want = load_golden(name)
    got = np_mctx.act(nets, inp["key"], obs=inp["obs"], invalid=inp["invalid"], noise=inp["noise"], **cfg)
    assert_same_search(got, want)


def test_root_supplied_equals_obs_path(c_oracle):
    nets, inp, cfg, want = load_golden("c1_muzero_seed1")
    model = np_mctx.Model(nets, np_mctx.ExactMath(), cfg["support_size"])
    root = model.root_inference(inp["obs"])
    got = c_oracle.search(nets, inp["key"], root=root, noise=inp["noise"], **cfg)
    assert_same_search(got, want)


def test_sharded_rows_equal_global_rows(c_oracle):
    """Trees are independent and PRNG draws are indexed by the GLOBAL row: any row range of a batch searched
    alone (global_batch / batch_offset) must reproduce the same rows of the full-batch search."""
    rng = np.random.default_rng(5)
    nets = make_nets(rng, 4, 8, 2, 21)
    obs = rng.standard_normal((10, 4)).astype(np.float32)
    key = threefry.PRNGKey(9)
    for policy, mode in ((0, 0), (0, 1), (1, 0)):
        kw = dict(policy=policy, prng_mode=mode, num_simulations=20, qtransform=policy)
        full = c_oracle.search(nets, key, obs=obs, **kw)
        part = c_oracle.search(nets, key, obs=obs[4:7], global_batch=10, batch_offset=4, **kw)
        for f in ("action", "action_weights", "node_visits", "children_index", "node_values", "root_noise"):
            assert np.array_equal(part[f], full[f][4:7]), (policy, mode, f)


def test_results_stable_under_libm_math():
    """Swap include/mz_math.h + sequential-fma dense for NumPy's libm + BLAS (what an XLA-CPU run would differ
    by).  Raw network values stay within 1e-5; tree values pass through muax/utils.py:70-76 `_inv_scaling`, whose
    sqrt(1 + 4e-3 x) - 1 cancellation amplifies a 1-ulp softmax difference ~100x in float32, so two *correct*
    float32 implementations agree only to ~5e-5 relative there (SURVEY.md §7 hard part 3).  Visit counts agree
    wherever no argmax near-tie flips."""
    nets, inp, cfg, want = load_golden("c1_muzero_seed0")
    got = np_mctx.act(nets, inp["key"], obs=inp["obs"], noise=inp["noise"], math=np_mctx.LibmMath(), **cfg)
    np.testing.assert_allclose(got["root_value"], want["root_value"], atol=1e-5, rtol=0)
    same = (got["children_index"] == want["children_index"]).all(axis=(1, 2))
    assert same.mean() >= 0.5
    for b in np.nonzero(same)[0]:
        np.testing.assert_allclose(got["node_values"][b], want["node_values"][b], atol=1e-5, rtol=5e-5)
        np.testing.assert_allclose(got["action_weights"][b], want["action_weights"][b], atol=1e-5, rtol=0)


def test_considered_visits_table(c_oracle):
    for m in range(0, 17):
        for n in (1, 5, 16, 32, 50, 200):
            assert c_oracle.considered_visits(m, n).tolist() == list(np_mctx.considered_visits_sequence(m, n))
    # hand-checked sequence: m=4, n=8 -> two rounds over 4 actions... (sequential halving)
    assert list(np_mctx.considered_visits_sequence(4, 8)) == [0, 0, 0, 0, 1, 1, 2, 2]


def test_dirichlet_sampler_is_a_distribution(c_oracle):
    d = c_oracle.dirichlet(threefry.PRNGKey(1), 0, 20000, 4, 0.3)
    assert np.allclose(d.sum(-1), 1.0, atol=1e-5) and (d >= 0).all()
    assert np.allclose(d.mean(0), 0.25, atol=0.01)
    # Var of a Dirichlet(alpha) marginal: (1/A)(1-1/A)/(A*alpha+1)
    assert np.allclose(d.var(0), 0.25 * 0.75 / (4 * 0.3 + 1), rtol=0.06)
    assert np.array_equal(c_oracle.dirichlet(threefry.PRNGKey(1), 100, 50, 4, 0.3), d[100:150])


def test_puct_pins():
    """The reference tree carries an independent scalar pUCT (muax/frameworks/acme/tf/mcts/search.py:463-497).  Its
    scores and choices on 512 random nodes (tests/golden/make_puct_pins.py executes the reference function in place)
    must be reproduced by the oracle's exploration term + the same raw Q: pb_c, sqrt(n), prior / (visits + 1) and the
    zero-prior masking are then pinned to the reference, not to this repo's reading of mctx."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "puct_pins.npz"))
    n, A = z["visits"].shape
    m = np_mctx.ExactMath()
    tree = np_mctx.Tree(n, 1, A, 1)
    tree.node_visits[:, 0] = z["node_visits"]
    tree.children_visits[:, 0] = z["visits"]
    with np.errstate(divide="ignore"):
        tree.children_prior_logits[:, 0] = np.log(z["priors"]).astype(np.float32)   # softmax(log p) = p
    rows, node = np.arange(n), np.zeros(n, np.int32)
    policy_score = np_mctx.muzero_policy_score(m, tree, rows, node, float(z["c_init"]), float(z["c_base"]))
    scores = z["values"].astype(np.float32) + policy_score
    legal = z["priors"] > 0
    np.testing.assert_allclose(scores[legal], z["scores"][legal], rtol=2e-6, atol=2e-6)
    assert (z["scores"][~legal] == np.finfo(np.float32).min).all()     # the reference masks zero-prior actions
    ours = np.where(legal, scores, -np.inf).argmax(-1)
    top2 = np.sort(np.where(legal, z["scores"], -np.inf), axis=-1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-5                          # skip float32-level near-ties
    assert clear.mean() > 0.95 and np.array_equal(ours[clear], z["actions"][clear])


def test_stochastic_restatement_invariants():
    """oracle.np_mctx.stochastic_muzero_policy (mctx.stochastic_muzero_policy, muax/policy.py:50-67): decision and
    chance nodes alternate by depth, chance nodes expand chance slots only, the visit summary covers the decision
    actions, and the walk is reproducible from the key."""
    from oracle import np_mctx
    rng = np.random.default_rng(5)
    B, A, C, NS, T = 12, 3, 4, 25, 53
    tab_cl, tab_av = rng.standard_normal((T, C)).astype(np.float32), rng.standard_normal(T).astype(np.float32)
    tab_al, tab_v, tab_r = (rng.standard_normal((T, A)).astype(np.float32), rng.standard_normal(T).astype(np.float32),
                            rng.standard_normal(T).astype(np.float32))

    def dec(action, state):
        idx = (state[:, 0].astype(np.int64) * 7 + action.astype(np.int64) * 3 + 1) % T
        return tab_cl[idx], tab_av[idx], idx.astype(np.float32)[:, None]

    def ch(outcome, after):
        idx = (after[:, 0].astype(np.int64) * 11 + outcome.astype(np.int64) * 5 + 2) % T
        return tab_al[idx], tab_v[idx], tab_r[idx], np.full(idx.shape, 0.97, np.float32), idx.astype(np.float32)[:, None]

    model = np_mctx.StochasticModel(np_mctx.ExactMath(), dec, ch, A, C, 1, 1)
    root = (rng.standard_normal((B, A)).astype(np.float32), rng.standard_normal(B).astype(np.float32),
            rng.integers(0, T, (B, 1)).astype(np.float32))
    noise = rng.dirichlet([0.3] * A, B).astype(np.float32)
    key = np.array([0, 9], np.uint32)
    out = np_mctx.stochastic_muzero_policy(model, key, root, NS, dirichlet_noise=noise)
    again = np_mctx.stochastic_muzero_policy(model, key, root, NS, dirichlet_noise=noise)
    t = out["tree"]
    assert np.array_equal(out["action"], again["action"]) and np.array_equal(t.children_visits, again["tree"].children_visits)
    assert (t.node_visits[:, 0] == NS + 1).all() and (t.children_visits[:, 0, :A].sum(-1) == NS).all()
    assert np.allclose(out["action_weights"].sum(-1), 1.0, atol=1e-6) and out["action_weights"].shape == (B, A)
    is_dec = t.embeddings[:, :, -1] != 0
    assert not ((t.children_index[:, :, A:] >= 0).any(-1) & is_dec).any()      # decision nodes expand actions
    assert not ((t.children_index[:, :, :A] >= 0).any(-1) & ~is_dec).any()     # chance nodes expand outcomes
    for b in range(B):
        for n in range(1, NS + 1):
            assert is_dec[b, n] != is_dec[b, t.parents[b, n]]
    # afterstate edges carry reward 0 and discount 1
    p, a = np.nonzero(t.children_index[0] >= 0)
    dec_edges = is_dec[0, p]
    assert (t.children_rewards[0, p[dec_edges], a[dec_edges]] == 0).all()
    assert (t.children_discounts[0, p[dec_edges], a[dec_edges]] == 1).all()
