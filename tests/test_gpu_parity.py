"""GPU parity tests proper: the CUDA search (through the C ABI) against the committed golden vectors and the
scalar C restatement on the same seeded inputs.  Integer tree state AND float fields must be bit-identical
(both sides share include/mz_math.h and the same accumulation orders)."""
import numpy as np
import pytest

from helpers import (OUT_FIELDS, assert_same_search, check_tree_invariants, golden_names, load_golden, make_nets)

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _engine(nets, B, cfg, num_sim, obs_dim=None):
    from muax_b200 import _lib
    from muax_b200.nn import pack_stacks
    from muax_b200.search import SearchEngine
    blob, cstacks = pack_stacks(nets)
    A = nets["pred_pi"][-1][0].shape[1]
    E = nets["pred_pi"][0][0].shape[0]
    eng = SearchEngine(cstacks, batch=B, num_actions=A, embed_dim=E,
                       obs_dim=nets["repr"][0][0].shape[0] if obs_dim is None else obs_dim,
                       support_size=cfg.get("support_size", 10), max_num_simulations=num_sim,
                       activation=cfg.get("activation", 0), repr_minmax=cfg.get("repr_minmax", 1),
                       dyn_minmax=cfg.get("dyn_minmax", 1), discount=cfg.get("discount", 0.99),
                       prng_mode=cfg.get("prng_mode", 0))
    eng.set_weights(blob)
    return eng


def _search_kwargs(cfg):
    kw = dict(policy=cfg.get("policy", 0), qtransform=cfg.get("qtransform", 0),
              num_simulations=cfg["num_simulations"], max_depth=cfg.get("max_depth") or None, want_tree=True)
    for k in ("temperature", "dirichlet_fraction", "dirichlet_alpha", "pb_c_init", "pb_c_base", "gumbel_scale",
              "global_batch", "batch_offset"):
        if k in cfg:
            kw[k] = cfg[k]
    if "max_considered" in cfg:
        kw["max_num_considered_actions"] = cfg["max_considered"]
    return kw


def _collect(eng, action, weights, root_value):
    torch.cuda.synchronize()
    got = {k: v.cpu().numpy() for k, v in eng.tree().items()}
    got.update(action=action.cpu().numpy(), action_weights=weights.cpu().numpy(), root_value=root_value.cpu().numpy())
    return got


ENGINES = [1, 2, 7, 8, 9]  # MZ_ENGINE_STEPWISE, _FUSED (best one-launch engine), _RESIDENT, _FUSED_WARP, _TREEWARP


def _fused_or_skip(eng, engine_id, run):
    try:
        return run()
    except RuntimeError as e:
        if engine_id != 1 and "one-launch engine does not support" in str(e):
            pytest.skip("this engine does not cover the configuration")
        raise


def test_device_math_is_bit_identical_to_host_build(c_oracle):
    from muax_b200.search import math_probe
    rng = np.random.default_rng(0)
    cases = {
        "expf": np.concatenate([rng.uniform(-104, 89, 200000), rng.uniform(-1, 1, 100000), [0, -200, 100]]),
        "logf": np.concatenate([rng.uniform(1e-38, 10, 100000), np.exp(rng.uniform(-87, 88, 200000)), [1e-45, 1.0]]),
        "expm1f": np.concatenate([-np.exp(rng.uniform(-30, 3, 200000)), rng.uniform(-1, 1, 100000)]),
        "inv_scaling": np.concatenate([rng.uniform(-10, 10, 200000), rng.uniform(-1e-2, 1e-2, 100000), [0.0]]),
    }
    host = {"expf": c_oracle.expf, "logf": c_oracle.logf, "expm1f": c_oracle.expm1f, "inv_scaling": c_oracle.inv_scaling}
    for kind, x in cases.items():
        x = x.astype(np.float32)
        y = math_probe(kind, torch.from_numpy(x).cuda()).cpu().numpy()
        assert np.array_equal(y.view(np.uint32), host[kind](x).view(np.uint32)), kind


def test_fast_division_matches_ieee():
    """The fused engines replace `__fdiv_rn` by its fast-path sequence issued in batches, with one cold IEEE branch
    when an operand is outside the safe exponent window.  Wherever the fast path is taken it must equal IEEE
    division (NumPy float32) bit for bit: random operands over the window, the ranges the search produces
    (probabilities, visit counts, value spreads), exact quotients and near-halfway quotients."""
    from muax_b200.search import math_probe
    rng = np.random.default_rng(0)
    n = 4_000_000
    a = np.concatenate([
        (rng.standard_normal(n) * np.exp2(rng.uniform(-29, 29, n))).astype(np.float32),
        rng.uniform(0, 1, n).astype(np.float32), (rng.uniform(0, 60, n) * rng.uniform(0, 1, n)).astype(np.float32),
        (rng.integers(1, 2**24, n) * rng.integers(1, 64, n)).astype(np.float32), np.zeros(16, np.float32)])
    b = np.concatenate([
        (rng.standard_normal(n) * np.exp2(rng.uniform(-29, 29, n))).astype(np.float32),
        rng.integers(1, 300, n).astype(np.float32), np.maximum(rng.uniform(0, 20, n), 1e-8).astype(np.float32),
        rng.integers(1, 64, n).astype(np.float32), np.arange(1, 17, dtype=np.float32)])
    b[b == 0] = 1.0
    pairs = np.stack([a, b], axis=1)
    got = math_probe("fast_div", torch.from_numpy(pairs).cuda()).cpu().numpy()
    slow = got.view(np.uint32) == 0x7FC00001
    assert slow.mean() < 0.05          # the window covers what the search feeds it
    with np.errstate(all="ignore"):
        want = (a / b).astype(np.float32)
    assert np.array_equal(got[~slow].view(np.uint32), want[~slow].view(np.uint32))


@pytest.mark.parametrize("engine_id", ENGINES)
@pytest.mark.parametrize("name", golden_names())
def test_golden_vectors(name, engine_id):
    nets, inp, cfg, want = load_golden(name)
    eng = _engine(nets, inp["obs"].shape[0], cfg, cfg["num_simulations"])
    out = _fused_or_skip(eng, engine_id, lambda: eng.search(
        inp["key"], obs=torch.from_numpy(inp["obs"]).cuda(), invalid_actions=inp["invalid"], noise=inp["noise"],
        engine=engine_id, **_search_kwargs(cfg)))
    got = _collect(eng, *out)
    assert_same_search(got, want)
    assert np.array_equal(got["sim_depth"], want["sim_depth"])
    assert np.array_equal(got["root_noise"], want["root_noise"])


@pytest.mark.parametrize("engine_id", ENGINES)
@pytest.mark.parametrize("policy,qt,A,E,B,NS", [(0, 0, 2, 8, 512, 50), (1, 1, 4, 16, 300, 32), (0, 1, 6, 24, 130, 40),
                                               (1, 0, 18, 32, 70, 24)])
def test_seeded_batches_match_c_restatement(c_oracle, engine_id, policy, qt, A, E, B, NS):
    rng = np.random.default_rng(100 + A)
    nets = make_nets(rng, 6, E, A, 21, bias_scale=0.05)
    obs = rng.standard_normal((B, 6)).astype(np.float32)
    invalid = (rng.random((B, A)) < 0.2).astype(np.uint8) if A > 2 else None
    if invalid is not None:
        invalid[:, 1] = 0
    key = np.array([3, 1000 + A], np.uint32)
    cfg = dict(policy=policy, qtransform=qt, num_simulations=NS, support_size=10)
    want = c_oracle.search(nets, key, obs=obs, invalid=invalid, **cfg)
    eng = _engine(nets, B, cfg, NS)
    out = _fused_or_skip(eng, engine_id, lambda: eng.search(
        key, obs=torch.from_numpy(obs).cuda(), invalid_actions=invalid, engine=engine_id, **_search_kwargs(cfg)))
    got = _collect(eng, *out)
    assert_same_search(got, want)
    check_tree_invariants(got, NS)


@pytest.mark.parametrize("engine_id", ENGINES)
def test_headline_size_every_row(c_oracle, engine_id):
    """BASELINE.json headline shapes (CartPole nets, B=4096, num_sim=50): size-independent tree invariants on
    every tree, and bit-exact parity of EVERY row (tree state, outputs, path depths) with the C restatement."""
    B, NS = 4096, 50
    rng = np.random.default_rng(0)
    nets = make_nets(rng, 4, 8, 2, 21)
    obs = rng.standard_normal((B, 4)).astype(np.float32)
    key = np.array([0, 0], np.uint32)
    cfg = dict(policy=0, qtransform=0, num_simulations=NS, support_size=10)
    eng = _engine(nets, B, cfg, NS)
    out = _fused_or_skip(eng, engine_id, lambda: eng.search(key, obs=torch.from_numpy(obs).cuda(), engine=engine_id,
                                                         **_search_kwargs(cfg)))
    got = _collect(eng, *out)
    check_tree_invariants(got, NS)
    assert np.allclose(got["action_weights"].sum(-1), 1.0, atol=1e-6)
    want = c_oracle.search(nets, key, obs=obs, **cfg)
    assert_same_search(got, want)
    assert np.array_equal(got["sim_depth"], want["sim_depth"])


@pytest.mark.parametrize("lanes,producers", [("8", "2"), ("16", "2"), ("8", "0"), ("16", "3")])
@pytest.mark.parametrize("A,S,B,NS,max_depth,table_depth", [(2, 10, 301, 50, 0, None), (4, 10, 37, 40, 0, None),
                                                            (3, 10, 129, 30, 5, None), (2, 5, 64, 24, 0, "2"),
                                                            (4, 10, 200, 33, 0, "0")])
def test_warp_engine_variants(c_oracle, monkeypatch, lanes, producers, A, S, B, NS, max_depth, table_depth):
    """The warp-autonomous engine (mz_warp.cuh) in both widths (8 / 16 lanes per tree), on every compiled shape
    variant, with ragged batches (surplus lane groups shadow the last tree), invalid actions at the root, a depth
    limit, and with the pre-computed tie-break table cut short (MZ_GROUP_K: the in-loop threefry path takes over at
    depth K; K = 0: no table at all), with the noise produced inside the kernel by dedicated warps (ring of
    mbarrier-guarded slots) or by the pre-pass kernel (producers = 0) — bit-identical to the C restatement."""
    monkeypatch.setenv("MZ_WARP_LANES", lanes)
    monkeypatch.setenv("MZ_WARP_PRODUCERS", producers)
    if table_depth is not None:
        monkeypatch.setenv("MZ_GROUP_K", table_depth)
    rng = np.random.default_rng(300 + A + S)
    nets = make_nets(rng, 5, 8, A, 2 * S + 1, bias_scale=0.05)
    obs = rng.standard_normal((B, 5)).astype(np.float32)
    invalid = (rng.random((B, A)) < 0.25).astype(np.uint8) if A > 2 else None
    if invalid is not None:
        invalid[:, 0] = 0
    key = np.array([9, 4000 + A], np.uint32)
    cfg = dict(policy=0, qtransform=0, num_simulations=NS, support_size=S)
    if max_depth:
        cfg["max_depth"] = max_depth
    want = c_oracle.search(nets, key, obs=obs, invalid=invalid, **cfg)
    eng = _engine(nets, B, cfg, NS)
    out = eng.search(key, obs=torch.from_numpy(obs).cuda(), invalid_actions=invalid, engine=8, **_search_kwargs(cfg))
    got = _collect(eng, *out)
    assert_same_search(got, want)
    assert np.array_equal(got["sim_depth"], want["sim_depth"])
    if not max_depth:  # a depth limit re-expands existing nodes: the A.8 invariants assume the default
        check_tree_invariants(got, NS)


@pytest.mark.parametrize("lanes", ["8", "16"])
@pytest.mark.parametrize("A,E,S,B,NS,max_depth", [(5, 8, 10, 70, 30, 0), (6, 8, 10, 45, 25, 4), (3, 8, 5, 50, 20, 0),
                                                  (2, 16, 5, 130, 40, 0), (4, 16, 5, 66, 30, 0)])
def test_warp_engine_wider_shapes(c_oracle, monkeypatch, lanes, A, E, S, B, NS, max_depth):
    """The warp engine's other compiled shapes: 5 and 6 actions (an 8-lane action subgroup), support_size 5 with 8- and
    16-wide embeddings — forced by name (engine = 8: the call fails if the shape is not compiled), bit-identical to
    the C restatement, with invalid actions at the root and a depth limit."""
    monkeypatch.setenv("MZ_WARP_LANES", lanes)
    rng = np.random.default_rng(500 + 10 * A + E + S)
    nets = make_nets(rng, 6, E, A, 2 * S + 1, bias_scale=0.05)
    obs = rng.standard_normal((B, 6)).astype(np.float32)
    invalid = (rng.random((B, A)) < 0.25).astype(np.uint8)
    invalid[:, 1] = 0
    key = np.array([11, 5000 + A + E], np.uint32)
    cfg = dict(policy=0, qtransform=0, num_simulations=NS, support_size=S)
    if max_depth:
        cfg["max_depth"] = max_depth
    want = c_oracle.search(nets, key, obs=obs, invalid=invalid, **cfg)
    eng = _engine(nets, B, cfg, NS)
    out = eng.search(key, obs=torch.from_numpy(obs).cuda(), invalid_actions=invalid, engine=8, **_search_kwargs(cfg))
    got = _collect(eng, *out)
    assert_same_search(got, want)
    assert np.array_equal(got["sim_depth"], want["sim_depth"])
    if not max_depth:
        check_tree_invariants(got, NS)


@pytest.mark.parametrize("lanes", ["8", "16", "32"])
@pytest.mark.parametrize("policy,qt,A,E,hidden,S,B,NS,max_depth,K", [
    (0, 0, 4, 64, (16,), 10, 301, 60, 0, None),        # C3 stock shapes, ragged batch
    (1, 0, 4, 64, (16,), 10, 130, 32, 0, None),        # C4: Gumbel with the qtransform MuZero.act forces
    (1, 1, 4, 64, (64, 64, 16), 20, 77, 24, 0, None),  # notebook nets, mctx's Gumbel default qtransform
    (0, 1, 7, 24, (20,), 5, 65, 40, 6, "3"),           # odd widths, depth limit, table cut short
    (0, 0, 18, 32, (48, 24), 10, 33, 30, 0, "0"),      # 18 actions (one tree per warp), no table at all
    (0, 0, 2, 8, (16,), 10, 64, 1, 0, None),           # a single simulation
])
def test_treewarp_engine_variants(c_oracle, monkeypatch, lanes, policy, qt, A, E, hidden, S, B, NS, max_depth, K):
    """The tree-warp engine (mz_treewarp.cu) with 8 / 16 / 32 lanes per tree (4 / 2 / 1 trees per warp) on generic
    shapes: both policies and qtransforms, invalid actions, a depth limit (re-expansion), ragged batches (surplus tree
    groups shadow the last tree), odd layer widths (scalar tails of the 128-bit loads), the tie-break table cut short
    or absent (inline threefry continuation), and an empty search — bit-identical to the C restatement."""
    monkeypatch.setenv("MZ_TREEWARP_LANES", lanes)
    if K is not None:
        monkeypatch.setenv("MZ_TREEWARP_K", K)
    rng = np.random.default_rng(500 + A + E)
    nets = make_nets(rng, 7, E, A, 2 * S + 1, hidden=hidden, bias_scale=0.05)
    obs = rng.standard_normal((B, 7)).astype(np.float32)
    invalid = (rng.random((B, A)) < 0.25).astype(np.uint8) if A > 2 else None
    if invalid is not None:
        invalid[:, 2] = 0
    key = np.array([11, 7000 + A], np.uint32)
    cfg = dict(policy=policy, qtransform=qt, num_simulations=NS, support_size=S)
    if max_depth:
        cfg["max_depth"] = max_depth
    want = c_oracle.search(nets, key, obs=obs, invalid=invalid, **cfg)
    eng = _engine(nets, B, cfg, max(NS, 1))
    out = eng.search(key, obs=torch.from_numpy(obs).cuda(), invalid_actions=invalid, engine=9, **_search_kwargs(cfg))
    got = _collect(eng, *out)
    assert_same_search(got, want)
    if NS > 0:
        assert np.array_equal(got["sim_depth"], want["sim_depth"])
    if not max_depth and NS > 0:
        check_tree_invariants(got, NS)


@pytest.mark.parametrize("policy,qt,A,E,H,B,NS", [(0, 0, 4, 64, (64,), 64, 16), (1, 1, 18, 64, (64, 32), 24, 12),
                                                  (0, 0, 3, 32, (48,), 9, 20)])
def test_resident_engine_streamed_weights(c_oracle, monkeypatch, policy, qt, A, E, H, B, NS):
    """Wide-net mode of the CTA-resident engine: weights stay in HBM/L2 and are streamed through the TMA ring
    (cp.async.bulk + mbarrier) — forced here on small nets with MZ_RESIDENT_GLOBAL_WEIGHTS — per CTA, shared by
    thread-block clusters of 4 / 8 CTAs through multicast copies (B = 64 and 24 give uniform grids; B = 9 does not and
    falls back to one ring per CTA), and with the ring disabled (plain read-only loads).  All must stay bit-identical
    to the C restatement."""
    rng = np.random.default_rng(7 + A)
    nets = make_nets(rng, 8, E, A, 21, hidden=H, bias_scale=0.05)
    obs = rng.standard_normal((B, 8)).astype(np.float32)
    key = np.array([5, 2000 + A], np.uint32)
    cfg = dict(policy=policy, qtransform=qt, num_simulations=NS, support_size=10)
    want = c_oracle.search(nets, key, obs=obs, **cfg)
    for no_tma, cluster in (("0", "4"), ("0", "8"), ("0", "1"), ("1", "1")):
        monkeypatch.setenv("MZ_RESIDENT_GLOBAL_WEIGHTS", "1")
        monkeypatch.setenv("MZ_RESIDENT_NO_TMA", no_tma)
        monkeypatch.setenv("MZ_RESIDENT_CLUSTER", cluster)
        eng = _engine(nets, B, cfg, NS)
        out = eng.search(key, obs=torch.from_numpy(obs).cuda(), engine=7, **_search_kwargs(cfg))
        got = _collect(eng, *out)
        assert_same_search(got, want)
        check_tree_invariants(got, NS)


@pytest.mark.parametrize("name,obs_dim,E,A,S,hidden,minmax,B,NS,policy", [
    ("C3 LunarLander stock", 8, 64, 4, 10, (16,), 1, 4096, 200, 0),
    ("C3 LunarLander notebook 64-64-16", 8, 64, 4, 20, (64, 64, 16), 0, 4096, 200, 0),
    ("C4 Gumbel", 8, 64, 4, 10, (16,), 1, 4096, 32, 1),
    ("C5 Atari-sized heads, one GPU's shard", 256, 256, 18, 10, (256,), 1, 1024, 50, 0)])
def test_baseline_full_sizes_every_row(c_oracle, name, obs_dim, E, A, S, hidden, minmax, B, NS, policy):
    """BASELINE.json configs 3-5 at their full sizes through AUTO (the tree-warp engine for the LunarLander nets, the
    CTA-resident engine with weights streamed through the TMA ring at the Atari-sized heads): size-independent tree
    invariants on every tree and bit-exact parity of EVERY row with the C restatement."""
    rng = np.random.default_rng(17)
    nets = make_nets(rng, obs_dim, E, A, 2 * S + 1, hidden=hidden)
    obs = rng.standard_normal((B, obs_dim)).astype(np.float32)
    key = np.array([0, 9], np.uint32)
    cfg = dict(policy=policy, qtransform=0, num_simulations=NS, support_size=S, repr_minmax=minmax, dyn_minmax=minmax)
    eng = _engine(nets, B, cfg, NS)
    out = eng.search(key, obs=torch.from_numpy(obs).cuda(), **_search_kwargs(cfg))
    got = _collect(eng, *out)
    check_tree_invariants(got, NS)
    assert np.allclose(got["action_weights"].sum(-1), 1.0, atol=1e-5)
    want = c_oracle.search(nets, key, obs=obs, **cfg)
    assert_same_search(got, want)
    assert np.array_equal(got["sim_depth"], want["sim_depth"])


def test_host_buffer_entry_point_equals_device_entry_point():
    nets, inp, cfg, want = load_golden("lunar_muzero_invalid_seed1")
    eng = _engine(nets, inp["obs"].shape[0], cfg, cfg["num_simulations"])
    a, w, v = eng.search_host(inp["key"], inp["obs"], invalid_actions=inp["invalid"], noise=inp["noise"],
                              **_search_kwargs(cfg))
    assert np.array_equal(a, want["action"]) and np.array_equal(w, want["action_weights"])
    assert np.array_equal(v, want["root_value"])


def test_root_supplied_by_caller():
    from oracle import np_mctx
    nets, inp, cfg, want = load_golden("c1_muzero_seed1")
    model = np_mctx.Model(nets, np_mctx.ExactMath(), cfg["support_size"])
    logits, value, emb = model.root_inference(inp["obs"])
    eng = _engine(nets, inp["obs"].shape[0], cfg, cfg["num_simulations"])
    for engine_id in (1, 7, 9):  # the engines that accept a caller-made root: stepwise, CTA-resident, tree-warp
        for root in ((logits, value, emb), (None, None, emb)):
            out = eng.search(inp["key"], root=root, noise=inp["noise"], engine=engine_id, **_search_kwargs(cfg))
            got = _collect(eng, *out)
            assert_same_search(got, want)


def test_callback_mode_reproduces_native_search():
    """mz_begin / mz_select / mz_expand_backup / mz_finish with the recurrent_fn played by the test (here: the
    NumPy restatement of muax/model.py:265-282) must rebuild exactly the golden tree."""
    from oracle import np_mctx
    for name in ("c1_muzero_seed42", "lunar_gumbel_invalid_seed0"):
        nets, inp, cfg, want = load_golden(name)
        model = np_mctx.Model(nets, np_mctx.ExactMath(), cfg["support_size"], cfg.get("discount", 0.99))
        root = model.root_inference(inp["obs"])
        eng = _engine(nets, inp["obs"].shape[0], cfg, cfg["num_simulations"])

        def recurrent_fn(action, emb):
            r, d, logits, v, ns = model.recurrent_inference(action.cpu().numpy(), emb.cpu().numpy())
            return r, d, logits, v, ns

        out = eng.search_with_callback(inp["key"], root, recurrent_fn, invalid_actions=inp["invalid"],
                                       noise=inp["noise"], **_search_kwargs(cfg))
        got = _collect(eng, *out)
        assert_same_search(got, want)


def test_sharded_engines_reproduce_the_global_batch():
    rng = np.random.default_rng(11)
    nets = make_nets(rng, 4, 8, 2, 21)
    obs = rng.standard_normal((96, 4)).astype(np.float32)
    key = np.array([0, 77], np.uint32)
    cfg = dict(policy=0, qtransform=0, num_simulations=30, support_size=10)
    full = _engine(nets, 96, cfg, 30)
    fa, fw, fv = full.search(key, obs=torch.from_numpy(obs).cuda(), **_search_kwargs(cfg))
    for lo, n in ((0, 32), (32, 64)):
        part = _engine(nets, n, cfg, 30)
        a, w, v = part.search(key, obs=torch.from_numpy(obs[lo:lo + n]).cuda(), global_batch=96, batch_offset=lo,
                              **_search_kwargs(cfg))
        assert torch.equal(a, fa[lo:lo + n]) and torch.equal(w, fw[lo:lo + n]) and torch.equal(v, fv[lo:lo + n])


def test_muzero_act_interface(c_oracle):
    """muax.MuZero.act contract (muax/model.py:82-179): return tuple shapes, int/float for single obs, arrays for
    batches, and agreement with the restatement on the same parameters."""
    import muax_b200
    from muax_b200 import nn
    net = nn.create_muzero_network(nn.Representation, nn.Prediction, nn.Dynamic, 8, 2, 21)
    for policy in ("muzero", "gumbel"):
        model = muax_b200.MuZero(net, policy=policy, discount=0.99, support_size=10)
        params = model.init(muax_b200.random.PRNGKey(0), np.zeros((1, 4), np.float32))
        key = muax_b200.random.PRNGKey(5)
        obs = np.array([0.1, -0.2, 0.03, 0.4], np.float32)
        a = model.act(key, obs, num_simulations=20)
        assert isinstance(a, int) and 0 <= a < 2
        a2, pi, v = model.act(key, obs, with_pi=True, with_value=True, num_simulations=20)
        assert a2 == a and pi.shape == (1, 2) and isinstance(v, float)
        batch = np.stack([obs, -obs, obs * 2])
        ab, pib, vb = model.act(key, batch, with_pi=True, with_value=True, obs_from_batch=True, num_simulations=20)
        assert ab.shape == (3,) and pib.shape == (3, 2) and vb.shape == (3,)
        stacks = model._spec.stacks(params)
        want = c_oracle.search(stacks, key, obs=batch, policy=0 if policy == "muzero" else 1, qtransform=0,
                               num_simulations=20)
        assert np.array_equal(ab, want["action"]) and np.array_equal(pib, want["action_weights"])
        assert np.array_equal(vb, want["root_value"])
        # device-tensor observations take the zero-copy path and agree with the host-buffer path
        ad = model.act(key, torch.from_numpy(batch).cuda(), obs_from_batch=True, num_simulations=20)
        assert np.array_equal(ad, ab)


def test_error_paths():
    import muax_b200
    from muax_b200 import nn
    nets, inp, cfg, _ = load_golden("c1_muzero_seed0")
    eng = _engine(nets, 4, cfg, 10)
    with pytest.raises(ValueError):
        eng.search(inp["key"], obs=torch.zeros(4, 4).cuda(), num_simulations=11)  # exceeds capacity
    with pytest.raises(ValueError):
        eng.search(inp["key"], obs=torch.zeros(3, 4).cuda(), num_simulations=5)   # wrong batch
    with pytest.raises(ValueError):
        muax_b200.policy.resolve_qtransform(lambda tree, node: None)
    with pytest.raises(ValueError):
        muax_b200.MuZero(nn.create_muzero_network(nn.Representation, nn.Prediction, nn.Dynamic, 8, 2, 21),
                         policy="alphazero")


def test_vector_actor_on_the_gpu_search():
    """§8(f) rank 1: the vectorised acting loop (CartPoleVec -> MuZero.act(obs_from_batch=True) -> BatchedPNStep ->
    TrajectoryStore) on the real search; the search inside it must equal a direct `act` on the same key and batch."""
    import muax_b200
    from muax_b200 import nn
    from muax_b200.actor import CartPoleVec, TrajectoryStore, VectorActor
    net = nn.create_muzero_network(nn.Representation, nn.Prediction, nn.Dynamic, 8, 2, 21)
    model = muax_b200.MuZero(net, policy="muzero", discount=0.997, support_size=10)
    model.init(muax_b200.random.PRNGKey(0), np.zeros((1, 4), np.float32))
    env = CartPoleVec(256, seed=3)
    store = TrajectoryStore(4096, random_seed=0)
    actor = VectorActor(model, env, store, n=5, gamma=0.997, k_steps=5, num_simulations=16)
    for t in range(40):
        obs = actor.obs.copy()
        key = muax_b200.random.PRNGKey(100 + t)
        a, pi, v, done = actor.step(key)
        a2, pi2, v2 = model.act(key, obs, with_pi=True, with_value=True, obs_from_batch=True, num_simulations=16)
        assert np.array_equal(a, a2) and np.array_equal(pi, pi2) and np.array_equal(v, v2)
        assert a.shape == (256,) and set(np.unique(a)) <= {0, 1} and np.allclose(pi.sum(-1), 1.0, atol=1e-6)
    assert actor.episodes > 0 and len(store) > 0
    batch = store.sample(batch_size=32, k_steps=5)
    assert batch.obs.shape == (32, 5, 4) and batch.pi.shape == (32, 5, 2) and np.all(batch.w >= 0)


@pytest.mark.parametrize("use_graph", [False, True])
def test_device_actor_equals_numpy_actor_data_path(use_graph):
    """Device-resident acting loop (CartPoleVecTorch + DevicePNStep on the GPU; eager, and with everything after the
    search replayed from a CUDA graph): the episodes it stores must be what the reference-pinned NumPy tracer produces
    from the same per-step (obs, a, r, done, v, pi) stream."""
    import muax_b200
    from muax_b200 import nn
    from muax_b200.actor import BatchedPNStep, TrajectoryStore
    from muax_b200.actor_device import CartPoleVecTorch, DeviceActor
    net = nn.create_muzero_network(nn.Representation, nn.Prediction, nn.Dynamic, 8, 2, 21)
    model = muax_b200.MuZero(net, policy="muzero", discount=0.997, support_size=10)
    model.init(muax_b200.random.PRNGKey(0), np.zeros((1, 4), np.float32))
    B = 128
    env = CartPoleVecTorch(B, seed=5)
    store = TrajectoryStore(100000, random_seed=0)
    actor = DeviceActor(model, env, store, n=5, gamma=0.997, k_steps=1, num_simulations=8, use_graph=use_graph)
    ref = BatchedPNStep(B, 5, 0.997, 0.5)
    ref_eps, pending = [], [[] for _ in range(B)]
    for t in range(45):
        obs = actor.obs.cpu().numpy()
        a, pi, v, done = actor.step(muax_b200.random.PRNGKey(t))
        a, pi, v, done = a.cpu().numpy(), pi.cpu().numpy(), v.cpu().numpy(), done.cpu().numpy()
        env_idx, tr = ref.add(obs, a, np.ones(B), done, v, pi)
        for i, e in enumerate(env_idx):
            pending[e].append(i)
        rows_by_env = pending
        pending = [[] for _ in range(B)]
        for e in range(B):
            if rows_by_env[e]:
                ref.__dict__.setdefault("_blocks", {}).setdefault(e, []).append(tr[np.array(rows_by_env[e])])
        for e in np.nonzero(done)[0]:
            blocks = ref.__dict__["_blocks"].pop(e)
            ref_eps.append(type(tr)(*(np.concatenate(c) for c in zip(*blocks))))
    assert len(store) == len(ref_eps) > 0 and actor.episodes == len(ref_eps)
    for have, want in zip(store, ref_eps):
        assert len(have) == len(want)
        for f in ("obs", "a", "r", "done", "v", "pi"):
            assert np.array_equal(np.asarray(getattr(have, f)).astype(getattr(want, f).dtype), getattr(want, f)), f
        np.testing.assert_allclose(have.Rn, want.Rn, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(have.w, want.w, rtol=1e-9, atol=1e-12)


def _run_engine(engine_id, nets, key, obs, cfg, invalid=None, noise=None):
    eng = _engine(nets, obs.shape[0], cfg, cfg["num_simulations"])
    out = _fused_or_skip(eng, engine_id, lambda: eng.search(
        key, obs=torch.from_numpy(obs).cuda(), invalid_actions=invalid, noise=noise, engine=engine_id,
        **_search_kwargs(cfg)))
    return _collect(eng, *out)


@pytest.mark.parametrize("engine_id", ENGINES)
def test_edge_cases_every_engine(c_oracle, engine_id):
    """The corners mctx defines but a batched run rarely meets, on every engine:
    a row whose actions are ALL invalid (finite min logit, masked_argmax -> 0), an empty search (num_simulations = 0:
    uniform visit_probs), a single action (A = 1), temperature = 0 with an exact 25 / 25 visit tie (the action is then
    decided by the final Gumbel draw alone), and a batch of one tree (BASELINE config 1's plumbing case)."""
    rng = np.random.default_rng(3)
    key = np.array([1, 2], np.uint32)
    nets = make_nets(rng, 6, 16, 4, 21)
    obs = rng.standard_normal((5, 6)).astype(np.float32)
    invalid = np.zeros((5, 4), np.uint8)
    invalid[0] = 1
    invalid[2, 1:] = 1
    for policy, qt in ((0, 0), (1, 0), (1, 1)):
        cfg = dict(policy=policy, qtransform=qt, num_simulations=20, support_size=10)
        want = c_oracle.search(nets, key, obs=obs, invalid=invalid, **cfg)
        assert want["action"][0] == 0
        assert_same_search(_run_engine(engine_id, nets, key, obs, cfg, invalid=invalid), want)
    for policy, qt in ((0, 0), (1, 1)):
        cfg = dict(policy=policy, qtransform=qt, num_simulations=0, support_size=10)
        want = c_oracle.search(nets, key, obs=obs, **cfg)
        assert_same_search(_run_engine(engine_id, nets, key, obs, cfg), want)
    one = make_nets(rng, 6, 16, 1, 21)
    for policy, qt in ((0, 0), (1, 1)):
        cfg = dict(policy=policy, qtransform=qt, num_simulations=12, support_size=10)
        want = c_oracle.search(one, key, obs=obs, **cfg)
        assert_same_search(_run_engine(engine_id, one, key, obs, cfg), want)
    flat = {k: [(w * 0, b * 0) for w, b in v] for k, v in make_nets(rng, 4, 8, 2, 21).items()}
    obs2 = rng.standard_normal((64, 4)).astype(np.float32)
    cfg = dict(policy=0, qtransform=0, num_simulations=50, support_size=10, temperature=0.0, dirichlet_fraction=0.0)
    want = c_oracle.search(flat, key, obs=obs2, **cfg)
    assert (want["children_visits"][:, 0] == 25).all() and 0 < want["action"].sum() < 64   # ties, both actions drawn
    assert_same_search(_run_engine(engine_id, flat, key, obs2, cfg), want)
    cfg = dict(policy=0, qtransform=0, num_simulations=50, support_size=10)
    want = c_oracle.search(nets, key, obs=obs[:1], **cfg)
    assert_same_search(_run_engine(engine_id, nets, key, obs[:1], cfg), want)
