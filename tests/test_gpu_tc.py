"""Throughput mode (mz_search_args.precision = bf16): the tcgen05 recurrent kernel (mz_recurrent_tc.cu).

bf16 operands cannot reproduce the fp32 trees bit for bit, so this file states the tolerances the mode is held to:
  * the kernel alone (`mz_recurrent`) against a float64 evaluation of the same network on bf16-rounded operands
    (weights, input embedding and every hidden activation rounded to bf16, exactly what the kernel feeds the tensor
    core): 5e-3 absolute on next-state / prior logits (different fp32 accumulation order can flip a bf16 rounding of a
    hidden unit, 2^-9 relative), 2e-2 on reward / value (the support transform amplifies logit differences);
  * against the library's own fp32 recurrent kernel: bf16-level agreement (5e-2);
  * a whole search: tree invariants hold exactly (the integer bookkeeping is the fp32 code), the root value (root
    inference runs on the tensor-core kernel too) agrees with the fp32 one to bf16 level, and the visit distributions
    stay close to the fp32 search on the same keys (mean total-variation distance <= 0.15, >= 70 % identical
    argmax-visit actions on random nets);
  * the throughput mode's own tree kernels (tree-warp walks on packed records, mz_treewarp.cu) against the generic
    SoA stepwise kernels around the SAME recurrent kernel, root supplied: every tree field bit for bit.
"""
import numpy as np
import pytest

from helpers import check_tree_invariants, make_nets

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _bf16(x):
    """Round float32 -> bf16 (nearest even) -> float64."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return u.astype(np.uint32).view(np.float32).astype(np.float64)


def _elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))


def _stack64(layers, x, onehot=None, A=0):
    """hk.Sequential on bf16-rounded operands in float64 (muax/nn.py:63-104)."""
    h = _bf16(x)
    if onehot is not None:
        h = np.concatenate([h, np.eye(A)[onehot]], axis=1)
    for i, (w, b) in enumerate(layers):
        h = h @ _bf16(w) + b.astype(np.float64)
        if i + 1 < len(layers):
            h = _bf16(_elu(h).astype(np.float32))
    return h


def _support_to_scalar64(logits, S):
    p = np.exp(logits - logits.max(-1, keepdims=True))
    p /= p.sum(-1, keepdims=True)
    x = (p * np.arange(-S, S + 1)).sum(-1)
    eps = 1e-3
    return np.sign(x) * (((np.sqrt(1 + 4 * eps * (np.abs(x) + 1 + eps)) - 1) / (2 * eps)) ** 2 - 1)


def _recurrent64(nets, action, emb, S, minmax):
    A = nets["pred_pi"][-1][0].shape[1]
    r = _stack64(nets["dyn_r"], emb, action, A)
    ns = _stack64(nets["dyn_ns"], emb, action, A)
    if minmax:
        lo, hi = ns.min(-1, keepdims=True), ns.max(-1, keepdims=True)
        scale = hi - lo
        scale = np.where(scale < 1e-5, scale + 1e-5, scale)
        ns = (ns - lo) / scale
    ns32 = ns.astype(np.float32)
    v = _stack64(nets["pred_v"], ns32)
    logits = _stack64(nets["pred_pi"], ns32)
    return _support_to_scalar64(r, S), _support_to_scalar64(v, S), logits, ns


def _engine(nets, B, S, num_sim, minmax=1):
    from muax_b200.nn import pack_stacks
    from muax_b200.search import SearchEngine
    blob, cstacks = pack_stacks(nets)
    eng = SearchEngine(cstacks, batch=B, num_actions=nets["pred_pi"][-1][0].shape[1],
                       embed_dim=nets["pred_pi"][0][0].shape[0], obs_dim=nets["repr"][0][0].shape[0], support_size=S,
                       max_num_simulations=num_sim, repr_minmax=minmax, dyn_minmax=minmax)
    eng.set_weights(blob)
    return eng


SHAPES = [
    # obs, E, A, S, hidden, minmax, B
    (8, 64, 4, 10, (16,), 1, 300),            # C3 stock: tiny layers, padded to 16
    (8, 64, 4, 20, (64, 64, 16), 0, 129),     # C3 notebook nets: three hidden layers (second hidden buffer), S = 20
    (32, 256, 18, 10, (256,), 1, 257),        # C5 heads: K = 274 -> 288, N = 256, weights streamed in 9 chunks per layer
    (5, 24, 7, 5, (40, 24), 1, 64),           # odd widths everywhere
]


@pytest.mark.parametrize("obs_dim,E,A,S,hidden,minmax,B", SHAPES)
def test_tcgen05_recurrent_kernel(obs_dim, E, A, S, hidden, minmax, B):
    rng = np.random.default_rng(E + A)
    nets = make_nets(rng, obs_dim, E, A, 2 * S + 1, hidden=hidden, bias_scale=0.1)
    eng = _engine(nets, B, S, 4, minmax)
    emb = rng.random((B, E)).astype(np.float32)      # embeddings are min-max normalised: [0, 1]
    action = rng.integers(0, A, B).astype(np.int32)
    got = [t.cpu().numpy() for t in eng.recurrent(action, emb, precision="bf16")]
    f32 = [t.cpu().numpy() for t in eng.recurrent(action, emb, precision="fp32")]
    want = _recurrent64(nets, action, emb, S, minmax)
    names = ("reward", "value", "prior_logits", "next_embedding")
    tol64 = (2e-2, 2e-2, 5e-3, 5e-3)
    for n, g, w, f, tol in zip(names, got, want, f32, tol64):
        assert np.isfinite(g).all(), n
        err = np.abs(g - w)
        # a flipped bf16 rounding of one hidden unit moves an output by ~|w| * 2^-9: allow a few such rows
        assert np.quantile(err, 0.99) <= tol, (n, float(np.quantile(err, 0.99)), float(err.max()))
        assert err.max() <= 10 * tol, (n, float(err.max()))
        assert np.abs(g - f).max() <= 5e-2 * max(1.0, float(np.abs(f).max())), (n, float(np.abs(g - f).max()))


@pytest.mark.parametrize("obs_dim,E,A,S,hidden,minmax,B,NS,policy", [
    (8, 64, 4, 10, (16,), 1, 512, 48, 0),
    (8, 64, 4, 20, (64, 64, 16), 0, 256, 32, 1),
    (32, 256, 18, 10, (256,), 1, 256, 24, 0),
])
def test_bf16_search_stays_close_to_the_fp32_search(obs_dim, E, A, S, hidden, minmax, B, NS, policy):
    rng = np.random.default_rng(7 * E + A)
    nets = make_nets(rng, obs_dim, E, A, 2 * S + 1, hidden=hidden)
    obs = torch.from_numpy(rng.standard_normal((B, obs_dim)).astype(np.float32)).cuda()
    key = np.array([0, 5], np.uint32)
    eng = _engine(nets, B, S, NS, minmax)
    kw = dict(policy=policy, qtransform=0, num_simulations=NS, want_tree=True)
    a32, w32, v32 = (t.cpu().numpy() for t in eng.search(key, obs=obs, **kw))
    a16, w16, v16 = (t.cpu().numpy() for t in eng.search(key, obs=obs, precision="bf16", **kw))
    tree = {k: v.cpu().numpy() for k, v in eng.tree().items()}
    check_tree_invariants(tree, NS)                       # integer bookkeeping is exact in either mode
    assert np.abs(v16 - v32).max() <= 5e-2 * max(1.0, float(np.abs(v32).max()))   # root inference on tcgen05 too
    assert np.allclose(w16.sum(-1), 1.0, atol=1e-5)
    tv = 0.5 * np.abs(w16 - w32).sum(-1)
    agree = (w16.argmax(-1) == w32.argmax(-1)).mean()
    print(f"bf16 vs fp32 search: mean TV {tv.mean():.4f}, max TV {tv.max():.3f}, argmax agreement {agree:.3f}")
    assert tv.mean() <= 0.15 and agree >= 0.7


@pytest.mark.parametrize("obs_dim,E,A,S,hidden,minmax,B,NS,policy,qt,max_depth", [
    (8, 64, 4, 10, (16,), 1, 300, 40, 0, 0, None),
    (8, 64, 4, 10, (16,), 1, 77, 20, 1, 1, None),          # Gumbel + completed_by_mix_value: the generic scores
    (32, 256, 18, 10, (256,), 1, 130, 24, 0, 0, None),
    (5, 24, 7, 5, (40, 24), 1, 64, 30, 0, 0, 4),           # max_depth: re-expansion of an existing node
])
def test_batched_tree_kernels_equal_the_stepwise_kernels(obs_dim, E, A, S, hidden, minmax, B, NS, policy, qt, max_depth):
    from muax_b200 import _lib
    rng = np.random.default_rng(11 * E + A)
    nets = make_nets(rng, obs_dim, E, A, 2 * S + 1, hidden=hidden)
    eng = _engine(nets, B, S, NS, minmax)
    root = (rng.standard_normal((B, A)).astype(np.float32), rng.standard_normal(B).astype(np.float32),
            rng.random((B, E)).astype(np.float32))
    invalid = (rng.random((B, A)) < 0.2).astype(np.uint8)
    invalid[:, 0] = 0
    key = np.array([3, 9], np.uint32)
    kw = dict(policy=policy, qtransform=qt, num_simulations=NS, want_tree=True, precision="bf16", max_depth=max_depth,
              invalid_actions=invalid)
    res = {}
    for name, engine in (("batched", _lib.ENGINE_AUTO), ("stepwise", _lib.ENGINE_STEPWISE)):
        out = [t.cpu().numpy() for t in eng.search(key, root=root, engine=engine, **kw)]
        res[name] = (out, {k: v.cpu().numpy() for k, v in eng.tree().items()})
    for g, w in zip(res["batched"][0], res["stepwise"][0]):
        assert np.array_equal(g, w)
    for k, w in res["stepwise"][1].items():
        if k == "embeddings":  # the throughput mode keeps its embeddings in bf16 (what the tensor core reads anyway)
            w = _bf16(w).astype(np.float32)
        assert np.array_equal(res["batched"][1][k], w), k
