"""Learner (muax_b200/learner.py) against the reference: tests/golden/loss_pins.npz holds parameters, a batch and
the value of the REFERENCE'S OWN `default_loss_fn` on them (tests/golden/make_loss_pins.py executes
muax/frameworks/coax/loss.py + utils.py); gradients are checked against finite differences, the optimiser against a
scalar restatement of optax's clip -> adam -> schedule chain."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from muax_b200 import nn  # noqa: E402
from muax_b200.learner import (Optimizer, TorchNets, default_loss_fn, scale_gradient,  # noqa: E402
                               warmup_exponential_decay_schedule)
from muax_b200.nn import MZNetworkParams, NetSpec  # noqa: E402
from muax_b200.utils import scalar_to_support  # noqa: E402

PINS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_pins.npz")


def _load(dtype=torch.float64, requires_grad=False):
    z = np.load(PINS)
    groups = {g: {} for g in MZNetworkParams._fields}
    for k in z.files:
        if k.startswith("param|"):
            _, g, m, leaf = k.split("|")
            groups[g].setdefault(m, {})[leaf] = torch.tensor(z[k], dtype=dtype, requires_grad=requires_grad)
    params = MZNetworkParams(**groups)
    batch = {f: torch.tensor(z[f"batch_{f}"]) for f in ("obs", "a", "r", "Rn", "pi")}
    for f in ("obs", "r", "Rn", "pi"):
        batch[f] = batch[f].to(dtype)
    spec = NetSpec(nn.Representation(8), nn.Prediction(2, 21), nn.Dynamic(8, 2, 21), obs_dim=4)
    return z, TorchNets(spec), params, batch


def test_scalar_to_support_matches_reference():
    z = np.load(PINS)
    got = scalar_to_support(torch.tensor(z["sts_x"]), int(z["support_size"])).numpy()
    np.testing.assert_allclose(got, z["sts_y"], rtol=0, atol=1e-12)
    assert np.allclose(got.sum(-1), 1.0)


def test_loss_value_matches_reference_default_loss_fn():
    z, nets, params, batch = _load(torch.float64)
    loss = default_loss_fn(nets, params, batch, int(z["support_size"]))
    assert abs(float(loss) - float(z["loss"])) < 1e-10, (float(loss), float(z["loss"]))
    z32, nets32, params32, batch32 = _load(torch.float32)
    loss32 = default_loss_fn(nets32, params32, batch32, int(z["support_size"]))
    assert abs(float(loss32) - float(z["loss"])) < 1e-4


def test_gradients_match_finite_differences_and_scale_gradient_halves_the_latent_path(monkeypatch):
    import muax_b200.learner as L
    z, nets, params, batch = _load(torch.float64, requires_grad=True)
    S = int(z["support_size"])
    leaves = [p for tree in params for mod in tree.values() for p in mod.values()]
    # (a) with scale_gradient = identity the autograd gradient is the plain derivative of the reference forward
    monkeypatch.setattr(L, "scale_gradient", lambda g, scale=1.0: g)
    grads = torch.autograd.grad(default_loss_fn(nets, params, batch, S), leaves)
    rng = np.random.default_rng(0)
    for _ in range(12):
        i = int(rng.integers(len(leaves)))
        idx = tuple(int(rng.integers(n)) for n in leaves[i].shape)
        eps = 1e-6
        with torch.no_grad():
            leaves[i][idx] += eps
            up = float(default_loss_fn(nets, params, batch, S))
            leaves[i][idx] -= 2 * eps
            dn = float(default_loss_fn(nets, params, batch, S))
            leaves[i][idx] += eps
        fd = (up - dn) / (2 * eps)
        assert abs(fd - float(grads[i][idx])) < 1e-6 * max(1.0, abs(fd)), (i, idx, fd, float(grads[i][idx]))
    monkeypatch.undo()
    # (b) scale_gradient: forward identity, backward scaled (muax/utils.py:54-56)
    x = torch.tensor([1.5, -2.0], dtype=torch.float64, requires_grad=True)
    y = scale_gradient(x, 0.5)
    assert torch.equal(y, x)
    (g,) = torch.autograd.grad((y ** 2).sum(), x)
    assert torch.allclose(g, 0.5 * 2 * x.detach())
    # and inside the loss it changes the representation gradient but not the loss value
    g_scaled = torch.autograd.grad(default_loss_fn(nets, params, batch, S), leaves)
    rep = [i for i, p in enumerate(leaves) if p.shape == (4, 8)][0]
    assert not torch.allclose(g_scaled[rep], grads[rep])


def test_schedule_and_optimizer_follow_optax_semantics():
    sch = warmup_exponential_decay_schedule(0.0, 0.02, 1000, 10000, 0.8, 0.001)
    assert sch(0) == 0.0 and abs(sch(500) - 0.01) < 1e-12 and abs(sch(1000) - 0.02) < 1e-12
    assert abs(sch(11000) - 0.02 * 0.8) < 1e-12 and sch(10 ** 7) == 0.001
    # three steps of clip(1.0) -> adam -> lr against a scalar restatement
    opt = Optimizer(init_value=0.01, peak_value=0.01, end_value=0.01, warmup_steps=0, transition_steps=1, decay_rate=1.0,
                    clip_by_global_norm=1.0)
    p = [torch.tensor([1.0, -2.0], dtype=torch.float64), torch.tensor([[0.5]], dtype=torch.float64)]
    ref = [x.clone().numpy() for x in p]
    mu = [np.zeros_like(x) for x in ref]
    nu = [np.zeros_like(x) for x in ref]
    for t in range(1, 4):
        grads = [3.0 * x.clone() for x in p]  # gradient of 1.5 * |p|^2: norm > 1, so the clip is active
        opt.step(p, grads)
        g_ref = [3.0 * x for x in ref]
        norm = np.sqrt(sum((g ** 2).sum() for g in g_ref))
        g_ref = [g * (1.0 / max(norm, 1.0)) for g in g_ref]
        for i, g in enumerate(g_ref):
            mu[i] = 0.9 * mu[i] + 0.1 * g
            nu[i] = 0.999 * nu[i] + 0.001 * g * g
            ref[i] = ref[i] - 0.01 * (mu[i] / (1 - 0.9 ** t)) / (np.sqrt(nu[i] / (1 - 0.999 ** t)) + 1e-8)
        for a, b in zip(p, ref):
            np.testing.assert_allclose(a.numpy(), b, rtol=1e-12, atol=1e-14)
