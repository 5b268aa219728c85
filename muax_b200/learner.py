"""Learner on the GPU: `MuZero.update` without JAX.

  default_loss_fn   muax/frameworks/coax/loss.py:10-78 (HEAD: muax/loss.py:9-88) — k-step unrolled model loss:
                    categorical cross-entropies of reward / n-step value / search policy, `scale_gradient(s, 0.5)`
                    on the latent between unroll steps, divided by the unroll length, + 1e-4 * L2
  optimizer         muax/frameworks/coax/model.py:23-70 — clip_by_global_norm -> scale_by_adam ->
                    warmup_exponential_decay_schedule -> descend (optax semantics, re-implemented on torch tensors)

The networks are the declarative stacks of muax_b200/nn.py evaluated with torch autograd on the haiku-shaped
parameter dict, so the parameters the learner updates are exactly the arrays the search engine packs
(`model.params = ...` re-packs the weight blob on the next `act`).  The forward value of the loss is pinned against
the reference's own `default_loss_fn` (tests/golden/make_loss_pins.py executes the reference source with NumPy
stand-ins for the jax / optax calls it makes); gradients are checked against finite differences of that loss.
"""
import math

import numpy as np
import torch

from .nn import MZNetworkParams, _haiku_name
from .utils import scalar_to_support


def scale_gradient(g, scale: float = 1.0):  # muax/utils.py:54-56
    return g * scale + g.detach() * (1.0 - scale)


def min_max_normalize(s):  # muax/nn.py:37-44
    s_min = s.min(dim=-1, keepdim=True).values
    s_max = s.max(dim=-1, keepdim=True).values
    scale = s_max - s_min
    scale = torch.where(scale < 1e-5, scale + 1e-5, scale)
    return (s - s_min) / scale


class TorchNets:
    """The three networks as pure functions of a `{module_path: {'w','b'}}` dict of torch tensors."""

    def __init__(self, spec):
        self.spec = spec
        self.act = torch.nn.functional.elu if spec.activation == 0 else torch.relu

    def _heads(self, module, params, x, in_dim):
        outs, i = [], 0
        for _, dims in module.heads(in_dim):
            h = x
            n = len(dims) - 1
            for l in range(n):
                p = params[_haiku_name(module.name, i)]
                h = h @ p["w"] + p["b"]
                if l < n - 1:
                    h = self.act(h)
                i += 1
            outs.append(h)
        return outs

    def representation(self, params, obs):  # muax/nn.py:67-70
        s = self._heads(self.spec.representation, params, obs.reshape(obs.shape[0], -1), self.spec.obs_dim)[0]
        return min_max_normalize(s) if self.spec.repr_minmax else s

    def prediction(self, params, s):  # muax/nn.py:86-90 -> (value logits, policy logits)
        v, logits = self._heads(self.spec.prediction, params, s, self.spec.embed_dim)
        return v, logits

    def dynamic(self, params, s, a):  # muax/nn.py:105-115 -> (reward logits, next state)
        onehot = torch.nn.functional.one_hot(a.to(torch.int64), self.spec.num_actions).to(s.dtype)
        sa = torch.cat([s, onehot], dim=-1)
        ns, r = self._heads(self.spec.dynamic, params, sa, self.spec.embed_dim + self.spec.num_actions)
        return r, (min_max_normalize(ns) if self.spec.dyn_minmax else ns)


def softmax_cross_entropy(logits, labels):  # optax.softmax_cross_entropy
    return -(labels * torch.log_softmax(logits, dim=-1)).sum(-1)


def default_loss_fn(nets, params, batch, support_size, c=1e-4):
    """muax/frameworks/coax/loss.py:10-78 on torch tensors.  batch fields are [B, L, ...]; pi may carry the
    reference's extra axis ([B, L, 1, A])."""
    B, L = batch["a"].shape
    r_t = scalar_to_support(batch["r"], support_size).reshape(B, L, -1)
    Rn_t = scalar_to_support(batch["Rn"], support_size).reshape(B, L, -1)
    pi_t = batch["pi"].reshape(B, L, -1)
    s = nets.representation(params.representation, batch["obs"][:, 0])
    loss = 0.0
    for i in range(L):
        v, logits = nets.prediction(params.prediction, s)
        s = scale_gradient(s, 0.5)
        r, ns = nets.dynamic(params.dynamic, s, batch["a"][:, i].flatten())
        loss = loss + softmax_cross_entropy(r, r_t[:, i]).mean() + softmax_cross_entropy(v, Rn_t[:, i]).mean() \
            + softmax_cross_entropy(logits, pi_t[:, i]).mean()
        s = ns
    loss = loss / L
    l2 = 0.5 * sum((p ** 2).sum() for tree in params for mod in tree.values() for p in mod.values())
    return loss + c * l2


def warmup_exponential_decay_schedule(init_value, peak_value, warmup_steps, transition_steps, decay_rate,
                                      end_value):
    """optax.warmup_exponential_decay_schedule: linear warm-up, then peak * decay_rate ** (t / transition_steps)
    bounded by end_value."""
    def schedule(count):
        if count < warmup_steps:
            return init_value + (peak_value - init_value) * count / max(warmup_steps, 1)
        t = count - warmup_steps
        value = peak_value * decay_rate ** (t / transition_steps)
        return max(value, end_value) if decay_rate < 1 else min(value, end_value)
    return schedule


class Optimizer:
    """muax/frameworks/coax/model.py:23-70: clip_by_global_norm(max) -> adam (b1=.9, b2=.999, eps=1e-8) ->
    learning-rate schedule -> descend."""

    def __init__(self, init_value=0.0, peak_value=2e-2, end_value=1e-3, warmup_steps=1000, transition_steps=10000,
                 decay_rate=0.8, clip_by_global_norm=1.0, b1=0.9, b2=0.999, eps=1e-8):
        self.schedule = warmup_exponential_decay_schedule(init_value, peak_value, warmup_steps, transition_steps,
                                                          decay_rate, end_value)
        self.clip, self.b1, self.b2, self.eps = float(clip_by_global_norm), b1, b2, eps
        self.count = 0
        self.mu = self.nu = None

    def step(self, leaves, grads):
        gnorm = torch.sqrt(sum((g ** 2).sum() for g in grads))
        scale = torch.clamp(self.clip / (gnorm + 1e-16), max=1.0) if self.clip > 0 else 1.0  # optax: g * clip / max(norm, clip)
        if self.mu is None:
            self.mu = [torch.zeros_like(g) for g in grads]
            self.nu = [torch.zeros_like(g) for g in grads]
        lr = self.schedule(self.count)
        self.count += 1
        c1, c2 = 1 - self.b1 ** self.count, 1 - self.b2 ** self.count
        with torch.no_grad():
            for p, g, m, v in zip(leaves, grads, self.mu, self.nu):
                g = g * scale
                m.mul_(self.b1).add_(g, alpha=1 - self.b1)
                v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                p.sub_(lr * (m / c1) / (torch.sqrt(v / c2) + self.eps))
        return float(lr), gnorm


def optimizer(**kw):
    """Factory with the reference's name and keyword arguments (coax/model.py:23)."""
    return Optimizer(**kw)


class Learner:
    """Holds the parameters as torch tensors on the search's device and applies `update` steps."""

    def __init__(self, model, opt=None, device=None):
        self.model = model
        self.device = torch.device(device or "cuda")
        self.nets = TorchNets(model._spec)
        self.opt = opt or Optimizer()
        self.params = MZNetworkParams(*[
            {mod: {k: torch.tensor(np.asarray(v), dtype=torch.float32, device=self.device, requires_grad=True)
                   for k, v in leaves.items()} for mod, leaves in tree.items()} for tree in model.params])
        self._leaves = [p for tree in self.params for mod in tree.values() for p in mod.values()]

    def _to_torch(self, batch):
        out = {}
        for f in ("obs", "a", "r", "Rn", "pi"):
            x = getattr(batch, f) if not isinstance(batch, dict) else batch[f]
            out[f] = torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x, device=self.device)
        out["obs"], out["r"], out["Rn"], out["pi"] = (out[k].to(torch.float32) for k in ("obs", "r", "Rn", "pi"))
        return out

    def loss(self, batch):
        return default_loss_fn(self.nets, self.params, self._to_torch(batch), self.model._support_size)

    def update(self, batch, sync_model=True):
        loss = self.loss(batch)
        grads = torch.autograd.grad(loss, self._leaves)
        lr, gnorm = self.opt.step(self._leaves, grads)
        if sync_model:
            self.push()
        return {"loss": loss.detach(), "lr": lr, "grad_norm": gnorm.detach()}

    def restore_optimizer(self, flat):
        """`opt|count`, `opt|mu|i`, `opt|nu|i` arrays of MuZero.save -> this learner's optimiser (leaf order = the
        order of `self._leaves`, which is the order save() wrote them in)."""
        n = len(self._leaves)
        if "opt|count" not in flat or any(f"opt|mu|{i}" not in flat for i in range(n)):
            return False
        mu = [torch.as_tensor(flat[f"opt|mu|{i}"], device=self.device, dtype=torch.float32) for i in range(n)]
        nu = [torch.as_tensor(flat[f"opt|nu|{i}"], device=self.device, dtype=torch.float32) for i in range(n)]
        if any(m.shape != p.shape for m, p in zip(mu, self._leaves)):
            return False
        self.opt.count, self.opt.mu, self.opt.nu = int(flat["opt|count"]), mu, nu
        return True

    def push(self):
        """Hands the current parameters to the acting side (`model.params`): the next `act` re-packs the weight blob."""
        self.model.params = MZNetworkParams(*[
            {mod: {k: v.detach().cpu().numpy() for k, v in leaves.items()} for mod, leaves in tree.items()}
            for tree in self.params])
