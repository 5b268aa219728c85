"""Policy plug-ins with the reference's interface (muax/policy.py:7-67), backed by the CUDA search.

`MuZero.__init__` takes `policy_class=` and calls `policy_class()(params, rng_key, root, recurrent_fn,
**kwargs)` (muax/model.py:56,232-242); the returned object must expose `.action` and `.action_weights`.
Here `recurrent_fn` is either the model's bound `_recurrent_inference` (declarative nets -> the search
runs natively end to end) or any torch callable, which is driven through the library's callback mode.
"""
from abc import ABC, abstractmethod
from typing import Any, NamedTuple

from . import _lib


class RootFnOutput(NamedTuple):  # mctx.RootFnOutput (muax/model.py:258-262)
    prior_logits: Any
    value: Any
    embedding: Any


class RecurrentFnOutput(NamedTuple):  # mctx.RecurrentFnOutput (muax/model.py:276-281)
    reward: Any
    discount: Any
    prior_logits: Any
    value: Any


class PolicyOutput(NamedTuple):  # mctx.PolicyOutput: muax reads .action / .action_weights (model.py:174-178)
    action: Any
    action_weights: Any
    search_tree: Any


class Policy(ABC):  # muax/policy.py:7-10
    @abstractmethod
    def __call__(self, params, rng_key, root, recurrent_fn=None, decision_recurrent_fn=None,
                 chance_recurrent_fn=None, **kwargs):
        pass


_QTRANSFORMS = {
    None: None,
    "qtransform_by_parent_and_siblings": _lib.QT_PARENT_AND_SIBLINGS,
    "qtransform_completed_by_mix_value": _lib.QT_COMPLETED_BY_MIX_VALUE,
    _lib.QT_PARENT_AND_SIBLINGS: _lib.QT_PARENT_AND_SIBLINGS,
    _lib.QT_COMPLETED_BY_MIX_VALUE: _lib.QT_COMPLETED_BY_MIX_VALUE,
}


def qtransform_by_parent_and_siblings():  # stands for mctx.qtransform_by_parent_and_siblings
    raise TypeError("marker for the qtransform= argument; it is evaluated inside the CUDA kernels")


def qtransform_completed_by_mix_value():  # stands for mctx.qtransform_completed_by_mix_value
    raise TypeError("marker for the qtransform= argument; it is evaluated inside the CUDA kernels")


def resolve_qtransform(q):
    """Maps the reference's `qtransform=` callables to the kernel's enum; unknown callables are rejected loudly
    (they would have to run inside the selection kernel)."""
    if callable(q):
        q = getattr(q, "__name__", None)
    try:
        return _QTRANSFORMS[q]
    except (KeyError, TypeError):
        raise ValueError(f"unsupported qtransform {q!r}: use qtransform_by_parent_and_siblings or "
                         "qtransform_completed_by_mix_value") from None


class _EnginePolicy(Policy):
    policy_kind = _lib.POLICY_MUZERO

    def _search_kwargs(self, kwargs):
        raise NotImplementedError

    def __call__(self, params, rng_key, root, recurrent_fn=None, decision_recurrent_fn=None,
                 chance_recurrent_fn=None, **kwargs):
        model = getattr(recurrent_fn, "__self__", None)
        kw = self._search_kwargs(kwargs)
        invalid = kwargs.get("invalid_actions")
        noise = kwargs.get("noise")
        if model is not None and getattr(model, "_native", False):
            engine = model._engine_for(root.embedding.shape[0], kw["num_simulations"], params)
            action, weights, _ = engine.search(rng_key, root=(root.prior_logits, root.value, root.embedding),
                                               invalid_actions=invalid, noise=noise, **kw)
        else:
            if model is None or not hasattr(model, "_callback_engine_for"):
                raise ValueError("recurrent_fn must be a bound method of a muax_b200.MuZero model")
            engine = model._callback_engine_for(root, kw["num_simulations"])

            def step(action, embedding):
                out, nxt = recurrent_fn(params, None, action, embedding)
                return out.reward, out.discount, out.prior_logits, out.value, nxt

            action, weights, _ = engine.search_with_callback(
                rng_key, (root.prior_logits, root.value, root.embedding), step, invalid_actions=invalid, noise=noise,
                **kw)
        return PolicyOutput(action=action, action_weights=weights, search_tree=engine)


class MuZeroPolicy(_EnginePolicy):  # muax/policy.py:13-30
    policy_kind = _lib.POLICY_MUZERO

    def _search_kwargs(self, kwargs):
        return dict(
            policy=self.policy_kind,
            num_simulations=kwargs.get("num_simulations", 5),
            temperature=kwargs.get("temperature", 1.0),
            max_depth=kwargs.get("max_depth"),
            qtransform=resolve_qtransform(kwargs.get("qtransform", _lib.QT_PARENT_AND_SIBLINGS)),
            dirichlet_fraction=kwargs.get("dirichlet_fraction", 0.25),
            dirichlet_alpha=kwargs.get("dirichlet_alpha", 0.3),
            pb_c_init=kwargs.get("pb_c_init", 1.25),
            pb_c_base=kwargs.get("pb_c_base", 19652),
            global_batch=kwargs.get("global_batch"), batch_offset=kwargs.get("batch_offset", 0),
            engine=kwargs.get("engine", _lib.ENGINE_AUTO), want_tree=kwargs.get("want_tree", False),
            precision=kwargs.get("precision", _lib.PRECISION_FP32))


class GumbelMuZeroPolicy(_EnginePolicy):  # muax/policy.py:33-47
    policy_kind = _lib.POLICY_GUMBEL

    def _search_kwargs(self, kwargs):
        return dict(
            policy=self.policy_kind,
            num_simulations=kwargs.get("num_simulations", 5),
            max_depth=kwargs.get("max_depth"),
            qtransform=resolve_qtransform(kwargs.get("qtransform", _lib.QT_COMPLETED_BY_MIX_VALUE)),
            max_num_considered_actions=kwargs.get("max_num_considered_actions", 16),
            gumbel_scale=kwargs.get("gumbel_scale", 1),
            global_batch=kwargs.get("global_batch"), batch_offset=kwargs.get("batch_offset", 0),
            engine=kwargs.get("engine", _lib.ENGINE_AUTO), want_tree=kwargs.get("want_tree", False),
            precision=kwargs.get("precision", _lib.PRECISION_FP32))


class DecisionRecurrentFnOutput(NamedTuple):  # mctx.DecisionRecurrentFnOutput
    chance_logits: Any
    afterstate_value: Any


class ChanceRecurrentFnOutput(NamedTuple):  # mctx.ChanceRecurrentFnOutput
    action_logits: Any
    value: Any
    reward: Any
    discount: Any


class StochasticMuZeroPolicy(Policy):  # muax/policy.py:50-67 (mctx.stochastic_muzero_policy)
    """Afterstate / chance-node search.  The tree kernels run in the library (callback mode) on A' = A + C
    pseudo-actions: decision nodes (even depth) select with pUCT over the A' slots, chance nodes (odd depth) with
    `argmax prior / (visits + 1)`; Dirichlet noise, the invalid-action mask, the visit summary and the final draw see
    the A decision actions only.  The two recurrent functions are torch callables with mctx's signatures:

        decision_recurrent_fn(params, rng_key, action i32[B], state_embedding[B, Es])
            -> (DecisionRecurrentFnOutput(chance_logits[B, C], afterstate_value[B]), afterstate_embedding[B, Ea])
        chance_recurrent_fn(params, rng_key, chance_outcome i32[B], afterstate_embedding[B, Ea])
            -> (ChanceRecurrentFnOutput(action_logits[B, A], value[B], reward[B], discount[B]), state_embedding[B, Es])

    As in mctx both are evaluated for every row of a simulation and one result is kept by the parent's node type (the
    index a row does not use is clamped into range instead of relying on XLA's out-of-range gather).  The node
    embedding is the flat row [state | afterstate | is_decision]."""

    def __init__(self, discount=1.0, prng_mode=_lib.PRNG_LEGACY):
        self._engines = {}
        self._discount, self._prng_mode = discount, prng_mode

    def _engine(self, B, A2, E2, num_simulations, device):
        import numpy as np

        from .nn import pack_stacks
        from .search import SearchEngine
        key = (B, A2, E2, str(device))
        eng = self._engines.get(key)
        if eng is None or eng.max_num_simulations < num_simulations:
            z = lambda i, o: [(np.zeros((i, o), np.float32), np.zeros(o, np.float32))]  # noqa: E731
            _, cstacks = pack_stacks(dict(pred_v=z(E2, 1), pred_pi=z(E2, A2), dyn_ns=z(E2 + A2, E2), dyn_r=z(E2 + A2, 1)))
            if eng is not None:
                eng.close()
            eng = SearchEngine(cstacks, batch=B, num_actions=A2, embed_dim=E2, obs_dim=0, support_size=0,
                               max_num_simulations=max(int(num_simulations), 1), discount=self._discount,
                               prng_mode=self._prng_mode, device=device)
            self._engines[key] = eng
        return eng

    def __call__(self, params, rng_key, root, recurrent_fn=None, decision_recurrent_fn=None,
                 chance_recurrent_fn=None, **kwargs):
        import torch
        if decision_recurrent_fn is None or chance_recurrent_fn is None:
            raise ValueError("StochasticMuZeroPolicy needs decision_recurrent_fn and chance_recurrent_fn")
        dev = kwargs.get("device")
        f32 = torch.float32

        def t(x, dtype=f32):
            x = x if isinstance(x, torch.Tensor) else torch.as_tensor(x)
            return x.to(device=dev if dev is not None else (x.device if x.is_cuda else "cuda"), dtype=dtype)

        logits, value, state = t(root.prior_logits), t(root.value), t(root.embedding)
        dev = logits.device
        B, A = logits.shape
        Es = state.shape[1]
        dummy, after0 = decision_recurrent_fn(params, None, torch.zeros(B, dtype=torch.int32, device=dev), state)
        C, Ea = int(dummy.chance_logits.shape[-1]), int(after0.shape[1])
        A2, E2 = A + C, Es + Ea + 1
        num_simulations = kwargs.get("num_simulations", 5)
        engine = self._engine(B, A2, E2, num_simulations, dev)
        ninf = lambda n: torch.full((B, n), float("-inf"), dtype=f32, device=dev)  # noqa: E731
        zeros = lambda n, dt=f32: torch.zeros((B, n), dtype=dt, device=dev)  # noqa: E731
        root2 = (torch.cat([logits, zeros(C)], 1), value,
                 torch.cat([state, t(after0), torch.ones((B, 1), dtype=f32, device=dev)], 1))
        invalid = kwargs.get("invalid_actions")
        if invalid is not None:
            invalid = torch.cat([t(invalid, torch.uint8), zeros(C, torch.uint8)], 1)
        noise = kwargs.get("noise")
        if noise is not None:
            noise = torch.cat([t(noise), zeros(C)], 1)

        def step(action, emb):
            st, af, is_dec = emb[:, :Es], emb[:, Es:Es + Ea], emb[:, -1] != 0
            dec, new_af = decision_recurrent_fn(params, None, action.clamp(max=A - 1), st)
            ch, new_st = chance_recurrent_fn(params, None, (action - A).clamp(min=0), af)
            p_logits = torch.where(is_dec[:, None], torch.cat([ninf(A), t(dec.chance_logits)], 1),
                                   torch.cat([t(ch.action_logits), ninf(C)], 1))
            val = torch.where(is_dec, t(dec.afterstate_value), t(ch.value))
            reward = torch.where(is_dec, torch.zeros_like(val), t(ch.reward))
            discount = torch.where(is_dec, torch.ones_like(val), t(ch.discount))
            new_emb = torch.cat([t(new_st), t(new_af), (~is_dec).to(f32)[:, None]], 1)
            return reward, discount, p_logits, val, new_emb

        action, weights, _ = engine.search_with_callback(
            rng_key, root2, step, invalid_actions=invalid, noise=noise, policy=_lib.POLICY_MUZERO,
            num_simulations=num_simulations, temperature=kwargs.get("temperature", 1.0),
            max_depth=kwargs.get("max_depth"),
            qtransform=resolve_qtransform(kwargs.get("qtransform", _lib.QT_PARENT_AND_SIBLINGS)),
            dirichlet_fraction=kwargs.get("dirichlet_fraction", 0.25),
            dirichlet_alpha=kwargs.get("dirichlet_alpha", 0.3), pb_c_init=kwargs.get("pb_c_init", 1.25),
            pb_c_base=kwargs.get("pb_c_base", 19652), engine=_lib.ENGINE_STEPWISE, want_tree=True,
            num_decision_actions=A)
        return PolicyOutput(action=action, action_weights=weights[:, :A].contiguous(), search_tree=engine)
