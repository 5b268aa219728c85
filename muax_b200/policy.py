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


class StochasticMuZeroPolicy(Policy):  # muax/policy.py:50-67 — SURVEY.md §8(f) rank 3, not built yet
    def __call__(self, params, rng_key, root, recurrent_fn=None, decision_recurrent_fn=None,
                 chance_recurrent_fn=None, **kwargs):
        raise NotImplementedError("StochasticMuZeroPolicy (afterstate / chance-node search) is outside the "
                                  "accelerated hot path; see DESIGN.md 'out of scope'")
