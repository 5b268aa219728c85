"""`MuZero` agent object with the reference's acting interface (muax/model.py:16-179), searching on a B200.

The acting half is the accelerated path: `init`, `act`, `_plan`, `_root_inference`, `_recurrent_inference` and
parameter plumbing.  `update` (learner, muax/model.py:181-201) runs the reference's default loss and optimiser on
torch autograd (muax_b200/learner.py); `save` / `load` / `save_load` read and write both this package's `.npz` and the
reference's pickled `.npy` (muax/model.py:203-212, muax_b200/checkpoint.py).
Both constructor shapes are accepted: HEAD's `MuZero(network, policy_class=...)` (model.py:43-50) and the
released `MuZero(repr_fn, pred_fn, dy_fn, policy='muzero'|'gumbel', ...)` (frameworks/coax/model.py:101-110).
"""
import numpy as np
import torch

from . import _lib
from .nn import MZNetwork, MZNetworkParams, NetFn, NetSpec, canonical_params, pack_stacks
from .policy import GumbelMuZeroPolicy, MuZeroPolicy, RecurrentFnOutput, RootFnOutput, resolve_qtransform
from .random import key_words
from .search import SearchEngine

_POLICIES = {"muzero": MuZeroPolicy, "gumbel": GumbelMuZeroPolicy}


class MuZero:
    def __init__(self, network, *args, policy_class=None, policy=None, optimizer=None, loss_fn=None,
                 discount: float = 0.99, support_size: int = 10, device=None, prng_mode: str = "legacy"):
        if isinstance(network, MZNetwork):
            fns = network
            if args:
                raise TypeError("MuZero(network, ...) takes no further positional arguments")
        else:  # released signature: MuZero(representation_fn, prediction_fn, dynamic_fn, policy='muzero', ...)
            if len(args) < 2:
                raise TypeError("MuZero(representation_fn, prediction_fn, dynamic_fn, ...) needs three functions")
            fns = MZNetwork(network, args[0], args[1])
            if len(args) > 2 and policy is None:
                policy = args[2]
        if policy_class is None:
            try:
                policy_class = _POLICIES[policy or "muzero"]
            except KeyError:
                raise ValueError(f"policy must be one of {sorted(_POLICIES)}, got {policy!r}") from None
        self._fns = fns
        self._native = all(isinstance(f, NetFn) for f in fns)
        # hybrid: a torch Representation (e.g. the conv torso of muax_b200.conv, run once per act at the root) in front
        # of declarative Prediction / Dynamic stacks, which are what the search loop evaluates natively
        self._hybrid = (not self._native and isinstance(fns.prediction_fn, NetFn) and isinstance(fns.dynamic_fn, NetFn)
                        and callable(fns.representation_fn))
        if not self._native and not self._hybrid and not all(callable(f) for f in fns):
            raise TypeError("network functions must be muax_b200.nn factories (native) or torch callables")
        self._policy = policy_class()
        self._optimizer = optimizer
        self.loss_fn = loss_fn
        self._discount = float(discount)
        self._support_size = int(support_size)
        self._device = device
        if prng_mode not in ("legacy", "partitionable"):
            raise ValueError("prng_mode must be 'legacy' or 'partitionable'")
        self._prng_mode = _lib.PRNG_LEGACY if prng_mode == "legacy" else _lib.PRNG_PARTITIONABLE
        self._params = None
        self._opt_state = None
        self._spec = None
        self._engines = {}
        self._weights_version = 0
        self._learner = None
        self._learner_version = -1
        self._pending_opt = None

    # ------------------------------------------------------------------ parameters
    def init(self, rng_key, sample_input):  # muax/model.py:62-80
        sample_input = np.asarray(sample_input)
        if not self._native and not self._hybrid:
            raise TypeError("init() builds parameters for muax_b200.nn modules; torch callables own theirs")
        if self._hybrid:  # the torch Representation owns its parameters; Prediction / Dynamic are initialised here
            pred, dyn = self._fns.prediction_fn.build(), self._fns.dynamic_fn.build()
            self._spec = NetSpec(None, pred, dyn, 0)
        else:
            rep, pred, dyn = (f.build() for f in self._fns)
            self._spec = NetSpec(rep, pred, dyn, int(np.prod(sample_input.shape[1:])))
        if self._spec.full_support_size != 2 * self._support_size + 1:
            raise ValueError("full_support_size of the networks must equal 2 * support_size + 1")
        k0, k1 = key_words(rng_key)
        self.params = self._spec.init(np.random.default_rng([k0, k1]))
        return self._params

    @property
    def params(self):
        return self._params

    @params.setter
    def params(self, value):
        value = MZNetworkParams(*value) if not isinstance(value, MZNetworkParams) else value
        if any("/~/" in mod for tree in value if tree for mod in tree):  # haiku's spelling of the reference's modules
            value = canonical_params(value)
        self._params = value
        self._weights_version += 1
        if self._spec is None and (self._native or self._hybrid):
            self._spec = self._spec_from_params(value)

    def _spec_from_params(self, params):
        """Parameters given without `init()` (a loaded checkpoint, reference params): the network spec is the three
        module factories plus the observation width read off the Representation's first layer."""
        if self._hybrid:
            spec = NetSpec(None, self._fns.prediction_fn.build(), self._fns.dynamic_fn.build(), 0)
        else:
            rep, pred, dyn = (f.build() for f in self._fns)
            first = (params.representation or {}).get(f"{rep.name}/linear")
            if first is None:
                return None
            spec = NetSpec(rep, pred, dyn, int(np.asarray(first["w"]).shape[0]))
        if spec.full_support_size != 2 * self._support_size + 1:
            raise ValueError("full_support_size of the networks must equal 2 * support_size + 1")
        return spec

    @property
    def optimizer_state(self):
        return self._opt_state

    def update(self, batch, *args, **kwargs):  # muax/model.py:181-201
        """One optimisation step on a `[B, L, ...]` transition batch (the reference's default loss and optimiser,
        re-hosted on torch autograd — muax_b200/learner.py).  Returns {'loss': ...} like the reference."""
        from .learner import Learner, Optimizer
        if self.loss_fn is not None:
            raise NotImplementedError("custom loss_fn callables are jax functions in the reference; only the default "
                                      "loss is re-hosted (muax_b200/learner.py)")
        if self._learner is None or self._learner_version != self._weights_version:
            opt = self._optimizer if isinstance(self._optimizer, Optimizer) else (
                self._learner.opt if self._learner is not None else Optimizer())
            self._learner = Learner(self, opt=opt, device=self._device)
            if self._pending_opt:  # a loaded checkpoint's Adam moments and schedule position
                self._learner.restore_optimizer(self._pending_opt)
                self._pending_opt = None
        out = self._learner.update(batch)
        self._learner_version = self._weights_version  # push() bumped it: the learner's copy is still current
        self._opt_state = self._learner.opt
        return out

    def save(self, file, reference_format=None):
        """Parameters + optimiser state.  Default: a flat .npz (`<group>|<module>|<w|b>`, `opt|...`) readable without
        JAX; `reference_format=True` (or a file name ending in .npy): the reference's pickled `.npy` dict
        (muax/model.py:203-207) through muax_b200.checkpoint, which `muax.MuZero.save_load(file, save=False)` opens."""
        if reference_format or (reference_format is None and str(file).endswith(".npy")):
            from .checkpoint import save_reference_checkpoint
            return save_reference_checkpoint(file, self._params, None)
        flat = {}
        for group, tree in zip(MZNetworkParams._fields, self._params):
            for mod, leaves in (tree or {}).items():
                for leaf, arr in leaves.items():
                    flat[f"{group}|{mod}|{leaf}"] = np.asarray(arr)
        opt = self._opt_state
        if opt is not None and getattr(opt, "mu", None) is not None:  # Adam moments + schedule count (learner.Optimizer)
            flat["opt|count"] = np.asarray(opt.count, np.int64)
            for i, (m, v) in enumerate(zip(opt.mu, opt.nu)):
                flat[f"opt|mu|{i}"] = m.detach().cpu().numpy()
                flat[f"opt|nu|{i}"] = v.detach().cpu().numpy()
        np.savez(file, **flat)

    def load(self, file):
        """Opens this package's .npz or the reference's .npy (NumPy leaves; see muax_b200/checkpoint.py).  A freshly
        constructed model needs no `init()` first: the network spec is rebuilt from the module factories."""
        import os
        name = str(file)
        if name.endswith(".npy") or (not name.endswith(".npz") and not os.path.exists(f"{name}.npz")
                                     and os.path.exists(f"{name}.npy")):
            from .checkpoint import load_reference_checkpoint
            self.params, self._opt_state = load_reference_checkpoint(name)
            self._pending_opt = None
            return
        if not name.endswith(".npz"):
            name = f"{name}.npz"
        groups = {g: {} for g in MZNetworkParams._fields}
        opt = {}
        with np.load(name) as z:
            for k in z.files:
                if k.startswith("opt|"):
                    opt[k] = z[k]
                    continue
                group, mod, leaf = k.split("|")
                groups[group].setdefault(mod, {})[leaf] = z[k]
        self.params = MZNetworkParams(**groups)
        self._pending_opt = opt or None  # restored into the learner's optimiser on the next update()
        self._learner = None

    def save_load(self, file, save=True):  # muax/model.py:203-212
        self.save(file) if save else self.load(file)

    # ------------------------------------------------------------------ engines
    def _engine_for(self, batch, num_simulations, params=None):
        if self._spec is None:
            raise RuntimeError("call init() (or load parameters) before act()")
        if params is not None and params is not self._params:
            self.params = params
        key = (int(batch), self._device)
        eng = self._engines.get(key)
        if eng is None or eng[0].max_num_simulations < num_simulations:
            blob, cstacks = self._spec.pack(self._params)
            spec = self._spec
            engine = SearchEngine(cstacks, batch=int(batch), num_actions=spec.num_actions, embed_dim=spec.embed_dim,
                                  obs_dim=spec.obs_dim, support_size=self._support_size,
                                  max_num_simulations=max(int(num_simulations), 1), activation=spec.activation,
                                  repr_minmax=spec.repr_minmax, dyn_minmax=spec.dyn_minmax, discount=self._discount,
                                  prng_mode=self._prng_mode, device=self._device)
            if eng is not None:
                eng[0].close()
            eng = [engine, -1]
            self._engines[key] = eng
        if eng[1] != self._weights_version:
            blob, _ = self._spec.pack(self._params)
            eng[0].set_weights(blob)
            eng[1] = self._weights_version
        return eng[0]

    def _callback_engine_for(self, root, num_simulations):
        """Engine without native nets: only the tree kernels run in the library, nets are torch callables."""
        B, A = root.prior_logits.shape
        E = root.embedding.shape[1]
        key = ("callback", int(B), int(A), int(E), self._device)
        eng = self._engines.get(key)
        if eng is None or eng[0].max_num_simulations < num_simulations:
            F = 2 * self._support_size + 1
            z = lambda i, o: [(np.zeros((i, o), np.float32), np.zeros(o, np.float32))]  # noqa: E731
            _, cstacks = pack_stacks(dict(pred_v=z(E, F), pred_pi=z(E, A), dyn_ns=z(E + A, E), dyn_r=z(E + A, F)))
            engine = SearchEngine(cstacks, batch=int(B), num_actions=int(A), embed_dim=int(E), obs_dim=0,
                                  support_size=self._support_size, max_num_simulations=max(int(num_simulations), 1),
                                  discount=self._discount, prng_mode=self._prng_mode, device=self._device)
            if eng is not None:
                eng[0].close()
            eng = [engine, 0]
            self._engines[key] = eng
        return eng[0]

    # ------------------------------------------------------------------ acting
    def act(self, rng_key, obs, with_pi: bool = False, with_value: bool = False, obs_from_batch: bool = False,
            num_simulations: int = 5, temperature: float = 1.0, invalid_actions=None, max_depth: int = None,
            loop_fn=None, qtransform=None, dirichlet_fraction: float = 0.25, dirichlet_alpha: float = 0.3,
            pb_c_init: float = 1.25, pb_c_base: float = 19652, **extra):
        """Same contract as muax/model.py:82-179 (`loop_fn` is accepted and ignored: there is no tracing)."""
        if not isinstance(obs, torch.Tensor):
            # uint8 frames for a torch (conv) Representation travel as uint8: a quarter of the H2D bytes and no host
            # conversion pass; everything else becomes float32 like in the reference
            keep_u8 = self._hybrid and isinstance(obs, np.ndarray) and obs.dtype == np.uint8
            obs = obs if keep_u8 else np.asarray(obs, dtype=np.float32)
        if not obs_from_batch:
            obs = obs[None]
        if invalid_actions is not None and not obs_from_batch and np.ndim(invalid_actions) == 1:
            invalid_actions = np.asarray(invalid_actions)[None]
        plan_output, root_value = self._plan(
            self._params, rng_key, obs, num_simulations=num_simulations, temperature=temperature,
            invalid_actions=invalid_actions, max_depth=max_depth, qtransform=qtransform,
            dirichlet_fraction=dirichlet_fraction, dirichlet_alpha=dirichlet_alpha, pb_c_init=pb_c_init,
            pb_c_base=pb_c_base, **extra)
        action, weights = plan_output.action, plan_output.action_weights
        if isinstance(action, torch.Tensor):  # the one mandatory host sync per act (model.py:173-174)
            action, weights, root_value = action.cpu().numpy(), weights.cpu().numpy(), root_value.cpu().numpy()
        if not obs_from_batch:
            root_value = float(root_value.reshape(-1)[0])
            action = int(action.reshape(-1)[0])
        if with_pi and with_value:
            return action, weights, root_value
        if with_value:
            return action, root_value
        if with_pi:
            return action, weights
        return action

    def act_device(self, rng_key, obs, num_simulations: int = 5, temperature: float = 1.0, **kw):
        """`act` for device-resident loops (muax_b200/actor_device.py): `obs` is a CUDA float32 tensor [B, ...];
        returns CUDA tensors (action i32[B], action_weights f32[B,A], root_value f32[B]) without synchronising.
        `out=(action, weights, root_value)`: preallocated tensors the search kernel writes into (native nets only)."""
        if not isinstance(obs, torch.Tensor) or not obs.is_cuda:
            raise ValueError("act_device expects a CUDA tensor; use act() for host observations")
        if not (self._hybrid and obs.dtype == torch.uint8):
            obs = obs.to(torch.float32)
        plan_output, root_value = self._plan(self._params, rng_key, obs, num_simulations=num_simulations,
                                             temperature=temperature, **kw)
        return plan_output.action, plan_output.action_weights, root_value

    def _plan(self, params, rng_key, obs, num_simulations=5, temperature=1.0, invalid_actions=None, max_depth=None,
              qtransform=None, dirichlet_fraction=0.25, dirichlet_alpha=0.3, pb_c_init=1.25, pb_c_base=19652,
              **extra):  # muax/model.py:222-243
        if qtransform is None:  # model.py:230-231 forces this default for every policy class
            qtransform = _lib.QT_PARENT_AND_SIBLINGS
        out = extra.pop("out", None)
        kwargs = dict(num_simulations=num_simulations, temperature=temperature, invalid_actions=invalid_actions,
                      max_depth=max_depth, qtransform=qtransform, dirichlet_fraction=dirichlet_fraction,
                      dirichlet_alpha=dirichlet_alpha, pb_c_init=pb_c_init, pb_c_base=pb_c_base, **extra)
        fast = self._native and type(self._policy) in (MuZeroPolicy, GumbelMuZeroPolicy)
        if self._hybrid and type(self._policy) in (MuZeroPolicy, GumbelMuZeroPolicy):
            # root Representation in torch (once per act), then the native search from root = (None, None, embedding):
            # the library runs Prediction on the embedding and everything inside the simulation loop
            dev = torch.device("cuda") if self._device is None else torch.device(self._device)
            rep = self._fns.representation_fn
            obs_t = torch.as_tensor(obs, device=dev)
            if getattr(rep, "supports_bf16", False) and extra.get("precision") in ("bf16", _lib.PRECISION_BF16):
                emb = rep(obs_t, bf16=True)
            else:
                emb = rep(obs_t)
            emb = emb.reshape(emb.shape[0], -1).to(torch.float32).contiguous()
            engine = self._engine_for(emb.shape[0], num_simulations, params)
            kw = self._policy._search_kwargs(kwargs)
            action, weights, root_value = engine.search(rng_key, root=(None, None, emb), invalid_actions=invalid_actions,
                                                        noise=extra.get("noise"), out=out, **kw)
            from .policy import PolicyOutput
            return PolicyOutput(action, weights, engine), root_value
        if fast and not isinstance(obs, torch.Tensor):
            # host observations: the whole act is one C-ABI call (H2D, search, D2H) — no torch on the path
            obs2 = np.ascontiguousarray(obs.reshape(obs.shape[0], -1), dtype=np.float32)
            engine = self._engine_for(obs2.shape[0], num_simulations, params)
            kw = self._policy._search_kwargs(kwargs)
            action, weights, root_value = engine.search_host(rng_key, obs2, invalid_actions=invalid_actions,
                                                             noise=extra.get("noise"), **kw)
            from .policy import PolicyOutput
            return PolicyOutput(action, weights, engine), root_value
        if fast:
            obs2 = obs.reshape(obs.shape[0], -1)
            engine = self._engine_for(obs2.shape[0], num_simulations, params)
            kw = self._policy._search_kwargs(kwargs)
            action, weights, root_value = engine.search(rng_key, obs=obs2, invalid_actions=invalid_actions,
                                                        noise=extra.get("noise"), out=out, **kw)
            from .policy import PolicyOutput
            return PolicyOutput(action, weights, engine), root_value
        root = self._root_inference(params, rng_key, obs)
        plan_output = self._policy(params, rng_key, root, self._recurrent_inference, **kwargs)
        if out is not None:  # callback path: results land in fresh tensors, copy them into the caller's buffers
            out[0].copy_(plan_output.action)
            out[1].copy_(plan_output.action_weights)
            out[2].copy_(root.value)
        return plan_output, root.value

    def _root_inference(self, params, rng_key, obs):  # muax/model.py:251-263 (torch-callable networks)
        from .utils import support_to_scalar
        dev = torch.device("cuda") if self._device is None else torch.device(self._device)
        obs = torch.as_tensor(obs, dtype=torch.float32, device=dev)
        s = self._fns.representation_fn(obs)
        v, logits = self._fns.prediction_fn(s)
        v = support_to_scalar(torch.softmax(v, dim=-1), self._support_size).flatten()
        return RootFnOutput(prior_logits=logits, value=v, embedding=s)

    def _recurrent_inference(self, params, rng_key, action, embedding):  # muax/model.py:265-282
        from .utils import support_to_scalar
        r, next_embedding = self._fns.dynamic_fn(embedding, action)
        v, logits = self._fns.prediction_fn(next_embedding)
        r = support_to_scalar(torch.softmax(r, dim=-1), self._support_size).flatten()
        v = support_to_scalar(torch.softmax(v, dim=-1), self._support_size).flatten()
        discount = torch.ones_like(r) * self._discount
        return RecurrentFnOutput(reward=r, discount=discount, prior_logits=logits, value=v), next_embedding
