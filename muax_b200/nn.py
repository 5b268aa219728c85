"""Declarative mirrors of the reference's haiku networks (muax/nn.py:59-115) and their factories.

The reference builds `hk.Module`s inside functions that `hk.transform` later traces.  Without JAX the
same information — layer widths, activation, min-max on/off — is carried by small declarative classes
with the SAME constructor signatures and the SAME haiku parameter-dict layout
(`{'representation/linear': {'w': [in,out], 'b': [out]}, ...}`), which is what the CUDA
engine consumes.  The reference creates its `hk.Linear`s inside the parent module's `__init__`
(muax/nn.py:63-65), for which haiku spells the path `representation/~/linear`; both spellings are accepted
everywhere and normalised to the short one (`_canon`, `canonical_params`).  Custom architectures subclass and override `hidden` / `normalize` / `activation`
(e.g. the LunarLander notebook's 64-64-16 stacks, examples/lunarlander.ipynb cell 2).
"""
from typing import Callable, NamedTuple, Optional

import numpy as np

from . import _lib


class MZNetworkParams(NamedTuple):  # muax/nn.py:11-14
    representation: Optional[dict] = None
    prediction: Optional[dict] = None
    dynamic: Optional[dict] = None


class MZNetwork(NamedTuple):  # muax/nn.py:17-20
    representation_fn: Callable
    prediction_fn: Callable
    dynamic_fn: Callable


def _haiku_name(prefix, i):
    return f"{prefix}/linear" if i == 0 else f"{prefix}/linear_{i}"


def _canon(path):
    """haiku module path -> this package's spelling: `representation/~/linear_1` == `representation/linear_1`."""
    return path.replace("/~/", "/")


def canonical_params(params):
    """MZNetworkParams (or a 3-tuple of haiku dicts) -> MZNetworkParams with canonical module paths, leaves untouched."""
    trees = []
    for tree in params:
        trees.append(None if tree is None else {_canon(mod): dict(leaves) for mod, leaves in tree.items()})
    return MZNetworkParams(*trees)


def _trunc_normal(rng, shape, stddev):
    w = rng.standard_normal(shape)
    bad = np.abs(w) > 2
    while bad.any():
        w[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(w) > 2
    return (w * stddev).astype(np.float32)


class Module:
    """Base of the declarative modules: an ordered list of MLP heads sharing one input."""
    hidden = (16,)
    activation = "elu"
    normalize = True

    def __init__(self, name):
        self.name = name

    def heads(self, in_dim):  # -> [(head name, [widths...])] in haiku construction order
        raise NotImplementedError

    def init(self, rng, in_dim):
        """hk.Linear default init: w ~ TruncatedNormal(0, 1/sqrt(fan_in)), b = 0."""
        params, i = {}, 0
        for _, dims in self.heads(in_dim):
            for fan_in, fan_out in zip(dims[:-1], dims[1:]):
                params[_haiku_name(self.name, i)] = {
                    "w": _trunc_normal(rng, (fan_in, fan_out), 1.0 / np.sqrt(fan_in)),
                    "b": np.zeros(fan_out, np.float32)}
                i += 1
        return params

    def stacks(self, params, in_dim):
        """-> {head name: [(w, b), ...]} picked out of a haiku-shaped param dict."""
        out, i = {}, 0
        for head, dims in self.heads(in_dim):
            layers = []
            for fan_in, fan_out in zip(dims[:-1], dims[1:]):
                name = _haiku_name(self.name, i)
                p = params[name] if name in params else params[name.replace("/", "/~/", 1)]  # haiku's own spelling
                w, b = _to_numpy(p["w"]), _to_numpy(p["b"])
                if w.shape != (fan_in, fan_out) or b.shape != (fan_out,):
                    raise ValueError(f"{_haiku_name(self.name, i)}: expected w{(fan_in, fan_out)}, got {w.shape}")
                layers.append((w, b))
                i += 1
            out[head] = layers
        return out


def _to_numpy(x):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


class Representation(Module):  # muax/nn.py:59-70
    hidden = ()

    def __init__(self, embedding_dim, name="representation"):
        super().__init__(name)
        self.embedding_dim = embedding_dim

    def heads(self, in_dim):
        return [("repr", [in_dim, *self.hidden, self.embedding_dim])]


class Prediction(Module):  # muax/nn.py:73-90
    def __init__(self, num_actions, full_support_size, name="prediction"):
        super().__init__(name)
        self.num_actions = num_actions
        self.full_support_size = full_support_size

    def heads(self, in_dim):
        return [("pred_v", [in_dim, *self.hidden, self.full_support_size]),
                ("pred_pi", [in_dim, *self.hidden, self.num_actions])]


class Dynamic(Module):  # muax/nn.py:93-115
    def __init__(self, embedding_dim, num_actions, full_support_size, name="dynamic"):
        super().__init__(name)
        self.embedding_dim = embedding_dim
        self.num_actions = num_actions
        self.full_support_size = full_support_size

    def heads(self, in_dim):  # in_dim = embedding_dim + num_actions (one-hot concat, nn.py:105-108)
        return [("dyn_ns", [in_dim, *self.hidden, self.embedding_dim]),
                ("dyn_r", [in_dim, *self.hidden, self.full_support_size])]


class NetFn:
    """What `_init_*_func` returns: stands where the reference has a to-be-transformed python function."""

    def __init__(self, module_cls, *args):
        self.module_cls, self.args = module_cls, args

    def build(self):
        return self.module_cls(*self.args)


def _init_representation_func(representation_module, embedding_dim):  # muax/nn.py:417-421
    return NetFn(representation_module, embedding_dim)


def _init_prediction_func(prediction_module, num_actions, full_support_size):  # muax/nn.py:423-427
    return NetFn(prediction_module, num_actions, full_support_size)


def _init_dynamic_func(dynamic_module, embedding_dim, num_actions, full_support_size):  # muax/nn.py:429-433
    return NetFn(dynamic_module, embedding_dim, num_actions, full_support_size)


def create_muzero_network(representation_module, prediction_module, dynamic_module, embedding_dim, num_actions,
                          full_support_size) -> MZNetwork:  # muax/nn.py:23-34
    return MZNetwork(_init_representation_func(representation_module, embedding_dim),
                     _init_prediction_func(prediction_module, num_actions, full_support_size),
                     _init_dynamic_func(dynamic_module, embedding_dim, num_actions, full_support_size))


_ACT = {"elu": _lib.ACT_ELU, "relu": _lib.ACT_RELU}
STACK_ORDER = ("repr", "pred_v", "pred_pi", "dyn_ns", "dyn_r")


class NetSpec:
    """The three modules resolved against concrete sizes: everything the engine needs except weights."""

    def __init__(self, representation, prediction, dynamic, obs_dim):
        self.representation, self.prediction, self.dynamic = representation, prediction, dynamic
        self.obs_dim = int(obs_dim)
        self.embed_dim = int(dynamic.embedding_dim)
        self.num_actions = int(prediction.num_actions)
        self.full_support_size = int(prediction.full_support_size)
        if self.full_support_size % 2 != 1:
            raise ValueError("full_support_size must be 2 * support_size + 1")
        if representation is not None and representation.embedding_dim != self.embed_dim:
            raise ValueError("Representation and Dynamic disagree on embedding_dim")
        if dynamic.num_actions != self.num_actions or dynamic.full_support_size != self.full_support_size:
            raise ValueError("Prediction and Dynamic disagree on num_actions / full_support_size")
        acts = {m.activation for m in (representation, prediction, dynamic) if m is not None}
        if len(acts) != 1 or next(iter(acts)) not in _ACT:
            raise ValueError(f"all modules must share one activation out of {sorted(_ACT)}, got {sorted(acts)}")
        self.activation = _ACT[next(iter(acts))]
        self.repr_minmax = int(bool(representation.normalize)) if representation is not None else 0
        self.dyn_minmax = int(bool(dynamic.normalize))

    def init(self, rng):
        rep = self.representation.init(rng, self.obs_dim) if self.representation is not None else None
        return MZNetworkParams(rep, self.prediction.init(rng, self.embed_dim),
                               self.dynamic.init(rng, self.embed_dim + self.num_actions))

    def stacks(self, params):
        out = {}
        if self.representation is not None:
            out.update(self.representation.stacks(params.representation, self.obs_dim))
        out.update(self.prediction.stacks(params.prediction, self.embed_dim))
        out.update(self.dynamic.stacks(params.dynamic, self.embed_dim + self.num_actions))
        return out

    def pack(self, params):
        """-> (float32 blob, {stack name: _lib.Stack}): W [in,out] row-major then b, in STACK_ORDER."""
        return pack_stacks(self.stacks(params))


def pack_stacks(stacks):
    chunks, off, cstacks = [], 0, {}
    for name in STACK_ORDER:
        st = _lib.Stack()
        layers = stacks.get(name, [])
        if len(layers) > _lib.MAX_LAYERS:
            raise ValueError(f"{name}: at most {_lib.MAX_LAYERS} layers are supported")
        st.n_layers = len(layers)
        for l, (w, b) in enumerate(layers):
            w, b = _to_numpy(w), _to_numpy(b)
            st.in_dim[l], st.out_dim[l] = w.shape
            pad = (-off) % 4  # weight matrices start on 16-byte boundaries: the kernels fetch 4 output units per load
            if pad:
                chunks.append(np.zeros(pad, np.float32))
                off += pad
            st.w_off[l] = off
            off += w.size
            st.b_off[l] = off
            off += b.size
            chunks += [w.ravel(), b.ravel()]
        cstacks[name] = st
    blob = np.ascontiguousarray(np.concatenate(chunks), dtype=np.float32)
    return blob, cstacks
