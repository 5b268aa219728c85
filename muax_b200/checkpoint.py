"""Checkpoint interop with the reference (`MuZero.save_load`, muax/model.py:203-212).

The reference writes `jnp.save(file, {'params': MZNetworkParams(...), 'optimizer_state': <optax state>})`: a NumPy
`.npy` whose payload is a pickle.  Two obstacles for a JAX-free reader: the pickle names `muax.nn.MZNetworkParams`
(and optax / haiku classes) by module path, and its leaves are `jax.Array`s, which only jax can unpickle.  So:

  * `load_reference_checkpoint(file)` reads such a file when its leaves are NumPy arrays — produced by
    `save_reference_checkpoint` here, or by `tools/export_reference_checkpoint.py` run once inside the JAX
    environment (it maps every jax.Array leaf to NumPy and re-saves).  Class references are resolved by NAME
    (`MZNetworkParams` -> this package's namedtuple, optax state namedtuples -> plain namespaces), haiku parameter
    paths are normalised (`representation/~/linear` == `representation/linear`, see nn._canon);
  * `save_reference_checkpoint(file, params, optimizer_state)` writes the reference's layout with NumPy leaves and
    `muax.nn.MZNetworkParams` as the container class, so `MuZero.save_load(file, save=False)` of the reference opens
    it (haiku accepts NumPy leaves; `optimizer_state=None` makes the reference start a fresh optimiser).
"""
import io
import pickle
import sys
import types
from collections import namedtuple
from contextlib import contextmanager

import numpy as np

from .nn import MZNetworkParams, canonical_params


class _Namespace(dict):
    """Stand-in for optax / haiku state classes that are not importable here: keeps fields by name."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None


def _stub_class(module, name):
    def build(*args, **kwargs):
        out = _Namespace(kwargs)
        out["__class__name__"] = f"{module}.{name}"
        out["args"] = args
        return out
    return build


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if name == "MZNetworkParams":
            return MZNetworkParams
        top = module.split(".")[0]
        if top in ("jax", "jaxlib"):
            raise RuntimeError("this checkpoint holds jax.Array leaves, which only jax can unpickle: run "
                               "tools/export_reference_checkpoint.py on it inside the JAX environment first")
        if top in ("optax", "haiku", "chex", "muax", "flax"):
            try:
                return super().find_class(module, name)
            except Exception:
                return _stub_class(module, name)
        return super().find_class(module, name)


def _read_npy_pickle(file):
    with open(file, "rb") as f:
        version = np.lib.format.read_magic(f)
        if version == (1, 0):
            shape, _, dtype = np.lib.format.read_array_header_1_0(f)
        else:
            shape, _, dtype = np.lib.format.read_array_header_2_0(f)
        if not dtype.hasobject:
            raise ValueError(f"{file}: not a pickled-object .npy (the reference saves a dict)")
        obj = _Unpickler(io.BytesIO(f.read())).load()
    if isinstance(obj, np.ndarray) and obj.shape == ():
        obj = obj.item()
    return obj


def load_reference_checkpoint(file):
    """-> (MZNetworkParams with NumPy float32 leaves and canonical haiku paths, optimizer_state or None)."""
    file = str(file)
    if not file.endswith(".npy"):
        file = f"{file}.npy"  # model.py:208-209
    saved = _read_npy_pickle(file)
    if not isinstance(saved, dict) or "params" not in saved:
        raise ValueError(f"{file}: expected the reference's {{'params', 'optimizer_state'}} dict")
    params = saved["params"]
    if not isinstance(params, MZNetworkParams):
        params = MZNetworkParams(*params)
    return canonical_params(params), saved.get("optimizer_state")


@contextmanager
def _reference_container():
    """The class the reference's pickle must name: muax.nn.MZNetworkParams.  When muax is not importable (this image),
    a module of that name holding an equivalent namedtuple is registered for the duration of the dump only."""
    try:
        from muax.nn import MZNetworkParams as ref_cls  # noqa: F401
        yield ref_cls
        return
    except Exception:
        pass
    cls = namedtuple("MZNetworkParams", ["representation", "prediction", "dynamic"], defaults=(None, None, None))
    cls.__module__ = "muax.nn"
    pkg, mod = types.ModuleType("muax"), types.ModuleType("muax.nn")
    mod.MZNetworkParams = cls
    pkg.nn = mod
    saved = {k: sys.modules.get(k) for k in ("muax", "muax.nn")}
    sys.modules["muax"], sys.modules["muax.nn"] = pkg, mod
    try:
        yield cls
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def save_reference_checkpoint(file, params, optimizer_state=None, haiku_style=True):
    """Writes `{'params', 'optimizer_state'}` the way muax/model.py:203-207 does, NumPy leaves.  `haiku_style`: module
    paths as haiku spells them for modules created inside a parent's __init__ (`representation/~/linear`)."""
    file = str(file)
    if not file.endswith(".npy"):
        file = f"{file}.npy"

    def tree(t):
        if t is None:
            return None
        out = {}
        for mod, leaves in t.items():
            name = mod
            if haiku_style and "/~/" not in mod and "/" in mod:
                head, tail = mod.split("/", 1)
                name = f"{head}/~/{tail}"
            out[name] = {k: np.asarray(v, np.float32) for k, v in leaves.items()}
        return out

    with _reference_container() as cls:
        payload = {"params": cls(*(tree(t) for t in params)), "optimizer_state": optimizer_state}
        np.save(file, np.array(payload, dtype=object), allow_pickle=True)
    return file
