"""SearchEngine — thin torch-facing wrapper over the C ABI (include/mzsearch.h).

PyTorch is plumbing here: it owns device memory and the stream; every computation of the search runs in
libmzsearch.so's own sm_100a kernels.  There is no CPU or eager fallback — constructing an engine without a
CUDA device raises.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .random import key_words

_TREE_FIELDS = {
    "node_visits": ("i4", "BN"), "parents": ("i4", "BN"), "action_from_parent": ("i4", "BN"),
    "children_index": ("i4", "BNA"), "children_visits": ("i4", "BNA"), "raw_values": ("f4", "BN"),
    "node_values": ("f4", "BN"), "children_prior_logits": ("f4", "BNA"), "children_values": ("f4", "BNA"),
    "children_rewards": ("f4", "BNA"), "children_discounts": ("f4", "BNA"), "embeddings": ("f4", "BNE"),
    "root_noise": ("f4", "BA"), "sim_depth": ("i4", "BS"),
}


class _DevArray:
    """Zero-copy view of library-owned device memory through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<" + typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NULL_CTX = _NullCtx()


def _raw_stream(index):
    """Handle of torch's current stream on device `index` as an int (torch.cuda.current_stream() builds a Stream
    object per call: 14 us of a 0.3 ms act)."""
    try:
        return torch._C._cuda_getCurrentRawStream(index)
    except AttributeError:  # private API moved: the public one
        return torch.cuda.current_stream(index).cuda_stream


def _current_device():
    try:
        return torch._C._cuda_getDevice()
    except AttributeError:
        return torch.cuda.current_device()


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _dev(x, device, dtype, shape=None, name="tensor"):
    if x is None:
        return None
    if isinstance(x, torch.Tensor) and x.dtype == dtype and x.device == device and x.is_contiguous():
        t = x  # already where the kernels want it: no torch dispatch on the hot call path
    else:
        t = torch.as_tensor(x).to(device=device, dtype=dtype, non_blocking=True).contiguous()
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    return t


class SearchEngine:
    """One handle = one (device, batch, network shape, max simulations) instance of the CUDA search."""

    def __init__(self, cstacks, *, batch, num_actions, embed_dim, obs_dim, support_size, max_num_simulations,
                 activation=_lib.ACT_ELU, repr_minmax=1, dyn_minmax=1, discount=0.99, prng_mode=_lib.PRNG_LEGACY,
                 device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("muax_b200 needs a CUDA device: the search has no CPU fallback")
        self.lib = _lib.load()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise ValueError("device must be a CUDA device")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        cfg = _lib.Config()
        cfg.batch, cfg.num_actions, cfg.embed_dim, cfg.obs_dim = batch, num_actions, embed_dim, obs_dim
        cfg.support_size, cfg.max_num_simulations = support_size, max_num_simulations
        cfg.activation, cfg.repr_minmax, cfg.dyn_minmax = activation, repr_minmax, dyn_minmax
        cfg.prng_mode, cfg.device, cfg.discount = prng_mode, self.device.index, discount
        for name, st in cstacks.items():
            setattr(cfg, name, st)
        self.cfg = cfg
        self.batch, self.A, self.E, self.obs_dim = batch, num_actions, embed_dim, obs_dim
        self.max_num_simulations = max_num_simulations
        self._h = ctypes.c_void_p()
        _lib.check(self.lib.mz_create(ctypes.byref(self._h), ctypes.byref(cfg)), "mz_create")
        self._last_num_sim = 0
        self._args_cache = {}

    def _on_device(self):
        """Context that makes `self.device` current — a no-op object when it already is (the common case)."""
        if _current_device() == self.device.index:
            return _NULL_CTX
        return torch.cuda.device(self.device)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.mz_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def set_weights(self, blob):
        """blob: float32 numpy array or CUDA tensor laid out by nn.pack_stacks."""
        with torch.cuda.device(self.device):
            stream = _raw_stream(self.device.index)
            if isinstance(blob, torch.Tensor) and blob.is_cuda:
                blob = blob.to(torch.float32).contiguous()
                rc = self.lib.mz_set_weights(self._h, _ptr(blob), blob.numel(), 1, stream)
            else:
                blob = np.ascontiguousarray(np.asarray(blob, dtype=np.float32))
                rc = self.lib.mz_set_weights(self._h, ctypes.c_void_p(blob.ctypes.data), blob.size, 0, stream)
        _lib.check(rc, "mz_set_weights")

    def set_peer_outputs(self, byte_deltas):
        """The next searches also store their outputs at (output pointer + delta) for every delta: the same slots of
        the peer GPUs' gather buffers (muax_b200/sharded.py).  Empty list = off.  Warp engine only."""
        deltas = [int(d) for d in byte_deltas]
        arr = (ctypes.c_int64 * max(len(deltas), 1))(*deltas)
        _lib.check(self.lib.mz_set_peer_outputs(self._h, len(deltas), arr), "mz_set_peer_outputs")

    def set_peer_flags(self, flags, rank=0, step=0):
        """Completion flags of the peer-store exchange (include/mzsearch.h: mz_set_peer_flags): `flags` = int32 CUDA
        tensor of W words inside this rank's symmetric gather buffer, or None to switch the signalling off."""
        ptr = None if flags is None else flags.data_ptr()
        _lib.check(self.lib.mz_set_peer_flags(self._h, ptr, int(rank), int(step)), "mz_set_peer_flags")

    def peer_wait(self, flags, world, step):
        """Enqueues, on the current stream, the wait for every rank's rows of act `step` (mz_peer_wait)."""
        _lib.check(self.lib.mz_peer_wait(self._h, flags.data_ptr(), int(world), int(step),
                                         _raw_stream(self.device.index)), "mz_peer_wait")

    # ------------------------------------------------------------------ arguments
    def make_args(self, rng_key, *, policy=_lib.POLICY_MUZERO, qtransform=None, num_simulations=5, temperature=1.0,
                  max_depth=None, dirichlet_fraction=0.25, dirichlet_alpha=0.3, pb_c_init=1.25, pb_c_base=19652,
                  max_num_considered_actions=16, gumbel_scale=1.0, value_scale=0.1, maxvisit_init=50.0,
                  global_batch=None, batch_offset=0, engine=_lib.ENGINE_AUTO, want_tree=False,
                  precision=_lib.PRECISION_FP32, num_decision_actions=0):
        # a 0.45 ms search makes the host call path matter: the argument struct is built once per distinct keyword
        # set and only the key words change from act to act
        ck = (policy, qtransform, num_simulations, temperature, max_depth, dirichlet_fraction, dirichlet_alpha,
              pb_c_init, pb_c_base, max_num_considered_actions, gumbel_scale, value_scale, maxvisit_init, global_batch,
              batch_offset, engine, want_tree, precision, num_decision_actions)
        cached = self._args_cache.get(ck)
        if cached is not None:
            cached.key0, cached.key1 = key_words(rng_key)
            return cached
        a = _lib.SearchArgs()
        self.lib.mz_default_args(ctypes.byref(a))
        a.policy = policy
        if qtransform is None:
            qtransform = _lib.QT_PARENT_AND_SIBLINGS if policy == _lib.POLICY_MUZERO else _lib.QT_COMPLETED_BY_MIX_VALUE
        a.qtransform = qtransform
        a.num_simulations = int(num_simulations)
        a.max_depth = 0 if max_depth is None else int(max_depth)
        a.max_considered = int(max_num_considered_actions)
        a.global_batch = self.batch if global_batch is None else int(global_batch)
        a.batch_offset = int(batch_offset)
        a.engine = engine
        a.flags = _lib.FLAG_WANT_TREE if want_tree else 0
        a.precision = {"fp32": _lib.PRECISION_FP32, "bf16": _lib.PRECISION_BF16}.get(precision, precision)
        a.num_decision_actions = int(num_decision_actions)
        a.temperature, a.dirichlet_fraction, a.dirichlet_alpha = temperature, dirichlet_fraction, dirichlet_alpha
        a.pb_c_init, a.pb_c_base, a.gumbel_scale = pb_c_init, pb_c_base, gumbel_scale
        a.value_scale, a.maxvisit_init = value_scale, maxvisit_init
        a.key0, a.key1 = key_words(rng_key)
        if len(self._args_cache) < 64:
            self._args_cache[ck] = a
        return a

    # ------------------------------------------------------------------ device-resident search
    def search(self, rng_key, obs=None, root=None, invalid_actions=None, noise=None, out=None, **kw):
        """One `act` worth of search on device tensors.  Returns (action i32[B], action_weights f32[B,A],
        root_value f32[B]) as CUDA tensors; nothing synchronises.  `out` = optional preallocated (action, weights,
        root_value) tensors the kernels write into (e.g. views of one all-gather send buffer)."""
        B, A, E = self.batch, self.A, self.E
        args = self.make_args(rng_key, **kw)
        with self._on_device():
            f32 = torch.float32
            obs_t = _dev(obs, self.device, f32, (B, self.obs_dim), "obs")
            r_logits = r_value = r_emb = None
            if obs_t is None:
                if root is None:
                    raise ValueError("give obs or root")
                logits, value, emb = root
                r_emb = _dev(emb, self.device, f32, (B, E), "root embedding")
                r_logits = _dev(logits, self.device, f32, (B, A), "root prior_logits")
                r_value = _dev(value, self.device, f32, (B,), "root value")
            inv_t = _dev(invalid_actions, self.device, torch.uint8, (B, A), "invalid_actions")
            noise_t = _dev(noise, self.device, f32, (B, A), "noise")
            if out is not None:
                action, weights, root_value = out
                for t, dt, shp in ((action, torch.int32, (B,)), (weights, f32, (B, A)), (root_value, f32, (B,))):
                    if t.dtype != dt or tuple(t.shape) != shp or not t.is_contiguous() or t.device != self.device:
                        raise ValueError("out: expected contiguous (int32[B], float32[B,A], float32[B]) on the engine's device")
            else:  # one allocation, three views
                flat = torch.empty(B * (A + 2), dtype=f32, device=self.device)
                weights, root_value = flat[:B * A].view(B, A), flat[B * A:B * A + B]
                action = flat[B * A + B:].view(torch.int32)
            stream = _raw_stream(self.device.index)
            rc = self.lib.mz_search(self._h, _ptr(obs_t), _ptr(r_logits), _ptr(r_value), _ptr(r_emb), _ptr(inv_t),
                                    _ptr(noise_t), ctypes.byref(args), _ptr(action), _ptr(weights), _ptr(root_value),
                                    stream)
        _lib.check(rc, "mz_search")
        self._last_num_sim = args.num_simulations
        return action, weights, root_value

    # ------------------------------------------------------------------ host-buffer search (MuZero.act's contract)
    def search_host(self, rng_key, obs, invalid_actions=None, noise=None, **kw):
        """numpy in / numpy out through mz_search_host: H2D, search, D2H and one stream sync inside the call."""
        B, A = self.batch, self.A
        args = self.make_args(rng_key, **kw)
        obs = np.ascontiguousarray(np.asarray(obs, dtype=np.float32))
        if obs.shape != (B, self.obs_dim):
            raise ValueError(f"obs: expected shape {(B, self.obs_dim)}, got {obs.shape}")
        inv = None if invalid_actions is None else np.ascontiguousarray(np.asarray(invalid_actions, dtype=np.uint8))
        nz = None if noise is None else np.ascontiguousarray(np.asarray(noise, dtype=np.float32))
        for name, arr in (("invalid_actions", inv), ("noise", nz)):
            if arr is not None and arr.shape != (B, A):
                raise ValueError(f"{name}: expected shape {(B, A)}, got {arr.shape}")
        action = np.empty(B, np.int32)
        weights = np.empty((B, A), np.float32)
        root_value = np.empty(B, np.float32)
        # the act is ~0.3 ms: `.ctypes.data` (1 us each) and torch.cuda.current_stream() (14 us) were 7 % of it
        vp = lambda a: None if a is None else a.__array_interface__["data"][0]  # noqa: E731
        with self._on_device():
            stream = _raw_stream(self.device.index)
            rc = self.lib.mz_search_host(self._h, vp(obs), vp(inv), vp(nz), ctypes.byref(args), vp(action),
                                         vp(weights), vp(root_value), stream)
        _lib.check(rc, "mz_search_host")
        self._last_num_sim = args.num_simulations
        return action, weights, root_value

    # ------------------------------------------------------------------ recurrent_fn on its own
    def recurrent(self, action, embedding, precision="fp32"):
        """`MuZero._recurrent_inference` (muax/model.py:265-282) for a batch: (action i32[B], embedding f32[B,E]) ->
        (reward[B], value[B], prior_logits[B,A], next_embedding[B,E]) as CUDA tensors.  precision "fp32" = the kernel
        the fp32 engines are checked against, "bf16" = the tcgen05 kernel of the throughput mode."""
        B, A, E = self.batch, self.A, self.E
        f32 = torch.float32
        with self._on_device():
            act = _dev(action, self.device, torch.int32, (B,), "action")
            emb = _dev(embedding, self.device, f32, (B, E), "embedding")
            reward = torch.empty(B, dtype=f32, device=self.device)
            value = torch.empty(B, dtype=f32, device=self.device)
            logits = torch.empty(B, A, dtype=f32, device=self.device)
            nxt = torch.empty(B, E, dtype=f32, device=self.device)
            stream = _raw_stream(self.device.index)
            prec = {"fp32": _lib.PRECISION_FP32, "bf16": _lib.PRECISION_BF16}.get(precision, precision)
            rc = self.lib.mz_recurrent(self._h, _ptr(act), _ptr(emb), prec, _ptr(reward), _ptr(value), _ptr(logits),
                                       _ptr(nxt), stream)
        _lib.check(rc, "mz_recurrent")
        return reward, value, logits, nxt

    # ------------------------------------------------------------------ callback mode (arbitrary recurrent_fn)
    def search_with_callback(self, rng_key, root, recurrent_fn, invalid_actions=None, noise=None, **kw):
        """mctx-style search where `recurrent_fn(action i32[B], embedding f32[B,E]) -> (reward[B], discount[B] or
        None, prior_logits[B,A], value[B], next_embedding[B,E])` is any torch callable (muax/model.py:265-282)."""
        B, A, E = self.batch, self.A, self.E
        args = self.make_args(rng_key, **kw)
        f32 = torch.float32
        with torch.cuda.device(self.device):
            logits, value, emb = root
            r_logits = _dev(logits, self.device, f32, (B, A), "root prior_logits")
            r_value = _dev(value, self.device, f32, (B,), "root value")
            r_emb = _dev(emb, self.device, f32, (B, E), "root embedding")
            inv_t = _dev(invalid_actions, self.device, torch.uint8, (B, A), "invalid_actions")
            noise_t = _dev(noise, self.device, f32, (B, A), "noise")
            stream = _raw_stream(self.device.index)
            _lib.check(self.lib.mz_begin(self._h, _ptr(r_logits), _ptr(r_value), _ptr(r_emb), _ptr(inv_t),
                                         _ptr(noise_t), ctypes.byref(args), stream), "mz_begin")
            act_buf = torch.empty(B, dtype=torch.int32, device=self.device)
            emb_buf = torch.empty(B, E, dtype=f32, device=self.device)
            for sim in range(args.num_simulations):
                _lib.check(self.lib.mz_select(self._h, sim, _ptr(act_buf), _ptr(emb_buf), stream), "mz_select")
                reward, discount, p_logits, val, nxt = recurrent_fn(act_buf, emb_buf)
                reward = _dev(reward, self.device, f32, (B,), "reward")
                discount = _dev(discount, self.device, f32, (B,), "discount")
                p_logits = _dev(p_logits, self.device, f32, (B, A), "prior_logits")
                val = _dev(val, self.device, f32, (B,), "value")
                nxt = _dev(nxt, self.device, f32, (B, E), "next embedding")
                _lib.check(self.lib.mz_expand_backup(self._h, sim, _ptr(reward), _ptr(discount), _ptr(p_logits),
                                                     _ptr(val), _ptr(nxt), stream), "mz_expand_backup")
            action = torch.empty(B, dtype=torch.int32, device=self.device)
            weights = torch.empty(B, A, dtype=f32, device=self.device)
            _lib.check(self.lib.mz_finish(self._h, _ptr(action), _ptr(weights), stream), "mz_finish")
        self._last_num_sim = args.num_simulations
        return action, weights, r_value

    # ------------------------------------------------------------------ introspection
    def tree(self):
        """Zero-copy CUDA tensor views of the last search tree (mctx.Tree field names).  The search must have run
        with `want_tree=True` (muax never reads `PolicyOutput.search_tree`, so `act` does not keep it by default)."""
        v = _lib.TreeView()
        _lib.check(self.lib.mz_get_tree(self._h, ctypes.byref(v)), "mz_get_tree")
        dims = {"B": v.batch, "N": v.num_nodes, "A": v.num_actions, "E": v.embed_dim}
        out = {}
        with torch.cuda.device(self.device):
            for name, (ts, shape) in _TREE_FIELDS.items():
                if name == "sim_depth":
                    if self._last_num_sim == 0:
                        continue
                    shp = (v.batch, self._last_num_sim)
                else:
                    shp = tuple(dims[c] for c in shape)
                out[name] = torch.as_tensor(_DevArray(getattr(v, name), shp, ts), device=self.device)
        return out

    def launch_count(self):
        n = ctypes.c_int64()
        _lib.check(self.lib.mz_launch_count(self._h, ctypes.byref(n)), "mz_launch_count")
        return n.value

    def last_kernel_ms(self):
        ms = ctypes.c_float()
        _lib.check(self.lib.mz_last_kernel_ms(self._h, ctypes.byref(ms)), "mz_last_kernel_ms")
        return ms.value


def math_probe(kind, x):
    """Evaluates include/mz_math.h on the device (bit-parity tests).  kind: expf/logf/expm1f/inv_scaling/gumbel."""
    kinds = {"expf": 0, "logf": 1, "expm1f": 2, "inv_scaling": 3, "gumbel": 4, "fast_div": 5}
    lib = _lib.load()
    x = x.contiguous()
    n = x.numel() // 2 if kind == "fast_div" else x.numel()   # fast_div: x is [n, 2] = (a, b) pairs
    y = torch.empty(n, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        stream = _raw_stream(x.device.index if x.device.index is not None else torch.cuda.current_device())
        _lib.check(lib.mz_math_probe(kinds[kind], _ptr(x), _ptr(y), n, stream), "mz_math_probe")
    return y if kind == "fast_div" else y.reshape(x.shape)
