"""TEST INFRASTRUCTURE — ctypes binding of oracle/mz_oracle.c (scalar C restatement, pthreads over trees).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
this; the product never does.  See the header of mz_oracle.c for the reference file:line it follows
and for the "parity unpinned" statement.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmzoracle.so")
_MAX_LAYERS = 8

STACK_NAMES = ("repr", "pred_v", "pred_pi", "dyn_ns", "dyn_r")


class _Stack(ctypes.Structure):
    _fields_ = [
        ("n_layers", ctypes.c_int32),
        ("in_dim", ctypes.c_int32 * _MAX_LAYERS),
        ("out_dim", ctypes.c_int32 * _MAX_LAYERS),
        ("w_off", ctypes.c_int64 * _MAX_LAYERS),
        ("b_off", ctypes.c_int64 * _MAX_LAYERS),
    ]


class _Config(ctypes.Structure):
    _fields_ = (
        [(n, ctypes.c_int32) for n in (
            "batch", "num_actions", "embed_dim", "obs_dim", "support_size", "policy", "qtransform", "prng_mode",
            "num_simulations", "max_depth", "max_considered", "activation", "repr_minmax", "dyn_minmax",
            "global_batch", "batch_offset", "noise_injected")]
        + [(n, ctypes.c_float) for n in (
            "temperature", "dirichlet_fraction", "dirichlet_alpha", "pb_c_init", "pb_c_base", "gumbel_scale",
            "discount", "value_scale", "maxvisit_init")]
        + [(n, _Stack) for n in STACK_NAMES]
    )


_I32P = ctypes.POINTER(ctypes.c_int32)
_F32P = ctypes.POINTER(ctypes.c_float)


class _TreeOut(ctypes.Structure):
    _fields_ = (
        [(n, _I32P) for n in ("node_visits", "parents", "action_from_parent", "children_index", "children_visits")]
        + [(n, _F32P) for n in ("raw_values", "node_values", "children_prior_logits", "children_values",
                                "children_rewards", "children_discounts", "embeddings", "root_noise")]
        + [("sim_depth", _I32P)]
    )


def _host_signature():
    """CPU feature flags of this host + the compile flags: the library is built -march=native, so a copy that travelled
    from another machine (the build container -> the GPU box) must be rebuilt, not executed."""
    flags = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                flags = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    import hashlib
    return hashlib.sha1((flags + "|" + build_flags()).encode()).hexdigest()


def build_flags():
    """The CFLAGS line of oracle/Makefile (reported beside the CPU baseline)."""
    for line in open(os.path.join(_HERE, "Makefile")):
        if line.startswith("CFLAGS"):
            return line.split("=", 1)[1].strip()
    return "?"


def build(force=False):
    """Compile the C restatement (gcc, -ffp-contract=off).  Building the checker is not using it."""
    src = os.path.join(_HERE, "mz_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "mz_math.h")
    mk = os.path.join(_HERE, "Makefile")
    sig_path = _LIB_PATH + ".host"
    sig = _host_signature()
    fresh = (os.path.exists(_LIB_PATH) and os.path.exists(sig_path) and open(sig_path).read() == sig
             and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(p) for p in (src, hdr, mk)))
    if fresh and not force:
        return _LIB_PATH
    import fcntl
    os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
    with open(_LIB_PATH + ".lock", "w") as lock:  # torchrun ranks / xdist workers: one builds, the others wait
        fcntl.flock(lock, fcntl.LOCK_EX)
        if force or not (os.path.exists(sig_path) and open(sig_path).read() == sig and os.path.exists(_LIB_PATH)
                         and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(p) for p in (src, hdr, mk))):
            subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
            with open(sig_path, "w") as f:
                f.write(sig)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.mzo_search.restype = ctypes.c_int
        _lib.mzo_max_threads.restype = ctypes.c_int
    return _lib


def _ptr(a, ty):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ty))


def pack_nets(nets):
    """nets: {stack name: [(W[in,out], b[out]), ...]} -> (float32 blob, {name: _Stack})."""
    chunks, off, stacks = [], 0, {}
    for name in STACK_NAMES:
        st = _Stack()
        layers = nets.get(name, [])
        st.n_layers = len(layers)
        assert len(layers) <= _MAX_LAYERS
        for l, (w, b) in enumerate(layers):
            w = np.ascontiguousarray(w, dtype=np.float32)
            b = np.ascontiguousarray(b, dtype=np.float32)
            st.in_dim[l], st.out_dim[l] = w.shape
            st.w_off[l] = off
            off += w.size
            st.b_off[l] = off
            off += b.size
            chunks += [w.ravel(), b.ravel()]
        stacks[name] = st
    blob = np.concatenate(chunks) if chunks else np.zeros(1, np.float32)
    return np.ascontiguousarray(blob, dtype=np.float32), stacks


DEFAULTS = dict(
    policy=0, qtransform=0, prng_mode=0, num_simulations=5, max_depth=0, max_considered=16, activation=0,
    repr_minmax=1, dyn_minmax=1, temperature=1.0, dirichlet_fraction=0.25, dirichlet_alpha=0.3, pb_c_init=1.25,
    pb_c_base=19652.0, gumbel_scale=1.0, discount=0.99, value_scale=0.1, maxvisit_init=50.0, support_size=10)


def search(nets, key, obs=None, root=None, invalid=None, noise=None, want_tree=True, nthreads=0,
           global_batch=None, batch_offset=0, **kw):
    """Run the scalar C restatement of one `act`.  Returns a dict of numpy arrays (mctx field names)."""
    cfgd = dict(DEFAULTS)
    cfgd.update(kw)
    blob, stacks = pack_nets(nets)
    if obs is not None:
        obs = np.ascontiguousarray(obs, dtype=np.float32)
        B = obs.shape[0]
        obs_dim = obs.shape[1]
    else:
        root_logits, root_value, root_emb = (np.ascontiguousarray(x, dtype=np.float32) for x in root)
        B = root_logits.shape[0]
        obs_dim = 0
    A = nets["pred_pi"][-1][0].shape[1]
    E = nets["pred_pi"][0][0].shape[0]
    NS = int(cfgd["num_simulations"])
    N = NS + 1
    c = _Config()
    c.batch, c.num_actions, c.embed_dim, c.obs_dim = B, A, E, obs_dim
    c.global_batch = B if global_batch is None else int(global_batch)
    c.batch_offset = int(batch_offset)
    c.noise_injected = 0 if noise is None else 1
    for k, v in cfgd.items():
        if k == "max_depth":
            v = 0 if v is None else v
        setattr(c, k, v)
    for name in STACK_NAMES:
        setattr(c, name, stacks[name])
    if invalid is not None:
        invalid = np.ascontiguousarray(invalid, dtype=np.uint8)
    if noise is not None:
        noise = np.ascontiguousarray(noise, dtype=np.float32)
    out = dict(
        action=np.zeros(B, np.int32), action_weights=np.zeros((B, A), np.float32), root_value=np.zeros(B, np.float32))
    tree = _TreeOut()
    if want_tree:
        shapes = dict(
            node_visits=((B, N), np.int32), parents=((B, N), np.int32), action_from_parent=((B, N), np.int32),
            children_index=((B, N, A), np.int32), children_visits=((B, N, A), np.int32),
            raw_values=((B, N), np.float32), node_values=((B, N), np.float32),
            children_prior_logits=((B, N, A), np.float32), children_values=((B, N, A), np.float32),
            children_rewards=((B, N, A), np.float32), children_discounts=((B, N, A), np.float32),
            embeddings=((B, N, E), np.float32), root_noise=((B, A), np.float32), sim_depth=((B, max(NS, 1)), np.int32))
        for name, (shape, dt) in shapes.items():
            out[name] = np.zeros(shape, dt)
            setattr(tree, name, _ptr(out[name], ctypes.c_int32 if dt == np.int32 else ctypes.c_float))
    key = np.asarray(key, dtype=np.uint32)
    rc = lib().mzo_search(
        ctypes.byref(c), _ptr(blob, ctypes.c_float), _ptr(obs, ctypes.c_float),
        _ptr(root_logits if obs is None else None, ctypes.c_float),
        _ptr(root_value if obs is None else None, ctypes.c_float),
        _ptr(root_emb if obs is None else None, ctypes.c_float),
        _ptr(invalid, ctypes.c_uint8), _ptr(noise, ctypes.c_float),
        ctypes.c_uint32(int(key[0])), ctypes.c_uint32(int(key[1])),
        _ptr(out["action"], ctypes.c_int32), _ptr(out["action_weights"], ctypes.c_float),
        _ptr(out["root_value"], ctypes.c_float), ctypes.byref(tree) if want_tree else None, int(nthreads))
    if rc != 0:
        raise RuntimeError(f"mzo_search failed with code {rc}")
    if want_tree:
        out["sim_depth"] = out["sim_depth"][:, :NS]
    return out


def max_threads():
    return int(lib().mzo_max_threads())


def _unary(name, x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    getattr(lib(), name)(_ptr(x, ctypes.c_float), _ptr(y, ctypes.c_float), ctypes.c_int64(x.size))
    return y


def expf(x):
    return _unary("mzo_expf_v", x)


def logf(x):
    return _unary("mzo_logf_v", x)


def expm1f(x):
    return _unary("mzo_expm1f_v", x)


def inv_scaling(x):
    return _unary("mzo_inv_scaling_v", x)


def fmaf(a, b, c):
    a, b, c = np.broadcast_arrays(np.asarray(a, np.float32), np.asarray(b, np.float32), np.asarray(c, np.float32))
    a, b, c = (np.ascontiguousarray(t) for t in (a, b, c))
    y = np.empty_like(a)
    lib().mzo_fmaf_v(_ptr(a, ctypes.c_float), _ptr(b, ctypes.c_float), _ptr(c, ctypes.c_float),
                     _ptr(y, ctypes.c_float), ctypes.c_int64(a.size))
    return y


def threefry(k0, k1, c0, c1):
    c0 = np.ascontiguousarray(c0, dtype=np.uint32)
    c1 = np.ascontiguousarray(c1, dtype=np.uint32)
    o0, o1 = np.empty_like(c0), np.empty_like(c1)
    lib().mzo_threefry_v(ctypes.c_uint32(int(k0)), ctypes.c_uint32(int(k1)), _ptr(c0, ctypes.c_uint32),
                         _ptr(c1, ctypes.c_uint32), _ptr(o0, ctypes.c_uint32), _ptr(o1, ctypes.c_uint32),
                         ctypes.c_int64(c0.size))
    return o0, o1


def dirichlet(key, row0, rows, A, alpha):
    out = np.empty((rows, A), np.float32)
    lib().mzo_dirichlet(ctypes.c_uint32(int(key[0])), ctypes.c_uint32(int(key[1])), ctypes.c_int64(row0),
                        ctypes.c_int64(rows), ctypes.c_int(A), ctypes.c_float(alpha), _ptr(out, ctypes.c_float))
    return out


def considered_visits(m, n):
    seq = np.zeros(max(n, 1), np.int32)
    lib().mzo_considered_visits(ctypes.c_int(m), ctypes.c_int(n), _ptr(seq, ctypes.c_int32))
    return seq[:n]


def support_from_probs(probs, S):
    probs = np.ascontiguousarray(probs, dtype=np.float32)
    out = np.empty(probs.shape[0], np.float32)
    lib().mzo_support_from_probs(_ptr(probs, ctypes.c_float), ctypes.c_int(probs.shape[0]), ctypes.c_int(S),
                                 _ptr(out, ctypes.c_float))
    return out


def min_max_normalize(s):
    s = np.array(s, dtype=np.float32, order="C", copy=True)
    lib().mzo_min_max_normalize(_ptr(s, ctypes.c_float), ctypes.c_int(s.shape[0]), ctypes.c_int(s.shape[1]))
    return s


def pb_c(visits, pb_c_init=1.25, pb_c_base=19652.0):
    visits = np.ascontiguousarray(visits, dtype=np.int32)
    out = np.empty(visits.shape[0], np.float32)
    lib().mzo_pb_c(_ptr(visits, ctypes.c_int32), ctypes.c_int(visits.shape[0]), ctypes.c_float(pb_c_init),
                   ctypes.c_float(pb_c_base), _ptr(out, ctypes.c_float))
    return out


def check_branch_free(lo, hi):
    """Counts inputs in the uint32 bit-pattern range [lo, hi] where the branch-free mz_expf/mz_expm1f differ
    from the early-return reference forms."""
    lib().mzo_check_branch_free.restype = ctypes.c_int64
    first = ctypes.c_uint32(0)
    n = lib().mzo_check_branch_free(ctypes.c_uint32(lo), ctypes.c_uint32(hi), ctypes.byref(first))
    return int(n), int(first.value)
