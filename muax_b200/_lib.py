"""ctypes binding of libmzsearch.so (C ABI: include/mzsearch.h).  No CPU fallback: if the library cannot be
built/loaded, or no CUDA device is usable, the calls raise."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MZ_LIB_PATH") or os.path.join(_HERE, "libmzsearch.so")  # MZ_LIB_PATH: A/B builds

MAX_LAYERS = 8
MAX_ACTIONS = 32
POLICY_MUZERO, POLICY_GUMBEL = 0, 1
QT_PARENT_AND_SIBLINGS, QT_COMPLETED_BY_MIX_VALUE = 0, 1
PRNG_LEGACY, PRNG_PARTITIONABLE = 0, 1
ACT_ELU, ACT_RELU = 0, 1
ENGINE_AUTO, ENGINE_STEPWISE, ENGINE_FUSED = 0, 1, 2
ENGINE_RESIDENT = 7
ENGINE_FUSED_WARP = 8
ENGINE_TREEWARP = 9
PRECISION_FP32, PRECISION_BF16 = 0, 1
FLAG_WANT_TREE = 1
OK, ERR_RUNTIME, ERR_INVALID_ARGUMENT = 0, 1, 2

EXPORTS = ("mz_last_error", "mz_default_args", "mz_create", "mz_destroy", "mz_set_weights", "mz_search",
           "mz_search_host", "mz_set_peer_outputs", "mz_set_peer_flags", "mz_peer_wait", "mz_recurrent", "mz_begin", "mz_select", "mz_expand_backup", "mz_finish", "mz_get_tree",
           "mz_launch_count", "mz_last_kernel_ms", "mz_math_probe")


class Stack(ctypes.Structure):
    _fields_ = [("n_layers", ctypes.c_int32), ("in_dim", ctypes.c_int32 * MAX_LAYERS),
                ("out_dim", ctypes.c_int32 * MAX_LAYERS), ("w_off", ctypes.c_int64 * MAX_LAYERS),
                ("b_off", ctypes.c_int64 * MAX_LAYERS)]


class Config(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_int32) for n in (
        "batch", "num_actions", "embed_dim", "obs_dim", "support_size", "max_num_simulations", "activation",
        "repr_minmax", "dyn_minmax", "prng_mode", "device")] + [("discount", ctypes.c_float)]
        + [(n, Stack) for n in ("repr", "pred_v", "pred_pi", "dyn_ns", "dyn_r")])


class SearchArgs(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_int32) for n in (
        "policy", "qtransform", "num_simulations", "max_depth", "max_considered", "global_batch", "batch_offset",
        "engine")] + [(n, ctypes.c_float) for n in (
            "temperature", "dirichlet_fraction", "dirichlet_alpha", "pb_c_init", "pb_c_base", "gumbel_scale",
            "value_scale", "maxvisit_init")] + [("key0", ctypes.c_uint32), ("key1", ctypes.c_uint32),
                                                ("flags", ctypes.c_uint32), ("precision", ctypes.c_int32),
                                               ("num_decision_actions", ctypes.c_int32)])


class TreeView(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_int32) for n in ("batch", "num_nodes", "num_actions", "embed_dim")]
                + [(n, ctypes.c_void_p) for n in (
                    "node_visits", "parents", "action_from_parent", "children_index", "children_visits",
                    "raw_values", "node_values", "children_prior_logits", "children_values", "children_rewards",
                    "children_discounts", "embeddings", "root_noise", "sim_depth")])


_lib = None


def load(build_if_missing=True):
    """Returns the loaded library; builds it in-tree with nvcc first if it is missing or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing and not os.environ.get("MZ_LIB_PATH"):
        from .csrc import build as _build
        _build.build()
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -m muax_b200.csrc.build` (needs nvcc)")
    lib = ctypes.CDLL(LIB_PATH)
    lib.mz_last_error.restype = ctypes.c_char_p
    for name in EXPORTS[2:]:
        getattr(lib, name).restype = ctypes.c_int
    lib.mz_default_args.restype = None
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    lib.mz_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(Config)]
    lib.mz_destroy.argtypes = [vp]
    lib.mz_default_args.argtypes = [ctypes.POINTER(SearchArgs)]
    lib.mz_set_weights.argtypes = [vp, vp, ctypes.c_size_t, ctypes.c_int, vp]
    lib.mz_search.argtypes = [vp, vp, vp, vp, vp, vp, vp, ctypes.POINTER(SearchArgs), vp, vp, vp, vp]
    lib.mz_search_host.argtypes = [vp, vp, vp, vp, ctypes.POINTER(SearchArgs), vp, vp, vp, vp]
    lib.mz_set_peer_outputs.argtypes = [vp, ctypes.c_int32, ctypes.POINTER(ctypes.c_int64)]
    lib.mz_set_peer_flags.argtypes = [vp, vp, i32, i32]
    lib.mz_peer_wait.argtypes = [vp, vp, i32, i32, vp]
    lib.mz_recurrent.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, vp]
    lib.mz_begin.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.POINTER(SearchArgs), vp]
    lib.mz_select.argtypes = [vp, i32, vp, vp, vp]
    lib.mz_expand_backup.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp]
    lib.mz_finish.argtypes = [vp, vp, vp, vp]
    lib.mz_get_tree.argtypes = [vp, ctypes.POINTER(TreeView)]
    lib.mz_launch_count.argtypes = [vp, ctypes.POINTER(i64)]
    lib.mz_last_kernel_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    lib.mz_math_probe.argtypes = [i32, vp, vp, i64, vp]
    _lib = lib
    return lib


def check(rc, what="libmzsearch call"):
    """MZ_ERR_INVALID_ARGUMENT -> ValueError (what the reference raises for bad arguments), anything else non-zero ->
    RuntimeError; the text comes from mz_last_error()."""
    if rc != OK:
        msg = load().mz_last_error().decode("utf-8", "replace")
        raise (ValueError if rc == ERR_INVALID_ARGUMENT else RuntimeError)(f"{what}: {msg}")
