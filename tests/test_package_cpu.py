"""CPU-side checks of the product package: the C ABI library builds for sm_100a, loads, exports every symbol
include/mzsearch.h declares (no compute calls without a GPU), and the host logic (params layout, PRNG keys,
argument plumbing, loud failure without CUDA) behaves."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from muax_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "mzsearch.h")).read()
    declared = set(re.findall(r"\b(mz_[a-z_]+)\s*\(", header))
    assert declared, "no declarations found"
    from muax_b200 import _lib
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_layouts_match_header(lib):
    from muax_b200 import _lib
    a = _lib.SearchArgs()
    lib.mz_default_args(ctypes.byref(a))
    assert (a.policy, a.qtransform, a.num_simulations, a.max_considered) == (0, 0, 5, 16)
    assert abs(a.pb_c_base - 19652.0) < 1e-3 and abs(a.dirichlet_alpha - 0.3) < 1e-6 and abs(a.temperature - 1) < 1e-6
    assert ctypes.sizeof(_lib.Stack) == 4 + 4 + 8 * 4 * 2 + 8 * 8 * 2  # n_layers (+pad), 2 int32[8], 2 int64[8]


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from muax_b200 import _lib
    cfg = _lib.Config()
    cfg.batch, cfg.num_actions, cfg.embed_dim, cfg.support_size = 1, 2, 8, 10
    h = ctypes.c_void_p()
    assert lib.mz_create(ctypes.byref(h), ctypes.byref(cfg)) != 0
    assert lib.mz_last_error()
    import muax_b200
    from muax_b200 import nn
    model = muax_b200.MuZero(nn.create_muzero_network(nn.Representation, nn.Prediction, nn.Dynamic, 8, 2, 21))
    model.init(muax_b200.random.PRNGKey(0), np.zeros((1, 4), np.float32))
    with pytest.raises(RuntimeError):
        model.act(muax_b200.random.PRNGKey(0), np.zeros(4, np.float32))


def test_host_prng_matches_restatement():
    from muax_b200 import random as mr
    from oracle import threefry as tf
    for seed in (0, 1, 42, 2**40 + 3):
        k = mr.PRNGKey(seed)
        assert np.array_equal(k, tf.PRNGKey(seed))
        for num in (2, 3, 7):
            for mode in (0, 1):
                assert np.array_equal(mr.split(k, num, mode), tf.split(k, num, mode))
    assert mr.key_words(7) == (0, 7)
    with pytest.raises(ValueError):
        mr.key_words(np.zeros(3, np.uint32))


def test_params_layout_is_haiku_shaped_and_packs_in_order():
    import muax_b200
    from muax_b200 import nn
    net = nn.create_muzero_network(nn.Representation, nn.Prediction, nn.Dynamic, 8, 2, 21)
    model = muax_b200.MuZero(net)
    p = model.init(muax_b200.random.PRNGKey(1), np.zeros((1, 4), np.float32))
    assert list(p.representation) == ["representation/linear"]
    assert [p.prediction[k]["w"].shape for k in p.prediction] == [(8, 16), (16, 21), (8, 16), (16, 2)]
    assert [p.dynamic[k]["w"].shape for k in p.dynamic] == [(10, 16), (16, 8), (10, 16), (16, 21)]
    assert all(np.all(v["b"] == 0) for v in p.dynamic.values())
    w = p.prediction["prediction/linear"]["w"]
    assert abs(w.std() - 0.88 / np.sqrt(8)) < 0.1 and np.abs(w).max() <= 2 / np.sqrt(8) + 1e-6
    blob, st = model._spec.pack(p)
    assert st["repr"].n_layers == 1 and st["dyn_r"].n_layers == 2
    assert st["repr"].w_off[0] == 0 and st["repr"].b_off[0] == 32 and st["pred_v"].w_off[0] == 40
    # weight matrices start on 16-byte boundaries (zero padding between layers), nothing else is added
    payload = sum(v["w"].size + v["b"].size for g in p for v in g.values())
    assert payload <= blob.size <= payload + 3 * 9
    for name in ("repr", "pred_v", "pred_pi", "dyn_ns", "dyn_r"):
        assert all(st[name].w_off[l] % 4 == 0 for l in range(st[name].n_layers))
    # stack order / content: dyn_ns is the first two dynamic linears, dyn_r the last two (muax/nn.py:97-104)
    stacks = model._spec.stacks(p)
    assert np.array_equal(stacks["dyn_r"][1][0], p.dynamic["dynamic/linear_3"]["w"])


def test_released_constructor_shape_and_custom_architecture(tmp_path):
    import muax_b200
    from muax_b200 import nn

    class Pred(nn.Prediction):
        hidden = (64, 64, 16)

    class Dyn(nn.Dynamic):
        hidden = (64, 64, 16)
        normalize = False

    class Rep(nn.Representation):
        normalize = False

    model = muax_b200.MuZero(nn._init_representation_func(Rep, 10), nn._init_prediction_func(Pred, 4, 41),
                             nn._init_dynamic_func(Dyn, 10, 4, 41), policy="gumbel", discount=0.999, support_size=20)
    p = model.init(muax_b200.random.PRNGKey(0), np.zeros((1, 8), np.float32))
    assert len(p.prediction) == 8 and len(p.dynamic) == 8
    assert model._spec.repr_minmax == 0 and model._spec.dyn_minmax == 0
    path = str(tmp_path / "params")
    model.save(path)
    other = muax_b200.MuZero(nn._init_representation_func(Rep, 10), nn._init_prediction_func(Pred, 4, 41),
                             nn._init_dynamic_func(Dyn, 10, 4, 41), policy="gumbel", support_size=20)
    other.init(muax_b200.random.PRNGKey(9), np.zeros((1, 8), np.float32))
    other.load(path)
    for g0, g1 in zip(p, other.params):
        for k in g0:
            assert np.array_equal(g0[k]["w"], g1[k]["w"])
    with pytest.raises(NotImplementedError):  # only the reference's default loss is re-hosted (muax_b200/learner.py)
        muax_b200.MuZero(nn._init_representation_func(Rep, 10), nn._init_prediction_func(Pred, 4, 41),
                         nn._init_dynamic_func(Dyn, 10, 4, 41), loss_fn=lambda *a: 0.0).update(None)


def test_support_transform_helpers_roundtrip():
    import torch
    from muax_b200 import utils
    x = torch.tensor([-7.3, -0.2, 0.0, 0.4, 3.9, 55.0])
    probs = utils.scalar_to_support(x, 10)
    assert probs.shape == (6, 21) and torch.allclose(probs.sum(-1), torch.ones(6))
    back = utils.support_to_scalar(probs, 10)
    assert torch.allclose(back, x, atol=2e-3, rtol=1e-3)


def test_checkpoint_interop_and_load_without_init(tmp_path):
    """muax/model.py:203-212 layout (pickled .npy of {'params', 'optimizer_state'}) both ways, haiku's `~` module
    paths, and the reference workflow `model = MuZero(...); model.load(path)` without `init()`."""
    import muax_b200
    from muax_b200 import checkpoint, nn
    net = nn.create_muzero_network(nn.Representation, nn.Prediction, nn.Dynamic, 8, 2, 21)
    model = muax_b200.MuZero(net, policy="muzero", support_size=10)
    params = model.init(muax_b200.random.PRNGKey(3), np.zeros((1, 4), np.float32))
    ref_file = checkpoint.save_reference_checkpoint(tmp_path / "model_params", params)
    assert ref_file.endswith(".npy")
    raw = open(ref_file, "rb").read()
    assert b"muax.nn" in raw and b"MZNetworkParams" in raw          # the class the reference's unpickler looks up
    assert b"representation/~/linear" in raw                        # haiku's spelling on disk
    assert "muax" not in __import__("sys").modules                  # ... and no fake module left behind
    loaded, opt = checkpoint.load_reference_checkpoint(tmp_path / "model_params")
    assert opt is None and isinstance(loaded, nn.MZNetworkParams)
    for a, b in zip(params, loaded):
        assert a.keys() == b.keys()
        for mod in a:
            assert np.array_equal(a[mod]["w"], b[mod]["w"]) and np.array_equal(a[mod]["b"], b[mod]["b"])
    fresh = muax_b200.MuZero(net, policy="muzero", support_size=10)
    fresh.load(str(tmp_path / "model_params"))                       # .npy found by name, no init()
    assert fresh._spec is not None and fresh._spec.obs_dim == 4
    blob_a, _ = model._spec.pack(params)
    blob_b, _ = fresh._spec.pack(fresh.params)
    assert np.array_equal(blob_a, blob_b)
    model.save(tmp_path / "own")                                     # this package's .npz
    other = muax_b200.MuZero(net, policy="muzero", support_size=10)
    other.load(tmp_path / "own")
    assert other._spec is not None and np.array_equal(other._spec.pack(other.params)[0], blob_a)
    tilde = nn.MZNetworkParams(*[{k.replace("/", "/~/", 1): v for k, v in t.items()} for t in params])
    third = muax_b200.MuZero(net, policy="muzero", support_size=10)
    third.params = tilde                                             # reference-spelled params straight in
    assert np.array_equal(third._spec.pack(third.params)[0], blob_a)


def test_sass_size_map_tool_reads_the_built_library(lib):
    """tools/sass_size_map.py (the map behind DESIGN.md §3.1b's instruction-fetch finding) parses the library built
    here: a small kernel is enough — every instruction lands on a source line of its own translation unit."""
    import shutil
    import subprocess
    import sys
    if not (shutil.which("cuobjdump") and shutil.which("nvdisasm")):
        pytest.skip("CUDA binary utilities not on PATH")
    from muax_b200 import _lib
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_size_map.py"), "recurrent_tc_bias_kernel",
                          _lib.LIB_PATH, "5"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-400:]
    head = out.stdout.splitlines()[0]
    m = re.search(r"recurrent_tc_bias_kernel\S*: (\d+) bytes of SASS \((\d+) instructions\)", head)
    assert m and int(m.group(1)) == 16 * int(m.group(2)) > 0
    assert "mz_recurrent_tc.cu:" in out.stdout
