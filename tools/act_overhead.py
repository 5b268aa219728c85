"""Host-side cost of one `MuZero.act` on a tiny search (16 trees x 1 simulation: the kernel is ~10 us), layer by layer:
the reference-shaped call, SearchEngine.search_host, and the bare C-ABI call with a prebuilt argument struct."""
import cProfile
import ctypes
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import muax_b200  # noqa: E402
from muax_b200 import nn  # noqa: E402


def main(reps=3000):
    net = nn.create_muzero_network(nn.Representation, nn.Prediction, nn.Dynamic, 8, 2, 21)
    model = muax_b200.MuZero(net, policy="muzero", discount=0.99, support_size=10, device="cuda:0")
    model.init(muax_b200.random.PRNGKey(0), np.zeros((1, 4), np.float32))
    B = int(os.environ.get("B", 16))
    NS = int(os.environ.get("NS", 1))
    obs = np.random.default_rng(0).standard_normal((B, 4)).astype(np.float32)
    key = muax_b200.random.PRNGKey(7)
    kw = dict(with_pi=True, with_value=True, obs_from_batch=True, num_simulations=NS)
    for _ in range(50):
        model.act(key, obs, **kw)
    t0 = time.perf_counter()
    for _ in range(reps):
        model.act(key, obs, **kw)
    t_act = (time.perf_counter() - t0) / reps * 1e6
    eng = model._engine_for(B, NS)
    skw = model._policy._search_kwargs(dict(num_simulations=NS, temperature=1.0, max_depth=None,
                                            qtransform=muax_b200._lib.QT_PARENT_AND_SIBLINGS, dirichlet_fraction=0.25,
                                            dirichlet_alpha=0.3, pb_c_init=1.25, pb_c_base=19652))
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.search_host(key, obs, **skw)
    t_host = (time.perf_counter() - t0) / reps * 1e6
    args = eng.make_args(key, **skw)
    action = np.empty(B, np.int32); weights = np.empty((B, 2), np.float32); value = np.empty(B, np.float32)
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)  # noqa: E731
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    po, pa, pw, pv, pargs = vp(obs), vp(action), vp(weights), vp(value), ctypes.byref(args)
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.lib.mz_search_host(eng._h, po, None, None, pargs, pa, pw, pv, stream)
    t_c = (time.perf_counter() - t0) / reps * 1e6
    print(f"B={B} NS={NS}: act {t_act:.1f} us | search_host {t_host:.1f} us | bare C call {t_c:.1f} us")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(reps):
        model.act(key, obs, **kw)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)


if __name__ == "__main__":
    main()
