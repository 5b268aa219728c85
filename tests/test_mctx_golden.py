"""Parity against REAL mctx output, when somebody has produced it: tests/golden/mctx_*.npz are written by
tools/dump_mctx_golden.py in an environment where `import jax, mctx` works (this image has neither, so the whole
module is skipped until such a file is committed).  Bar (BASELINE.json north_star): integer tree state exact,
values / logits within 1e-5."""
import glob
import os

import numpy as np
import pytest

from helpers import GOLDEN_DIR, INT_FIELDS, OUT_FIELDS, load_golden

FILES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "mctx_*.npz")))
pytestmark = pytest.mark.skipif(not FILES, reason="no tests/golden/mctx_*.npz (run tools/dump_mctx_golden.py where "
                                                  "jax + mctx are installed)")


def _compare(got, want):
    for f in OUT_FIELDS:
        if f not in want or f not in got:
            continue
        g, w = np.asarray(got[f]), np.asarray(want[f])
        assert g.shape == w.shape, (f, g.shape, w.shape)
        if f in INT_FIELDS:
            assert np.array_equal(g, w), f"{f}: {np.sum(g != w)} of {g.size} entries differ from mctx"
        else:
            np.testing.assert_allclose(g, w, rtol=0, atol=1e-5, err_msg=f)


@pytest.mark.parametrize("name", FILES or ["none"])
def test_cpu_restatements_match_mctx(name, c_oracle):
    from oracle import np_mctx
    nets, inp, cfg, want = load_golden(name)
    got = c_oracle.search(nets, inp["key"], obs=inp["obs"], invalid=inp["invalid"], noise=inp["noise"], **cfg)
    _compare(got, want)
    got = np_mctx.act(nets, inp["key"], obs=inp["obs"], invalid=inp["invalid"], noise=inp["noise"], **cfg)
    _compare(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("name", FILES or ["none"])
def test_cuda_search_matches_mctx(name):
    import torch
    from test_gpu_parity import _collect, _engine, _search_kwargs
    nets, inp, cfg, want = load_golden(name)
    eng = _engine(nets, inp["obs"].shape[0], cfg, cfg["num_simulations"])
    out = eng.search(inp["key"], obs=torch.from_numpy(inp["obs"]).cuda(), invalid_actions=inp["invalid"],
                     noise=inp["noise"], **_search_kwargs(cfg))
    _compare(_collect(eng, *out), want)
