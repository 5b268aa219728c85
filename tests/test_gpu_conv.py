"""Conv networks: the reference's all-conv MuZero (muax/nn.py:313-395) as torch modules
(`muax_b200.conv`) — root Representation in torch, Prediction / Dynamic as torch callables INSIDE the simulation loop
through the library's callback mode (tree kernels native)."""
import numpy as np
import pytest

from helpers import check_tree_invariants

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def test_resnet_muzero_searches_through_the_callback_mode():
    import muax_b200
    from muax_b200.conv import create_resnet_muzero_network
    torch.manual_seed(0)
    A, S, B, NS = 6, 10, 12, 16
    net = create_resnet_muzero_network(A, 2 * S + 1, input_channels=8, height=32, width=32)
    model = muax_b200.MuZero(net, discount=0.99, support_size=S, device="cuda")
    frames = np.random.default_rng(0).integers(0, 256, (B, 32, 32, 4), dtype=np.uint8)
    key = np.array([0, 7], np.uint32)
    a, w, v = model.act(key, frames, with_pi=True, with_value=True, obs_from_batch=True, num_simulations=NS,
                        want_tree=True)
    assert a.shape == (B,) and w.shape == (B, A) and v.shape == (B,)
    assert np.allclose(w.sum(-1), 1.0, atol=1e-5) and ((0 <= a) & (a < A)).all()
    # root value = the torch root inference (raw network value, muax/model.py:243)
    root = model._root_inference(None, None, torch.from_numpy(frames).cuda())
    assert np.allclose(v, root.value.cpu().numpy(), atol=1e-6)
    eng = next(iter(model._engines.values()))[0]
    tree = {k: t.cpu().numpy() for k, t in eng.tree().items()}
    check_tree_invariants(tree, NS)
    h, ww, c = net.representation_fn.shape
    assert tree["embeddings"].shape == (B, NS + 1, h * ww * c)
    # the stored next states are min-max normalised per channel (muax/nn.py:48-56)
    emb = tree["embeddings"][:, 1:].reshape(B, NS, h * ww, c)
    assert emb.min() >= 0.0 and emb.max() <= 1.0 + 1e-6
    # same key, same frames -> same search
    a2, w2, _ = model.act(key, frames, with_pi=True, with_value=True, obs_from_batch=True, num_simulations=NS)
    assert np.array_equal(a, a2) and np.allclose(w, w2)
