"""Actor-side data path (muax_b200/actor.py) against the reference's own tracer / trajectory / replay buffer:
tests/golden/tracer_pins.npz was produced by EXECUTING /root/reference/muax/episode_tracer.py and
replay_buffer.py (tests/golden/make_tracer_pins.py), so these are reference-pinned, not restatement-pinned."""
import os

import numpy as np
import pytest

from muax_b200.actor import BatchedPNStep, CartPoleVec, TrajectoryStore, Transitions, VectorActor

PINS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tracer_pins.npz")
FIELDS = ("obs", "a", "r", "done", "Rn", "v", "pi", "w")


@pytest.fixture(scope="module")
def pins():
    return np.load(PINS)


def _run_batched(p):
    T, B = p["a"].shape
    tracer = BatchedPNStep(B, int(p["n"]), float(p["gamma"]), float(p["alpha"]))
    got = {f: [[] for _ in range(B)] for f in FIELDS + ("t",)}
    episodes, pending = [], [[] for _ in range(B)]
    for t in range(T):
        env, tr = tracer.add(p["obs"][t], p["a"][t], p["r"][t], p["done"][t], p["v"][t], p["pi"][t])
        for i, e in enumerate(env):
            for f in FIELDS:
                got[f][e].append(getattr(tr, f)[i])
            got["t"][e].append(t)
            pending[e].append(i)
        for e in range(B):
            rows, pending[e] = pending[e], []
            if rows:
                episodes_rows = tr[np.array(rows)]
                got.setdefault("_blocks", {}).setdefault(e, []).append(episodes_rows)
            if p["done"][t, e]:
                blocks = got["_blocks"].pop(e, [])
                if blocks:
                    episodes.append((t, e, Transitions(*(np.concatenate(c) for c in zip(*blocks)))))
    return got, episodes


def test_batched_pnstep_reproduces_reference_pops(pins):
    """Every transition the reference's `while tracer: tracer.pop()` loop emits — same step, same order, same values
    (Rn, w in float64, bit for bit) — for 6 environments with ragged episode lengths (one-step episodes, episodes
    shorter than n, an environment that never terminates)."""
    got, _ = _run_batched(pins)
    B = pins["a"].shape[1]
    for b in range(B):
        assert np.array_equal(np.asarray(got["t"][b]), pins[f"pop_t_{b}"]), b
        for f in FIELDS:
            want = pins[f"pop_{f}_{b}"]
            have = np.asarray(got[f][b]).reshape(want.shape)
            if f in ("Rn", "w"):
                assert np.array_equal(have, want), (b, f, np.abs(have - want).max())
            else:
                assert np.array_equal(have.astype(want.dtype), want), (b, f)


def test_trajectory_store_reproduces_reference_sampling(pins):
    """Same episodes, same ring capacity, same seed -> the reference's `buffer.sample(num_trajectory=7,
    sample_per_trajectory=2, k_steps=3)` batches, three calls in a row (the reference restores its saved `random`
    state before each episode draw, replay_buffer.py:225-229)."""
    _, episodes = _run_batched(pins)
    k_steps = int(pins["k_steps"])
    store = TrajectoryStore(int(pins["buffer_capacity"]), random_seed=int(pins["buffer_seed"]))
    kept = []
    for t, e, ep in sorted(episodes, key=lambda x: (x[0], x[1])):
        if len(ep) >= k_steps:
            store.add(ep)
            kept.append((t, e))
    assert np.array_equal(np.array(kept), pins["kept"])
    for call in range(3):
        s = store.sample(batch_size=None, num_trajectory=7, sample_per_trajectory=2, k_steps=k_steps)
        for f in FIELDS:
            want = pins[f"sample{call}_{f}"]
            have = np.asarray(getattr(s, f))
            if f == "pi":  # the reference keeps a leading batch axis of 1 on pi: [B, k, 1, A]
                want = want.reshape(have.shape)
            assert have.shape == want.shape, (call, f, have.shape, want.shape)
            assert np.array_equal(have.astype(want.dtype), want), (call, f)


def test_store_edge_cases():
    store = TrajectoryStore(3, random_seed=0)
    assert not store and len(store) == 0
    with pytest.raises(ValueError):
        store.sample(batch_size=None, num_trajectory=None)
    ep = Transitions(obs=np.zeros((2, 4)), a=np.zeros(2, int), r=np.ones(2), done=np.zeros(2, bool), Rn=np.ones(2),
                     v=np.zeros(2), pi=np.ones((2, 2)) / 2, w=np.ones(2))
    store.add(ep)
    assert store.sample(batch_size=4, k_steps=5) is None  # every episode shorter than the window: nothing to draw
    for _ in range(5):
        store.add(ep)
    assert len(store) == 3  # ring


class _RandomModel:
    """Stand-in for MuZero.act on CPU (the GPU test drives the real one)."""

    def __init__(self, seed=0):
        self.rng = np.random.default_rng(seed)

    def act(self, key, obs, with_pi, with_value, obs_from_batch, num_simulations, temperature):
        B = obs.shape[0]
        pi = self.rng.dirichlet(np.ones(2), B).astype(np.float32)
        return (self.rng.random(B) < pi[:, 1]).astype(np.int32), pi, self.rng.standard_normal(B).astype(np.float32)


def test_vector_actor_loop_fills_the_store():
    env = CartPoleVec(32, seed=1)
    store = TrajectoryStore(500, random_seed=1)
    actor = VectorActor(_RandomModel(), env, store, n=5, gamma=0.99, k_steps=4, num_simulations=8)
    for t in range(120):
        actor.step(np.array([0, t], np.uint32))
    assert actor.env_steps == 120 * 32 and actor.episodes > 32 and len(store) > 0
    lengths = [len(ep) for ep in store]
    assert min(lengths) >= 4 and max(lengths) <= 120
    for ep in store:
        assert ep.done[-1] and ep.obs.shape == (len(ep), 4) and np.all(ep.r == 1.0)
        # flushed tail: truncated returns, no bootstrap -> Rn of the last transition is its own reward
        assert ep.Rn[-1] == 1.0
    batch = store.sample(batch_size=16, k_steps=4)
    assert batch.obs.shape == (16, 4, 4) and batch.pi.shape == (16, 4, 2)


def test_device_pnstep_matches_reference_pins_on_cpu_torch(pins):
    """The torch tracer of the device-resident loop (fixed-shape outputs + mask), run here on torch's CPU device:
    every valid slot, in slot order, must be the reference's pop — exact fields exactly, Rn / w to 1e-12 (the device
    version sums the discounted rewards in a different association order than np.sum)."""
    torch = pytest.importorskip("torch")
    from muax_b200.actor_device import DevicePNStep
    p = pins
    T, B = p["a"].shape
    tracer = DevicePNStep(B, int(p["n"]), float(p["gamma"]), float(p["alpha"]), device="cpu")
    got = {f: [[] for _ in range(B)] for f in FIELDS + ("t",)}
    for t in range(T):
        tr, mask = tracer.add(torch.from_numpy(p["obs"][t]), torch.from_numpy(p["a"][t]), torch.from_numpy(p["r"][t]),
                              torch.from_numpy(p["done"][t]), torch.from_numpy(p["v"][t]), torch.from_numpy(p["pi"][t]))
        mask = mask.numpy()
        for b in range(B):
            for k in np.nonzero(mask[b])[0]:
                for f in FIELDS:
                    got[f][b].append(getattr(tr, f)[b, k].numpy())
                got["t"][b].append(t)
    for b in range(B):
        assert np.array_equal(np.asarray(got["t"][b]), p[f"pop_t_{b}"]), b
        for f in FIELDS:
            want = p[f"pop_{f}_{b}"]
            have = np.asarray(got[f][b]).reshape(want.shape)
            if f in ("Rn", "w"):
                np.testing.assert_allclose(have, want, rtol=1e-12, atol=1e-12, err_msg=f"{b} {f}")
            else:
                assert np.array_equal(have.astype(want.dtype), want), (b, f)
