"""`muax_b200.train.fit` / `test` (the vectorised mirror of muax/train.py + muax/test.py).  CPU: the loop's host logic
with a stand-in agent; GPU: a short real run — search-driven acting, n-step tracer, replay buffer, learner — whose
loss must fall and whose updated parameters must reach the search engine."""
import numpy as np
import pytest

from muax_b200.actor import CartPoleVec
from muax_b200.train import _temperature_fn, fit, test as run_test


class _StubAgent:
    """act/update/save with the MuZero signatures; the 'policy' pushes the cart towards the pole's fall."""
    params = {}

    def __init__(self):
        self.updates, self.temps, self.saved = 0, [], []

    def act(self, key, obs, with_pi=False, with_value=False, obs_from_batch=False, num_simulations=5, temperature=1.0):
        assert obs_from_batch
        self.temps.append(temperature)
        a = (obs[:, 2] > 0).astype(np.int32)
        if not with_pi:
            return a
        pi = np.eye(2, dtype=np.float32)[a]
        return a, pi, np.zeros(len(a), np.float32)

    def update(self, batch):
        assert batch.obs.shape[1] == 5 and batch.a.shape == batch.r.shape
        self.updates += 1
        return {"loss": 1.0 / self.updates}

    def save(self, path):
        self.saved.append(path)


def test_temperature_schedule_matches_reference():  # muax/train.py:14-23
    assert [_temperature_fn(100, s) for s in (0, 49, 50, 74, 75, 99)] == [1.0, 1.0, 0.5, 0.5, 0.25, 0.25]


def test_fit_loop_host_logic(tmp_path):
    agent = _StubAgent()
    path, hist = fit(agent, CartPoleVec(16, seed=1), test_env=CartPoleVec(8, seed=2), k_steps=5, buffer_warm_up=8,
                     steps_per_iteration=40, max_iterations=4, max_training_steps=12, num_update_per_episode=4,
                     num_trajectory=4, sample_per_trajectory=2, test_interval=2, model_save_path=str(tmp_path),
                     num_simulations=3)
    assert agent.updates == 12 and len(hist) == 3  # stops when training_step reaches max_training_steps
    assert hist[0]["test_G"] is not None and hist[1]["test_G"] is None and hist[2]["test_G"] is not None
    assert path is not None and agent.saved and path == agent.saved[-1]
    assert 0.0 in agent.temps and 1.0 in agent.temps and 0.5 in agent.temps  # greedy tests + the schedule
    assert hist[-1]["loss"] < hist[0]["loss"]
    # the stub balances for a while: the greedy test return is a real episode length
    assert 8 <= run_test(agent, CartPoleVec(8, seed=3), np.array([0, 1], np.uint32), 3) <= 500


@pytest.mark.gpu
def test_fit_trains_on_the_gpu(tmp_path):
    import muax_b200
    from muax_b200 import nn
    model = muax_b200.MuZero(nn.create_muzero_network(nn.Representation, nn.Prediction, nn.Dynamic, 8, 2, 21), discount=0.997, support_size=10)  # stock nets
    model.init(muax_b200.random.PRNGKey(0), np.zeros((1, 4), np.float32))
    before = {k: v["w"].copy() for k, v in model.params.prediction.items()}
    env, test_env = CartPoleVec(256, seed=0), CartPoleVec(64, seed=1)
    path, hist = fit(model, env, test_env=test_env, n_steps=10, k_steps=5, buffer_warm_up=64, steps_per_iteration=40,
                     max_iterations=3, max_training_steps=10 ** 6, num_update_per_episode=25, num_trajectory=16,
                     sample_per_trajectory=4, test_interval=1, num_simulations=16, model_save_path=str(tmp_path))
    assert len(hist) == 3 and all(np.isfinite(h["loss"]) for h in hist)
    assert hist[-1]["loss"] < hist[0]["loss"], [h["loss"] for h in hist]
    assert hist[-1]["training_step"] == 75 and hist[-1]["env_steps"] >= 3 * 40 * 256
    assert all(h["test_G"] is not None and h["test_G"] >= 8 for h in hist)
    after = model.params.prediction
    assert any(not np.array_equal(before[k], after[k]["w"]) for k in before)  # the learner's step reached the agent
    assert path is not None
    # the same loop with the acting phase resident on the GPU (search kernel + CUDA graph per step)
    from muax_b200.actor_device import CartPoleVecTorch, DeviceActor
    _, hist_dev = fit(model, CartPoleVecTorch(256, seed=3), n_steps=10, k_steps=5, buffer_warm_up=32,
                      steps_per_iteration=40, max_iterations=2, max_training_steps=10 ** 6, num_update_per_episode=5,
                      num_trajectory=16, sample_per_trajectory=4, num_simulations=16, actor_cls=DeviceActor)
    assert len(hist_dev) == 2 and all(np.isfinite(h["loss"]) for h in hist_dev) and hist_dev[-1]["episodes"] > 0
    other = muax_b200.MuZero(nn.create_muzero_network(nn.Representation, nn.Prediction, nn.Dynamic, 8, 2, 21), discount=0.997, support_size=10)
    other.init(muax_b200.random.PRNGKey(5), np.zeros((1, 4), np.float32))
    other.load(path)
    obs = env.reset()
    key = muax_b200.random.PRNGKey(3)
    a0 = model.act(key, obs, obs_from_batch=True, num_simulations=16, temperature=0.0)
    # the saved checkpoint is the best-in-test model; loading the CURRENT parameters must reproduce the same search
    model.save(str(tmp_path / "now"))
    other.load(str(tmp_path / "now"))
    a1 = other.act(key, obs, obs_from_batch=True, num_simulations=16, temperature=0.0)
    assert np.array_equal(np.asarray(a0), np.asarray(a1))
