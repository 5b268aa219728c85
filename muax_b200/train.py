"""`muax.fit` / `muax.test` over a vector environment (SURVEY.md §8(f) ranks 1 + 2 put together).

The reference drives ONE gym environment (`muax/train.py:26-260`, `muax/test.py:5-50`): every `model.act` is a B = 1
search.  The B200 search only pays off on a batch, so this loop keeps the reference's structure and defaults —
buffer warm-up, act -> env.step -> PNStep -> Trajectory -> TrajectoryReplayBuffer, `num_update_per_episode`
optimisation steps between acting phases, the `_temperature_fn` schedule, periodic greedy tests, best-model
checkpoints — while `env` is a vector environment (`CartPoleVec`-style: `batch`, `reset()`, `step(actions) ->
(obs, reward, done)` with auto-reset) and one "episode" of the reference becomes one acting phase of
`steps_per_iteration` vector steps.  The tracer / buffer semantics are the reference's (pinned in
tests/test_actor_cpu.py), the learner is `MuZero.update` (muax_b200/learner.py).
"""
import os
import time

import numpy as np

from . import random as mz_random
from .actor import TrajectoryStore, VectorActor


def _temperature_fn(max_training_steps, training_steps):  # muax/train.py:14-23
    if training_steps < 0.5 * max_training_steps:
        return 1.0
    if training_steps < 0.75 * max_training_steps:
        return 0.5
    return 0.25


def test(model, env, key, num_simulations, num_test_episodes=None, max_steps=None):
    """muax/test.py:5-50 on a vector environment: every environment plays ONE episode with temperature 0; returns
    the mean total reward over the first `num_test_episodes` environments (default: all of them)."""
    obs = env.reset()
    B = env.batch
    alive = np.ones(B, bool)
    total = np.zeros(B)
    on_device = hasattr(obs, "is_cuda")  # a torch vector environment (CartPoleVecTorch) steps on tensors
    for _ in range(int(max_steps or getattr(env, "MAX_STEPS", 1000))):
        key, sub = mz_random.split(key)
        a = model.act(sub, obs, obs_from_batch=True, num_simulations=num_simulations, temperature=0.0)
        if on_device:
            import torch
            a = torch.as_tensor(np.asarray(a), device=obs.device)
        else:
            a = np.asarray(a)
        obs, r, done = env.step(a)
        r, done = (x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x) for x in (r, done))
        total += np.where(alive, r, 0.0)
        alive &= ~done.astype(bool)
        if not alive.any():
            break
    n = B if num_test_episodes is None else min(int(num_test_episodes), B)
    return float(total[:n].mean())


def fit(model, env, test_env=None, n_steps=10, gamma=0.997, alpha=0.5, buffer=None, buffer_capacity=500,
        max_iterations=1000, steps_per_iteration=None, test_interval=10, num_test_episodes=None,
        max_training_steps=10000, num_simulations=50, k_steps=10, buffer_warm_up=128, num_trajectory=32,
        sample_per_trajectory=10, model_save_path=None, save_name="model_params", random_seed=42,
        temperature_fn=_temperature_fn, num_update_per_episode=50, log=None, actor_cls=VectorActor):
    """Fits `model` on the vector environment `env`.  Keyword names and defaults follow muax/train.py:26-52
    (`tracer=PNStep(n_steps, gamma, alpha)` is spelled out because the batched tracer is built per environment
    batch).  Returns `(model_path, history)`: the path of the best parameters seen in testing (None when nothing was
    saved) and a list of per-iteration dicts {iteration, training_step, loss, episodes, env_steps, test_G, seconds}.
    `actor_cls=muax_b200.actor_device.DeviceActor` with a torch vector environment (`CartPoleVecTorch`) keeps the
    acting phase on the GPU (search kernel + one CUDA graph per step)."""
    if env is None:
        raise ValueError("You must provide a vector `env` (gym is not a dependency of this package).")
    buffer = buffer if buffer is not None else TrajectoryStore(buffer_capacity, random_seed=random_seed)
    key = mz_random.PRNGKey(random_seed)
    key, test_key, sub = mz_random.split(key, 3)
    if model.params is None:
        model.init(sub, np.zeros((1, env.obs_dim), np.float32))
    actor = actor_cls(model, env, buffer, n=n_steps, gamma=gamma, alpha=alpha, k_steps=k_steps,
                      num_simulations=num_simulations)
    steps_per_iteration = int(steps_per_iteration or max(1, getattr(env, "MAX_STEPS", 500) // 10))
    model_dir = model_save_path
    training_step, best_test_G, model_path, history = 0, -float("inf"), None, []

    def act_phase(steps):
        nonlocal key
        actor.temperature = float(temperature_fn(max_training_steps=max_training_steps, training_steps=training_step))
        for _ in range(steps):
            key, k = mz_random.split(key)
            actor.step(k)

    warm_up_phases = 0
    while len(buffer) < buffer_warm_up:  # buffer warm up (train.py:148-173)
        act_phase(steps_per_iteration)
        warm_up_phases += 1
        if warm_up_phases > 10000:
            raise RuntimeError(f"buffer warm-up never reached {buffer_warm_up} episodes: no episode of at least "
                               f"k_steps = {k_steps} transitions ended in {warm_up_phases * steps_per_iteration} steps")
    for it in range(max_iterations):
        t0 = time.perf_counter()
        act_phase(steps_per_iteration)
        train_loss = 0.0
        for _ in range(num_update_per_episode):  # train.py:206-214
            batch = buffer.sample(num_trajectory=num_trajectory, sample_per_trajectory=sample_per_trajectory,
                                  k_steps=k_steps)
            if batch is None:  # every drawn episode was shorter than k_steps + 1 (replay_buffer.py:85): draw again later
                continue
            train_loss += float(model.update(batch)["loss"])
            training_step += 1
        rec = {"iteration": it, "training_step": training_step, "loss": train_loss / max(num_update_per_episode, 1),
               "episodes": actor.episodes, "env_steps": actor.env_steps, "test_G": None}
        if test_env is not None and it % test_interval == 0:  # train.py:230-243
            rec["test_G"] = test(model, test_env, test_key, num_simulations, num_test_episodes)
            if rec["test_G"] >= best_test_G:
                best_test_G = rec["test_G"]
                if model_dir is not None:
                    folder = os.path.join(model_dir, f"epoch_{it:04d}_test_G_{rec['test_G']:.8f}")
                    os.makedirs(folder, exist_ok=True)
                    model_path = os.path.join(folder, save_name)
                    model.save(model_path)
        rec["seconds"] = time.perf_counter() - t0
        history.append(rec)
        if log is not None:
            log(rec)
        if training_step >= max_training_steps:
            break
    return model_path, history
