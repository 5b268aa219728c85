"""Pins the actor-side data path (n-step tracer, trajectory, replay buffer) against the REFERENCE'S OWN code.

`muax/episode_tracer.py` and `muax/replay_buffer.py` are plain Python + NumPy apart from `import jax` (pytree
registration, `tree_transpose`).  jax is not installable here, so this script executes the reference sources in
place (nothing is copied into this repo) with a ten-line stand-in for the three `jax.tree_util` calls they make and
NumPy for `jax.numpy`, drives them with seeded random episodes exactly like `muax.fit` does (muax/train.py:160-173:
`tracer.add(...)`, `while tracer: trajectory.add(tracer.pop())`, `buffer.add(trajectory, w.mean())`,
`buffer.sample(...)`), and stores inputs and outputs:

    muax/episode_tracer.py:114-249   NStep / PNStep
    muax/replay_buffer.py:39-119     Trajectory
    muax/replay_buffer.py:154-254    TrajectoryReplayBuffer

Run in the build container only (needs /root/reference):  python tests/golden/make_tracer_pins.py
"""
import ast
import dataclasses
import os
import sys
import types

import numpy as np

REF = "/root/reference/muax"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tracer_pins.npz")


def _jax_stand_in():
    jax = types.ModuleType("jax")
    tu = types.ModuleType("jax.tree_util")
    tu.register_pytree_node = lambda *a, **k: None
    tu.tree_structure = lambda x: x

    def tree_transpose(outer_treedef, inner_treedef, pytree_to_transpose):
        cls = type(inner_treedef)
        items = list(pytree_to_transpose)
        return cls(*[[getattr(t, f.name) for t in items] for f in dataclasses.fields(cls)])

    tu.tree_transpose = tree_transpose
    jax.tree_util = tu
    jax.numpy = np
    return jax, tu


def load_reference():
    jax, tu = _jax_stand_in()
    sys.modules.update({"jax": jax, "jax.tree_util": tu, "jax.numpy": np})
    utils = types.ModuleType("muax.utils")
    tree = ast.parse(open(os.path.join(REF, "utils.py")).read())
    picked = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "sliceable_deque"]
    ns = {"deque": __import__("collections").deque, "islice": __import__("itertools").islice}
    exec(compile(ast.Module(body=picked, type_ignores=[]), "utils.py", "exec"), ns)
    utils.sliceable_deque = ns["sliceable_deque"]
    utils.n_step_bootstrapped_returns = None
    pkg = types.ModuleType("muax")
    sys.modules.update({"muax": pkg, "muax.utils": utils})
    mods = {}
    for name in ("episode_tracer", "replay_buffer"):
        m = types.ModuleType("muax." + name)
        sys.modules["muax." + name] = m
        exec(compile(open(os.path.join(REF, name + ".py")).read(), name + ".py", "exec"), m.__dict__)
        mods[name] = m
    return mods["episode_tracer"], mods["replay_buffer"]


def main():
    et, rb = load_reference()
    rng = np.random.default_rng(11)
    B, T, A, n, gamma, alpha, k_steps = 6, 90, 3, 4, 0.97, 0.5, 3
    obs = rng.standard_normal((T, B, 4)).astype(np.float32)
    a = rng.integers(0, A, (T, B))
    r = rng.standard_normal((T, B))
    v = rng.standard_normal((T, B)).astype(np.float32)
    pi = rng.dirichlet(np.ones(A), (T, B)).astype(np.float32)
    done = rng.random((T, B)) < 0.08
    done[7, 0] = done[8, 0] = True          # back-to-back one-step episodes
    done[2, 1] = True                       # episode shorter than n
    done[:, 5] = False                      # one environment never ends inside the window
    done[T - 1, :5] = True
    # reference: one PNStep + Trajectory per environment, episodes added to the buffer per environment in time order
    pops = {f: [[] for _ in range(B)] for f in ("obs", "a", "r", "done", "Rn", "v", "pi", "w", "t")}
    episodes = []  # (end step, env, Trajectory)
    for b in range(B):
        tracer = et.PNStep(n, gamma, alpha)
        traj = rb.Trajectory()
        for t in range(T):
            tracer.add(obs[t, b], int(a[t, b]), float(r[t, b]), bool(done[t, b]), v=float(v[t, b]), pi=pi[t, b])
            while tracer:
                tr = tracer.pop()
                traj.add(tr)
                for f in ("obs", "a", "r", "done", "Rn", "v", "pi", "w"):
                    pops[f][b].append(getattr(tr, f))
                pops["t"][b].append(t)
            if done[t, b]:
                traj.finalize()
                episodes.append((t, b, traj))
                tracer.reset()
                traj = rb.Trajectory()
    episodes.sort(key=lambda e: (e[0], e[1]))  # the order a vectorised actor finishes them in
    buf = rb.TrajectoryReplayBuffer(8, random_seed=5)  # smaller than the number of episodes: the ring wraps
    kept = []
    for t, b, traj in episodes:
        if len(traj) >= k_steps:
            buf.add(traj, traj.batched_transitions.w.mean())
            kept.append((t, b))
    samples = []
    for call in range(3):
        s = buf.sample(num_trajectory=7, sample_per_trajectory=2, k_steps=k_steps)
        samples.append({f.name: np.asarray(getattr(s, f.name)) for f in dataclasses.fields(s)})
    out = dict(obs=obs, a=a, r=r, v=v, pi=pi, done=done, n=n, gamma=gamma, alpha=alpha, k_steps=k_steps,
               kept=np.array(kept), buffer_seed=5, buffer_capacity=8)
    for b in range(B):
        for f in pops:
            out[f"pop_{f}_{b}"] = np.asarray(pops[f][b])
    for i, s in enumerate(samples):
        for f, val in s.items():
            out[f"sample{i}_{f}"] = val
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(episodes), "episodes,", len(kept), "kept")


if __name__ == "__main__":
    main()
