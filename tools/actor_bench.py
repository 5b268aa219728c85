#!/usr/bin/env python
"""End-to-end acting throughput of the vectorised loop (muax_b200/actor.py): B CartPole environments, one
`MuZero.act(obs_from_batch=True)` per step on the B200, n-step tracer and trajectory store on the host.
Prints env-steps/s and the split between the search (incl. H2D/D2H) and the host-side NumPy work."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--num-simulations", type=int, default=50)
    args = ap.parse_args()
    import muax_b200
    from muax_b200 import nn
    from muax_b200.actor import CartPoleVec, TrajectoryStore, VectorActor
    net = nn.create_muzero_network(nn.Representation, nn.Prediction, nn.Dynamic, 8, 2, 21)
    model = muax_b200.MuZero(net, policy="muzero", discount=0.997, support_size=10)
    model.init(muax_b200.random.PRNGKey(0), np.zeros((1, 4), np.float32))
    env = CartPoleVec(args.batch, seed=0)
    store = TrajectoryStore(100000, random_seed=0)
    actor = VectorActor(model, env, store, n=10, gamma=0.997, k_steps=5, num_simulations=args.num_simulations)
    act_s = [0.0]
    inner = model.act

    def timed_act(*a, **k):
        t0 = time.perf_counter()
        out = inner(*a, **k)
        act_s[0] += time.perf_counter() - t0
        return out

    model.act = timed_act
    for t in range(5):
        actor.step(muax_b200.random.PRNGKey(t))
    act_s[0] = 0.0
    t0 = time.perf_counter()
    for t in range(args.steps):
        actor.step(muax_b200.random.PRNGKey(1000 + t))
    wall = time.perf_counter() - t0
    steps = args.steps * args.batch
    print(json.dumps({"metric": "env_steps_per_sec", "value": steps / wall, "batch": args.batch,
                      "num_simulations": args.num_simulations, "steps": args.steps, "ms_per_step": wall / args.steps * 1e3,
                      "act_ms_per_step": act_s[0] / args.steps * 1e3,
                      "host_ms_per_step": (wall - act_s[0]) / args.steps * 1e3,
                      "sims_per_sec_end_to_end": steps * args.num_simulations / wall,
                      "episodes": actor.episodes, "stored_episodes": len(store)}))


if __name__ == "__main__":
    main()
