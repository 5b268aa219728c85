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
    ap.add_argument("--device-loop", action="store_true", help="tracer + environment on the GPU (actor_device.py)")
    ap.add_argument("--no-graph", action="store_true", help="device loop without the CUDA graph (eager torch kernels)")
    ap.add_argument("--long-episodes", action="store_true",
                    help="no early termination: episodes run the full 500 steps, as with a trained CartPole agent "
                         "(random weights end ~10% of the environments every step)")
    args = ap.parse_args()
    lim = dict(x_limit=1e9, theta_limit=1e9) if args.long_episodes else {}
    import muax_b200
    from muax_b200 import nn
    from muax_b200.actor import CartPoleVec, TrajectoryStore, VectorActor
    net = nn.create_muzero_network(nn.Representation, nn.Prediction, nn.Dynamic, 8, 2, 21)
    model = muax_b200.MuZero(net, policy="muzero", discount=0.997, support_size=10)
    model.init(muax_b200.random.PRNGKey(0), np.zeros((1, 4), np.float32))
    store = TrajectoryStore(100000, random_seed=0)
    if args.device_loop:
        import torch
        from muax_b200.actor_device import CartPoleVecTorch, DeviceActor
        env = CartPoleVecTorch(args.batch, seed=0, **lim)
        actor = DeviceActor(model, env, store, n=10, gamma=0.997, k_steps=5, num_simulations=args.num_simulations,
                            use_graph=not args.no_graph)
        for t in range(5):
            actor.step(muax_b200.random.PRNGKey(t))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in range(args.steps):
            actor.step(muax_b200.random.PRNGKey(1000 + t))
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        steps = args.steps * args.batch
        print(json.dumps({"metric": "env_steps_per_sec", "loop": "device", "value": steps / wall, "batch": args.batch,
                          "num_simulations": args.num_simulations, "steps": args.steps,
                          "ms_per_step": wall / args.steps * 1e3, "sims_per_sec_end_to_end": steps * args.num_simulations / wall,
                          "episodes": actor.episodes, "stored_episodes": len(store)}))
        return
    env = CartPoleVec(args.batch, seed=0, **lim)
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
    print(json.dumps({"metric": "env_steps_per_sec", "loop": "host", "value": steps / wall, "batch": args.batch,
                      "num_simulations": args.num_simulations, "steps": args.steps, "ms_per_step": wall / args.steps * 1e3,
                      "act_ms_per_step": act_s[0] / args.steps * 1e3,
                      "host_ms_per_step": (wall - act_s[0]) / args.steps * 1e3,
                      "sims_per_sec_end_to_end": steps * args.num_simulations / wall,
                      "episodes": actor.episodes, "stored_episodes": len(store)}))


if __name__ == "__main__":
    main()
