#!/usr/bin/env python
"""bench.py — MCTS simulations/sec of the batched MuZero search (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the reference path's CPU restatement

A "step" is one `act` over one batch of synthetic observations: root inference, `num_simulations` rounds of
select -> recurrent_fn -> backup, and the final action draw.  Workload at N=1 = the configuration the metric
is quoted on: CartPole-v1 MLP (obs 4, embed 8, hidden 16, A=2, support 21), batch 4096, num_sim 50.  With N>1
every rank owns 4096 trees of a 4096*N global batch (weak scaling; PRNG draws indexed by global row) and the
step ends with one all-gather of (action, action_weights, root_value) — the only exchange on this path.

`value`  : whole-job sims/s with observations already resident in HBM (CUDA events, max over ranks).
`e2e`    : the same metric through the host-buffer C-ABI call `mz_search_host` (pinned H2D of the observations
           and D2H of action/action_weights/root_value inside the timed region) — what MuZero.act does.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: obs_dim, E, A, S, hidden, batch per GPU, num_sim, policy, qtransform
    "cartpole_mlp_e8_b4096_sim50": dict(obs_dim=4, E=8, A=2, S=10, hidden=(16,), batch=4096, num_sim=50, policy=0,
                                        qtransform=0, minmax=1),
    "cartpole_mlp_e8_b1024_sim50": dict(obs_dim=4, E=8, A=2, S=10, hidden=(16,), batch=1024, num_sim=50, policy=0,
                                        qtransform=0, minmax=1),
    "lunarlander_mlp_e64_b4096_sim200": dict(obs_dim=8, E=64, A=4, S=10, hidden=(16,), batch=4096, num_sim=200,
                                             policy=0, qtransform=0, minmax=1),
    "lunarlander_gumbel_e64_b4096_sim32": dict(obs_dim=8, E=64, A=4, S=10, hidden=(16,), batch=4096, num_sim=32,
                                               policy=1, qtransform=0, minmax=1),
    # examples/lunarlander.ipynb cell 2-3: 64-64-16 ELU stacks, no min-max, support 20
    "lunarlander_notebook_e64_b4096_sim200": dict(obs_dim=8, E=64, A=4, S=20, hidden=(64, 64, 16), batch=4096,
                                                  num_sim=200, policy=0, qtransform=0, minmax=0),
    # C5 per-GPU shard: 1024 trees, A=18, E=H=256; the flat 256-wide "observation" stands in for the conv torso's
    # output (the conv root representation runs once per act outside the search loop)
    "atari_mlp_e256_b1024_sim50": dict(obs_dim=256, E=256, A=18, S=10, hidden=(256,), batch=1024, num_sim=50,
                                       policy=0, qtransform=0, minmax=1),
}
DEFAULT_WORKLOAD = "cartpole_mlp_e8_b4096_sim50"


def haiku_linear(rng, fan_in, fan_out):
    w = rng.standard_normal((fan_in, fan_out))
    bad = np.abs(w) > 2
    while bad.any():
        w[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(w) > 2
    return (w / np.sqrt(fan_in)).astype(np.float32), np.zeros(fan_out, np.float32)


def make_nets(wl, seed=0):
    """haiku default init (w ~ TruncNormal(0, 1/sqrt(fan_in)), b = 0) — BASELINE.md §4."""
    rng = np.random.default_rng(seed)
    E, A, F = wl["E"], wl["A"], 2 * wl["S"] + 1

    def mlp(i, o):
        dims = [i, *wl["hidden"], o]
        return [haiku_linear(rng, a, b) for a, b in zip(dims[:-1], dims[1:])]

    return dict(repr=[haiku_linear(rng, wl["obs_dim"], E)], pred_v=mlp(E, F), pred_pi=mlp(E, A),
                dyn_ns=mlp(E + A, E), dyn_r=mlp(E + A, F))


def bytes_per_sim(wl, mean_depth):
    """fp32 SoA tree traffic per simulation = D*(56 + 24A) + 8E + 4A + 36."""
    return mean_depth * (56 + 24 * wl["A"]) + 8 * wl["E"] + 4 * wl["A"] + 36


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx, self.rows, self.proc = device_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(wl, nets, obs, key, steps, warmup, global_batch, threads=0):
    """Times the scalar C restatement of the reference path (oracle/mz_oracle.c) on all host cores."""
    from oracle import c_oracle
    c_oracle.build()
    threads = threads or c_oracle.max_threads()
    kw = dict(policy=wl["policy"], qtransform=wl["qtransform"], num_simulations=wl["num_sim"],
              support_size=wl["S"], repr_minmax=wl["minmax"], dyn_minmax=wl["minmax"], want_tree=False,
              nthreads=threads, global_batch=global_batch)
    for _ in range(warmup):
        c_oracle.search(nets, key, obs=obs, **kw)
    t0 = time.perf_counter()
    for _ in range(steps):
        c_oracle.search(nets, key, obs=obs, **kw)
    dt = time.perf_counter() - t0
    return obs.shape[0] * wl["num_sim"] * steps / dt, dt / steps, threads


def run_reference(args, wl, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nets = make_nets(wl)
    B = wl["batch"]
    obs = np.random.default_rng(1).standard_normal((B, wl["obs_dim"])).astype(np.float32)
    key = np.array([0, 0], np.uint32)
    value, sec, threads = cpu_reference_run(wl, nets, obs, key, args.steps, max(args.warmup, 1), B)
    sample = f"full step: {B} trees x {wl['num_sim']} simulations per step, {args.steps} steps"
    line = {
        "impl": "reference", "metric": "mcts_simulations_per_sec", "value": value, "unit": "sims/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "batch": B, "num_simulations": wl["num_sim"],
                   "note": "CPU restatement of the mctx path (JAX/mctx are not installable on this image); "
                           "scalar C, one tree per task, pthreads over all host cores"},
        "cpu_baseline": {"value": value, "unit": "sims/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args, wl, name):
    import torch
    import torch.distributed as dist

    from muax_b200 import _lib
    from muax_b200.nn import pack_stacks
    from muax_b200.search import SearchEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, NS, A = wl["batch"], wl["num_sim"], wl["A"]
    GB = B * world
    nets = make_nets(wl)
    blob, cstacks = pack_stacks(nets)
    eng = SearchEngine(cstacks, batch=B, num_actions=A, embed_dim=wl["E"], obs_dim=wl["obs_dim"],
                       support_size=wl["S"], max_num_simulations=NS, repr_minmax=wl["minmax"],
                       dyn_minmax=wl["minmax"], discount=0.99, device=dev)
    eng.set_weights(blob)
    engine_id = {"auto": _lib.ENGINE_AUTO, "stepwise": _lib.ENGINE_STEPWISE, "fused": _lib.ENGINE_FUSED,
                 "fused_warp": _lib.ENGINE_FUSED_WARP, "treewarp": _lib.ENGINE_TREEWARP,
                 "resident": _lib.ENGINE_RESIDENT}[args.engine]
    obs_all = np.random.default_rng(1).standard_normal((GB, wl["obs_dim"])).astype(np.float32)
    obs_host = np.ascontiguousarray(obs_all[rank * B:(rank + 1) * B])
    obs_dev = torch.from_numpy(obs_host).to(dev)
    kw = dict(policy=wl["policy"], qtransform=wl["qtransform"], num_simulations=NS, global_batch=GB,
              batch_offset=rank * B, engine=engine_id)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2
    from muax_b200.sharded import ShardedSearch
    kw_local = {k: v for k, v in kw.items() if k not in ("global_batch", "batch_offset")}
    peer = {"auto": None, "on": True, "off": False}[os.environ.get("MZ_PEER_STORES", "off")]
    sharded = ShardedSearch(eng.search, GB, A, writes_into_out=True, peer_stores=peer)

    def step_device(i):
        # rows [rank*B, (rank+1)*B) of the global batch; with world > 1 this ends with the one all-gather of
        # (action_weights, root_value, action) for the shared replay buffer — done by the search kernel's own NVLink
        # peer stores + a cross-rank barrier when symmetric memory is available, else one NCCL all-gather
        key = np.array([0, i], np.uint32)
        return sharded.act(key, obs_dev, **kw_local)

    def step_host(i):
        key = np.array([0, i], np.uint32)
        return eng.search_host(key, obs_host, **kw)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(step_fn, steps, events=True):
        per_step, kernel_ms = [], []
        sync_all()
        wall0 = time.perf_counter()
        for i in range(steps):
            flush.fill_(float(i))  # L2 flush between timed iterations, outside the timed events
            if events:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                step_fn(1000 + i)
                e1.record()
                e1.synchronize()
                per_step.append(e0.elapsed_time(e1))
                kernel_ms.append(eng.last_kernel_ms())
            else:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                step_fn(2000 + i)  # synchronises internally (D2H of the outputs)
                per_step.append((time.perf_counter() - t0) * 1e3)
        sync_all()
        wall = time.perf_counter() - wall0
        return per_step, kernel_ms, wall

    for i in range(max(args.warmup, 3)):
        step_device(i)
        step_host(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    dev_ms, kern_ms, wall = timed(step_device, args.steps, events=True)
    launches = eng.launch_count() - launches0
    host_ms, _, _ = timed(step_host, args.steps, events=False)
    clocks = sampler.stop() if rank == 0 else None

    tot = torch.tensor([sum(dev_ms), sum(host_ms), sum(kern_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    dev_total_ms, host_total_ms, kern_total_ms = (float(x) for x in tot.tolist())
    eng.search(np.array([0, 1000], np.uint32), obs=obs_dev, want_tree=True, **kw)  # untimed: the tree view for D
    depth = float(eng.tree()["sim_depth"].float().mean().item())
    if rank == 0:
        sims = GB * NS * args.steps
        value = sims / (dev_total_ms * 1e-3)
        e2e = sims / (host_total_ms * 1e-3)
        peak, peak_src = measured_peak_hbm()
        alg_bytes = B * NS * bytes_per_sim(wl, depth) + 4.0 * B * (wl["obs_dim"] + A + 2)
        kernel_ms = kern_total_ms / args.steps
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        nets1 = make_nets(wl)
        cpu_val, cpu_sec, threads = cpu_reference_run(wl, nets1, obs_host, np.array([0, 0], np.uint32), 3, 1, GB)
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        traffic = None
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get(name, {}).get(args.engine)
            except Exception:
                traffic = None
        line = {
            "metric": "mcts_simulations_per_sec", "value": value, "unit": "sims/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "batch_per_gpu": B, "global_batch": GB, "num_simulations": NS,
                       "num_actions": A, "embed_dim": wl["E"], "policy": "muzero" if wl["policy"] == 0 else "gumbel",
                       "engine": args.engine, "mean_path_depth": depth, "parallelism": f"dp{world}",
                       "exchange": sharded.exchange,
                       "l2": "256 MB buffer written between timed steps (outside the CUDA-event window)",
                       "timing": "CUDA events per step on the launching stream, summed, max over ranks",
                       "wall_s_incl_flush": wall},
            "e2e": {"value": e2e, "unit": "sims/s", "h2d_bytes_per_step": int(obs_host.nbytes),
                    "d2h_bytes_per_step": int(B * (4 + 4 * A + 4)), "ms_per_step": host_total_ms / args.steps,
                    "api": "mz_search_host (host buffers -> pinned staging -> the search kernel reads the observations and "
                           "writes action / action_weights / root_value over PCIe in place (mapped pinned memory; "
                           "copy-engine H2D for observation batches > 256 KB) -> stream sync -> caller's buffers)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "search is latency-bound (50 dependent simulations per tree); see DESIGN.md"},
            "cpu_baseline": {"value": cpu_val, "unit": "sims/s", "cores": threads, "kind": "port",
                             "sample": f"3 full steps of {B} trees x {NS} simulations ({cpu_sec * 1e3:.0f} ms each)"},
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default="auto", choices=["auto", "stepwise", "fused", "fused_warp", "treewarp", "resident"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl, args.workload)
    else:
        run_ours(args, wl, args.workload)


if __name__ == "__main__":
    main()
