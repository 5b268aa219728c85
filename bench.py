#!/usr/bin/env python
"""bench.py — MCTS simulations/sec of the batched MuZero search (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the reference path's CPU restatement

A "step" is one `act` over one batch of synthetic observations: root inference, `num_simulations` rounds of
select -> recurrent_fn -> backup, and the final action draw.  Workload at N=1 = the configuration the metric
is quoted on: CartPole-v1 MLP (obs 4, embed 8, hidden 16, A=2, support 21), batch 4096, num_sim 50.

N > 1 (`--scaling weak`, default): every rank owns `batch` trees of a `batch * N` global batch; `--scaling strong`:
the global batch stays `batch` and every rank owns batch / N rows ("rows sharded 4096/G").  PRNG draws
are indexed by global row, so any split reproduces the single-GPU result.  The only exchange on the path is one
all-gather per act of (action_weights, root_value, action) for the shared replay buffer; nothing a rank needs to step
its own environments depends on it, so it runs on a side stream and overlaps the NEXT act's search
(muax_b200/sharded.py `act_async`): a timed step = search of act t + whatever of act t-1's gather is still in flight.

`value`  : whole-job sims/s with observations already resident in HBM (CUDA events per step, max over ranks).
`e2e`    : the same metric through the reference-shaped plug-in call `MuZero.act(key, obs, with_pi=True,
           with_value=True, obs_from_batch=True)` with HOST (NumPy) observations in and NumPy results out — pinned
           H2D / D2H and the stream sync inside the timed region; at N > 1 followed by the (synchronous) gather.
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
    # name: obs_dim, E, A, S, hidden, batch per GPU (weak) / global batch (strong), num_sim, policy, qtransform
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
    # C5 per-GPU shard with a flat 256-wide observation standing in for the conv torso's output
    "atari_mlp_e256_b1024_sim50": dict(obs_dim=256, E=256, A=18, S=10, hidden=(256,), batch=1024, num_sim=50,
                                       policy=0, qtransform=0, minmax=1),
    # C5 as specified: 84x84x4 uint8 frames -> conv root representation (torch / cuDNN, once per act) -> embedding 256
    # -> search with MLP Dynamic / Prediction heads of width 256, 18 actions; 1024 trees per GPU (8192 over 8 GPUs)
    "atari_conv_e256_b1024_sim50": dict(obs_dim=0, E=256, A=18, S=10, hidden=(256,), batch=1024, num_sim=50, policy=0,
                                        qtransform=0, minmax=1, conv=(84, 84, 4)),
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
    """haiku default init (w ~ TruncNormal(0, 1/sqrt(fan_in)), b = 0) — BASELINE.md §4.  Draw order: Representation
    (when the workload has one), value head, policy head, next-state head, reward head — unchanged since round 1, so
    the synthetic networks and their mean path depths are the same."""
    rng = np.random.default_rng(seed)
    E, A, F = wl["E"], wl["A"], 2 * wl["S"] + 1

    def mlp(i, o):
        dims = [i, *wl["hidden"], o]
        return [haiku_linear(rng, a, b) for a, b in zip(dims[:-1], dims[1:])]

    nets = {}
    if wl["obs_dim"] > 0:
        nets["repr"] = [haiku_linear(rng, wl["obs_dim"], E)]
    nets.update(pred_v=mlp(E, F), pred_pi=mlp(E, A), dyn_ns=mlp(E + A, E), dyn_r=mlp(E + A, F))
    return nets


def bytes_per_sim(wl, mean_depth):
    """fp32 SoA tree traffic per simulation = D*(56 + 24A) + 8E + 4A + 36."""
    return mean_depth * (56 + 24 * wl["A"]) + 8 * wl["E"] + 4 * wl["A"] + 36


def flops_per_sim(wl):
    """recurrent_fn FLOPs per simulation (2 x MACs of the four heads)."""
    E, A, F = wl["E"], wl["A"], 2 * wl["S"] + 1

    def macs(i, o):
        dims = [i, *wl["hidden"], o]
        return sum(a * b for a, b in zip(dims[:-1], dims[1:]))

    return 2 * (macs(E + A, E) + macs(E + A, F) + macs(E, F) + macs(E, A))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", 1400.0)), "measured"
        except Exception:
            pass
    return 6650.0, 1400.0, "fallback"


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


def shard(args, wl, world):
    """-> (rows per rank, global batch)."""
    if args.scaling == "strong":
        if wl["batch"] % world:
            raise SystemExit(f"--scaling strong needs the batch ({wl['batch']}) to divide by the GPU count")
        return wl["batch"] // world, wl["batch"]
    return wl["batch"], wl["batch"] * world


def common_config(args, wl, name, world):
    """The `config` object BOTH arms print (the driver compares them key by key)."""
    B, GB = shard(args, wl, world)
    return {"workload": name, "batch_per_gpu": B, "global_batch": GB, "num_simulations": wl["num_sim"],
            "num_actions": wl["A"], "embed_dim": wl["E"], "policy": "muzero" if wl["policy"] == 0 else "gumbel",
            "parallelism": f"dp{world}", "scaling": args.scaling, "precision": args.precision}


def root_inputs(wl, rows, seed=1):
    """Synthetic root inputs for `rows` trees: flat observations, or (conv workload) the embeddings the search starts
    from on the CPU arm (the conv torso is outside the search path the CPU restatement covers)."""
    rng = np.random.default_rng(seed)
    if wl.get("conv"):
        return rng.random((rows, wl["E"])).astype(np.float32)
    return rng.standard_normal((rows, wl["obs_dim"])).astype(np.float32)


def cpu_c_port(wl, nets, x, key, steps, warmup, global_batch, threads=0):
    """Times the scalar C restatement of the reference path (oracle/mz_oracle.c) on all host cores."""
    from oracle import c_oracle
    c_oracle.build()
    threads = threads or c_oracle.max_threads()
    kw = dict(policy=wl["policy"], qtransform=wl["qtransform"], num_simulations=wl["num_sim"],
              support_size=wl["S"], repr_minmax=wl["minmax"], dyn_minmax=wl["minmax"], want_tree=False,
              nthreads=threads, global_batch=global_batch)
    if wl.get("conv"):  # root = (prior logits, value, embedding) from the NumPy model on the given embeddings
        from oracle import np_mctx
        m = np_mctx.ExactMath()
        v = np_mctx.stack(m, nets["pred_v"], x, 0)
        logits = np_mctx.stack(m, nets["pred_pi"], x, 0)
        value = np_mctx.support_to_scalar(m, np_mctx.softmax(m, v), wl["S"])
        run = lambda: c_oracle.search(nets, key, root=(logits, value, x), **kw)  # noqa: E731
    else:
        run = lambda: c_oracle.search(nets, key, obs=x, **kw)  # noqa: E731
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = time.perf_counter() - t0
    return x.shape[0] * wl["num_sim"] * steps / dt, dt / steps, threads, c_oracle.build_flags()


def cpu_numpy_batched(wl, nets, x, key, budget_s=12.0):
    """CPU-A of BASELINE.md §3: the batched NumPy restatement (oracle/np_mctx.py, libm math + BLAS) — the same array
    program mctx runs ([B, N, A] arrays, masked loops), the closest stand-in for "JAX on CPU".  Bounded sample: as many
    trees of the batch as fit the time budget."""
    from oracle import np_mctx
    if wl.get("conv"):
        return None
    rows = min(x.shape[0], 256)
    kw = dict(math=np_mctx.LibmMath(), policy=wl["policy"], qtransform=wl["qtransform"], num_simulations=wl["num_sim"],
              support_size=wl["S"], repr_minmax=wl["minmax"], dyn_minmax=wl["minmax"])
    t0 = time.perf_counter()
    np_mctx.act(nets, key, obs=x[:rows], global_batch=x.shape[0], **kw)
    dt = time.perf_counter() - t0
    if dt < budget_s / 4 and rows < x.shape[0]:  # cheap enough: take a larger sample once
        rows = min(x.shape[0], int(rows * budget_s / (2 * dt)))
        t0 = time.perf_counter()
        np_mctx.act(nets, key, obs=x[:rows], global_batch=x.shape[0], **kw)
        dt = time.perf_counter() - t0
    return {"value": rows * wl["num_sim"] / dt, "unit": "sims/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"one act of {rows} trees x {wl['num_sim']} simulations ({dt:.1f} s), NumPy batched restatement "
                      f"(BLAS threads as configured by the host)"}


def run_reference(args, wl, name):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    nets = make_nets(wl)
    B, GB = shard(args, wl, world)
    # bounded sample of the job: one rank's rows (the CPU arm has no ranks; its sims/s does not depend on the row count
    # beyond filling the cores)
    x = root_inputs(wl, B)
    key = np.array([0, 0], np.uint32)
    value, sec, threads, flags = cpu_c_port(wl, nets, x, key, args.steps, max(args.warmup, 1), GB)
    sample = f"{B} trees x {wl['num_sim']} simulations per step, {args.steps} steps"
    line = {
        "impl": "reference", "metric": "mcts_simulations_per_sec", "value": value, "unit": "sims/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": common_config(args, wl, name, world),
        "details": {"note": "CPU restatement of the mctx path (JAX / mctx are not installable on this image); scalar C, "
                            "one tree per task, pthreads over all host cores", "cc_flags": flags,
                    "host_cpus": os.cpu_count()},
        "cpu_baseline": {"value": value, "unit": "sims/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def build_model(wl, nets, dev):
    """The workload's networks behind the reference-shaped API (muax_b200.MuZero)."""
    import torch

    import muax_b200
    from muax_b200 import nn

    class Pred(nn.Prediction):
        hidden, normalize = wl["hidden"], bool(wl["minmax"])

    class Dyn(nn.Dynamic):
        hidden, normalize = wl["hidden"], bool(wl["minmax"])

    class Rep(nn.Representation):
        normalize = bool(wl["minmax"])

    E, A, F = wl["E"], wl["A"], 2 * wl["S"] + 1
    if wl.get("conv"):
        from muax_b200.conv import ResNetRepresentation
        torch.backends.cudnn.benchmark = True  # fixed shapes: let cuDNN pick its fastest convolution algorithms
        torch.manual_seed(0)
        H, W, C = wl["conv"]
        conv = ResNetRepresentation(E, frame_channels=C, height=H, width=W).to(dev).eval()
        network = nn.MZNetwork(conv, nn._init_prediction_func(Pred, A, F), nn._init_dynamic_func(Dyn, E, A, F))
    else:
        network = nn.create_muzero_network(Rep, Pred, Dyn, E, A, F)
    model = muax_b200.MuZero(network, policy="muzero" if wl["policy"] == 0 else "gumbel", discount=0.99,
                             support_size=wl["S"], device=dev)

    def haiku(prefix, stacks):
        out, i = {}, 0
        for layers in stacks:
            for w, b in layers:
                out[f"{prefix}/linear" if i == 0 else f"{prefix}/linear_{i}"] = {"w": w, "b": b}
                i += 1
        return out

    model.params = nn.MZNetworkParams(haiku("representation", [nets["repr"]]) if "repr" in nets else None,
                                      haiku("prediction", [nets["pred_v"], nets["pred_pi"]]),
                                      haiku("dynamic", [nets["dyn_ns"], nets["dyn_r"]]))
    return model


def run_ours(args, wl, name):
    import torch
    import torch.distributed as dist

    from muax_b200 import _lib
    from muax_b200.sharded import ShardedSearch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # The exchange is one 64 KB-per-rank all-gather per act, overlapped with the next act's search kernel, which
        # occupies 147 of the 148 SMs (one CTA of 210 KB shared memory each).  With NCCL's default channel count its
        # kernel's CTAs take SMs the next search kernel is waiting for (+19 us per act measured at N = 2); one channel
        # is one CTA and fits the free SM.  A user's own setting wins.
        os.environ.setdefault("NCCL_MAX_NCHANNELS", os.environ.get("MZ_BENCH_NCCL_CHANNELS", "1"))
        os.environ.setdefault("NCCL_MIN_NCHANNELS", "1")
        if os.environ.get("MZ_BENCH_NCCL_PRIO", "0") != "0":
            opts = dist.ProcessGroupNCCL.Options()
            opts.is_high_priority_stream = True
            dist.init_process_group("nccl", device_id=dev, pg_options=opts)
        else:
            dist.init_process_group("nccl", device_id=dev)
    B, GB = shard(args, wl, world)
    NS, A = wl["num_sim"], wl["A"]
    nets = make_nets(wl)
    model = build_model(wl, nets, dev)
    engine_id = {"auto": _lib.ENGINE_AUTO, "stepwise": _lib.ENGINE_STEPWISE, "fused": _lib.ENGINE_FUSED,
                 "fused_warp": _lib.ENGINE_FUSED_WARP, "treewarp": _lib.ENGINE_TREEWARP,
                 "resident": _lib.ENGINE_RESIDENT}[args.engine]
    act_kw = dict(num_simulations=NS, engine=engine_id, precision=args.precision)
    if wl["policy"] == 1:
        act_kw["qtransform"] = wl["qtransform"]
    my_rows = dict(global_batch=GB, batch_offset=rank * B)

    # ---- inputs: rank r owns global rows [r * B, (r + 1) * B)
    if wl.get("conv"):
        H, W, C = wl["conv"]
        obs_all = np.random.default_rng(1).integers(0, 256, (GB, H, W, C), dtype=np.uint8)
    else:
        obs_all = root_inputs(wl, GB)
    obs_host = np.ascontiguousarray(obs_all[rank * B:(rank + 1) * B])
    obs_dev = torch.from_numpy(obs_host).to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def search_fn(rng_key, obs_local, out=None, global_batch=None, batch_offset=0, **kw):
        return model.act_device(rng_key, obs_local, out=out, global_batch=global_batch, batch_offset=batch_offset,
                                **act_kw, **kw)

    # N > 1: the exchange is the NCCL all-gather of act t - 1 on a side stream, overlapped with act t.  MZ_BENCH_PEER=1
    # selects the search kernel's own NVLink peer stores + completion flags instead (no NCCL, no barrier kernel;
    # bit-identical, tests/test_sharded_gpu.py) — measured slower at the headline shapes on 2 x B200 (0.257 against
    # 0.244 ms per step: the system-scope fences at the end of the kernel cost what the overlapped all-gather saves)
    peer_mode = None if os.environ.get("MZ_BENCH_PEER", "0") != "0" else False
    sharded = ShardedSearch(search_fn, GB, A, writes_into_out=True, peer_stores=peer_mode,
                            engine=(lambda: model._engine_for(B, NS)) if not wl.get("conv") else None)

    def engine():
        return next(iter(model._engines.values()))[0]

    pending = []
    # diagnostic only (default "full" = the contract): "nowait" leaves the exchange unawaited inside the timed steps,
    # "none" does not launch it — to attribute the per-step cost of the exchange at N > 1
    exchange_mode = os.environ.get("MZ_BENCH_EXCHANGE", "full")
    exchange_lag = 2 if exchange_mode == "lag2" else 1

    def step_device(i):
        """act t on the main stream; the all-gather of act t - 1 starts at the same moment on the side stream and the
        step ends when both are done (N = 1: just the search)."""
        key = np.array([0, i], np.uint32)
        if pending and exchange_mode != "none":
            pending[-1].launch()          # the exchange of act t - 1 starts together with act t's search
        handle = sharded.act_async(key, obs_dev)
        if exchange_mode in ("full", "lag2") and len(pending) >= exchange_lag:
            pending.pop(0).wait()         # ... and the step ends when the exchange of act t - lag is in
        elif pending and exchange_mode not in ("full", "lag2"):
            pending.pop(0)
        if world > 1:
            pending.append(handle)
        return handle

    def drain():
        while pending:
            h = pending.pop(0)
            h.launch()
            h.wait()

    def step_host(i):
        key = np.array([0, i], np.uint32)
        if world > 1:  # NumPy in, NumPy out, results of every rank on every rank: search + all-gather + one D2H
            return sharded.act_host(key, obs_host)[0]
        return model.act(key, obs_host, with_pi=True, with_value=True, obs_from_batch=True, **act_kw, **my_rows)[0]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(step_fn, steps, events=True, tail=None):
        """events=True: device time.  Every step is bracketed by its own CUDA event pair on the launching stream, with
        the L2 flush between steps outside the pairs; the host does NOT wait between steps (it queues flush, events
        and launches ahead of the GPU, as an acting loop that keeps its observations on the device does), so a pair
        measures the step's device work and not the host's launch path.  One synchronisation + barrier on both sides
        of the K steps.  events=False: host wall clock per step (the step synchronises itself: NumPy in, NumPy out)."""
        per_step, pairs = [], []
        sync_all()
        wall0 = time.perf_counter()
        for i in range(steps + (1 if tail else 0)):
            flush.fill_(float(i))  # L2 flush between timed iterations, outside the timed events
            if events:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if i < steps:
                    step_fn(1000 + i)
                else:
                    tail()  # the last act's exchange
                e1.record()
                pairs.append((e0, e1))
            else:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                step_fn(2000 + i)  # synchronises internally (D2H of the outputs)
                per_step.append((time.perf_counter() - t0) * 1e3)
        sync_all()
        wall = time.perf_counter() - wall0
        if events:
            per_step = [a.elapsed_time(b) for a, b in pairs]
        return per_step, wall

    def kernel_times(steps):
        """Device time of the search kernels alone (the library's own event pair around its launches), one act at a
        time with the same L2 flush: the roofline's denominator."""
        out = []
        for i in range(steps):
            flush.fill_(float(i))
            step_device(1000 + i)
            drain()
            out.append(engine().last_kernel_ms())  # waits for the act
        return out

    for i in range(max(args.warmup, 3)):
        step_device(i)
        drain()
        step_host(i)
    eng = engine()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    dev_ms, wall = timed(step_device, args.steps, events=True, tail=drain if world > 1 else None)
    launches = eng.launch_count() - launches0
    kern_ms = kernel_times(args.steps)
    host_ms, _ = timed(step_host, args.steps, events=False)
    clocks = sampler.stop() if rank == 0 else None

    tot = torch.tensor([sum(dev_ms), sum(host_ms), sum(kern_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    dev_total_ms, host_total_ms, kern_total_ms = (float(x) for x in tot.tolist())
    # untimed: one more search keeping its tree, for the mean selected-path depth D of the roofline formula
    model.act_device(np.array([0, 1000], np.uint32), obs_dev, want_tree=True, **act_kw, **my_rows)
    depth = float(eng.tree()["sim_depth"].float().mean().item())
    if rank == 0:
        sims = GB * NS * args.steps
        value = sims / (dev_total_ms * 1e-3)
        e2e = sims / (host_total_ms * 1e-3)
        hbm_peak, tensor_peak, peak_src = measured_peaks()
        kernel_ms = kern_total_ms / args.steps
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        traffic = None
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get(name, {}).get(args.engine if args.precision == "fp32" else "bf16")
            except Exception:
                traffic = None
        if args.precision == "bf16":
            alg = B * NS * flops_per_sim(wl)
            achieved = alg / (kernel_ms * 1e-3) / 1e12
            roofline = {"bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
                        "frac": achieved / tensor_peak, "traffic": traffic, "peak_source": peak_src + " (sustained bf16)",
                        "kernel_ms": kernel_ms, "algorithmic_flops_per_launch": alg,
                        "note": "recurrent_fn FLOPs of one act / device time of the whole act (select + tcgen05 "
                                "recurrent kernel + backup per simulation); the recurrent kernel's own share is in "
                                "profiles/"}
        else:
            alg = B * NS * bytes_per_sim(wl, depth) + 4.0 * B * (wl["obs_dim"] + A + 2)
            achieved = alg / (kernel_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                        "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src, "kernel_ms": kernel_ms,
                        "algorithmic_bytes_per_launch": alg,
                        "note": "speed-of-data yardstick: the search is latency- / issue-bound (dependent simulations "
                                "per tree), see DESIGN.md"}
        x_cpu = root_inputs(wl, B)
        cpu_val, cpu_sec, threads, flags = cpu_c_port(wl, nets, x_cpu, np.array([0, 0], np.uint32), 3, 1, GB)
        line = {
            "metric": "mcts_simulations_per_sec", "value": value, "unit": "sims/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_total_ms / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": common_config(args, wl, name, world),
            "details": {"engine": args.engine, "mean_path_depth": depth, "exchange": sharded.exchange,
                        "l2": "256 MB buffer written between timed steps (outside the CUDA-event window)",
                        "timing": "one CUDA event pair per step on the launching stream, queued ahead by the host "
                                  "(no host wait between steps; the step waits for the previous act's exchange), "
                                  "summed, plus the last exchange; max over ranks; kernel_ms from a second pass of "
                                  "the same steps",
                        "wall_s_incl_flush": wall},
            "e2e": {"value": e2e, "unit": "sims/s", "h2d_bytes_per_step": int(obs_host.nbytes),
                    "d2h_bytes_per_step": int(GB * (4 + 4 * A + 4)), "ms_per_step": host_total_ms / args.steps,
                    "api": ("MuZero.act(key, obs, with_pi=True, with_value=True, obs_from_batch=True): NumPy observations "
                            "in, NumPy (action, action_weights, root_value) out, one stream sync") if world == 1 else
                           ("ShardedSearch.act_host(key, obs): NumPy observations of this rank's rows in, NumPy (action, "
                            "action_weights, root_value) of ALL ranks' rows out: pinned H2D, MuZero.act_device, NCCL "
                            "all-gather, one D2H, one stream sync")},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": {"value": cpu_val, "unit": "sims/s", "cores": threads, "kind": "port",
                             "sample": f"3 full steps of {B} trees x {NS} simulations ({cpu_sec * 1e3:.0f} ms each), "
                                       f"scalar C port, cc flags: {flags}, host cpus: {os.cpu_count()}"},
            "cpu_baseline_numpy": cpu_numpy_batched(wl, nets, x_cpu, np.array([0, 0], np.uint32)),
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
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl, args.workload)
    else:
        run_ours(args, wl, args.workload)


if __name__ == "__main__":
    main()
