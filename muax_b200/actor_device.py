"""Device-resident acting loop: the batched tracer and the vector environment as torch ops on the search's GPU.

`muax_b200/actor.py` (NumPy, reference-pinned) spends ~9 ms of host time per step next to a 0.5 ms search
(tools/actor_bench.py).  Here observations, the n-step caches and the episodes in progress never leave the device:
per step the host enqueues the search, ~60 small fixed-shape torch kernels and ONE tiny D2H (which environments
ended); only finished episodes are copied out (one gather + one D2H per field) into the same `TrajectoryStore`.

  DevicePNStep        == actor.BatchedPNStep  (muax/episode_tracer.py:114-249), float64, fixed-shape outputs + mask
  CartPoleVecTorch    == actor.CartPoleVec    on device
  DeviceActor         == actor.VectorActor    (muax/train.py:148-173)

Parity: tests feed DevicePNStep the reference pins (tests/golden/tracer_pins.npz); Rn / w agree with the reference to
1e-12 (the device sums the n discounted rewards in a different association order than `np.sum`), everything else
exactly.
"""
import numpy as np
import torch

from .actor import Transitions


class DevicePNStep:
    """`batch` PNStep caches as ring tensors.  `add` returns fixed-shape blocks `[batch, n + 1, ...]` plus a validity
    mask: slot k of environment b is the k-th transition popped for b at this step (slot 0 in steady state, the whole
    flushed cache when the episode ended) — no data-dependent shapes, so nothing synchronises."""

    def __init__(self, batch, n, gamma, alpha=0.5, device="cuda"):
        self.batch, self.n, self.gamma, self.alpha = int(batch), int(n), float(gamma), float(alpha)
        self.device = torch.device(device)
        f64 = torch.float64
        self._cap = self.n + 1
        self._gammas = torch.pow(torch.tensor(self.gamma, dtype=f64), torch.arange(self.n, dtype=f64)).to(self.device)
        self._gamman = float(np.power(self.gamma, self.n))
        B, C = self.batch, self._cap
        self._obs = self._pi = None
        self._a = torch.zeros(B, C, dtype=torch.int64, device=self.device)
        self._r = torch.zeros(B, C, dtype=f64, device=self.device)
        self._v = torch.zeros(B, C, dtype=f64, device=self.device)
        self._len = torch.zeros(B, dtype=torch.int64, device=self.device)
        self._head = torch.zeros(B, dtype=torch.int64, device=self.device)
        self._rows = torch.arange(B, device=self.device)
        self._k = torch.arange(C, device=self.device)[None, :]            # [1, n + 1] slot index
        self._i = torch.arange(self.n, device=self.device)[None, None, :]  # [1, 1, n] reward offset

    def add(self, obs, a, r, done, v, pi):
        """All arguments are device tensors with leading dimension `batch`.  Returns (Transitions of `[B, n + 1, ...]`
        tensors, mask `[B, n + 1]`)."""
        B, C, n = self.batch, self._cap, self.n
        if self._obs is None:
            self._obs = torch.zeros((B, C) + tuple(obs.shape[1:]), dtype=obs.dtype, device=self.device)
            self._pi = torch.zeros((B, C) + tuple(pi.shape[1:]), dtype=pi.dtype, device=self.device)
        tail = (self._head + self._len) % C
        self._obs[self._rows, tail] = obs
        self._pi[self._rows, tail] = pi
        self._a[self._rows, tail] = a.to(torch.int64)
        self._r[self._rows, tail] = r.to(torch.float64)
        self._v[self._rows, tail] = v.to(torch.float64)
        self._len += 1
        done = done.to(torch.bool)
        length = self._len[:, None]                               # [B, 1]
        steady = (~done & (self._len > n))[:, None]               # one pop (slot 0)
        mask = torch.where(done[:, None], self._k < length, steady & (self._k == 0))
        remaining = length - self._k                              # cached steps from slot k on
        m = torch.clamp(remaining, max=n)                         # rewards in the partial return
        slot = (self._head[:, None] + self._k) % C                # [B, C]
        ridx = (slot[:, :, None] + self._i) % C                   # [B, C, n]
        rs = torch.gather(self._r[:, None, :].expand(B, C, C), 2, ridx)
        Rn = (rs * self._gammas * (self._i < m[:, :, None])).sum(-1)
        boot = (remaining - 1) >= n
        v_next = torch.gather(self._v, 1, (slot + n) % C)
        Rn = Rn + torch.where(boot, v_next * self._gamman, torch.zeros_like(v_next))
        v_slot = torch.gather(self._v, 1, slot)
        w = torch.abs(v_slot - Rn) ** self.alpha
        rows = self._rows[:, None].expand(B, C)
        trans = Transitions(obs=self._obs[rows, slot], a=torch.gather(self._a, 1, slot), r=torch.gather(self._r, 1, slot),
                            done=~boot, Rn=Rn, v=v_slot, pi=self._pi[rows, slot], w=w)
        popped = steady[:, 0].to(torch.int64)
        # in place: the step may be replayed from a CUDA graph, which sees the state through fixed addresses
        self._head.copy_(torch.where(done, torch.zeros_like(self._head), (self._head + popped) % C))
        self._len.copy_(torch.where(done, torch.zeros_like(self._len), self._len - popped))
        return trans, mask


class CartPoleVecTorch:
    """actor.CartPoleVec on the device (float64 state, float32 observations); finished environments auto-reset."""
    GRAVITY, MASSCART, MASSPOLE, LENGTH, FORCE, TAU = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02
    X_LIMIT, THETA_LIMIT, MAX_STEPS = 2.4, 12 * 2 * np.pi / 360, 500
    obs_dim, num_actions = 4, 2

    def __init__(self, batch, seed=0, device="cuda", x_limit=None, theta_limit=None):
        self.batch, self.device = int(batch), torch.device(device)
        if x_limit is not None:
            self.X_LIMIT = float(x_limit)
        if theta_limit is not None:
            self.THETA_LIMIT = float(theta_limit)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(seed)
        self.state = torch.zeros(self.batch, 4, dtype=torch.float64, device=self.device)
        self.t = torch.zeros(self.batch, dtype=torch.int64, device=self.device)

    def _fresh(self):
        return torch.rand(self.batch, 4, dtype=torch.float64, device=self.device, generator=self.gen) * 0.1 - 0.05

    def reset(self):
        self.state.copy_(self._fresh())
        self.t.zero_()
        return self.state.to(torch.float32)

    def step(self, action):
        x, x_dot, th, th_dot = self.state.unbind(1)
        force = torch.where(action == 1, self.FORCE, -self.FORCE).to(torch.float64)
        total_mass, pml = self.MASSPOLE + self.MASSCART, self.MASSPOLE * self.LENGTH
        cos, sin = torch.cos(th), torch.sin(th)
        temp = (force + pml * th_dot ** 2 * sin) / total_mass
        th_acc = (self.GRAVITY * sin - cos * temp) / (self.LENGTH * (4.0 / 3.0 - self.MASSPOLE * cos ** 2 / total_mass))
        x_acc = temp - pml * th_acc * cos / total_mass
        state = torch.stack([x + self.TAU * x_dot, x_dot + self.TAU * x_acc, th + self.TAU * th_dot,
                             th_dot + self.TAU * th_acc], dim=1)
        self.t += 1
        terminated = (state[:, 0].abs() > self.X_LIMIT) | (state[:, 2].abs() > self.THETA_LIMIT)
        done = terminated | (self.t >= self.MAX_STEPS)
        self.state.copy_(torch.where(done[:, None], self._fresh(), state))  # in place (CUDA-graph replay)
        self.t.copy_(torch.where(done, torch.zeros_like(self.t), self.t))
        reward = torch.ones(self.batch, dtype=torch.float64, device=self.device)
        return self.state.to(torch.float32), reward, done


class DeviceActor:
    """actor.VectorActor with everything but the finished episodes resident on the GPU.

    Per step: the search (one kernel, written into fixed output buffers), then environment step + tracer + episode
    scatter — ~90 small fixed-shape torch kernels that are launch-bound next to a 0.3 ms search, so on CUDA they are
    captured once into a **CUDA graph** and replayed (`use_graph`; all state is updated in place) — then ONE tiny D2H
    (which environments ended).  The tensors `step` returns are the fixed buffers: valid until the next step."""

    def __init__(self, model, env, store, n=10, gamma=0.997, alpha=0.5, k_steps=5, num_simulations=50,
                 temperature=1.0, act_kwargs=None, max_episode_steps=None, use_graph=None):
        self.model, self.env, self.store = model, env, store
        self.device = env.device
        self.tracer = DevicePNStep(env.batch, n, gamma, alpha, device=self.device)
        self.k_steps, self.num_simulations, self.temperature = int(k_steps), int(num_simulations), float(temperature)
        self.act_kwargs = dict(act_kwargs or {})
        self.L = int(max_episode_steps or getattr(env, "MAX_STEPS", 1000))
        self._episode = None  # Transitions of [batch, L + 1, ...] tensors; column L swallows masked-off writes
        B = env.batch
        self._ep_len = torch.zeros(B, dtype=torch.int64, device=self.device)
        self._rows = torch.arange(B, device=self.device)
        self.obs = env.reset().clone()
        self._out = (torch.zeros(B, dtype=torch.int32, device=self.device),
                     torch.zeros(B, env.num_actions, dtype=torch.float32, device=self.device),
                     torch.zeros(B, dtype=torch.float32, device=self.device))
        self._done = torch.zeros(B, dtype=torch.bool, device=self.device)
        self.use_graph = (self.device.type == "cuda") if use_graph is None else bool(use_graph)
        self._graph, self._eager_steps = None, 0
        self.env_steps = 0
        self.episodes = 0

    def _after_search(self):
        """Environment step, tracer, scatter into the episodes in progress: fixed shapes, state updated in place."""
        action, weights, value = self._out
        obs_next, r, done = self.env.step(action)
        trans, mask = self.tracer.add(self.obs, action, r, done, value, weights)
        B, C = mask.shape
        if self._episode is None:
            self._episode = Transitions(*(torch.zeros((B, self.L + 1) + tuple(x.shape[2:]), dtype=x.dtype,
                                                      device=self.device) for x in trans))
        # slot k of environment b lands at position ep_len[b] + (number of valid slots before k); masked-off slots
        # are written to the spare column L
        pos = self._ep_len[:, None] + torch.cumsum(mask.to(torch.int64), 1) - 1
        pos = torch.where(mask, torch.clamp(pos, max=self.L - 1), torch.full_like(pos, self.L))
        rows = self._rows[:, None].expand(B, C)
        for dst, src in zip(self._episode, trans):
            dst[rows, pos] = src
        self._ep_len.add_(mask.sum(1))
        self._done.copy_(done)
        self.obs.copy_(obs_next)

    def _capture(self):
        graph = torch.cuda.CUDAGraph()
        gen = getattr(self.env, "gen", None)
        if gen is not None:
            graph.register_generator_state(gen)  # the environment draws its reset states inside the graph
        torch.cuda.synchronize(self.device)
        with torch.cuda.graph(graph):
            self._after_search()
        return graph

    def step(self, rng_key):
        self.model.act_device(rng_key, self.obs, num_simulations=self.num_simulations, temperature=self.temperature,
                              out=self._out, **self.act_kwargs)
        if self._graph is not None:
            self._graph.replay()
        elif self.use_graph and self._eager_steps >= 2:  # two eager steps allocate the lazily-shaped buffers first
            self._graph = self._capture()                # capture only records: nothing has run yet
            self._graph.replay()
        else:
            self._after_search()
            self._eager_steps += 1
        action, weights, value = self._out
        B = self.env.batch
        ended = torch.nonzero(self._done)[:, 0]  # the one host sync of the step
        if len(ended):
            lengths = self._ep_len[ended].cpu().numpy()
            longest = int(lengths.max())
            blocks = [x[ended, :longest].cpu().numpy() for x in self._episode]  # one gather + D2H per field
            for j, length in enumerate(lengths):
                if length >= self.k_steps:
                    self.store.add(Transitions(*(blk[j, :length] for blk in blocks)))
            self.episodes += len(lengths)
            self._ep_len[ended] = 0
        self.env_steps += B
        return action, weights, value, self._done
