"""Env-batch sharding across GPUs (one process per GPU, torch.distributed).

Trees are independent, so the search needs no collective: rank r owns global rows
[offset_r, offset_r + count_r) and runs them with `global_batch` / `batch_offset` set, which makes every PRNG
draw (per-tree simulate keys, root Dirichlet / Gumbel, final categorical) identical to the single-process run
for ANY world size (legacy threefry `split(key, B)[b]` mixes b with B — SURVEY.md §8e).  The only exchange on
the path is one all-gather per act of (action_weights, root_value, action) so that every rank can feed the
shared replay buffer (the reference's `TrajectoryReplayBuffer`, muax/replay_buffer.py:154).

Nothing a rank needs to step its OWN environments depends on that exchange: `act_async` returns the local results at
once and a handle whose all-gather runs on a side stream, double-buffered, so that act t + 1's search overlaps act t's
exchange (round 1 issued it synchronously on the search stream: 0.81 weak-scaling efficiency at 8 GPUs, all of it
NCCL latency on a 64 KB message).
"""
import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, world_size: int, rank: int):
    """Contiguous, as-even-as-possible row ranges: the first `global_batch % world_size` ranks get one extra."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, extra = divmod(int(global_batch), int(world_size))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def pack_outputs(action, weights, value):
    """[n] i32, [n, A] f32, [n] f32 -> one [n, A + 2] f32 tensor (actions < 2^24 are exact in float32)."""
    return torch.cat([weights, value[:, None], action[:, None].to(torch.float32)], dim=1)


def unpack_outputs(packed):
    A = packed.shape[1] - 2
    return packed[:, A + 1].to(torch.int32), packed[:, :A].contiguous(), packed[:, A].contiguous()


class GatherHandle:
    """The exchange of one act: `launch()` enqueues the all-gather on the side stream (it waits for the search that
    produced the send buffer), `wait()` makes the CURRENT stream wait for it and returns the global
    (action, weights, value) views.  `launch` is separate from `act_async` so that a caller can place it (bench.py
    starts it together with the next act's search)."""

    def __init__(self, owner, slot, local, peer_step=None):
        self._owner, self._slot, self.local = owner, slot, local
        self._work = None
        self._launched = False
        self._peer_step = peer_step  # peer-store exchange: the act's number (the flags count acts)

    def launch(self):
        if self._launched or self._owner.world == 1 or self._peer_step is not None:
            self._launched = True
            return self
        o = self._owner
        send, recv, ready = o._async_send[self._slot], o._async_recv[self._slot], o._async_ready[self._slot]
        if o._side is None:  # CPU tensors (gloo): no streams to overlap on
            self._work = dist.all_gather_into_tensor(recv, send, group=o.group, async_op=True)
        else:
            with torch.cuda.stream(o._side):
                o._side.wait_event(ready)  # the search kernel that filled `send`
                self._work = dist.all_gather_into_tensor(recv, send, group=o.group, async_op=True)
        self._launched = True
        return self

    def wait(self):
        o = self._owner
        if o.world == 1:
            return self.local
        n, A, W = o.count, o.A, o.world
        if self._peer_step is not None:
            # the rows were stored here by the peers' own search kernels; wait for their completion flags
            engine, bufs, flags = o._peer_async
            engine.peer_wait(flags[self._slot], W, self._peer_step)
            r = bufs[self._slot][:W * n * (A + 2)].view(W, n * (A + 2))
        else:
            self.launch()
            self._work.wait()  # the current stream waits; the host does not
            r = o._async_recv[self._slot].view(W, n * (A + 2))
        return (r[:, n * A + n:].view(torch.int32).reshape(W * n), r[:, :n * A].reshape(W * n, A),
                r[:, n * A:n * A + n].reshape(W * n))


class ShardedSearch:
    """Runs `search_fn` on this rank's rows and all-gathers the per-row outputs in global row order.

    search_fn(rng_key, obs_local, global_batch=..., batch_offset=..., **kw) -> (action, weights, value) tensors;
    with a `SearchEngine` pass `engine.search` (observations via `obs=`)."""

    def __init__(self, search_fn, global_batch, num_actions, group=None, writes_into_out=False, peer_stores=False,
                 engine=None):
        """writes_into_out: `search_fn` accepts `out=(action, weights, value)` and writes its results there
        (SearchEngine.search does).  With even shards the three outputs are then views of ONE flat send buffer, so
        the step is: search kernels -> one all-gather -> three strided copies, no packing kernels."""
        self.search_fn = search_fn
        self._engine = engine  # the SearchEngine behind search_fn (or a callable returning it) when search_fn is a wrapper
        self.writes_into_out = bool(writes_into_out)
        self._send = self._recv = None
        self.global_batch = int(global_batch)
        self.A = int(num_actions)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.offset, self.count = shard_bounds(self.global_batch, self.world, self.rank)
        self.max_count = shard_bounds(self.global_batch, self.world, 0)[1]
        self._gather_buf = None
        # peer_stores: the search kernel writes its outputs straight into every rank's gather buffer over NVLink
        # (torch symmetric memory + SearchEngine.set_peer_outputs) and the step ends with a cross-rank barrier instead
        # of an NCCL all-gather.  Opt-in: measured on 2 x B200 at the headline shapes the barrier kernel costs what the
        # 64 KB all-gather costs (0.336 vs 0.324 ms per act, profiles/r01_bench_n2_peer_*.json).  True = require it,
        # None = try it and fall back to NCCL, False (default) = NCCL.
        self.peer_stores = peer_stores
        self.exchange = "none" if self.world == 1 else "nccl all-gather"
        self._peer = None
        self._step = 0
        self.async_slots = 3  # act_async: a handle stays valid until async_slots - 1 further acts have been issued
        self.peer_slots = 4   # ... with the peer-store exchange (see _setup_peer_async)

    def local_rows(self, global_tensor):
        return global_tensor[self.offset:self.offset + self.count]

    def act(self, rng_key, obs_local, **kw):
        if obs_local.shape[0] != self.count:
            raise ValueError(f"rank {self.rank} owns {self.count} rows, got {obs_local.shape[0]}")
        if self.world > 1 and self.writes_into_out and self.global_batch % self.world == 0:
            return self._act_flat(rng_key, obs_local, **kw)
        action, weights, value = self.search_fn(rng_key, obs_local, global_batch=self.global_batch,
                                                batch_offset=self.offset, **kw)
        if self.world == 1:
            return action, weights, value
        packed = pack_outputs(action, weights, value)
        if self.count < self.max_count:  # ragged shards: pad to the largest one for the fixed-size collective
            pad = torch.zeros(self.max_count - self.count, self.A + 2, dtype=packed.dtype, device=packed.device)
            packed = torch.cat([packed, pad], dim=0)
        if (self._gather_buf is None or self._gather_buf.device != packed.device):
            self._gather_buf = torch.empty(self.world * self.max_count, self.A + 2, dtype=torch.float32,
                                           device=packed.device)
        dist.all_gather_into_tensor(self._gather_buf, packed.contiguous(), group=self.group)
        parts = []
        for r in range(self.world):
            _, cnt = shard_bounds(self.global_batch, self.world, r)
            parts.append(self._gather_buf[r * self.max_count:r * self.max_count + cnt])
        return unpack_outputs(torch.cat(parts, dim=0))

    def act_async(self, rng_key, obs_local, **kw):
        """Search this rank's rows now; exchange later.  Returns a GatherHandle: `.local` = this rank's
        (action, weights, value), `.wait()` = everybody's.  `async_slots` (3) send / receive buffer pairs rotate, so a
        handle stays valid until two further acts have been issued.  Needs even shards and a `search_fn` that writes into `out`."""
        if obs_local.shape[0] != self.count:
            raise ValueError(f"rank {self.rank} owns {self.count} rows, got {obs_local.shape[0]}")
        if self.world > 1 and (not self.writes_into_out or self.global_batch % self.world):
            raise ValueError("act_async needs even shards and a search_fn that writes into `out`")
        n, A, W = self.count, self.A, self.world
        row = n * (A + 2)
        dev = obs_local.device
        if W > 1 and self.peer_stores is not False and dev.type == "cuda":
            if getattr(self, "_peer_async", None) is not None or self._setup_peer_async(dev, row):
                return self._act_async_peer(rng_key, obs_local, row, **kw)
        if getattr(self, "_async_send", None) is None or self._async_send[0].device != dev:
            self._async_send = [torch.empty(row, dtype=torch.float32, device=dev) for _ in range(self.async_slots)]
            self._async_recv = [torch.empty(W * row, dtype=torch.float32, device=dev) for _ in range(self.async_slots)]
            self._async_ready = [torch.cuda.Event() if dev.type == "cuda" else None for _ in range(self.async_slots)]
            self._side = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
            self.exchange = "none" if W == 1 else "nccl all-gather on a side stream, overlapped with the next act"
        slot = self._step % self.async_slots
        self._step += 1
        send = self._async_send[slot]
        out = (send[n * A + n:].view(torch.int32), send[:n * A].view(n, A), send[n * A:n * A + n])
        self.search_fn(rng_key, obs_local, global_batch=self.global_batch, batch_offset=self.offset, out=out, **kw)
        if self._async_ready[slot] is not None:
            self._async_ready[slot].record()
        return GatherHandle(self, slot, out)

    def act_host(self, rng_key, obs_host, **kw):
        """The end-to-end sharded act for host-side observations (NumPy in, NumPy out): H2D of this rank's rows from a
        pinned staging buffer, the search (its kernels write into the all-gather send buffer), one all-gather on the
        device, ONE D2H of everybody's results into pinned memory, one stream synchronisation.  Returns the global
        (action i32[GB], action_weights f32[GB, A], root_value f32[GB]) on every rank.  Compared with `MuZero.act`
        followed by `gather_host` this saves a D2H + H2D round trip of the local results."""
        import numpy as np
        if not (self.writes_into_out and self.global_batch % max(self.world, 1) == 0):
            raise ValueError("act_host needs even shards and a search_fn that writes into `out`")
        n, A, W = self.count, self.A, self.world
        row = n * (A + 2)
        dev = torch.device("cuda", torch.cuda.current_device())
        obs_host = np.ascontiguousarray(obs_host)
        st = getattr(self, "_host_state", None)
        if st is None or st["obs"].shape != obs_host.shape or st["obs"].numpy().dtype != obs_host.dtype:
            t = torch.from_numpy(obs_host)
            st = dict(obs=torch.empty(t.shape, dtype=t.dtype).pin_memory(), obs_dev=torch.empty(t.shape, dtype=t.dtype, device=dev),
                      send=torch.empty(row, dtype=torch.float32, device=dev),
                      recv=torch.empty(W * row, dtype=torch.float32, device=dev),
                      out=torch.empty(W * row, dtype=torch.float32).pin_memory())
            self._host_state = st
        st["obs"].numpy()[...] = obs_host
        st["obs_dev"].copy_(st["obs"], non_blocking=True)
        send = st["send"]
        out = (send[n * A + n:].view(torch.int32), send[:n * A].view(n, A), send[n * A:n * A + n])
        self.search_fn(rng_key, st["obs_dev"], global_batch=self.global_batch, batch_offset=self.offset, out=out, **kw)
        if W > 1:
            dist.all_gather_into_tensor(st["recv"], send, group=self.group)
            st["out"].copy_(st["recv"], non_blocking=True)
        else:
            st["out"].copy_(send, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        r = st["out"].numpy().reshape(W, row)
        return (np.ascontiguousarray(r[:, n * A + n:]).view(np.int32).reshape(W * n),
                r[:, :n * A].reshape(W * n, A).copy(), r[:, n * A:n * A + n].reshape(W * n).copy())

    def gather_host(self, action, weights, value):
        """The synchronous exchange for host-side results (NumPy arrays of this rank's rows): returns the global
        arrays on every rank.  Used by the end-to-end path (`MuZero.act` returns NumPy)."""
        if self.world == 1:
            return action, weights, value
        import numpy as np
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        packed = torch.from_numpy(np.concatenate([np.asarray(weights, np.float32).reshape(self.count, -1),
                                                  np.asarray(value, np.float32)[:, None],
                                                  np.asarray(action, np.int32).view(np.float32)[:, None]], axis=1)).to(dev)
        full = torch.empty(self.world * self.count, self.A + 2, dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(full, packed, group=self.group)
        full = full.cpu().numpy()
        return (np.ascontiguousarray(full[:, self.A + 1]).view(np.int32), full[:, :self.A].copy(), full[:, self.A].copy())

    def _setup_peer(self, dev, row):
        """Two symmetric gather buffers (alternating per act, so that a fast rank's stores of act t+1 never land in a
        buffer a slow rank still reads from act t) + the byte offsets from this rank's buffer to each peer's."""
        engine = getattr(self.search_fn, "__self__", None)
        if self.peer_stores is False or engine is None or not hasattr(engine, "set_peer_outputs") or self.world > 8:
            return False
        try:
            import torch.distributed._symmetric_memory as symm
            group = self.group if self.group is not None else dist.group.WORLD
            bufs, hdls, deltas = [], [], []
            for _ in range(2):
                buf = symm.empty(self.world * row, dtype=torch.float32, device=dev)
                hdl = symm.rendezvous(buf, group)
                ptrs = [int(p) for p in hdl.buffer_ptrs]
                bufs.append(buf)
                hdls.append(hdl)
                deltas.append([ptrs[q] - ptrs[self.rank] for q in range(self.world) if q != self.rank])
            self._peer = (engine, bufs, hdls, deltas)
            self.exchange = "peer stores from the search kernel + symmetric-memory barrier"
            return True
        except Exception as e:  # no P2P / symmetric memory on this system: keep the NCCL path, say so
            if self.peer_stores is True:
                raise
            self.exchange = f"nccl all-gather (peer stores unavailable: {type(e).__name__})"
            self.peer_stores = False
            return False

    def _resolve_engine(self):
        e = self._engine() if callable(self._engine) else self._engine
        return e if e is not None else getattr(self.search_fn, "__self__", None)

    def _setup_peer_async(self, dev, row):
        """The overlapped exchange without NCCL and without a barrier kernel: `peer_slots` symmetric gather buffers of
        W rows + W completion flags each.  The search kernel stores its rows into every rank's buffer (NVLink peer
        stores) and its last CTA sets this rank's flag everywhere; a consumer waits for W flags (mz_peer_wait).
        Four slots: when rank A issues act s + 4 into slot s % 4 it has waited for act s + 2 of every peer, whose
        stream-ordered consumption of act s (after its own wait, before its act s + 2) is therefore over."""
        engine = self._resolve_engine()
        if engine is None or not hasattr(engine, "set_peer_flags") or self.world > 8:
            if self.peer_stores is True:
                raise RuntimeError("peer stores need the SearchEngine behind search_fn (engine=)")
            self.peer_stores = False
            return False
        try:
            import torch.distributed._symmetric_memory as symm
            group = self.group if self.group is not None else dist.group.WORLD
            W = self.world
            bufs, flags, deltas = [], [], []
            for _ in range(self.peer_slots):
                buf = symm.empty(W * row + W, dtype=torch.float32, device=dev)
                hdl = symm.rendezvous(buf, group)
                ptrs = [int(p) for p in hdl.buffer_ptrs]
                buf.zero_()
                bufs.append(buf)
                flags.append(buf[W * row:].view(torch.int32))
                deltas.append([ptrs[q] - ptrs[self.rank] for q in range(W) if q != self.rank])
            torch.cuda.current_stream().synchronize()
            dist.barrier(group=self.group)  # every rank's flags are zero before anybody's kernel may set them
            self._peer_async = (engine, bufs, flags)
            self._peer_async_deltas = deltas
            self.exchange = "peer stores + completion flags from the search kernel (no NCCL, no barrier kernel)"
            return True
        except Exception as e:  # no P2P / symmetric memory on this system: keep the NCCL path, say so
            if self.peer_stores is True:
                raise
            self.exchange = f"nccl all-gather on a side stream (peer stores unavailable: {type(e).__name__})"
            self.peer_stores = False
            return False

    def _act_async_peer(self, rng_key, obs_local, row, **kw):
        n, A = self.count, self.A
        engine, bufs, flags = self._peer_async
        slot = self._step % self.peer_slots
        self._step += 1
        step_no = self._step  # 1, 2, ...: the flags count acts
        mine = bufs[slot][self.rank * row:(self.rank + 1) * row]
        out = (mine[n * A + n:].view(torch.int32), mine[:n * A].view(n, A), mine[n * A:n * A + n])
        engine.set_peer_outputs(self._peer_async_deltas[slot])
        engine.set_peer_flags(flags[slot], self.rank, step_no)
        try:
            self.search_fn(rng_key, obs_local, global_batch=self.global_batch, batch_offset=self.offset, out=out, **kw)
        except RuntimeError as e:
            if self.peer_stores is True or "warp engine only" not in str(e):
                raise
            # this configuration runs on another engine: back to NCCL for good
            self.peer_stores, self._peer_async = False, None
            self._step -= 1
            return self.act_async(rng_key, obs_local, **kw)
        finally:
            engine.set_peer_outputs([])
            engine.set_peer_flags(None)
        return GatherHandle(self, slot, out, peer_step=step_no)

    def _act_peer(self, rng_key, obs_local, row, **kw):
        n, A, W = self.count, self.A, self.world
        engine, bufs, hdls, deltas = self._peer
        k = self._step & 1
        self._step += 1
        mine = bufs[k][self.rank * row:(self.rank + 1) * row]
        out = (mine[n * A + n:].view(torch.int32), mine[:n * A].view(n, A), mine[n * A:n * A + n])
        engine.set_peer_outputs(deltas[k])
        try:
            self.search_fn(rng_key, obs_local, global_batch=self.global_batch, batch_offset=self.offset, out=out, **kw)
        finally:
            engine.set_peer_outputs([])
        hdls[k].barrier(channel=0)  # every rank's kernel (and its NVLink stores) is complete past this point
        r = bufs[k].view(W, row)
        weights = r[:, :n * A].reshape(W * n, A)
        value = r[:, n * A:n * A + n].reshape(W * n)
        action = r[:, n * A + n:].view(torch.int32).reshape(W * n)
        return action, weights, value

    def _act_flat(self, rng_key, obs_local, **kw):
        n, A, W = self.count, self.A, self.world
        row = n * (A + 2)  # floats per rank: weights [n, A] | value [n] | action [n] (int32 bits)
        dev = obs_local.device
        if self.peer_stores is not False and dev.type == "cuda":
            if self._peer is not None or self._setup_peer(dev, row):
                try:
                    return self._act_peer(rng_key, obs_local, row, **kw)
                except RuntimeError as e:
                    if self.peer_stores is True or "warp engine only" not in str(e):
                        raise
                    self.peer_stores, self._peer = False, None  # this configuration runs on another engine
                    self.exchange = "nccl all-gather (the configuration does not run on the warp engine)"
        if self._send is None or self._send.device != dev:
            self._send = torch.empty(row, dtype=torch.float32, device=dev)
            self._recv = torch.empty(W * row, dtype=torch.float32, device=dev)
        send = self._send
        out = (send[n * A + n:].view(torch.int32), send[:n * A].view(n, A), send[n * A:n * A + n])
        self.search_fn(rng_key, obs_local, global_batch=self.global_batch, batch_offset=self.offset, out=out, **kw)
        dist.all_gather_into_tensor(self._recv, send, group=self.group)
        r = self._recv.view(W, row)
        weights = r[:, :n * A].reshape(W * n, A)
        value = r[:, n * A:n * A + n].reshape(W * n)
        action = r[:, n * A + n:].view(torch.int32).reshape(W * n)
        return action, weights, value
