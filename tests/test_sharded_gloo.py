"""world_size-2 `gloo` test of the multi-GPU host logic: row partitioning, global-index PRNG and
the single all-gather of (action_weights, root_value, action).  No GPU here, so each rank's search is played by
the C restatement (tests may use the checker); the gathered result must equal the single-process search of
the whole batch bit for bit, for even and ragged shards."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import make_nets


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, global_batch, policy, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from muax_b200.sharded import ShardedSearch
        from oracle import c_oracle
        rng = np.random.default_rng(3)
        nets = make_nets(rng, 4, 8, 3, 21, bias_scale=0.05)
        obs = rng.standard_normal((global_batch, 4)).astype(np.float32)
        key = np.array([0, 11], np.uint32)

        def search_fn(rng_key, obs_local, global_batch, batch_offset, out=None, **kw):
            r = c_oracle.search(nets, rng_key, obs=obs_local.numpy(), policy=policy, qtransform=policy,
                                num_simulations=16, global_batch=global_batch, batch_offset=batch_offset,
                                want_tree=False, nthreads=1)
            res = (torch.from_numpy(r["action"]), torch.from_numpy(r["action_weights"]),
                   torch.from_numpy(r["root_value"]))
            if out is not None:  # the engine's `out=` contract: results land in the caller's (flat-buffer) views
                for dst, src in zip(out, res):
                    dst.copy_(src)
                return out
            return res

        # even shards take the flat path (outputs written straight into the all-gather send buffer)
        sh = ShardedSearch(search_fn, global_batch, 3, writes_into_out=True)
        a, w, v = sh.act(key, sh.local_rows(torch.from_numpy(obs)))
        extra = {}
        if global_batch % world == 0:
            # the overlapped exchange (act_async): two acts in flight with alternating buffers, then the host-side
            # gather the end-to-end path uses — all must reproduce the synchronous result
            key2 = np.array([0, 12], np.uint32)
            h1 = sh.act_async(key, sh.local_rows(torch.from_numpy(obs)))
            h2 = sh.act_async(key2, sh.local_rows(torch.from_numpy(obs)))
            a1, w1, v1 = (t.clone() for t in h1.wait())
            a2, w2, v2 = (t.clone() for t in h2.wait())
            assert torch.equal(a1, a) and torch.equal(w1, w) and torch.equal(v1, v)
            la, lw, lv = h2.local
            ga, gw, gv = sh.gather_host(la.numpy(), lw.numpy(), lv.numpy())
            assert np.array_equal(ga, a2.numpy()) and np.array_equal(gw, w2.numpy()) and np.array_equal(gv, v2.numpy())
            extra = dict(a2=a2.numpy(), w2=w2.numpy())
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), a=a.numpy(), w=w.numpy(), v=v.numpy(),
                 offset=sh.offset, count=sh.count, **extra)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("global_batch,policy", [(10, 0), (7, 0), (9, 1)])
def test_two_rank_sharded_search_equals_single_process(tmp_path, global_batch, policy):
    from oracle import c_oracle
    c_oracle.build()
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), global_batch, policy, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(3)
    nets = make_nets(rng, 4, 8, 3, 21, bias_scale=0.05)
    obs = rng.standard_normal((global_batch, 4)).astype(np.float32)
    want = c_oracle.search(nets, np.array([0, 11], np.uint32), obs=obs, policy=policy, qtransform=policy,
                           num_simulations=16, want_tree=False)
    counts = 0
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(z["a"], want["action"])
        assert np.array_equal(z["w"], want["action_weights"])
        assert np.array_equal(z["v"], want["root_value"])
        counts += int(z["count"])
    assert counts == global_batch


def test_shard_bounds_cover_the_batch():
    from muax_b200.sharded import pack_outputs, shard_bounds, unpack_outputs
    for gb in (1, 7, 4096, 8191):
        for world in (1, 2, 3, 8):
            rows = []
            for r in range(world):
                off, cnt = shard_bounds(gb, world, r)
                rows += list(range(off, off + cnt))
            assert rows == list(range(gb))
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)
    a = torch.tensor([3, 0, 17], dtype=torch.int32)
    w = torch.rand(3, 18)
    v = torch.randn(3)
    a2, w2, v2 = unpack_outputs(pack_outputs(a, w, v))
    assert torch.equal(a, a2) and torch.equal(w, w2) and torch.equal(v, v2)
