"""Multi-GPU check of the sharded act (2, 4 or 8 GPUs — whatever the box has; skipped on one): the overlapped exchange
(`act_async`: NCCL all-gather on a side stream), and the search kernel's own NVLink peer stores +
symmetric-memory barrier (`ShardedSearch(peer_stores=True)`) must deliver exactly what the NCCL all-gather path
delivers, act after act, and both must equal the single-GPU search of the whole batch (PRNG draws are indexed by
global row)."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from helpers import make_nets
    from muax_b200.nn import pack_stacks
    from muax_b200.search import SearchEngine
    from muax_b200.sharded import ShardedSearch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    GB, NS, A = 192, 24, 2
    n = GB // world
    rng = np.random.default_rng(11)
    nets = make_nets(rng, 4, 8, A, 21)
    obs = rng.standard_normal((GB, 4)).astype(np.float32)
    blob, cstacks = pack_stacks(nets)

    def engine(batch):
        e = SearchEngine(cstacks, batch=batch, num_actions=A, embed_dim=8, obs_dim=4, support_size=10,
                         max_num_simulations=NS, device=torch.device("cuda", rank))
        e.set_weights(blob)
        return e

    eng = engine(n)
    obs_local = torch.from_numpy(obs[rank * n:(rank + 1) * n]).cuda()
    res = {}
    for name, peer in (("peer", True), ("nccl", False)):
        sh = ShardedSearch(eng.search, GB, A, writes_into_out=True, peer_stores=peer)
        outs = []
        for step in range(5):  # several acts back to back: the two gather buffers alternate
            a, w, v = sh.act(np.array([7, step], np.uint32), obs_local, num_simulations=NS)
            outs.append((a.clone(), w.clone(), v.clone()))
        torch.cuda.synchronize()
        res[name] = [[t.cpu().numpy() for t in o] for o in outs]
        res[name + "_exchange"] = sh.exchange
    # the overlapped exchange: act t + 1 is issued before act t's all-gather (side stream) is waited for
    sh = ShardedSearch(eng.search, GB, A, writes_into_out=True)
    handles, outs = [], []
    for step in range(5):
        handles.append(sh.act_async(np.array([7, step], np.uint32), obs_local, num_simulations=NS))
        if len(handles) == 2:
            outs.append([t.clone() for t in handles.pop(0).wait()])
    outs.append([t.clone() for t in handles.pop(0).wait()])
    torch.cuda.synchronize()
    res["async"] = [[t.cpu().numpy() for t in o] for o in outs]
    res["async_exchange"] = sh.exchange
    # the same overlap with the kernel's own peer stores + completion flags (no NCCL, no barrier kernel); more acts
    # than gather slots, so that the slots are reused
    sh = ShardedSearch(eng.search, GB, A, writes_into_out=True, peer_stores=True)
    handles, outs = [], []
    for step in range(9):
        handles.append(sh.act_async(np.array([7, step % 5], np.uint32), obs_local, num_simulations=NS))
        if len(handles) == 2:
            outs.append([t.clone() for t in handles.pop(0).wait()])
    outs.append([t.clone() for t in handles.pop(0).wait()])
    torch.cuda.synchronize()
    res["async_peer"] = [[t.cpu().numpy() for t in o] for o in outs]
    res["async_peer_exchange"] = sh.exchange
    # the end-to-end host path: NumPy rows in, everybody's results out (search + all-gather + one D2H)
    obs_local_np = obs[rank * n:(rank + 1) * n]
    res["host"] = [list(sh.act_host(np.array([7, step], np.uint32), obs_local_np, num_simulations=NS)) for step in range(5)]
    if rank == 0:
        full = engine(GB)
        ref = []
        for step in range(5):
            a, w, v = full.search(np.array([7, step], np.uint32), obs=torch.from_numpy(obs).cuda(), num_simulations=NS)
            torch.cuda.synchronize()
            ref.append([t.cpu().numpy() for t in (a, w, v)])
        ok = all(np.array_equal(x, y) and np.array_equal(x, z) and np.array_equal(x, u) and np.array_equal(x, h)
                 for p, q, r, s, hh in zip(res["peer"], res["nccl"], ref, res["async"], res["host"])
                 for x, y, z, u, h in zip(p, q, r, s, hh))
        ok = ok and len(res["async_peer"]) == 9 and all(
            np.array_equal(x, y) for i, got in enumerate(res["async_peer"]) for x, y in zip(got, ref[i % 5]))
        with open(os.path.join(out_dir, "result.txt"), "w") as f:
            f.write(f"{int(ok)}|{res['peer_exchange']}|{res['nccl_exchange']}|{res['async_exchange']}|"
                    f"{res['async_peer_exchange']}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_peer_stores_equal_nccl_all_gather_and_the_single_gpu_search(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    world = 8 if torch.cuda.device_count() >= 8 else (4 if torch.cuda.device_count() >= 4 else 2)
    mp.spawn(_worker, args=(world, 29533, str(tmp_path)), nprocs=world, join=True)
    ok, peer_exchange, nccl_exchange, async_exchange, async_peer = open(tmp_path / "result.txt").read().split("|")
    assert peer_exchange.startswith("peer stores"), peer_exchange
    assert nccl_exchange.startswith("nccl"), nccl_exchange
    assert "side stream" in async_exchange, async_exchange
    assert "completion flags" in async_peer, async_peer
    assert ok == "1"
