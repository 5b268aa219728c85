"""Times `mz_recurrent` (muax/model.py:265-282 for a batch) on its own: the tcgen05 kernel (bf16) against the fp32 SIMT
kernel, per BASELINE shape.  CUDA events around `reps` back-to-back calls (a memset + one kernel each)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import make_nets  # noqa: E402

from muax_b200.nn import pack_stacks  # noqa: E402
from muax_b200.search import SearchEngine  # noqa: E402

SHAPES = {
    "lunar_e64_h16_b4096": (8, 64, 4, 10, (16,), 1, 4096),
    "notebook_e64_h64x64x16_b4096": (8, 64, 4, 20, (64, 64, 16), 0, 4096),
    "atari_e256_h256_b1024": (256, 256, 18, 10, (256,), 1, 1024),
    "atari_e256_h256_b8192": (256, 256, 18, 10, (256,), 1, 8192),
}


def flops(E, A, F, hidden):
    def macs(i, o):
        dims = [i, *hidden, o]
        return sum(a * b for a, b in zip(dims[:-1], dims[1:]))
    return 2 * (macs(E + A, E) + macs(E + A, F) + macs(E, F) + macs(E, A))


def main(reps=200):
    out = {}
    for name, (obs_dim, E, A, S, hidden, minmax, B) in SHAPES.items():
        rng = np.random.default_rng(0)
        nets = make_nets(rng, obs_dim, E, A, 2 * S + 1, hidden=hidden)
        blob, cstacks = pack_stacks(nets)
        eng = SearchEngine(cstacks, batch=B, num_actions=A, embed_dim=E, obs_dim=obs_dim, support_size=S,
                           max_num_simulations=4, repr_minmax=minmax, dyn_minmax=minmax)
        eng.set_weights(blob)
        emb = torch.rand(B, E, device="cuda")
        action = torch.randint(0, A, (B,), device="cuda", dtype=torch.int32)
        row = {}
        for prec in ("bf16", "fp32"):
            for _ in range(5):
                eng.recurrent(action, emb, precision=prec)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                eng.recurrent(action, emb, precision=prec)
            e1.record()
            e1.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / reps
            row[prec] = {"us_per_call": us, "tflops": B * flops(E, A, 2 * S + 1, hidden) / us / 1e6}
        out[name] = row
        print(name, json.dumps(row), flush=True)
    return out


if __name__ == "__main__":
    main()
