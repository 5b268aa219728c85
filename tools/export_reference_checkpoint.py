#!/usr/bin/env python
"""Run INSIDE the reference's JAX environment: rewrites a `MuZero.save_load` checkpoint (muax/model.py:203-207, a
pickled .npy whose leaves are jax.Arrays) with NumPy leaves, so that muax_b200.checkpoint.load_reference_checkpoint /
`muax_b200.MuZero.load` can open it without jax.

    python tools/export_reference_checkpoint.py model_params.npy model_params_numpy.npy

The optimiser state is kept (optax namedtuples with NumPy leaves); pass --no-optimizer to drop it."""
import sys

import numpy as np


def main():
    import jax
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if len(args) != 2:
        raise SystemExit(__doc__)
    saved = np.load(args[0], allow_pickle=True).item()
    to_numpy = lambda tree: jax.tree_util.tree_map(lambda x: np.asarray(x), tree)  # noqa: E731
    out = {"params": to_numpy(saved["params"]),
           "optimizer_state": None if "--no-optimizer" in sys.argv else to_numpy(saved.get("optimizer_state"))}
    np.save(args[1], np.array(out, dtype=object), allow_pickle=True)
    print("wrote", args[1])


if __name__ == "__main__":
    main()
