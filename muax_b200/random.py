"""Host-side PRNG keys with jax.random's threefry semantics, so that user loops written as
`key, subkey = jax.random.split(key)` (muax/train.py:138,154,184) keep producing the same key stream
without JAX.  NumPy only; the device draws the rest (muax_b200/csrc/mz_device.cuh)."""
import numpy as np

LEGACY, PARTITIONABLE = 0, 1
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_M = 0xFFFFFFFF


def _threefry(k0, k1, c0, c1):
    ks = (k0, k1, k0 ^ k1 ^ 0x1BD11BDA)
    x0, x1 = (c0 + ks[0]) & _M, (c1 + ks[1]) & _M
    for i in range(5):
        for r in _ROT[i % 2]:
            x0 = (x0 + x1) & _M
            x1 = ((x1 << r) | (x1 >> (32 - r))) & _M
            x1 ^= x0
        x0 = (x0 + ks[(i + 1) % 3]) & _M
        x1 = (x1 + ks[(i + 2) % 3] + i + 1) & _M
    return x0, x1


def PRNGKey(seed):
    seed = int(seed)
    return np.array([(seed >> 32) & _M, seed & _M], dtype=np.uint32)


def split(key, num=2, mode=LEGACY):
    """jax.random.split(key, num) -> uint32[num, 2]."""
    k0, k1 = int(key[0]), int(key[1])
    out = np.zeros((num, 2), np.uint32)
    if mode == LEGACY:
        words = []
        for m in range(2 * num):
            i = m if m < num else m - num
            y0, y1 = _threefry(k0, k1, i, num + i)
            words.append(y0 if m < num else y1)
        out[:] = np.array(words, dtype=np.uint64).astype(np.uint32).reshape(num, 2)
    else:
        for j in range(num):
            out[j] = _threefry(k0, k1, 0, j)
    return out


def key_words(key):
    """Accepts a uint32[2] array (jax legacy key layout), a (hi, lo) tuple or an int seed."""
    if isinstance(key, (int, np.integer)):
        key = PRNGKey(key)
    if hasattr(key, "detach"):
        key = key.detach().cpu().numpy()
    key = np.asarray(key)
    if key.shape != (2,):
        raise ValueError(f"rng_key must have shape (2,), got {key.shape}")
    return int(key[0]) & _M, int(key[1]) & _M
