"""TEST INFRASTRUCTURE — CPU restatement of the `jax.random` pieces the mctx search consumes.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import this
module; the product (muax_b200/) never does.

Upstream: `jax._src.prng` / `jax._src.random` (jax is an unpinned, un-vendored dependency of
bwfbowen/muax: setup.py:13-20; not importable in this image).  Restated from SURVEY.md Appendix A.7.
Call sites in the reference that fix which functions matter: muax/train.py:138,154,184
(PRNGKey / split), and — inside mctx — `split`, `uniform`, `gumbel`, `categorical` (Appendix A.2-A.5).

Two layouts exist upstream (`jax_threefry_partitionable`; False before JAX 0.5.0, True after):
mode 0 = "legacy", mode 1 = "partitionable".  Both are restated; legacy is the default because the
reference's dependency window (Nov 2024, jax 0.4.x) predates the flip.

Parity status: pinned only by the Random123 threefry2x32 known-answer vectors and the two widely
quoted JAX values split(PRNGKey(0)) / uniform(PRNGKey(0)) (tests/test_threefry.py).
"""
import numpy as np

LEGACY = 0
PARTITIONABLE = 1

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return ((x << np.uint32(r)) | (x >> np.uint32(32 - r))).astype(np.uint32)


def threefry2x32(k0, k1, c0, c1):
    """Threefry-2x32, 20 rounds.  All arguments broadcastable uint32 arrays; returns (x0, x1)."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, dtype=np.uint32)
        k1 = np.asarray(k1, dtype=np.uint32)
        x0 = np.asarray(c0, dtype=np.uint32)
        x1 = np.asarray(c1, dtype=np.uint32)
        ks = (k0, k1, (k0 ^ k1 ^ np.uint32(0x1BD11BDA)).astype(np.uint32))
        x0 = (x0 + ks[0]).astype(np.uint32)
        x1 = (x1 + ks[1]).astype(np.uint32)
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = (x0 + x1).astype(np.uint32)
                x1 = _rotl(x1, r)
                x1 = (x1 ^ x0).astype(np.uint32)
            x0 = (x0 + ks[(i + 1) % 3]).astype(np.uint32)
            x1 = (x1 + ks[(i + 2) % 3] + np.uint32(i + 1)).astype(np.uint32)
    return x0, x1


def PRNGKey(seed):
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def random_bits(key, n, mode=LEGACY):
    """32-bit draws for a flat vector of n elements. `key` is uint32[..., 2]; returns uint32[..., n]."""
    key = np.asarray(key, dtype=np.uint32)
    k0 = key[..., 0:1]
    k1 = key[..., 1:2]
    if n == 0:
        return np.zeros(key.shape[:-1] + (0,), dtype=np.uint32)
    if mode == LEGACY:
        n_pad = n + (n & 1)
        half = n_pad // 2
        counts = np.arange(n_pad, dtype=np.uint32)
        counts[n:] = 0  # odd sizes are padded with one zero counter
        y0, y1 = threefry2x32(k0, k1, counts[:half], counts[half:])
        return np.concatenate([y0, y1], axis=-1)[..., :n]
    y0, y1 = threefry2x32(k0, k1, np.uint32(0), np.arange(n, dtype=np.uint32))
    return (y0 ^ y1).astype(np.uint32)


def split(key, num=2, mode=LEGACY):
    """jax.random.split: key uint32[..., 2] -> uint32[..., num, 2]."""
    key = np.asarray(key, dtype=np.uint32)
    if mode == LEGACY:
        bits = random_bits(key, 2 * num, LEGACY)
        return bits.reshape(key.shape[:-1] + (num, 2))
    y0, y1 = threefry2x32(key[..., 0:1], key[..., 1:2], np.uint32(0), np.arange(num, dtype=np.uint32))
    return np.stack([y0, y1], axis=-1)


def bits_to_unit(bits):
    f = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).astype(np.uint32).view(np.float32)
    return (f - np.float32(1.0)).astype(np.float32)


def uniform(key, n, mode=LEGACY):
    """jax.random.uniform(key, (n,)) with minval=0, maxval=1."""
    return np.maximum(np.float32(0.0), bits_to_unit(random_bits(key, n, mode)))


def uniform_tiny(key, n, mode=LEGACY):
    """uniform(minval=tiny, maxval=1): the argument of the double log in jax.random.gumbel."""
    tiny = np.finfo(np.float32).tiny
    u = bits_to_unit(random_bits(key, n, mode))
    scale = np.float32(np.float32(1.0) - tiny)
    return np.maximum(tiny, (u * scale + tiny).astype(np.float32))
