"""Known-answer tests for the PRNG restatement: Random123 threefry2x32 vectors and
the widely quoted jax.random values for PRNGKey(0)."""
import numpy as np

from oracle import threefry as tf


def test_random123_kats():
    assert [int(x) for x in tf.threefry2x32(0, 0, 0, 0)] == [0x6B200159, 0x99BA4EFE]
    ones = 0xFFFFFFFF
    assert [int(x) for x in tf.threefry2x32(ones, ones, ones, ones)] == [0x1CB996FC, 0xBB002BE7]
    assert [int(x) for x in tf.threefry2x32(0x13198A2E, 0x03707344, 0x243F6A88, 0x85A308D3)] == [0xC4923A9C, 0x483DF7A0]


def test_jax_split_and_uniform_values():
    k = tf.PRNGKey(0)
    assert tf.split(k, 2).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert tf.split(k, 2, tf.PARTITIONABLE).tolist() == [[1797259609, 2579123966], [928981903, 3453687069]]
    assert abs(float(tf.uniform(k, 1)[0]) - 0.41845703) < 1e-8
    assert tf.PRNGKey(42).tolist() == [0, 42]
    assert tf.PRNGKey((7 << 32) | 5).tolist() == [7, 5]


def test_split_shapes_and_odd_sizes():
    k = tf.PRNGKey(3)
    ks = tf.split(k, 5)
    assert ks.shape == (5, 2)
    batched = tf.split(ks, 2)
    assert batched.shape == (5, 2, 2)
    for i in range(5):
        assert np.array_equal(batched[i], tf.split(ks[i], 2))
    for mode in (tf.LEGACY, tf.PARTITIONABLE):
        u3 = tf.uniform(ks, 3, mode)
        assert u3.shape == (5, 3) and (u3 >= 0).all() and (u3 < 1).all()
        assert np.array_equal(u3[2], tf.uniform(ks[2], 3, mode))


def test_c_threefry_matches_numpy(c_oracle):
    rng = np.random.default_rng(0)
    c0 = rng.integers(0, 2**32, 1000, dtype=np.uint32)
    c1 = rng.integers(0, 2**32, 1000, dtype=np.uint32)
    a0, a1 = c_oracle.threefry(123, 456, c0, c1)
    b0, b1 = tf.threefry2x32(123, 456, c0, c1)
    assert np.array_equal(a0, b0) and np.array_equal(a1, b1)
