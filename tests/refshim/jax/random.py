"""jax.random.{PRNGKey, split, bernoulli, uniform} on the threefry2x32 generator, restated from JAX's published
implementation (jax/_src/prng.py, the classic layout with jax_threefry_partitionable=False, the default up to
jax 0.4.x):

  threefry_2x32(key, counts): counts (padded to even length) are cut into two halves x0 | x1, the block cipher maps
      every pair (x0[i], x1[i]) -> (y0[i], y1[i]), the result is concat(y0, y1)
  split(key, n)      = threefry_2x32(key, arange(2n)).reshape(n, 2)
  random_bits(key,s) = threefry_2x32(key, arange(prod(s))).reshape(s)
  uniform(key, s)    = bitcast_f32((bits >> 9) | 0x3F800000) - 1.0
  bernoulli(key,p,s) = uniform(key, s) < p

Pinned in tests/test_refshim_rng.py against the Random123 known-answer vectors of Threefry-2x32-20 and against the
outputs JAX's documentation prints for PRNGKey(0).  A recorder hook lets the golden generator capture every mask the
reference code draws.
"""
import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
RECORD = None   # set to a list to capture (shape, p, mask) of every bernoulli draw


def _rotl(x, r):
    return ((x << np.uint32(r)) | (x >> np.uint32(32 - r))).astype(np.uint32)


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds.  k0,k1: uint32 scalars; x0,x1: uint32 arrays."""
    with np.errstate(over="ignore"):
        k0 = np.uint32(k0)
        k1 = np.uint32(k1)
        ks = (k0, k1, np.uint32(0x1BD11BDA) ^ k0 ^ k1)
        x0 = (np.asarray(x0, np.uint32) + ks[0]).astype(np.uint32)
        x1 = (np.asarray(x1, np.uint32) + ks[1]).astype(np.uint32)
        for blk in range(5):
            for r in _ROT[blk & 1]:
                x0 = (x0 + x1).astype(np.uint32)
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = (x0 + ks[(blk + 1) % 3]).astype(np.uint32)
            x1 = (x1 + ks[(blk + 2) % 3] + np.uint32(blk + 1)).astype(np.uint32)
    return x0, x1


def threefry_2x32(key, counts):
    counts = np.asarray(counts, np.uint32).ravel()
    odd = counts.size % 2
    if odd:
        counts = np.concatenate([counts, np.zeros(1, np.uint32)])
    h = counts.size // 2
    y0, y1 = threefry2x32(key[0], key[1], counts[:h], counts[h:])
    out = np.concatenate([y0, y1])
    return out[:-1] if odd else out


def PRNGKey(seed):
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], np.uint32)


def split(key, num=2):
    key = np.asarray(key, np.uint32)
    return threefry_2x32(key, np.arange(2 * num, dtype=np.uint32)).reshape(num, 2)


def random_bits(key, shape):
    n = int(np.prod(shape)) if len(shape) else 1
    return threefry_2x32(np.asarray(key, np.uint32), np.arange(n, dtype=np.uint32)).reshape(shape)


def uniform(key, shape=(), dtype=np.float32):
    bits = random_bits(key, tuple(shape))
    f = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    return np.maximum(np.float32(0.0), f)


def bernoulli(key, p=0.5, shape=None):
    m = uniform(key, tuple(shape)) < np.float32(p)
    if RECORD is not None:
        RECORD.append((tuple(shape), float(p), m.copy()))
    return m
