"""Pins the shim's threefry / split / uniform to published outputs, so that the dropout and zoneout masks recorded in
tests/golden/nat_ref_*.npz are the masks JAX itself would draw from the checkpoint's rng (classic threefry layout,
jax_threefry_partitionable=False):

  * Threefry-2x32-20 known-answer vectors of the Random123 distribution (the same three JAX's own test-suite uses);
  * the values JAX's documentation prints for `random.split(random.PRNGKey(0))` and `random.uniform(random.PRNGKey(0))`.
"""
import sys
from pathlib import Path

import numpy as np

SHIM = str(Path(__file__).resolve().parent / "refshim")


def _rand():
    sys.path.insert(0, SHIM)
    try:
        for m in [k for k in sys.modules if k == "jax" or k.startswith("jax.")]:
            del sys.modules[m]
        import jax
        return jax.random
    finally:
        sys.path.remove(SHIM)
        for m in [k for k in sys.modules if k == "jax" or k.startswith("jax.")]:
            del sys.modules[m]


def test_threefry_known_answers():
    r = _rand()
    kat = [((0, 0), (0, 0), (0x6B200159, 0x99BA4EFE)),
           ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
           ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]
    for key, ctr, out in kat:
        y0, y1 = r.threefry2x32(key[0], key[1], [ctr[0]], [ctr[1]])
        assert (int(y0[0]), int(y1[0])) == out


def test_split_and_uniform_match_jax_documentation():
    r = _rand()
    k = r.PRNGKey(0)
    assert k.tolist() == [0, 0]
    assert r.split(k).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert abs(float(r.uniform(k, ())) - 0.41845703) < 1e-8


def test_product_stream_matches_shim():
    """viettts_b200.jaxrng (host side of the JAX-compatible dropout mode) draws the same chain as the shim running the
    reference: hk.next_rng_key() per draw, bernoulli over [B,256] with the counter halves layout."""
    r = _rand()
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    from viettts_b200 import jaxrng
    key = np.array([0, 42], np.uint32)
    subs = jaxrng.subkey_chain(key, 6)
    k = key.copy()
    for i in range(6):
        nk = r.split(k, 2)
        k = nk[0]
        assert subs[i].tolist() == nk[1].tolist()
        m = r.bernoulli(nk[1], 0.5, (2, 256))
        assert np.array_equal(jaxrng.bernoulli_bits(subs[i], 2 * 256).reshape(2, 256) < 0x80000000, m)
