"""CPU tests of the DurationModel restatement (oracle/nat_oracle.py; model.py:49-70) and of the blob packing.

Pinned to the reference's own source since round 2 (tests/test_reference_goldens.py::test_duration_matches_reference_source);
the earlier structural checks stay: the reference's own shape test
(tests/test_nat_duration.py:9-15), the dm-haiku / jax function definitions against independent torch
implementations, and structural properties of the restatement."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import nat_oracle as no
from viettts_b200 import _lib, synthetic, weights


@pytest.fixture(scope="module")
def duration_ckpt():
    return synthetic.duration_ckpt(1234)


def test_reference_shape_test(duration_ckpt):
    """tests/test_nat_duration.py: tokens zeros (2,10), lengths zeros -> one value per position."""
    out = no.duration_model(duration_ckpt, np.zeros((2, 10), np.int32), np.zeros(2, np.int32))
    assert out.shape == (2, 10)
    assert np.all(out > 0) and np.all(np.isfinite(out))      # softplus range
    np.testing.assert_allclose(out[0], out[1], rtol=0, atol=0)


def test_activation_definitions():
    x = torch.linspace(-12, 12, 4001, dtype=torch.float64)
    np.testing.assert_allclose(no.gelu_tanh(x).numpy(), F.gelu(x, approximate="tanh").numpy(), atol=1e-12)
    np.testing.assert_allclose(no.softplus(x).numpy(), F.softplus(x, beta=1, threshold=1e9).numpy(), atol=1e-12)
    big = torch.tensor([-800.0, 800.0], dtype=torch.float64)
    assert no.softplus(big).tolist() == [0.0, 800.0]        # logaddexp form: no overflow


def test_padding_rows_do_not_change_a_row(duration_ckpt):
    """lengths only move the backward core's reset point (model.py:37): a row evaluated alone with
    lengths=len equals the same row inside a wider batch on the first len positions only if nothing leaks from
    the padding -- which the reference does NOT guarantee (conv SAME padding reads the pad tokens).  The library's
    batch contract (row b == row b alone) is tested on the GPU; here we pin the reference semantics."""
    tk, _ = synthetic.utterance(5, 24, None)
    tk = np.asarray(tk, np.int32)
    alone = no.predict_duration(duration_ckpt, tk)
    assert alone.shape == (1, 24)
    padded = np.zeros((1, 30), np.int32)
    padded[0, :24] = tk
    wide = no.duration_model(duration_ckpt, padded, np.array([24], np.int32))
    assert np.abs(wide[0, :18] - alone[0, :18]).max() > 0     # the bwd core starts from a different state
    f32 = no.duration_model(duration_ckpt, tk[None], np.array([24], np.int32), dtype=torch.float32)
    f64 = no.duration_model(duration_ckpt, tk[None], np.array([24], np.int32), dtype=torch.float64)
    assert np.abs(f32 - f64).max() < 1e-5


def test_durations_are_plausible(duration_ckpt):
    tk, _ = synthetic.utterance(0, 100, None)
    d = no.predict_duration(duration_ckpt, tk)[0]
    assert 0.005 < d.min() and d.max() < 2.0 and 2.0 < d.sum() < 30.0


def test_blob_layout(duration_ckpt):
    blob = weights.pack_duration(duration_ckpt)
    assert blob.dtype == np.float32
    assert blob.size == synthetic.n_params(duration_ckpt["params"]) + 3 * 2 * 256   # + BN eval statistics
    lib = _lib.load()
    assert blob.size == lib.vtts_duration_blob_floats()
    # the encoder block is laid out exactly like the acoustic blob's encoder block
    ac = synthetic.acoustic_ckpt(1234)
    enc_floats = 256 * 256 + 3 * (3 * 256 * 256 + 5 * 256) + 2 * (512 * 1024 + 1024)
    assert blob.size - enc_floats == 512 * 256 + 256 + 256 + 1
    assert weights.pack_acoustic(ac)[:enc_floats].size == enc_floats
    with pytest.raises(KeyError):
        weights.pack_duration(dict(params={}, aux={}))
