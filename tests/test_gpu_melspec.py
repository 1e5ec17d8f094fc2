"""GPU parity: STFT/log-mel kernel vs the CPU restatement of MelFilter (dsp.py:104-128).
Tolerance: log-mel L-inf <= 5e-4 against the float64 arbiter (fp32 FFT noise near the 1e-5 clip)."""
import numpy as np
import pytest

from oracle import mel_oracle as mo

pytestmark = pytest.mark.gpu
TOL = 5e-4


@pytest.fixture(scope="module")
def eng():
    from viettts_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("B,S", [(1, 512), (2, 4096), (3, 256 * 7), (1, 79872)])
def test_melspec_vs_oracle(eng, B, S):
    rng = np.random.default_rng(S)
    y = (rng.standard_normal((B, S)) * 0.1).astype(np.float32)
    y[0, : S // 3] = 0.0   # a silent span exercises the 1e-9 / 1e-5 floors (dsp.py:125-127)
    got = eng.melspec(y)
    ref64 = mo.mel_filter(y, dtype=np.float64)
    assert got.shape == ref64.shape == (B, S // 256, 80)
    err = np.abs(got - ref64).max()
    print(f"B={B} S={S} Linf={err:.3e}")
    assert err < TOL


def test_melfilter_dropin_class(eng):
    from viettts_b200.nat.dsp import MelFilter
    rng = np.random.default_rng(1)
    y = (rng.standard_normal((2, 2048)) * 0.3).astype(np.float32)
    mf = MelFilter(16000, 1024, 80, fmin=0.0, fmax=8000, engine=eng)
    assert np.abs(mf(y) - mo.mel_filter(y, dtype=np.float64)).max() < TOL
    with pytest.raises(AssertionError):
        mf(y[0])


def test_large_batch_linearity_property(eng):
    """Full-size property (B=32 x 5 s): magnitude spectra are homogeneous, so scaling the
    waveform by 2 shifts every unclipped log-mel bin by log 2."""
    rng = np.random.default_rng(2)
    y = (rng.standard_normal((32, 79872)) * 0.05).astype(np.float32)
    a = eng.melspec(y)
    b = eng.melspec(2 * y)
    assert np.isfinite(a).all()
    assert np.abs((b - a) - np.log(2.0)).max() < 1e-3
