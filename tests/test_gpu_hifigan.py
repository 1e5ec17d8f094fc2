"""GPU parity: HiFiGAN generator (CUDA, through the C ABI) vs the reference-pinned oracle.

Tolerances (fp32 path; differences are summation order only):
    waveform L-inf <= 1e-4, RMS <= 1e-5   (full scale 1.0)"""
import numpy as np
import pytest
import torch

from oracle import hifigan_oracle as ho
from viettts_b200 import synthetic

pytestmark = pytest.mark.gpu

WAV_LINF, WAV_RMS = 1e-4, 1e-5


@pytest.fixture(scope="module")
def eng(hifigan_params):
    from viettts_b200.engine import Engine
    e = Engine(0)
    e.load_hifigan(hifigan_params)
    e.set_precision("fp32")      # this file checks the strict fp32 path; test_gpu_tc_conv.py the bf16x3 one
    yield e
    e.close()


def _cmp(got, ref, tag=""):
    err = np.abs(got - ref)
    rms = float(np.sqrt(np.mean(err**2)))
    print(f"[{tag}] wav Linf={err.max():.3e} rms={rms:.3e}")
    assert err.max() <= WAV_LINF and rms <= WAV_RMS, (tag, err.max(), rms)


@pytest.mark.parametrize("tag", ["small", "t32"])
def test_golden_reference_vectors(eng, golden_dir, tag):
    g = np.load(golden_dir / f"hifigan_ref_{tag}.npz")
    wav = eng.mel2wave(g["mel"])
    assert wav.shape == g["wav"].shape
    _cmp(wav, g["wav"], f"golden-{tag}")


@pytest.mark.parametrize("T", [1, 2, 5, 33])
def test_short_inputs_vs_oracle(eng, hifigan_params, T):
    mel = synthetic.mel_input(T, 2, T)
    ref = ho.mel2wave(hifigan_params, mel).reshape(2, -1)
    _cmp(eng.mel2wave(mel), ref, f"T={T}")


def test_ragged_batch_rows_equal_single_runs(eng, hifigan_params):
    B, T = 3, 40
    nf = np.array([40, 23, 1], np.int32)
    mel = synthetic.mel_input(3, B, T)
    wav = eng.mel2wave(mel, n_frames=nf)
    for b in range(B):
        ref = ho.mel2wave(hifigan_params, mel[b : b + 1, : nf[b]]).reshape(-1)
        _cmp(wav[b, : nf[b] * 256], ref, f"ragged-row{b}")
        assert np.all(wav[b, nf[b] * 256 :] == 0.0)


def test_device_pointer_api_matches_host_api(eng):
    mel = synthetic.mel_input(5, 2, 21)
    host = eng.mel2wave(mel)
    t = torch.from_numpy(mel).cuda()
    out = eng.hifigan_forward(t)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), host)


def test_config2_full_size_vs_oracle(eng, hifigan_params):
    """BASELINE config 2: 80x400 mel, batch 1."""
    mel = synthetic.mel_input(0, 1, 400)
    ref = ho.mel2wave(hifigan_params, mel).reshape(1, -1)
    wav = eng.mel2wave(mel)
    assert wav.shape == (1, 102400)
    _cmp(wav, ref, "config2")


def test_full_batch_properties(eng, hifigan_params):
    """B=32 x 312 frames (config 3 size): batch rows must equal the same rows run alone
    (size-independent property: rows are independent), output finite and in (-1,1)."""
    mel = synthetic.mel_input(11, 32, 312)
    wav = eng.mel2wave(mel)
    assert wav.shape == (32, 79872) and np.isfinite(wav).all() and np.abs(wav).max() < 1.0
    for b in (0, 17, 31):
        alone = eng.mel2wave(mel[b : b + 1])
        assert np.array_equal(alone[0], wav[b])
    ref = ho.mel2wave(hifigan_params, mel[5:6]).reshape(-1)
    _cmp(wav[5], ref, "config3-row5")


def test_bad_arguments_raise(eng):
    from viettts_b200._lib import VttsError
    with pytest.raises(ValueError):
        eng.mel2wave(np.zeros((1, 4, 79), np.float32))
    from viettts_b200.engine import Engine
    e2 = Engine(0)
    with pytest.raises(VttsError):
        e2.mel2wave(np.zeros((1, 4, 80), np.float32))   # weights not loaded
    e2.close()


def test_long_utterance_vs_oracle(hifigan_params):
    """Maximum-size style case: 1 000 mel frames (16 s) in one row, both arithmetic paths."""
    from viettts_b200.engine import Engine
    e = Engine(0)
    e.load_hifigan(hifigan_params)
    mel = synthetic.mel_input(99, 1, 1000)
    ref = ho.mel2wave(hifigan_params, mel).reshape(1, -1)
    for mode in ("fp32", "bf16x3"):
        e.set_precision(mode)
        _cmp(e.mel2wave(mel), ref, f"T=1000 {mode}")
    e.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_streaming_chunks_equal_full_utterance(hifigan_params, precision):
    """Chunked vocoding with recomputed context (Engine.mel2wave_stream) reproduces the one-shot waveform:
    exact receptive-field bookkeeping, checked for chunk sizes that do and do not divide T.  A halo one frame
    short of the receptive field must NOT reproduce it (the bound is tight enough to matter)."""
    from viettts_b200.engine import Engine
    e = Engine(0)
    try:
        e.load_hifigan(hifigan_params)
        e.set_precision(precision)
        mel = synthetic.mel_input(21, 1, 150)
        full = e.mel2wave(mel)[0]
        for chunk in (32, 47, 150, 400):
            parts = list(e.mel2wave_stream(mel, chunk_frames=chunk))
            assert all(p.size == min(chunk, 150 - i * chunk) * 256 for i, p in enumerate(parts))
            got = np.concatenate(parts)
            assert got.shape == full.shape
            assert np.abs(got - full).max() <= 1e-6, (chunk, np.abs(got - full).max())
        short = np.concatenate(list(e.mel2wave_stream(mel[0], chunk_frames=32, halo=6)))
        assert np.abs(short - full).max() > 1e-5
    finally:
        e.close()
