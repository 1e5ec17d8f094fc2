"""CPU: self-consistency of the (unpinned) NAT and mel restatements."""
import numpy as np
import torch

from oracle import mel_oracle as mo
from oracle import nat_oracle as no
from viettts_b200 import synthetic


def test_mel_filterbank_vs_torchaudio():
    import torchaudio.functional as AF
    fb = mo.librosa_mel_filterbank()
    ref = AF.melscale_fbanks(513, 0.0, 8000.0, 80, 16000, norm="slaney", mel_scale="slaney").T.numpy()
    assert fb.shape == (80, 513)
    assert np.abs(fb - ref).max() < 1e-6


def test_mel_filter_vs_torch_stft():
    rng = np.random.default_rng(0)
    y = (rng.standard_normal((2, 4096)) * 0.1).astype(np.float32)
    got = mo.mel_filter(y)
    assert got.shape == (2, 16, 80)
    yt = torch.from_numpy(y)
    yp = torch.nn.functional.pad(yt[:, None], (384, 384), mode="reflect")[:, 0]
    st = torch.stft(yp, 1024, 256, 1024, torch.hann_window(1024, periodic=True), center=False, return_complex=True)
    mag = torch.sqrt(st.real**2 + st.imag**2 + 1e-9)
    mel = torch.einsum("ms,bsf->bfm", torch.from_numpy(mo.librosa_mel_filterbank()), mag)
    ref = torch.log(torch.clip(mel, 1e-5)).numpy()
    assert np.abs(got - ref).max() < 2e-4
    got64 = mo.mel_filter(y, dtype=np.float64)
    assert np.abs(got - got64).max() < 2e-4


def test_nat_shapes_and_determinism(acoustic_ckpt):
    tokens, dur = synthetic.utterance(0, n_phonemes=20, seconds=0.5)
    d, n = no.seconds_to_frames(dur)
    assert n == 31
    masks = synthetic.dropout_masks(42, 1, n)
    taps = {}
    mel = no.inference(acoustic_ckpt, np.asarray(tokens)[None], d, n, masks, taps=taps).numpy()
    assert mel.shape == (1, n, 80)
    assert taps["enc"].shape == (1, 20, 512) and taps["cond"].shape == (1, n, 512)
    assert np.allclose(taps["attn"].sum(-1).numpy(), 1.0, atol=1e-5)
    mel64 = no.inference(acoustic_ckpt, np.asarray(tokens)[None], d, n, masks, dtype=torch.float64).numpy()
    assert np.abs(mel - mel64).max() < 5e-4
    assert 0.05 < np.sqrt(np.mean((mel64 - mel64.mean()) ** 2)) < 20
    # masks matter (dropout is live at inference, model.py:132)
    mel_nomask = no.inference(acoustic_ckpt, np.asarray(tokens)[None], d, n, None).numpy()
    assert np.abs(mel - mel_nomask).max() > 1e-2


def test_nat_ragged_definition(acoustic_ckpt):
    # row semantics: batch row == single-utterance run
    t0, d0 = synthetic.utterance(1, 12, 0.3)
    t1, d1 = synthetic.utterance(2, 17, 0.4)
    f0, _ = no.seconds_to_frames(d0)
    f1, _ = no.seconds_to_frames(d1)
    outs = no.inference_ragged(acoustic_ckpt, [t0, t1], [f0[0], f1[0]])
    single = no.predict_mel(acoustic_ckpt, t1, d1)
    assert np.array_equal(outs[1], single)
