"""CPU checks of the teacher-forced / GTA restatement (oracle/nat_oracle.py; model.py:146-169, gta.py:28-41).
Since round 2 the restatement is pinned to the reference's own source (tests/test_reference_goldens.py).  The
structural pins that came first are kept: (i) the reference's own shape test (tests/test_nat_acoustic.py),
(ii) teacher forcing on the autoregressive path's own output reproduces it when dropout and zoneout are off --
the two restatements share no decoder code, (iii) zoneout semantics on hand-made masks."""
import numpy as np
import torch

from oracle import nat_oracle as no
from viettts_b200 import synthetic


def _setup(seed, L, seconds):
    tk, dur = synthetic.utterance(seed, L, seconds)
    d, n = no.seconds_to_frames(dur)
    return np.asarray(tk, np.int32)[None], d, n


def test_reference_shape_test(acoustic_ckpt):
    """tests/test_nat_acoustic.py: tokens zeros (2,10), mel zeros (2,20,*) -> two outputs shaped like the mel."""
    tok = np.zeros((2, 10), np.int32)
    m1, m2 = no.teacher_forced(acoustic_ckpt, tok, np.array([10, 10]), np.zeros((2, 10), np.float32), np.zeros((2, 20, 80), np.float32))
    assert m1.shape == m2.shape == (2, 20, 80) and np.isfinite(m2).all()


def test_teacher_forcing_reproduces_the_autoregressive_path(acoustic_ckpt):
    tok, d, n = _setup(0, 14, 0.4)
    P, S = acoustic_ckpt["params"], acoustic_ckpt["aux"]
    with torch.no_grad():
        cond, _ = no.upsample(no.token_encoder(P, S, tok, np.array([14])), torch.as_tensor(d), n)
        ar = no.decode(P, cond).numpy()
    shifted = np.concatenate([np.zeros_like(ar[:, :1]), ar[:, :-1]], axis=1)
    m1, m2 = no.teacher_forced(acoustic_ckpt, tok, np.array([14]), d, shifted)
    assert np.abs(m1 - ar).max() < 1e-4
    full = no.inference(acoustic_ckpt, tok, d, n).numpy()
    assert np.abs(m2 - full).max() < 1e-4


def test_zoneout_semantics(acoustic_ckpt):
    tok, d, n = _setup(1, 10, 0.25)
    mi = synthetic.mel_input(2, 1, n)
    base1, _ = no.teacher_forced(acoustic_ckpt, tok, np.array([10]), d, mi)
    zero = np.zeros((1, n, 4, 512), np.uint8)
    z1, _ = no.teacher_forced(acoustic_ckpt, tok, np.array([10]), d, mi, np.ones((1, n, 2, 256), np.uint8), zero)
    # keep-mask all ones = every unit kept and scaled by 2 -> differs from dropout-off
    assert np.abs(z1 - base1).max() > 1e-3
    # zoning every state at frame 3 changes nothing at frames <= 3 (outputs are the un-zoned states) but does afterwards
    zone = zero.copy()
    zone[:, 3] = 1
    a, _ = no.teacher_forced(acoustic_ckpt, tok, np.array([10]), d, mi, None, None)
    b, _ = no.teacher_forced(acoustic_ckpt, tok, np.array([10]), d, mi, np.zeros((1, n, 2, 256), np.uint8) + 1, zone)
    c, _ = no.teacher_forced(acoustic_ckpt, tok, np.array([10]), d, mi, np.zeros((1, n, 2, 256), np.uint8) + 1, zero)
    assert np.array_equal(b[:, :4], c[:, :4]) and np.abs(b[:, 4:] - c[:, 4:]).max() > 1e-4
    assert a.shape == b.shape


def test_gta_forward_shapes_and_shift(acoustic_ckpt):
    tok, d, n = _setup(3, 12, 20 * 256 / 16000)
    rng = np.random.default_rng(0)
    wav = (np.tanh(rng.standard_normal((1, 256 * 20)) * 0.4) * 20000).astype(np.int16)
    dur_sec = synthetic.utterance(3, 12, 20 * 256 / 16000)[1]
    gt, m2 = no.gta_forward(acoustic_ckpt, wav, tok, np.array([12]), dur_sec)
    assert gt.shape == m2.shape == (1, 20, 80)
    # causality of the pre-postnet output: perturbing the teacher-forcing input at frame t leaves mel1[:, :t] alone
    mi = synthetic.mel_input(4, 1, 20)
    a, _ = no.teacher_forced(acoustic_ckpt, tok, np.array([12]), d, mi)
    mi2 = mi.copy()
    mi2[:, 9] += 1.0
    b, _ = no.teacher_forced(acoustic_ckpt, tok, np.array([12]), d, mi2)
    assert np.array_equal(a[:, :9], b[:, :9]) and np.abs(a[:, 9:] - b[:, 9:]).max() > 1e-3
