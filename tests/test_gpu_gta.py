"""GPU parity: teacher-forced acoustic pass with zoneout (AcousticModel.__call__, model.py:146-169) and the GTA
forward (gta.py:28-41) vs the CPU restatement.

Oracle status: pinned to the reference's own source (tests/test_reference_goldens.py, tests/test_gpu_reference_goldens.py).  Dropout and zoneout masks are explicit inputs shared by both
sides.  Tolerance: mel L-inf <= 1e-3 (log-mel units), the same bar as the autoregressive path."""
import numpy as np
import pytest
import torch

from oracle import nat_oracle as no
from viettts_b200 import synthetic

pytestmark = pytest.mark.gpu
MEL_LINF = 1e-3


@pytest.fixture(scope="module", params=["fp32", "bf16x3"])
def eng(acoustic_ckpt, request):
    from viettts_b200.engine import Engine
    e = Engine(0)
    e.load_acoustic(acoustic_ckpt)
    e.load_mel_filterbank()
    e.set_precision(request.param)
    yield e
    e.close()


def _utt(seed, L, seconds):
    tokens, dur = synthetic.utterance(seed, L, seconds)
    d, n = no.seconds_to_frames(dur)
    return np.asarray(tokens, np.int32), d[0], n


def _masks(seed, B, N):
    rng = np.random.default_rng(seed)
    keep = (rng.random((B, N, 2, 256)) < 0.5).astype(np.uint8)
    zone = (rng.random((B, N, 4, 512)) < 0.1).astype(np.uint8)
    return keep, zone


def test_teacher_forced_vs_oracle(eng, acoustic_ckpt):
    tk, d, n = _utt(0, 40, 1.5)
    mels_in = synthetic.mel_input(5, 1, n)
    keep, zone = _masks(1, 1, n)
    m1, m2 = eng.teacher_forced(tk[None], d[None], mels_in, keep_masks=keep, zone_masks=zone)
    r1, r2 = no.teacher_forced(acoustic_ckpt, tk[None], np.array([40]), d[None], mels_in, keep, zone, dtype=torch.float64)
    print(f"teacher forced: mel1 {np.abs(m1-r1).max():.3e} mel2 {np.abs(m2-r2).max():.3e}")
    assert np.abs(m1 - r1).max() < MEL_LINF and np.abs(m2 - r2).max() < MEL_LINF
    # zoneout really is in the path: without the masks the result differs
    r1_off, _ = no.teacher_forced(acoustic_ckpt, tk[None], np.array([40]), d[None], mels_in, keep, np.zeros_like(zone))
    assert np.abs(r1_off - r1).max() > 10 * MEL_LINF


def test_masks_off_and_c1_size(eng, acoustic_ckpt):
    tk, d, n = _utt(2, 100, 5.0)
    assert n == 312
    mels_in = synthetic.mel_input(6, 1, n)
    m1, m2 = eng.teacher_forced(tk[None], d[None], mels_in)
    r1, r2 = no.teacher_forced(acoustic_ckpt, tk[None], np.array([100]), d[None], mels_in, dtype=torch.float64)
    assert np.abs(m1 - r1).max() < MEL_LINF and np.abs(m2 - r2).max() < MEL_LINF


def test_ragged_batch_rows_equal_single_runs(eng, acoustic_ckpt):
    """34 rows (two scan launches) of mixed length: row b == that row alone, unpadded."""
    B = 34
    rng = np.random.default_rng(8)
    Ls = rng.integers(8, 30, size=B)
    utts = [_utt(50 + b, int(Ls[b]), None) for b in range(B)]
    L, N = int(Ls.max()), max(u[2] for u in utts)
    tok = np.zeros((B, L), np.int32)
    dur = np.zeros((B, L), np.float32)
    nf = np.array([u[2] for u in utts], np.int32)
    for b, (tk, d, n) in enumerate(utts):
        tok[b, : len(tk)] = tk
        dur[b, : len(tk)] = d
    mels_in = synthetic.mel_input(9, B, N)
    keep, zone = _masks(2, B, N)
    m1, m2 = eng.teacher_forced(tok, dur, mels_in, lengths=Ls.astype(np.int32), n_frames=nf, keep_masks=keep, zone_masks=zone)
    for b in (0, 17, 32, 33):
        tk, d, n = utts[b]
        r1, r2 = no.teacher_forced(acoustic_ckpt, tk[None], np.array([len(tk)]), d[None], mels_in[b : b + 1, :n], keep[b : b + 1, :n],
                                   zone[b : b + 1, :n], dtype=torch.float64)
        assert np.abs(m2[b, :n] - r2[0]).max() < MEL_LINF, b
        assert np.abs(m1[b, :n] - r1[0]).max() < MEL_LINF, b
        assert np.all(m2[b, n:] == 0) and np.all(m1[b, n:] == 0)


def test_seed_mode_is_deterministic_and_active(eng):
    tk, d, n = _utt(4, 20, 0.6)
    mels_in = synthetic.mel_input(7, 1, n)
    a1, a2 = eng.teacher_forced(tk[None], d[None], mels_in, seed=11)
    b1, b2 = eng.teacher_forced(tk[None], d[None], mels_in, seed=11)
    c1, c2 = eng.teacher_forced(tk[None], d[None], mels_in, seed=12)
    off1, _ = eng.teacher_forced(tk[None], d[None], mels_in)
    assert np.array_equal(a2, b2) and not np.array_equal(a2, c2) and not np.array_equal(a1, off1)


def test_gta_forward_vs_oracle(eng, acoustic_ckpt):
    """gta.py:28-41: int16 wav -> MelFilter -> shift -> teacher-forced model, one library call."""
    B, L, S = 2, 24, 256 * 60
    rng = np.random.default_rng(3)
    wav = (np.tanh(rng.standard_normal((B, S)) * 0.4) * 20000).astype(np.int16)
    tok = np.stack([_utt(70 + b, L, None)[0] for b in range(B)])
    dur_sec = np.stack([synthetic.utterance(70 + b, L, 60 * 256 / 16000)[1][0] for b in range(B)])
    keep, zone = _masks(4, B, S // 256)
    out, gt = eng.gta(wav, tok, dur_sec, keep_masks=keep, zone_masks=zone, return_gt=True)
    for b in range(B):
        gt_ref, ref = no.gta_forward(acoustic_ckpt, wav[b : b + 1], tok[b : b + 1], np.array([L]), dur_sec[b : b + 1], keep[b : b + 1],
                                     zone[b : b + 1], dtype=torch.float64)
        print(f"gta row {b}: mel_gt {np.abs(gt[b]-gt_ref[0]).max():.3e} mel2 {np.abs(out[b]-ref[0]).max():.3e}")
        assert np.abs(gt[b] - gt_ref[0]).max() < 2e-3
        assert np.abs(out[b] - ref[0]).max() < 2e-3      # includes the fp32 STFT's error on the teacher-forcing input
    # wav_lengths: frames past wav_length // 256 are zero, earlier ones unchanged
    out2 = eng.gta(wav, tok, dur_sec, wav_lengths=[S, 256 * 41 + 100], keep_masks=keep, zone_masks=zone)
    assert np.all(out2[1, 41:] == 0) and np.abs(out2[0] - out[0]).max() < 1e-5


def test_bad_arguments_raise(eng):
    from viettts_b200._lib import VttsError
    tk, d, n = _utt(4, 10, 0.3)
    with pytest.raises(ValueError):
        eng.teacher_forced(tk[None], d[None], np.zeros((1, n, 79), np.float32))
    with pytest.raises(ValueError):
        eng.teacher_forced(tk[None], d[None], np.zeros((1, n, 80), np.float32), keep_masks=np.zeros((1, n, 2, 256), np.uint8))
    with pytest.raises(ValueError):
        eng.gta(np.zeros((1, 1000), np.float32), tk[None], d[None])
    with pytest.raises(VttsError):
        eng.gta(np.zeros((1, 1000), np.int16), tk[None], d[None])        # S not a multiple of 256


def test_gta_dropin_writes_reference_format(eng, acoustic_ckpt, tmp_path, monkeypatch):
    """nat/gta.py drop-in: checkpoint from the reference's cwd-relative path, one <name>.npy = mel[:l].T per utterance."""
    import pickle
    from viettts_b200.engine import get_engine
    from viettts_b200.nat import gta
    (tmp_path / "assets/infore/nat").mkdir(parents=True)
    with open(tmp_path / "assets/infore/nat/acoustic_latest_ckpt.pickle", "wb") as f:
        pickle.dump(acoustic_ckpt, f)
    monkeypatch.chdir(tmp_path)
    get_engine(0).set_precision(eng.lib.vtts_get_precision(eng.h))
    B, L, S = 2, 16, 256 * 30
    rng = np.random.default_rng(5)
    wav = (np.tanh(rng.standard_normal((B, S)) * 0.4) * 20000).astype(np.int16)
    tok = np.stack([_utt(90 + b, L, None)[0] for b in range(B)])
    dur = np.stack([synthetic.utterance(90 + b, L, 30 * 256 / 16000)[1][0] for b in range(B)])
    wl = np.array([S, 256 * 21 + 17], np.int32)
    keep, zone = _masks(6, B, 30)
    files = gta.generate_gta([(["a", "b"], wav, wl, tok, np.array([L, L], np.int32), dur)], tmp_path / "gta", keep_masks=keep, zone_masks=zone)
    a, b = np.load(files[0]), np.load(files[1])
    assert a.shape == (80, 30) and b.shape == (80, 21) and a.dtype == np.float32
    # like the reference's forward_fn_ (gta.py:28-41) the model runs over the FULL padded length; wav_lengths only
    # slices the saved array (gta.py:70-76)
    ref = eng.gta(wav, tok, dur, lengths=[L, L], wav_lengths=None, keep_masks=keep, zone_masks=zone)
    assert np.array_equal(a, ref[0].T) and np.array_equal(b, ref[1, :21].T)
    # and the saved short row equals the oracle's forward over the full padded N, sliced (the last ~10 frames before
    # the cut see postnet context from frames >= l, exactly as in the reference)
    _, mel2 = no.gta_forward(acoustic_ckpt, wav, tok, np.array([L, L]), dur, keep, zone, dtype=torch.float64)
    assert np.abs(b - mel2[1, :21].T).max() < 2 * MEL_LINF
