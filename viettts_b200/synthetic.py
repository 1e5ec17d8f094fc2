"""Seeded synthetic checkpoints in the reference's Haiku layout.

There is no network, so the pretrained pickles that scripts/quick_start.sh:4-6
downloads are unavailable. Everything (tests, bench, golden fixtures) runs on
weights drawn here, in exactly the layout the reference stores:

  * hk_hifi.pickle  : {module_path: {"w","b"}} of numpy arrays, as written by
    vietTTS/hifigan/convert_torch_model_to_haiku.py:33-62 (Conv1D w is
    [K, C_in, C_out]; Conv1DTranspose w is [K, C_out, C_in]).
  * acoustic_latest_ckpt.pickle : {"step","params","aux","rng","optim_state"}
    (vietTTS/nat/text2mel.py:63-71) with the Haiku module names that
    vietTTS/nat/model.py:76-93 creates.

Scales are chosen so every layer's activation RMS stays O(0.1..3): N(0,0.01)
(the torch init, torch_model.py:16-19) would make every parity test vacuous.
"""
from __future__ import annotations

import numpy as np

from . import config as C

_RB_K = C.HIFIGAN["resblock_kernel_sizes"]
_UPS = list(zip(C.HIFIGAN["upsample_rates"], C.HIFIGAN["upsample_kernel_sizes"]))


def _normal(rng, shape, std):
    return (rng.standard_normal(shape, dtype=np.float32) * np.float32(std)).astype(np.float32)


def hifigan_params(seed: int = 1234) -> dict:
    """Haiku-layout parameter dict of the Generator (13 926 017 parameters)."""
    rng = np.random.default_rng(seed)
    p = {}
    c0 = C.HIFIGAN["upsample_initial_channel"]
    p["generator/~/conv1_d"] = dict(
        w=_normal(rng, (7, C.MEL_DIM, c0), 1.0 / np.sqrt(7 * C.MEL_DIM) / 5.0),  # mel values are O(5)
        b=_normal(rng, (c0,), 0.05),
    )
    ch = c0
    for i, (u, k) in enumerate(_UPS):
        cout = ch // 2
        # each output sample sees k/u taps of C_in channels
        p[f"generator/~/ups_{i}"] = dict(
            w=_normal(rng, (k, cout, ch), 1.3 / np.sqrt((k // u) * ch)),
            b=_normal(rng, (cout,), 0.05),
        )
        for j, kk in enumerate(_RB_K):
            n = i * len(_RB_K) + j
            for m in range(3):
                p[f"generator/~/res_block1_{n}/~/convs1_{m}"] = dict(
                    w=_normal(rng, (kk, cout, cout), 1.3 / np.sqrt(kk * cout)),
                    b=_normal(rng, (cout,), 0.05),
                )
                p[f"generator/~/res_block1_{n}/~/convs2_{m}"] = dict(
                    w=_normal(rng, (kk, cout, cout), 0.7 / np.sqrt(kk * cout)),
                    b=_normal(rng, (cout,), 0.05),
                )
        ch = cout
    p["generator/~/conv1_d_1"] = dict(
        w=_normal(rng, (7, ch, 1), 0.4 / np.sqrt(7 * ch)),
        b=_normal(rng, (1,), 0.02),
    )
    return p


def n_params(tree: dict) -> int:
    return int(sum(int(np.prod(a.shape)) for d in tree.values() for a in d.values()))


def _bn(rng, prefix, c, params, state):
    params[prefix] = dict(
        scale=(1.0 + _normal(rng, (1, 1, c), 0.1)).astype(np.float32),
        offset=_normal(rng, (1, 1, c), 0.1),
    )
    mean = _normal(rng, (1, 1, c), 0.1)
    var = rng.uniform(0.5, 1.5, (1, 1, c)).astype(np.float32)
    # hk.ExponentialMovingAverage state: counter / hidden / average; eval reads "average"
    state[prefix + "/~/mean_ema"] = dict(counter=np.array(1000, np.int32), hidden=mean.copy(), average=mean)
    state[prefix + "/~/var_ema"] = dict(counter=np.array(1000, np.int32), hidden=var.copy(), average=var)


def acoustic_ckpt(seed: int = 1234) -> dict:
    """Checkpoint dict with the keys predict_mel reads (text2mel.py:63-71)."""
    rng = np.random.default_rng(seed + 1)
    P, S = {}, {}
    A = "acoustic_model/~/"
    T = A + "token_encoder/~/"
    D = C.ACOUSTIC_ENCODER_DIM
    H = C.ACOUSTIC_DECODER_DIM
    P[T + "embed"] = dict(embeddings=_normal(rng, (C.VOCAB_SIZE, D), 1.0))
    for i in range(3):
        sfx = "" if i == 0 else f"_{i}"
        P[T + "conv1_d" + sfx] = dict(w=_normal(rng, (3, D, D), 1.4 / np.sqrt(3 * D)), b=_normal(rng, (D,), 0.05))
        _bn(rng, T + "batch_norm" + sfx, D, P, S)
    for name in ("lstm", "lstm_1"):
        P[T + name + "/linear"] = dict(w=_normal(rng, (2 * D, 4 * D), 1.0 / np.sqrt(2 * D)), b=_normal(rng, (4 * D,), 0.05))
    x_dim = C.ENC_OUT_DIM + C.PRENET_DIM  # 768
    P[A + "lstm/linear"] = dict(w=_normal(rng, (x_dim + H, 4 * H), 1.0 / np.sqrt(x_dim + H)), b=_normal(rng, (4 * H,), 0.05))
    P[A + "lstm_1/linear"] = dict(w=_normal(rng, (x_dim + 2 * H, 4 * H), 1.0 / np.sqrt(x_dim + 2 * H)), b=_normal(rng, (4 * H,), 0.05))
    # projection: mel-ish range (log-mel lives in [-11.5, 2])
    pb = (-4.0 + _normal(rng, (C.MEL_DIM,), 1.0)).astype(np.float32)
    P[A + "linear"] = dict(w=_normal(rng, (2 * H, C.MEL_DIM), 3.0 / np.sqrt(2 * H)), b=pb)
    P[A + "linear_1"] = dict(w=_normal(rng, (C.MEL_DIM, C.PRENET_DIM), 1.0 / np.sqrt(C.MEL_DIM) / 3.0))
    P[A + "linear_2"] = dict(w=_normal(rng, (C.PRENET_DIM, C.PRENET_DIM), 1.6 / np.sqrt(C.PRENET_DIM)))
    dims = [C.MEL_DIM] + [C.POSTNET_DIM] * 4 + [C.MEL_DIM]
    for i in range(5):
        sfx = "" if i == 0 else f"_{i}"
        gain = (1.0 / 4.0 if i == 0 else 1.2) if i < 4 else 0.5
        P[A + "conv1_d" + sfx] = dict(
            w=_normal(rng, (5, dims[i], dims[i + 1]), gain / np.sqrt(5 * dims[i])),
            b=_normal(rng, (dims[i + 1],), 0.05),
        )
        if i < 4:
            _bn(rng, A + "batch_norm" + sfx, C.POSTNET_DIM, P, S)
    S["acoustic_model"] = dict(attn=np.zeros((1, 1), np.float32))
    return dict(step=0, params=P, aux=S, rng=np.array([0, 42], np.uint32), optim_state=None)


def duration_ckpt(seed: int = 1234) -> dict:
    """Checkpoint dict with the keys predict_duration reads (text2mel.py:27-34): params / aux / rng.
    DurationModel (model.py:49-70): TokenEncoder(vocab, 256) + Sequential([Linear(256), gelu, Linear(1)])."""
    rng = np.random.default_rng(seed + 2)
    P, S = {}, {}
    M = "duration_model/~/"
    T = M + "token_encoder/~/"
    D = C.DURATION_LSTM_DIM
    P[T + "embed"] = dict(embeddings=_normal(rng, (C.VOCAB_SIZE, D), 1.0))
    for i in range(3):
        sfx = "" if i == 0 else f"_{i}"
        P[T + "conv1_d" + sfx] = dict(w=_normal(rng, (3, D, D), 1.4 / np.sqrt(3 * D)), b=_normal(rng, (D,), 0.05))
        _bn(rng, T + "batch_norm" + sfx, D, P, S)
    for name in ("lstm", "lstm_1"):
        P[T + name + "/linear"] = dict(w=_normal(rng, (2 * D, 4 * D), 1.0 / np.sqrt(2 * D)), b=_normal(rng, (4 * D,), 0.05))
    P[M + "linear"] = dict(w=_normal(rng, (2 * D, D), 2.0 / np.sqrt(2 * D)), b=_normal(rng, (D,), 0.1))
    # head bias -2.5: softplus(-2.5) = 0.079 s, a plausible phoneme duration; the weights spread it over ~0.02..0.3 s
    P[M + "linear_1"] = dict(w=_normal(rng, (D, 1), 3.0 / np.sqrt(D)), b=np.array([-2.5], np.float32))
    return dict(step=0, params=P, aux=S, rng=np.array([0, 43], np.uint32), optim_state=None)


# ---------------------------------------------------------------------------
# synthetic inputs (C1..C5)
# ---------------------------------------------------------------------------

def utterance(seed: int, n_phonemes: int = 100, seconds: float | None = 5.0):
    """One C1-style utterance: tokens list[int] and durations f32[1,L] in seconds.

    tokens[0]=tokens[-1]=sil, every 5th token is a word end (duration 0, as
    text2mel.py:95-97 forces), the rest uniform in the 89-letter alphabet.
    If `seconds` is None every non-word-end token lasts 0.05 s (C5 style).
    """
    rng = np.random.default_rng(seed)
    L = n_phonemes
    tokens = rng.integers(4, C.ALPHABET_SIZE, size=L)
    tokens[4::5] = C.WORD_END_INDEX
    tokens[0] = tokens[-1] = C.SIL_INDEX
    dur = np.where(tokens == C.WORD_END_INDEX, 0.0, 0.05).astype(np.float64)
    dur *= rng.uniform(0.6, 1.4, size=L)
    dur[tokens == C.WORD_END_INDEX] = 0.0
    if seconds is not None:
        dur *= seconds / dur.sum()
        # keep int(sum(frames)) stable against f32 summation order
        dur[0] += 0.3 / 62.5
    return [int(t) for t in tokens], dur.astype(np.float32)[None, :]


def mel_input(seed: int, batch: int, n_frames: int) -> np.ndarray:
    """C2-style random log-mel: clip(N(-5, 2^2), log 1e-5, 2)."""
    rng = np.random.default_rng(seed)
    m = rng.standard_normal((batch, n_frames, C.MEL_DIM), dtype=np.float32) * 2.0 - 5.0
    return np.clip(m, np.log(1e-5), 2.0).astype(np.float32)


def dropout_masks(seed: int, batch: int, n_frames: int) -> np.ndarray:
    """uint8 keep-masks [B, N, 2, 256] for the two prenet dropouts (rate 0.5)."""
    rng = np.random.default_rng(seed)
    return (rng.random((batch, n_frames, 2, C.PRENET_DIM)) < 0.5).astype(np.uint8)
