"""Drop-in for `predict_mel` of vietTTS/nat/text2mel.py:61-82.

`predict_mel(tokens, durations)`: tokens list[int] (length L), durations f32
[1,L] in SECONDS; reads assets/infore/nat/acoustic_latest_ckpt.pickle (keys
step/params/aux/rng/optim_state) and returns mel f32 [1,N,80] with
N = int(sum(durations*16000/256)).

Dropout: the reference applies prenet dropout at inference with the
checkpoint's `rng` through JAX's threefry stream (nat/model.py:95-100,132).
That stream cannot be reproduced without JAX, so by default the masks are drawn
on the device from a threefry2x32 counter stream keyed by the checkpoint's rng
words; pass `masks=` (uint8 [1,N,2,256], e.g. dumped from the real JAX run) for
bit-identical masks, or `dropout=False` for the deterministic mode."""
from __future__ import annotations

import os

import numpy as np

from .. import config
from ..engine import get_engine
from ..weights import load_pickle

CKPT_FILE = config.ACOUSTIC_CKPT


def _file_key(path):
    st = os.stat(path)
    return (str(path), st.st_mtime_ns, st.st_size)


_rng_words = {}


def load_acoustic(engine=None, ckpt_file=None):
    engine = engine or get_engine()
    ckpt_file = ckpt_file or CKPT_FILE
    key = _file_key(ckpt_file)
    if engine._acoustic_key != key:
        dic = load_pickle(ckpt_file)
        engine.load_acoustic(dic, key=key)
        rng = np.asarray(dic.get("rng", [0, 42])).astype(np.uint64).ravel()
        _rng_words[key] = int((int(rng[0]) << 32) | int(rng[-1]))
    return engine, _rng_words.get(key, 42)


def seconds_to_frames(durations):
    """text2mel.py:78-79 in float32: durations * sample_rate / (n_fft // 4)."""
    d = (np.asarray(durations, np.float32) * np.float32(config.SAMPLE_RATE)) / np.float32(config.N_FFT // 4)
    return d, int(np.sum(d, dtype=np.float32))


def predict_mel(tokens, durations, masks=None, dropout=True):
    engine, seed = load_acoustic()
    d, n_frames = seconds_to_frames(durations)
    tokens = np.array(tokens, dtype=np.int32)[None, :]
    d = d.reshape(1, -1)
    return engine.predict_mel(tokens, d, n_frames=[n_frames], masks=masks, seed=(seed if (dropout and masks is None) else None))
