"""Checkpoint handling: the reference's pickles -> the library's canonical blobs.

  * hk_hifi.pickle (vietTTS/hifigan/mel2wave.py:35-36) is a plain dict of numpy
    arrays in the layout convert_torch_model_to_haiku.py:34-58 writes.
  * acoustic_latest_ckpt.pickle (vietTTS/nat/text2mel.py:62-71) holds Haiku
    param/state trees whose leaves are jax arrays; `load_pickle` unpickles it
    WITHOUT jax by mapping array reconstruction onto numpy.

The blob order is the one csrc/api.cu (vtts_hifigan_specs / vtts_acoustic_specs)
expects and INTEGRATION.md documents.
"""
from __future__ import annotations

import io
import pickle
from pathlib import Path

import numpy as np

from . import config as C

_RB_K = C.HIFIGAN["resblock_kernel_sizes"]


# ---------------------------------------------------------------------------
# jax-free unpickling
# ---------------------------------------------------------------------------
class _Opaque:
    """Stand-in for classes we do not need (optimizer state etc.)."""

    def __init__(self, *a, **k):
        self.args, self.kwargs = a, k

    def __setstate__(self, state):
        self.state = state


def _np_from_jax(*args, **kwargs):
    """Stand-in for jax's array reconstruction functions.  Two pickled layouts exist:
      * `reconstruct(fun, args, arr_state, aval_state)` (jax.Array / DeviceArray.__reduce__ since jax 0.2.10):
        `fun, args, arr_state` are the numpy array's own `__reduce__` triple -> rebuild the ndarray from them;
      * an ndarray passed directly (or a plain numpy pickle, which never reaches this function)."""
    if len(args) >= 3 and callable(args[0]) and isinstance(args[1], tuple):
        try:
            arr = args[0](*args[1])
            arr.__setstate__(args[2])
            return np.asarray(arr)
        except Exception:
            pass
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, np.ndarray):
            return a
    return _Opaque(*args, **kwargs)


class _FlatMapping(dict):
    """Stand-in for haiku's FlatMapping (hk.Params / hk.State).  Haiku pickles it as `FlatMapping(dict)`
    (its `__reduce__` goes through a plain dict); very old versions used the default object protocol with the
    mapping in `_mapping`.  Both restore to a plain dict of dicts."""

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[0], dict):   # (dict_state, slots)
            state = state[0]
        m = state.get("_mapping") if isinstance(state, dict) else None
        if m is None:
            raise pickle.UnpicklingError("FlatMapping pickled as (treedef, leaves): re-save the checkpoint with "
                                         "hk.data_structures.to_mutable_dict / jax.device_get, or load it once with haiku installed")
        self.update(m)


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        top = module.split(".")[0]
        if top in ("jax", "jaxlib"):
            return _np_from_jax
        if top == "haiku" or module.startswith("haiku."):
            if "FlatMap" in name:
                return _FlatMapping
            return _Opaque
        if top in ("optax", "chex", "flax"):
            return _Opaque
        return super().find_class(module, name)


def _plain(x):
    """_FlatMapping -> dict, recursively (pack_* index plain dicts)."""
    if isinstance(x, dict):
        return {k: _plain(v) for k, v in x.items()}
    return x


def load_pickle(path):
    with open(path, "rb") as f:
        data = f.read()
    try:
        return _plain(_Unpickler(io.BytesIO(data)).load())
    except (ModuleNotFoundError, AttributeError, ImportError) as first:
        # a class the jax-free unpickler does not know: only the real modules can resolve it.  Any other failure
        # (e.g. the explicit UnpicklingError of _FlatMapping.__setstate__) propagates unchanged.
        try:
            return pickle.loads(data)
        except Exception as second:
            raise second from first


# ---------------------------------------------------------------------------
# HiFiGAN
# ---------------------------------------------------------------------------
def _get(tree: dict, path: str):
    if path in tree:
        return tree[path]
    alt = path.replace("/~/", "/")
    for k in tree:
        if k.replace("/~/", "/") == alt:
            return tree[k]
    raise KeyError(f"module {path!r} not found in checkpoint (have e.g. {sorted(tree)[:3]})")


def _arr(a, shape, what):
    a = np.asarray(a, dtype=np.float32)
    if tuple(a.shape) != tuple(shape):
        if a.size == int(np.prod(shape)) and a.squeeze().shape == tuple(s for s in shape if s != 1):
            a = a.reshape(shape)
        else:
            raise ValueError(f"{what}: expected shape {tuple(shape)}, checkpoint has {tuple(a.shape)}")
    return np.ascontiguousarray(a).ravel()


def pack_hifigan(params: dict) -> np.ndarray:
    """Haiku-layout generator parameters -> float32 blob (13 926 017 floats)."""
    g = "generator/~/"
    parts = []
    c0 = C.HIFIGAN["upsample_initial_channel"]
    m = _get(params, g + "conv1_d")
    parts += [_arr(m["w"], (7, C.MEL_DIM, c0), "conv_pre.w"), _arr(m["b"], (c0,), "conv_pre.b")]
    ch = c0
    for i, k in enumerate(C.HIFIGAN["upsample_kernel_sizes"]):
        m = _get(params, g + f"ups_{i}")
        parts += [_arr(m["w"], (k, ch // 2, ch), f"ups_{i}.w"), _arr(m["b"], (ch // 2,), f"ups_{i}.b")]
        ch //= 2
    for n in range(12):
        k = _RB_K[n % 3]
        ch = 256 >> (n // 3)
        for which in ("convs1", "convs2"):
            for mi in range(3):
                m = _get(params, g + f"res_block1_{n}/~/{which}_{mi}")
                parts += [_arr(m["w"], (k, ch, ch), f"rb{n}.{which}_{mi}.w"), _arr(m["b"], (ch,), f"rb{n}.{which}_{mi}.b")]
    m = _get(params, g + "conv1_d_1")
    parts += [_arr(m["w"], (7, 32, 1), "conv_post.w"), _arr(m["b"], (1,), "conv_post.b")]
    blob = np.concatenate(parts)
    assert blob.size == C.HIFIGAN_PARAMS, blob.size
    return blob


# ---------------------------------------------------------------------------
# acoustic model
# ---------------------------------------------------------------------------
def pack_acoustic(ckpt: dict) -> np.ndarray:
    """{"params","aux",...} (text2mel.py:63-71) -> float32 blob in api.cu order."""
    P, S = ckpt["params"], ckpt["aux"]
    A = "acoustic_model/~/"
    T = A + "token_encoder/~/"
    D = C.ACOUSTIC_ENCODER_DIM
    H = C.ACOUSTIC_DECODER_DIM
    parts = [_arr(_get(P, T + "embed")["embeddings"], (C.VOCAB_SIZE, D), "embed")]

    def bn(prefix, c, what):
        m = _get(P, prefix)
        return [
            _arr(m["scale"], (c,), what + ".scale"),
            _arr(m["offset"], (c,), what + ".offset"),
            _arr(_get(S, prefix + "/~/mean_ema")["average"], (c,), what + ".mean"),
            _arr(_get(S, prefix + "/~/var_ema")["average"], (c,), what + ".var"),
        ]

    for i in range(3):
        sfx = "" if i == 0 else f"_{i}"
        m = _get(P, T + "conv1_d" + sfx)
        parts += [_arr(m["w"], (3, D, D), f"enc.conv{i}.w"), _arr(m["b"], (D,), f"enc.conv{i}.b")]
        parts += bn(T + "batch_norm" + sfx, D, f"enc.bn{i}")
    for name in ("lstm", "lstm_1"):
        m = _get(P, T + name + "/linear")
        parts += [_arr(m["w"], (2 * D, 4 * D), f"enc.{name}.w"), _arr(m["b"], (4 * D,), f"enc.{name}.b")]
    x_dim = C.ENC_OUT_DIM + C.PRENET_DIM
    m = _get(P, A + "lstm/linear")
    parts += [_arr(m["w"], (x_dim + H, 4 * H), "dec.lstm0.w"), _arr(m["b"], (4 * H,), "dec.lstm0.b")]
    m = _get(P, A + "lstm_1/linear")
    parts += [_arr(m["w"], (x_dim + 2 * H, 4 * H), "dec.lstm1.w"), _arr(m["b"], (4 * H,), "dec.lstm1.b")]
    m = _get(P, A + "linear")
    parts += [_arr(m["w"], (2 * H, C.MEL_DIM), "proj.w"), _arr(m["b"], (C.MEL_DIM,), "proj.b")]
    parts += [_arr(_get(P, A + "linear_1")["w"], (C.MEL_DIM, C.PRENET_DIM), "prenet.fc1.w")]
    parts += [_arr(_get(P, A + "linear_2")["w"], (C.PRENET_DIM, C.PRENET_DIM), "prenet.fc2.w")]
    dims = [C.MEL_DIM] + [C.POSTNET_DIM] * 4 + [C.MEL_DIM]
    for i in range(5):
        sfx = "" if i == 0 else f"_{i}"
        m = _get(P, A + "conv1_d" + sfx)
        parts += [_arr(m["w"], (5, dims[i], dims[i + 1]), f"postnet.conv{i}.w"), _arr(m["b"], (dims[i + 1],), f"postnet.conv{i}.b")]
        if i < 4:
            parts += bn(A + "batch_norm" + sfx, C.POSTNET_DIM, f"postnet.bn{i}")
    return np.concatenate(parts)


def pack_duration(ckpt: dict) -> np.ndarray:
    """duration_latest_ckpt.pickle contents {"params","aux",...} (text2mel.py:27-34) -> float32 blob in
    vtts_duration_specs order: the TokenEncoder tensors (same order as in the acoustic blob), then the
    projection head Linear(512->256), Linear(256->1) (model.py:60-62)."""
    P, S = ckpt["params"], ckpt["aux"]
    M = "duration_model/~/"
    T = M + "token_encoder/~/"
    D = C.DURATION_LSTM_DIM
    parts = [_arr(_get(P, T + "embed")["embeddings"], (C.VOCAB_SIZE, D), "embed")]
    for i in range(3):
        sfx = "" if i == 0 else f"_{i}"
        m = _get(P, T + "conv1_d" + sfx)
        parts += [_arr(m["w"], (3, D, D), f"enc.conv{i}.w"), _arr(m["b"], (D,), f"enc.conv{i}.b")]
        bn = _get(P, T + "batch_norm" + sfx)
        parts += [
            _arr(bn["scale"], (D,), f"enc.bn{i}.scale"),
            _arr(bn["offset"], (D,), f"enc.bn{i}.offset"),
            _arr(_get(S, T + "batch_norm" + sfx + "/~/mean_ema")["average"], (D,), f"enc.bn{i}.mean"),
            _arr(_get(S, T + "batch_norm" + sfx + "/~/var_ema")["average"], (D,), f"enc.bn{i}.var"),
        ]
    for name in ("lstm", "lstm_1"):
        m = _get(P, T + name + "/linear")
        parts += [_arr(m["w"], (2 * D, 4 * D), f"enc.{name}.w"), _arr(m["b"], (4 * D,), f"enc.{name}.b")]
    m = _get(P, M + "linear")
    parts += [_arr(m["w"], (2 * D, D), "proj.fc1.w"), _arr(m["b"], (D,), "proj.fc1.b")]
    m = _get(P, M + "linear_1")
    parts += [_arr(m["w"], (D, 1), "proj.fc2.w"), _arr(m["b"], (1,), "proj.fc2.b")]
    return np.concatenate(parts)


# ---------------------------------------------------------------------------
# librosa-compatible mel filterbank (MelFilter.__init__, dsp.py:107-113)
# ---------------------------------------------------------------------------
def mel_filterbank(sample_rate=C.SAMPLE_RATE, n_fft=C.N_FFT, n_mels=C.MEL_DIM, fmin=C.FMIN, fmax=C.FMAX) -> np.ndarray:
    """Slaney-scale, area-normalised triangular filterbank, float32 [n_mels, 1+n_fft//2]
    (what librosa.filters.mel returns with its defaults htk=False, norm='slaney')."""
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp

    def hz2mel(f):
        f = np.asarray(f, np.float64)
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, f / f_sp)

    def mel2hz(m):
        m = np.asarray(m, np.float64)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    n_bins = 1 + n_fft // 2
    fft_f = np.linspace(0.0, sample_rate / 2.0, n_bins)
    mel_f = mel2hz(np.linspace(hz2mel(fmin), hz2mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fft_f[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2:] - mel_f[:-2]))[:, None]
    return w.astype(np.float32)
