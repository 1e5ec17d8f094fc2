"""Stand-in for the slice of dm-haiku that vietTTS/nat/{model,text2mel,gta}.py touch.  See ../README.md.

Restated from dm-haiku's published source (module.py, basic.py, conv.py, batch_norm.py, moving_averages.py,
recurrent.py, stateful.py, base.py).  Only "apply" mode exists: parameters and state must already be in the
trees passed to `apply`; a missing entry or a shape mismatch raises, which is what pins the reference's Haiku
parameter names and shapes to the checkpoint layout.

Module naming (module.py `unique_and_canonical_name`): default name = snake_case(class name); the n-th module with
that name created during ONE method call of its parent gets the suffix `_n`; full name = parent + "/~/" + name when
created inside the parent's __init__, parent + "/" + name inside __call__, parent + "/~method/" + name elsewhere.
"""
from __future__ import annotations

import collections
import functools
import re
from typing import NamedTuple

import numpy as np

import jax
import jax.numpy as jnp

F64 = np.float64


# ------------------------------------------------------------------------------------------------------------
# frame (transform_with_state.apply)
# ------------------------------------------------------------------------------------------------------------
class _Frame:
    def __init__(self, params, state, rng):
        self.params = params
        self.state = {k: dict(v) for k, v in (state or {}).items()}
        self.key = None if rng is None else np.asarray(rng, np.uint32)
        self.module_stack = []                       # (module, method_name)
        self.counter_stack = [collections.Counter()]


_frames: list[_Frame] = []


def _frame() -> _Frame:
    if not _frames:
        raise RuntimeError("haiku shim: must be called inside transform_with_state(...).apply")
    return _frames[-1]


class _Transformed(NamedTuple):
    init: object
    apply: object


def transform_with_state(f):
    def apply(params, state, rng, *args, **kwargs):
        _frames.append(_Frame(params, state, rng))
        try:
            out = f(*args, **kwargs)
            return out, _frames[-1].state
        finally:
            _frames.pop()

    def init(*a, **k):
        raise NotImplementedError("haiku shim: apply-only")

    return _Transformed(init, apply)


def next_rng_key():
    """PRNGSequence.__next__ (base.py): key, subkey = split(key) -- the first half is carried, the second returned."""
    fr = _frame()
    if fr.key is None:
        raise ValueError("haiku shim: next_rng_key without an rng")
    new = jax.random.split(fr.key, 2)
    fr.key = new[0]
    return new[1]


# ------------------------------------------------------------------------------------------------------------
# modules
# ------------------------------------------------------------------------------------------------------------
_CAMEL_TO_SNAKE_R = re.compile(r"((?<=[a-z0-9])[A-Z]|(?!^)[A-Z](?=[a-z]))")


def _snake(name: str) -> str:
    """utils.camel_to_snake: Conv1D -> conv1_d, BatchNorm -> batch_norm, LSTM -> lstm."""
    return _CAMEL_TO_SNAKE_R.sub(r"_\1", name.lstrip("_")).lower()


def _wrap_method(method_name, fn):
    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        fr = _frame()
        fr.module_stack.append((self, method_name))
        fr.counter_stack.append(collections.Counter())
        try:
            return fn(self, *a, **k)
        finally:
            fr.counter_stack.pop()
            fr.module_stack.pop()

    return wrapped


class _ModuleMeta(type):
    def __new__(mcs, name, bases, dct):
        for key, val in list(dct.items()):
            if key == "__init__" or not callable(val) or isinstance(val, (staticmethod, classmethod, type)):
                continue
            if key.startswith("__") and key != "__call__":
                continue
            dct[key] = _wrap_method(key, val)
        return super().__new__(mcs, name, bases, dct)

    def __call__(cls, *a, **k):
        fr = _frame()
        obj = cls.__new__(cls)
        fr.module_stack.append((obj, "__init__"))
        fr.counter_stack.append(collections.Counter())
        try:
            obj.__init__(*a, **k)
        finally:
            fr.counter_stack.pop()
            fr.module_stack.pop()
        if not hasattr(obj, "module_name"):
            raise ValueError(f"{cls.__name__}: super().__init__() was not called")
        return obj


class Module(metaclass=_ModuleMeta):
    def __init__(self, name=None):
        fr = _frame()
        if name is None:
            name = _snake(type(self).__name__)
        # the constructor frame of THIS module is on top; the creator (if any) is one below
        if len(fr.module_stack) > 1:
            parent, method = fr.module_stack[-2]
            if method == "__init__":
                name = "~/" + name
            elif method != "__call__":
                name = "~" + method + "/" + name
            name = parent.module_name + "/" + name
        counters = fr.counter_stack[-2]
        n = counters[name]
        counters[name] += 1
        self.module_name = f"{name}_{n}" if n else name
        self.name = self.module_name.split("/")[-1]


def _current_module_name() -> str:
    fr = _frame()
    if not fr.module_stack:
        raise ValueError("haiku shim: parameters / state must be accessed inside a module method")
    return fr.module_stack[-1][0].module_name


def get_parameter(name, shape, dtype=None, init=None):
    fr = _frame()
    mod = _current_module_name()
    if mod not in fr.params or name not in fr.params[mod]:
        raise KeyError(f"haiku shim: parameter {mod!r}/{name!r} missing from the checkpoint")
    v = np.asarray(fr.params[mod][name])
    if tuple(v.shape) != tuple(shape):
        raise ValueError(f"haiku shim: {mod}/{name} has shape {v.shape}, the module asks for {tuple(shape)}")
    return v.astype(F64)


def get_state(name, shape=None, dtype=None, init=None):
    fr = _frame()
    mod = _current_module_name()
    if mod not in fr.state or name not in fr.state[mod]:
        raise KeyError(f"haiku shim: state {mod!r}/{name!r} missing from the checkpoint")
    v = np.asarray(fr.state[mod][name])
    return v.astype(F64) if v.dtype.kind == "f" else v


def set_state(name, value):
    fr = _frame()
    fr.state.setdefault(_current_module_name(), {})[name] = value


# ---- basic.py ----------------------------------------------------------------------------------------------
class Linear(Module):
    def __init__(self, output_size, with_bias=True, w_init=None, b_init=None, name=None):
        super().__init__(name=name)
        self.output_size = output_size
        self.with_bias = with_bias

    def __call__(self, inputs):
        in_size = inputs.shape[-1]
        w = get_parameter("w", [in_size, self.output_size])
        out = jnp.dot(inputs, w)
        if self.with_bias:
            out = out + get_parameter("b", [self.output_size])
        return out


class Sequential(Module):
    def __init__(self, layers, name=None):
        super().__init__(name=name)
        self.layers = tuple(layers)

    def __call__(self, inputs):
        out = inputs
        for layer in self.layers:
            out = layer(out)
        return out


def dropout(rng, rate, x):
    """basic.py dropout: keep = bernoulli(rng, 1 - rate, x.shape); keep * x / (1 - rate)."""
    if rate < 0 or rate >= 1:
        raise ValueError("rate")
    if rate == 0.0:
        return x
    keep_rate = 1.0 - rate
    keep = jax.random.bernoulli(rng, keep_rate, shape=x.shape)
    return keep * x / keep_rate


class Embed(Module):
    def __init__(self, vocab_size=None, embed_dim=None, name=None):
        super().__init__(name=name)
        self.vocab_size = vocab_size
        self.embed_dim = embed_dim

    def __call__(self, ids):
        emb = get_parameter("embeddings", [self.vocab_size, self.embed_dim])
        return emb[np.asarray(ids)]


# ---- conv.py -----------------------------------------------------------------------------------------------
class Conv1D(Module):
    """ConvND(num_spatial_dims=1), data_format NWC, kernel [K, C_in, C_out], stride 1, padding "SAME" (default)."""

    def __init__(self, output_channels, kernel_shape, stride=1, rate=1, padding="SAME", with_bias=True, name=None):
        super().__init__(name=name)
        assert stride == 1 and padding == "SAME"
        self.output_channels = output_channels
        self.k = int(kernel_shape)
        self.rate = rate
        self.with_bias = with_bias

    def __call__(self, inputs):
        k, d = self.k, self.rate
        cin = inputs.shape[-1]
        w = get_parameter("w", [k, cin, self.output_channels])
        total = (k - 1) * d                           # XLA "SAME": total padding split low = total // 2
        lo = total // 2
        x = np.pad(inputs, ((0, 0), (lo, total - lo), (0, 0)))
        T = inputs.shape[1]
        out = 0.0
        for j in range(k):
            out = out + jnp.einsum("bti,io->bto", x[:, j * d : j * d + T], w[j])
        if self.with_bias:
            out = out + get_parameter("b", [self.output_channels])
        return out


# ---- moving_averages.py / batch_norm.py --------------------------------------------------------------------
class ExponentialMovingAverage(Module):
    def __init__(self, decay, zero_debias=True, warmup_length=0, name=None):
        super().__init__(name=name)
        self.decay = decay

    @property
    def average(self):
        return self._average()

    def _average(self):
        return get_state("average")


class BatchNorm(Module):
    def __init__(self, create_scale, create_offset, decay_rate, eps=1e-5, name=None):
        super().__init__(name=name)
        self.create_scale = create_scale
        self.create_offset = create_offset
        self.eps = eps
        self.mean_ema = ExponentialMovingAverage(decay_rate, name="mean_ema")
        self.var_ema = ExponentialMovingAverage(decay_rate, name="var_ema")

    def __call__(self, inputs, is_training, test_local_stats=False):
        if is_training or test_local_stats:
            raise NotImplementedError("haiku shim: eval-mode BatchNorm only")
        mean = self.mean_ema.average
        var = self.var_ema.average
        w_shape = [1] * (inputs.ndim - 1) + [inputs.shape[-1]]
        scale = get_parameter("scale", w_shape) if self.create_scale else 1.0
        offset = get_parameter("offset", w_shape) if self.create_offset else 0.0
        inv = scale / jnp.sqrt(var + self.eps)
        return (inputs - mean) * inv + offset


# ---- recurrent.py ------------------------------------------------------------------------------------------
class LSTMState(NamedTuple):
    hidden: np.ndarray
    cell: np.ndarray


class RNNCore(Module):
    pass


class LSTM(RNNCore):
    def __init__(self, hidden_size, name=None):
        super().__init__(name=name)
        self.hidden_size = hidden_size

    def __call__(self, inputs, prev_state):
        x_and_h = jnp.concatenate([inputs, prev_state.hidden], axis=-1)
        gated = Linear(4 * self.hidden_size)(x_and_h)
        i, g, f, o = jnp.split(gated, indices_or_sections=4, axis=-1)
        f = jax.nn.sigmoid(f + 1)  # forget-gate bias
        c = f * prev_state.cell + jax.nn.sigmoid(i) * jnp.tanh(g)
        h = jax.nn.sigmoid(o) * jnp.tanh(c)
        return h, LSTMState(h, c)

    def initial_state(self, batch_size):
        z = jnp.zeros([batch_size, self.hidden_size], dtype=F64)
        return LSTMState(hidden=z, cell=z.copy())


class ResetCore(RNNCore):
    def __init__(self, core, name=None):
        super().__init__(name=name)
        self.core = core

    def __call__(self, inputs, state):
        inputs, should_reset = inputs
        initial = self.initial_state(np.asarray(should_reset).shape[0])

        def sel(s, i):
            r = np.asarray(should_reset).reshape(should_reset.shape + (1,) * (s.ndim - should_reset.ndim))
            return jnp.where(r, i, s)

        state = jax.tree_map(sel, state, initial)
        return self.core(inputs, state)

    def initial_state(self, batch_size):
        return self.core.initial_state(batch_size)


class _DeepRNN(RNNCore):
    def __init__(self, layers, skip_connections, name=None):
        super().__init__(name=name)
        self.layers = layers
        self.skip_connections = skip_connections

    def __call__(self, inputs, state):
        current_inputs = inputs
        next_states = []
        outputs = []
        state_idx = 0
        concat = lambda *args: jnp.concatenate(args, axis=-1)  # noqa: E731
        for idx, layer in enumerate(self.layers):
            if self.skip_connections and idx > 0:
                current_inputs = jax.tree_map(concat, inputs, current_inputs)
            if isinstance(layer, RNNCore):
                current_inputs, next_state = layer(current_inputs, state[state_idx])
                outputs.append(current_inputs)
                next_states.append(next_state)
                state_idx += 1
            else:
                current_inputs = layer(current_inputs)
        out = jax.tree_map(concat, *outputs) if self.skip_connections else current_inputs
        return out, tuple(next_states)

    def initial_state(self, batch_size):
        return tuple(layer.initial_state(batch_size) for layer in self.layers if isinstance(layer, RNNCore))


def deep_rnn_with_skip_connections(layers, name=None):
    return _DeepRNN(layers, skip_connections=True, name=name or "deep_rnn")


def dynamic_unroll(core, input_sequence, initial_state, time_major=True, reverse=False, return_all_states=False):
    """recurrent.py dynamic_unroll = hk.scan over the leading (time) axis.  The Haiku rng sequence is carried through
    the steps (stateful.py scan threads internal_state), i.e. keys are drawn exactly as if the loop were unrolled."""
    assert not reverse and not return_all_states
    if not time_major:
        input_sequence = jax.tree_map(lambda x: np.swapaxes(x, 0, 1), input_sequence)
    first = input_sequence
    while isinstance(first, (tuple, list)):
        first = first[0]
    T = first.shape[0]
    state = initial_state
    outs = []
    for t in range(T):
        x_t = jax.tree_map(lambda x: x[t], input_sequence)
        out, state = core(x_t, state)
        outs.append(out)
    output_sequence = jax.tree_map(lambda *xs: np.stack(xs, axis=0), *outs)
    if not time_major:
        output_sequence = jax.tree_map(lambda x: np.swapaxes(x, 0, 1), output_sequence)
    return output_sequence, state
