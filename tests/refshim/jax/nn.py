"""jax.nn activations used by vietTTS/nat/model.py (published definitions)."""
import numpy as np


def relu(x):
    return np.maximum(x, 0)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def softmax(x, axis=-1):
    e = np.exp(x - np.max(x, axis=axis, keepdims=True))
    return e / np.sum(e, axis=axis, keepdims=True)


def gelu(x, approximate=True):
    assert approximate
    return 0.5 * x * (1.0 + np.tanh(np.sqrt(2.0 / np.pi) * (x + 0.044715 * x ** 3)))


def softplus(x):
    return np.logaddexp(x, 0.0)


def leaky_relu(x, negative_slope=0.01):
    return np.where(x >= 0, x, negative_slope * x)
