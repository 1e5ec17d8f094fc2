"""Stand-in for the slice of `jax` that vietTTS/nat/{model,text2mel,gta,dsp}.py touch.  See ../README.md."""
import functools

import numpy as _np

from . import nn, numpy, random  # noqa: F401


def jit(fun=None, static_argnums=None, **_kw):
    """jax.jit: tracing is irrelevant for values; the function runs eagerly."""
    if fun is None:
        return functools.partial(jit, static_argnums=static_argnums)
    return fun


def device_put(x):
    return _np.asarray(x)


def device_get(x):
    return _np.asarray(x)


def _is_leaf(x):
    return not isinstance(x, (tuple, list, dict)) or (isinstance(x, tuple) and False)


def tree_map(f, tree, *rest):
    """jax.tree_map over tuples / lists / namedtuples / dicts (leaves: arrays, scalars, None is a leaf-less node)."""
    if tree is None:
        return None
    if isinstance(tree, tuple) and hasattr(tree, "_fields"):   # namedtuple
        return type(tree)(*[tree_map(f, t, *[r[i] for r in rest]) for i, t in enumerate(tree)])
    if isinstance(tree, (tuple, list)):
        return type(tree)(tree_map(f, t, *[r[i] for r in rest]) for i, t in enumerate(tree))
    if isinstance(tree, dict):
        return {k: tree_map(f, v, *[r[k] for r in rest]) for k, v in tree.items()}
    return f(tree, *rest)


class tree_util:  # noqa: N801
    tree_map = staticmethod(tree_map)
