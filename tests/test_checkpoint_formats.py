"""The NAT checkpoints (`acoustic_latest_ckpt.pickle`, `duration_latest_ckpt.pickle`, text2mel.py:27-28,62-71) are
pickles of Haiku FlatMappings whose leaves are jax arrays, plus an optax optimizer state.  `weights.load_pickle`
must read them WITHOUT jax / haiku / optax installed.  These tests write pickles with the layouts those libraries
produce (through stand-in modules that exist only while dumping) and load them after the modules are gone."""
import collections
import pickle
import sys
import types

import numpy as np
import pytest

from viettts_b200 import synthetic, weights


def _install(name, **attrs):
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        sub = ".".join(parts[:i])
        if sub not in sys.modules:
            sys.modules[sub] = types.ModuleType(sub)
    mod = sys.modules[name]
    for k, v in attrs.items():
        v.__module__ = name
        v.__qualname__ = k
        setattr(mod, k, v)
    return mod


@pytest.fixture
def fake_libs():
    before = set(sys.modules)

    def _reconstruct_array(fun, args, arr_state, aval_state):   # jax/_src/array.py: pickled by jax.Array.__reduce__
        raise AssertionError("must not run: the loader substitutes it")

    def reconstruct_device_array(fun, args, arr_state, aval_state):   # older location: jax/interpreters/xla.py
        raise AssertionError("must not run")

    class JaxArray:                         # what jax.Array / DeviceArray.__reduce__ emits
        _recon = None

        def __init__(self, a):
            self.a = np.asarray(a)

        def __reduce__(self):
            fun, args, arr_state = self.a.__reduce__()
            return (type(self)._recon, (fun, args, arr_state, {"weak_type": False, "named_shape": {}}))

    class FlatMapping(dict):                # haiku/_src/data_structures.py: __reduce__ goes through a plain dict
        def __reduce__(self):
            return (FlatMapping, (dict(self),))

    class OldFlatMapping:                   # default object protocol with the mapping in _mapping
        def __init__(self, m):
            self._mapping = dict(m)

    ScaleByAdamState = collections.namedtuple("ScaleByAdamState", "count mu nu")
    _install("jax._src.array", _reconstruct_array=_reconstruct_array)
    _install("jax.interpreters.xla", reconstruct_device_array=reconstruct_device_array)
    _install("haiku._src.data_structures", FlatMapping=FlatMapping, OldFlatMapping=OldFlatMapping)
    _install("optax._src.transform", ScaleByAdamState=ScaleByAdamState)
    yield dict(JaxArray=JaxArray, new=_reconstruct_array, old=reconstruct_device_array, FlatMapping=FlatMapping,
               OldFlatMapping=OldFlatMapping, Adam=ScaleByAdamState)
    for k in set(sys.modules) - before:
        del sys.modules[k]


def _dump_like_jax(ckpt, libs, recon, mapping_cls, path):
    libs["JaxArray"]._recon = recon
    J = libs["JaxArray"]

    def tree(d):
        return mapping_cls({mod: mapping_cls({k: J(v) for k, v in leaves.items()}) for mod, leaves in d.items()})

    dic = dict(step=123, params=tree(ckpt["params"]), aux=tree(ckpt["aux"]), rng=J(ckpt["rng"]),
               optim_state=(libs["Adam"](count=J(np.zeros((), np.int32)), mu=tree(ckpt["params"]), nu=tree(ckpt["params"])),))
    with open(path, "wb") as f:
        pickle.dump(dic, f)


@pytest.mark.parametrize("recon", ["new", "old"])
def test_jax_array_and_flatmapping_layout(tmp_path, fake_libs, acoustic_ckpt, recon):
    f = tmp_path / "acoustic_latest_ckpt.pickle"
    _dump_like_jax(acoustic_ckpt, fake_libs, fake_libs[recon], fake_libs["FlatMapping"], f)
    for k in [k for k in sys.modules if k.split(".")[0] in ("jax", "haiku", "optax")]:
        del sys.modules[k]                  # the libraries are NOT importable at load time
    with pytest.raises(Exception):
        pickle.load(open(f, "rb"))
    dic = weights.load_pickle(f)
    assert dic["step"] == 123 and type(dic["params"]) is dict
    assert isinstance(dic["rng"], np.ndarray) and dic["rng"].tolist() == acoustic_ckpt["rng"].tolist()
    assert np.array_equal(weights.pack_acoustic(dic), weights.pack_acoustic(acoustic_ckpt))
    leaf = dic["params"]["acoustic_model/~/linear"]["w"]
    assert isinstance(leaf, np.ndarray) and leaf.dtype == np.float32 and leaf.shape == (1024, 80)


def test_old_style_flatmapping_and_duration_checkpoint(tmp_path, fake_libs):
    ck = synthetic.duration_ckpt(1234)
    f = tmp_path / "duration_latest_ckpt.pickle"
    _dump_like_jax(ck, fake_libs, fake_libs["new"], fake_libs["OldFlatMapping"], f)
    # OldFlatMapping pickles under its own name; the loader keys on "FlatMap" in the class name
    for k in [k for k in sys.modules if k.split(".")[0] in ("jax", "haiku", "optax")]:
        del sys.modules[k]
    dic = weights.load_pickle(f)
    assert np.array_equal(weights.pack_duration(dic), weights.pack_duration(ck))


def test_plain_numpy_pickle_still_loads(tmp_path, hifigan_params):
    f = tmp_path / "hk_hifi.pickle"          # convert_torch_model_to_haiku.py:60-61 writes plain numpy dicts
    with open(f, "wb") as fh:
        pickle.dump(hifigan_params, fh)
    assert np.array_equal(weights.pack_hifigan(weights.load_pickle(f)), weights.pack_hifigan(hifigan_params))
