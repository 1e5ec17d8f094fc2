"""CPU: the C-ABI library loads and exports every symbol include/viettts_b200.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]


def _declared():
    text = (REPO / "include" / "viettts_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vtts_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from viettts_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    # and the ctypes table binds exactly the declared functions
    assert sorted(_lib.SIGNATURES) == names


def test_blob_sizes_agree_with_packers(hifigan_params, acoustic_ckpt):
    from viettts_b200 import _lib, weights
    lib = _lib.load()
    assert lib.vtts_version() == 1
    assert weights.pack_hifigan(hifigan_params).size == lib.vtts_hifigan_blob_floats() == 13_926_017
    assert weights.pack_acoustic(acoustic_ckpt).size == lib.vtts_acoustic_blob_floats()


def test_no_cpu_fallback():
    """Without a GPU, creating a context must fail loudly (no silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from viettts_b200 import _lib
    from viettts_b200.engine import Engine
    with pytest.raises(_lib.VttsError):
        Engine(0)


def test_product_does_not_import_oracle():
    for p in (REPO / "viettts_b200").rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p


def test_jax_free_unpickler_roundtrip(tmp_path, acoustic_ckpt):
    import pickle
    import numpy as np
    from viettts_b200 import weights
    f = tmp_path / "ck.pickle"
    with open(f, "wb") as fh:
        pickle.dump(acoustic_ckpt, fh)
    back = weights.load_pickle(f)
    assert np.array_equal(weights.pack_acoustic(back), weights.pack_acoustic(acoustic_ckpt))
