"""torch checkpoint -> Haiku layout converter (viettts_b200/hifigan/convert.py) against the output of the
reference's own convert_to_haiku (convert_torch_model_to_haiku.py:27-62) on a weight-normed checkpoint with random
gains; fixture: tests/golden/hifigan_converter_ref.npz, written by tests/golden/make_golden.py."""
import pickle

import numpy as np
import pytest

from viettts_b200 import synthetic, weights
from viettts_b200.hifigan import convert


@pytest.fixture(scope="module")
def fx(golden_dir):
    return dict(np.load(golden_dir / "hifigan_converter_ref.npz"))


def _torch_sd(fx):
    return {k[len("torch/"):]: v for k, v in fx.items() if k.startswith("torch/")}


def test_matches_reference_converter(fx):
    hk = convert.state_dict_to_haiku(_torch_sd(fx))
    names = {k[len("haiku/"):].rsplit("/", 1)[0] for k in fx if k.startswith("haiku/")}
    assert set(hk) == names
    for name in names:
        w_ref, b_ref = fx[f"haiku/{name}/w"], fx[f"haiku/{name}/b"]
        assert hk[name]["w"].shape == w_ref.shape and hk[name]["w"].dtype == np.float32
        assert np.array_equal(hk[name]["b"], b_ref)
        # the layout must be exact; the weight-norm product may differ by the rounding of |v| (<= 2 ulp)
        np.testing.assert_allclose(hk[name]["w"], w_ref, rtol=3e-7, atol=0)


def test_parametrization_names_and_plain_weights(fx):
    sd = _torch_sd(fx)
    alt = {}
    for k, v in sd.items():
        alt[k.replace(".weight_g", ".parametrizations.weight.original0").replace(".weight_v", ".parametrizations.weight.original1")] = v
    a, b = convert.state_dict_to_haiku(sd), convert.state_dict_to_haiku(alt)
    assert all(np.array_equal(a[k]["w"], b[k]["w"]) for k in a)
    plain = {"ups.3.weight": convert.fold_weight_norm(sd["ups.3.weight_g"], sd["ups.3.weight_v"]), "ups.3.bias": sd["ups.3.bias"]}
    c = convert.state_dict_to_haiku(plain)
    assert np.array_equal(c["generator/~/ups_3"]["w"], a["generator/~/ups_3"]["w"])
    with pytest.raises(KeyError):
        convert.state_dict_to_haiku({"ups.3.weight_g": sd["ups.3.weight_g"], "ups.3.bias": sd["ups.3.bias"]})
    with pytest.raises(KeyError):
        convert.state_dict_to_haiku({"mpd.0.weight": np.zeros((2, 2, 2), np.float32)})


def test_full_generator_round_trip(tmp_path):
    """Synthetic Haiku params -> torch layout -> converter -> pickle -> pack_hifigan: bit-identical blob."""
    import torch
    hk = synthetic.hifigan_params(1234)
    sd = {}
    for name, d in hk.items():
        short = name.split("generator/~/")[1]
        if short == "conv1_d":
            key, up = "conv_pre", False
        elif short == "conv1_d_1":
            key, up = "conv_post", False
        elif short.startswith("ups_"):
            key, up = f"ups.{short[4:]}", True
        else:
            rb, cv = short.split("/~/")
            key, up = f"resblocks.{rb.split('_')[-1]}.{cv.replace('_', '.')}", False
        w = d["w"][::-1] if up else d["w"]
        sd[key + ".weight"] = torch.from_numpy(np.ascontiguousarray(np.transpose(w, (2, 1, 0))))
        sd[key + ".bias"] = torch.from_numpy(d["b"].copy())
    ck = tmp_path / "g_synth"
    torch.save({"generator": sd}, ck)
    out = tmp_path / "hk_hifi.pickle"
    assert convert.main(["--checkpoint-file", str(ck), "--output", str(out)]) == 0
    with open(out, "rb") as f:
        back = pickle.load(f)
    assert np.array_equal(weights.pack_hifigan(back), weights.pack_hifigan(hk))
