"""CPU: the HiFiGAN oracle restatement is pinned against vectors produced by the
reference's own torch implementation (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import hifigan_oracle as ho
from viettts_b200 import synthetic


def _checksum(hk):
    s = 0.0
    for k in sorted(hk):
        for kk in sorted(hk[k]):
            a = hk[k][kk].astype(np.float64).ravel()
            s += float(np.dot(a, np.cos(np.arange(a.size) * 1e-3)))
    return s


def test_param_tree_matches_reference_layout(hifigan_params):
    # 1 + 4 + 72 + 1 modules, 13 926 017 parameters
    assert len(hifigan_params) == 78
    assert synthetic.n_params(hifigan_params) == 13_926_017
    assert hifigan_params["generator/~/ups_0"]["w"].shape == (16, 256, 512)
    assert hifigan_params["generator/~/res_block1_11/~/convs2_2"]["w"].shape == (11, 32, 32)


@pytest.mark.parametrize("tag", ["small", "t32"])
def test_oracle_matches_reference_torch_model(hifigan_params, golden_dir, tag):
    g = np.load(golden_dir / f"hifigan_ref_{tag}.npz")
    assert abs(_checksum(hifigan_params) - float(g["weight_checksum"])) < 1e-6 * abs(float(g["weight_checksum"])) + 1e-9
    taps = {}
    with torch.no_grad():
        wav = ho.generator_forward(hifigan_params, g["mel"], torch.float32, taps).numpy()
    assert wav.shape == g["wav"].shape
    err = np.abs(wav - g["wav"])
    # same math, same library conv kernels in a different layout: fp32 reassociation noise only
    assert err.max() < 2e-5, err.max()
    assert np.sqrt(np.mean(err**2)) < 2e-6
    if tag == "small":
        for k in ("pre", "ups_0", "stage_0"):
            e = np.abs(taps[k].numpy() - g[k]).max()
            assert e < 5e-5, (k, e)


def test_oracle_fp64_arbiter(hifigan_params, golden_dir):
    g = np.load(golden_dir / "hifigan_ref_small.npz")
    with torch.no_grad():
        w64 = ho.generator_forward(hifigan_params, g["mel"], torch.float64).numpy()
    assert np.abs(w64 - g["wav"]).max() < 2e-5


def test_conv_transpose_restatement_vs_torch():
    # hk.Conv1DTranspose("SAME") restatement == F.conv_transpose1d with the converter's rot90
    rng = np.random.default_rng(0)
    for (k, u, cin, cout) in [(16, 8, 12, 6), (4, 2, 8, 4)]:
        wt = rng.standard_normal((cin, cout, k)).astype(np.float32)   # torch layout
        b = rng.standard_normal(cout).astype(np.float32)
        x = rng.standard_normal((2, 9, cin)).astype(np.float32)
        hk_w = np.ascontiguousarray(np.rot90(wt, k=1, axes=(0, 2)))   # convert_...:53-54
        y = ho.conv1d_transpose_nwc(torch.from_numpy(x), torch.from_numpy(hk_w), torch.from_numpy(b), u)
        yt = torch.nn.functional.conv_transpose1d(torch.from_numpy(x).transpose(1, 2), torch.from_numpy(wt),
                                                  torch.from_numpy(b), stride=u, padding=(k - u) // 2).transpose(1, 2)
        assert y.shape == yt.shape == (2, 9 * u, cout)
        assert (y - yt).abs().max() < 1e-5
