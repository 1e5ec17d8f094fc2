"""torch HiFiGAN generator checkpoint -> Haiku-layout parameter dict (hk_hifi.pickle).

Does what vietTTS/hifigan/convert_torch_model_to_haiku.py:27-62 does, without instantiating the torch model:
the checkpoint's `generator` state dict is folded (weight norm) and re-laid-out tensor by tensor.

  torch name                       Haiku module                               layout change
  conv_pre / conv_post             generator/~/conv1_d / conv1_d_1            Conv1d  [Cout,Cin,K] -> w[K,Cin,Cout]
  ups.i                            generator/~/ups_i                          ConvT1d [Cin,Cout,K] -> w[K,Cout,Cin], taps reversed
  resblocks.n.convs{1,2}.m         generator/~/res_block1_n/~/convs{1,2}_m    Conv1d
  *.bias                           b

Weight norm (torch.nn.utils.weight_norm, dim 0): w = v * (g / ||v||), the norm taken over all axes but the
first -- `remove_weight_norm()` in the reference (torch_model.py) materialises exactly this product.  Both the
classic `weight_g / weight_v` names and the parametrization names (`parametrizations.weight.original0/1`) are accepted.
"""
from __future__ import annotations

import pickle
from argparse import ArgumentParser
from pathlib import Path

import numpy as np

from .. import config


def _np(t) -> np.ndarray:
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.asarray(t, dtype=np.float32)


def fold_weight_norm(g, v) -> np.ndarray:
    v = _np(v)
    g = _np(g).reshape((v.shape[0],) + (1,) * (v.ndim - 1))
    norm = np.sqrt(np.sum(np.square(v.reshape(v.shape[0], -1)), axis=1, dtype=np.float32)).reshape(g.shape)
    return (v * (g / norm)).astype(np.float32)


def _haiku_name(torch_prefix: str, resblock: str = "1") -> str:
    parts = torch_prefix.split(".")
    if parts[0] == "conv_pre":
        return "generator/~/conv1_d"
    if parts[0] == "conv_post":
        return "generator/~/conv1_d_1"
    if parts[0] == "ups":
        return f"generator/~/ups_{parts[1]}"
    if parts[0] == "resblocks":
        return f"generator/~/res_block{resblock}_{parts[1]}/~/{parts[2]}_{parts[3]}"
    raise KeyError(f"unexpected generator tensor {torch_prefix!r}")


def state_dict_to_haiku(state_dict: dict, resblock: str = "1") -> dict:
    """`checkpoint["generator"]` (weight-normed or plain) -> {"generator/~/...": {"w": ..., "b": ...}}."""
    weights, biases = {}, {}
    pending = {}
    for name, t in state_dict.items():
        if name.endswith(".bias"):
            biases[name[: -len(".bias")]] = _np(t)
        elif name.endswith(".weight"):
            weights[name[: -len(".weight")]] = _np(t)
        elif name.endswith(".weight_g") or name.endswith(".parametrizations.weight.original0"):
            pending.setdefault(name.rsplit(".weight_g", 1)[0].rsplit(".parametrizations", 1)[0], {})["g"] = t
        elif name.endswith(".weight_v") or name.endswith(".parametrizations.weight.original1"):
            pending.setdefault(name.rsplit(".weight_v", 1)[0].rsplit(".parametrizations", 1)[0], {})["v"] = t
        else:
            raise KeyError(f"unexpected generator tensor {name!r}")
    for prefix, gv in pending.items():
        if set(gv) != {"g", "v"}:
            raise KeyError(f"{prefix}: weight norm needs both g and v")
        weights[prefix] = fold_weight_norm(gv["g"], gv["v"])
    out = {}
    for prefix, w in weights.items():
        name = _haiku_name(prefix, resblock)
        if prefix.startswith("ups."):
            w = np.transpose(w, (2, 1, 0))[::-1]      # [Cin,Cout,K] -> [K,Cout,Cin], tap order reversed
        else:
            w = np.transpose(w, (2, 1, 0))            # [Cout,Cin,K] -> [K,Cin,Cout]
        out[name] = {"w": np.ascontiguousarray(w), "b": biases[prefix]}
    return out


def convert_checkpoint(checkpoint_file, out_file=None) -> dict:
    """Read a torch generator checkpoint (`{"generator": state_dict}`) and write the pickle mel2wave reads."""
    import torch
    sd = torch.load(checkpoint_file, map_location="cpu")["generator"]
    hk = state_dict_to_haiku(sd, config.HIFIGAN["resblock"])
    out_file = Path(out_file) if out_file is not None else config.HIFIGAN_CKPT
    out_file.parent.mkdir(parents=True, exist_ok=True)
    with open(out_file, "wb") as f:
        pickle.dump(hk, f)
    return hk


def main(argv=None) -> int:
    p = ArgumentParser()
    p.add_argument("--checkpoint-file", required=True)
    p.add_argument("--config-file", default=None, help="generator config.json; checked against the built-in architecture")
    p.add_argument("--output", default=None, type=Path)
    a = p.parse_args(argv)
    if a.config_file:
        import json
        config.check_hifigan_config(json.loads(Path(a.config_file).read_text()))
    hk = convert_checkpoint(a.checkpoint_file, a.output)
    print(f"wrote {len(hk)} modules to {a.output or config.HIFIGAN_CKPT}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
