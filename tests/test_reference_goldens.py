"""CPU: the oracle restatements against golden vectors produced by executing the reference's own NAT / duration /
GTA / MelFilter source files (tests/golden/make_nat_golden.py, tests/refshim/README.md).

These pin the WIRING of oracle/nat_oracle.py and oracle/mel_oracle.py to vietTTS/nat/{model,text2mel,gta,dsp}.py:
float64 oracle vs float64 execution of the reference source, stored as float32 -> agreement to float32 storage
rounding (tolerance 2e-6 relative to values of O(10)).
"""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from oracle import mel_oracle, nat_oracle  # noqa: E402
from viettts_b200 import synthetic  # noqa: E402

G = REPO / "tests" / "golden"
TOL = 3e-6


def unpack(z, name):
    shape = tuple(int(s) for s in z[name + "_shape"])
    n = int(np.prod(shape))
    return np.unpackbits(z[name + "_bits"])[:n].reshape(shape)


@pytest.fixture(scope="module")
def ack():
    return synthetic.acoustic_ckpt(1234)


def test_predict_mel_matches_reference_source(ack):
    z = np.load(G / "nat_ref_predict_mel.npz")
    keep = unpack(z, "keep")
    mel = nat_oracle.predict_mel(ack, z["tokens"].tolist(), z["durations_sec"], keep, dtype=torch.float64)
    assert mel.shape == z["mel"][0].shape
    err = np.abs(mel - z["mel"][0]).max()
    assert err < TOL * 10, err
    # and the float32 oracle (what the GPU tests compare against) stays inside the stated mel tolerance
    mel32 = nat_oracle.predict_mel(ack, z["tokens"].tolist(), z["durations_sec"], keep, dtype=torch.float32)
    assert np.abs(mel32 - z["mel"][0]).max() < 1e-3


def test_inference_batch2_matches_reference_source(ack):
    z = np.load(G / "nat_ref_inference_b2.npz")
    keep = unpack(z, "keep")
    taps = {}
    mel = nat_oracle.inference(ack, z["tokens"], z["durations_frames"], int(z["n_frames"]), keep, dtype=torch.float64, taps=taps)
    assert np.abs(mel.numpy() - z["mel"]).max() < TOL * 10
    # hk.set_state("attn", w[0]) (model.py:109): the upsampling weights of row 0
    assert np.abs(taps["attn"][0].numpy() - z["attn"]).max() < TOL


def test_duration_matches_reference_source():
    z = np.load(G / "nat_ref_duration.npz")
    dk = synthetic.duration_ckpt(1234)
    d = nat_oracle.predict_duration(dk, z["tokens"].tolist(), dtype=torch.float64)
    assert np.abs(d - z["durations_sec"]).max() < TOL


def test_text2mel_chain_matches_reference_source(ack):
    """text2tokens -> predict_duration -> fix-ups -> predict_mel -> trailing-silence trim (text2mel.py:85-103)."""
    z = np.load(G / "nat_ref_text2mel.npz")
    dk = synthetic.duration_ckpt(1234)
    tokens = z["tokens"].tolist()
    # host front end of the product (pinned separately against the reference's own functions)
    from viettts_b200.nat.text2mel import text2tokens
    assert text2tokens(str(z["text"]), G / "lexicon_small.txt") == tokens
    d = nat_oracle.predict_duration(dk, tokens, dtype=torch.float64)
    d = nat_oracle.adjust_durations(tokens, d, float(z["silence_duration"]))
    keep = unpack(z, "keep")
    fr, n = nat_oracle.seconds_to_frames(d)
    assert n == int(z["n_frames_model"])
    mel = nat_oracle.inference(ack, np.asarray(tokens, np.int32)[None], fr, n, keep, dtype=torch.float64).numpy()
    mel = nat_oracle.trim_end_silence(tokens, d, mel)
    assert mel.shape == z["mel"].shape, (mel.shape, z["mel"].shape)
    assert np.abs(mel - z["mel"]).max() < 2e-4     # durations pass through float32 in adjust_durations (like the reference)


def test_melfilter_matches_reference_source():
    z = np.load(G / "nat_ref_gta.npz")
    wav = z["wav_i16"].astype(np.float32) / np.float32(2 ** 15)
    m = mel_oracle.mel_filter(wav, dtype=np.float64)
    assert np.abs(m - z["logmel"]).max() < 2e-5     # the reference's filterbank is rounded to float32 (librosa default dtype)
    m32 = mel_oracle.mel_filter(wav, dtype=np.float32)
    assert np.abs(m32 - z["logmel"]).max() < 5e-4


def test_gta_teacher_forced_matches_reference_source(ack):
    z = np.load(G / "nat_ref_gta.npz")
    keep, zone = unpack(z, "keep"), unpack(z, "zone")
    logmel, mel2 = nat_oracle.gta_forward(ack, z["wav_i16"], z["tokens"], z["lengths"], z["durations_sec"], keep, zone, dtype=torch.float64)
    assert np.abs(mel2 - z["mel2"]).max() < 2e-4    # the float32 oracle mel front end feeds it; see the next assert for exact wiring
    inp = np.concatenate([np.zeros_like(z["logmel"][:, :1]), z["logmel"][:, :-1]], axis=1)
    frames = (z["durations_sec"] * np.float32(16000)) / np.float32(256)
    mel1, mel2 = nat_oracle.teacher_forced(ack, z["tokens"], z["lengths"], frames, inp, keep, zone, dtype=torch.float64)
    assert np.abs(mel1 - z["mel1"]).max() < TOL * 10
    assert np.abs(mel2 - z["mel2"]).max() < TOL * 10


def test_product_mask_stream_equals_the_masks_the_reference_drew():
    """viettts_b200.jaxrng (what the drop-in predict_mel / GTA forward_fn feed the device by default) against the
    masks recorded while the reference's own code drew them from the checkpoint rng (hk.next_rng_key chain +
    jax.random.bernoulli, executed through tests/refshim)."""
    from viettts_b200 import jaxrng
    z = np.load(G / "nat_ref_predict_mel.npz")
    keep = unpack(z, "keep")
    assert np.array_equal(jaxrng.inference_keep_masks(z["rng"], 1, keep.shape[1]), keep)
    z = np.load(G / "nat_ref_inference_b2.npz")
    keep = unpack(z, "keep")
    assert np.array_equal(jaxrng.inference_keep_masks(z["rng"], 2, keep.shape[1]), keep)      # [B,256] draws of a batched call
    z = np.load(G / "nat_ref_gta.npz")
    keep, zone = unpack(z, "keep"), unpack(z, "zone")
    k2, z2 = jaxrng.teacher_forced_masks(z["rng"], keep.shape[0], keep.shape[1])
    assert np.array_equal(k2, keep) and np.array_equal(z2, zone)
    assert 0.08 < zone.mean() < 0.12 and 0.45 < keep.mean() < 0.55
