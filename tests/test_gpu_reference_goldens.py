"""GPU parity against golden vectors produced by executing the reference's OWN source files
(vietTTS/nat/{model,text2mel,gta,dsp}.py through tests/refshim; tests/golden/make_nat_golden.py).

Unlike the oracle comparisons these do not depend on any restatement of the reference's wiring: inputs, the masks
the reference drew from the checkpoint rng, and its outputs are all in the fixtures.  Tolerances are the stated ones:
mel L-inf <= 1e-3 (log-mel units), durations <= 2e-5 s, log-mel of MelFilter <= 5e-4."""
import pickle

import numpy as np
import pytest

from viettts_b200 import jaxrng, synthetic

pytestmark = pytest.mark.gpu
MEL_LINF = 1e-3


def unpack(z, name):
    shape = tuple(int(s) for s in z[name + "_shape"])
    return np.unpackbits(z[name + "_bits"])[: int(np.prod(shape))].reshape(shape)


@pytest.fixture(scope="module", params=["fp32", "bf16x3"])
def eng(acoustic_ckpt, request):
    from viettts_b200.engine import Engine
    e = Engine(0)
    e.load_acoustic(acoustic_ckpt)
    e.load_duration(synthetic.duration_ckpt(1234))
    e.load_mel_filterbank()
    e.set_precision(request.param)
    yield e
    e.close()


def test_predict_mel_vs_reference_source(eng, golden_dir):
    z = np.load(golden_dir / "nat_ref_predict_mel.npz")
    keep = unpack(z, "keep")
    d = (z["durations_sec"] * np.float32(16000)) / np.float32(256)
    N = z["mel"].shape[1]
    mel = eng.predict_mel(z["tokens"][None], d, n_frames=[N], masks=keep)
    err = np.abs(mel - z["mel"]).max()
    print(f"predict_mel vs reference source: {err:.3e}")
    assert err < MEL_LINF
    # the masks the product derives from the checkpoint rng are the masks the reference drew
    assert np.array_equal(jaxrng.inference_keep_masks(z["rng"], 1, N), keep)


def test_inference_batch2_vs_reference_source(eng, golden_dir):
    z = np.load(golden_dir / "nat_ref_inference_b2.npz")
    keep = unpack(z, "keep")
    N = int(z["n_frames"])
    mel = eng.predict_mel(z["tokens"], z["durations_frames"], n_frames=[N, N], masks=keep)
    err = np.abs(mel - z["mel"]).max()
    print(f"inference B=2 vs reference source: {err:.3e}")
    assert err < MEL_LINF


def test_duration_vs_reference_source(eng, golden_dir):
    z = np.load(golden_dir / "nat_ref_duration.npz")
    d = eng.predict_duration(z["tokens"][None])
    err = np.abs(d - z["durations_sec"]).max()
    print(f"predict_duration vs reference source: {err:.3e}")
    assert err < 2e-5


def test_gta_vs_reference_source(eng, golden_dir):
    z = np.load(golden_dir / "nat_ref_gta.npz")
    keep, zone = unpack(z, "keep"), unpack(z, "zone")
    mel2, gt = eng.gta(z["wav_i16"], z["tokens"], z["durations_sec"], lengths=z["lengths"], keep_masks=keep, zone_masks=zone, return_gt=True)
    e_mel = np.abs(gt - z["logmel"]).max()
    e0 = np.abs(mel2[0] - z["mel2"][0]).max()
    print(f"GTA vs reference source: MelFilter {e_mel:.3e}  mel2 row 0 {e0:.3e}")
    assert e_mel < 5e-4
    assert e0 < 2e-3          # the GTA bar (float32 mel front end feeds the model)
    # row 1 is shorter than the padded L: the reference lets the padding tokens take part (encoder convs, upsampling
    # softmax); the product defines batched rows as "the row alone" (DESIGN.md §2), so only its frames that are at
    # least a receptive field away from the padding influence are compared
    assert np.abs(mel2[1, :20] - z["mel2"][1, :20]).max() < 5e-2
    # masks from the checkpoint rng == the masks the reference drew (whole-batch draws, state-tree order)
    k2, z2 = jaxrng.teacher_forced_masks(z["rng"], keep.shape[0], keep.shape[1])
    assert np.array_equal(k2, keep) and np.array_equal(z2, zone)


def test_dropin_callables_reproduce_reference_run(golden_dir, tmp_path, monkeypatch):
    """The drop-in `predict_mel(tokens, durations)` and `text2mel(text, lexicon, silence)` with NO extra arguments:
    checkpoint pickles on disk, masks from the checkpoint rng -- the call a user of the reference makes."""
    from viettts_b200.nat import text2mel as t2m
    for name, ck in (("acoustic_latest_ckpt.pickle", synthetic.acoustic_ckpt(1234)), ("duration_latest_ckpt.pickle", synthetic.duration_ckpt(1234))):
        with open(tmp_path / name, "wb") as f:
            pickle.dump(ck, f)
    monkeypatch.setattr(t2m, "CKPT_FILE", tmp_path / "acoustic_latest_ckpt.pickle")
    monkeypatch.setattr(t2m, "DURATION_CKPT_FILE", tmp_path / "duration_latest_ckpt.pickle")
    z = np.load(golden_dir / "nat_ref_predict_mel.npz")
    mel = t2m.predict_mel(z["tokens"].tolist(), z["durations_sec"])
    assert mel.shape == z["mel"].shape
    err = np.abs(mel - z["mel"]).max()
    print(f"drop-in predict_mel: {err:.3e}")
    assert err < MEL_LINF
    z = np.load(golden_dir / "nat_ref_text2mel.npz")
    mel = t2m.text2mel(str(z["text"]), golden_dir / "lexicon_small.txt", float(z["silence_duration"]))
    assert mel.shape == z["mel"].shape, (mel.shape, z["mel"].shape)
    err = np.abs(mel - z["mel"]).max()
    print(f"drop-in text2mel: {err:.3e}")
    assert err < 2e-3        # predicted durations (float32 on the device) feed the upsampler
