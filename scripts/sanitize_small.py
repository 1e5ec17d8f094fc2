"""Small invocations of the new kernels for compute-sanitizer (memcheck): fused pairs (both forms), the multi-group
decoder scan (40 rows = two row groups), the teacher-forced scan."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from viettts_b200 import synthetic
from viettts_b200.engine import Engine
eng = Engine(0)
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
for kind in ("smem2", "tmem"):
    eng.set_fused_pairs(False, kind=kind)
    for C, k, dil in ((64, 7, 3), (32, 11, 5)):
        x = torch.from_numpy(rng.standard_normal((2, 700, C)).astype(np.float32)).to(dev)
        w1 = torch.from_numpy((rng.standard_normal((k, C, C)) / np.sqrt(k * C)).astype(np.float32)).to(dev)
        w2 = torch.from_numpy((rng.standard_normal((k, C, C)) / np.sqrt(k * C)).astype(np.float32)).to(dev)
        b = torch.zeros(C, device=dev)
        lens = torch.tensor([700, 257], dtype=torch.int32, device=dev)
        out = eng.debug_pair(x, w1, b, w2, b, k, dil, 0.1, lens)
        assert torch.isfinite(out).all()
eng.load_acoustic(synthetic.acoustic_ckpt(1234))
eng.load_mel_filterbank()
B = 40
tok = np.stack([np.asarray(synthetic.utterance(i, 12, 0.3)[0], np.int32) for i in range(B)])
dur = np.stack([(synthetic.utterance(i, 12, 0.3)[1][0] * np.float32(62.5)) for i in range(B)]).astype(np.float32)
mel = eng.predict_mel(tok, dur, seed=3)
assert np.isfinite(mel).all()
m1, m2 = eng.teacher_forced(tok[:3], dur[:3], synthetic.mel_input(1, 3, 18), seed=5)
assert np.isfinite(m2).all()
print("sanitize_small ok", mel.shape)
