"""Launches the kernels of one workload once (for ncu captures).  MODE: step | gta | melspec | fp32."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from viettts_b200 import config as C, synthetic  # noqa: E402
from viettts_b200.engine import Engine  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "step"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
eng = Engine(0)
eng.load_hifigan(synthetic.hifigan_params(1234))
eng.load_acoustic(synthetic.acoustic_ckpt(1234))
B = 32
tokens, durs, nfs = bench.make_batch(B, 100, 5.0, 1000)
N = int(nfs.max())
if mode == "step":
    job = bench.Job(eng, dev, tokens, durs, nfs, 0xC0FFEE)
    for _ in range(reps):
        job.step()
elif mode == "gta":
    eng.load_mel_filterbank()
    wav_i16 = (np.random.default_rng(1).standard_normal((B, N * C.HOP)) * 3000).astype(np.int16)
    dur_sec = durs * np.float32(C.HOP / C.SAMPLE_RATE)
    for _ in range(reps):
        eng.gta(wav_i16, tokens, dur_sec, seed=7)
elif mode == "melspec":
    wav = torch.rand((512, 79872), dtype=torch.float32, device=dev) - 0.5
    for _ in range(reps):
        eng.melspec_forward(wav)
elif mode == "fp32":
    eng.set_precision("fp32")
    mel = torch.from_numpy(synthetic.mel_input(3, 8, 312)).to(dev)
    for _ in range(reps):
        eng.hifigan_forward(mel)
torch.cuda.synchronize()
print("done", mode)
