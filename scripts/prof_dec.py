import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from viettts_b200 import synthetic
from viettts_b200.engine import Engine
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench
eng = Engine(0)
eng.load_acoustic(synthetic.acoustic_ckpt(1234))
eng.set_precision("fp32")   # keep the tensor-core launches (which share the counter buffer) out of the way
for B in (1, 8, 32, 64):
    tokens, durs, nfs = bench.make_batch(B, 100, 5.0, 0)
    eng.predict_mel(tokens, durs, n_frames=nfs, seed=1)
    eng.tc_stats(True)
    eng.predict_mel(tokens, durs, n_frames=nfs, seed=1)
    ms = eng.last_stage_ms(1)
    raw = eng.tc_stats(False).astype(np.float64) / 1.9e3 / 312  # us per step
    names = ["EA", "sync", "B", "sync", "C", "sync", "D", "sync"]
    for role, sl in (("lstm", slice(0, 128)), ("prenet", slice(128, 144)), ("proj", slice(144, 148))):
        st = raw[sl, :8]
        print(f"B={B} {role:6s}: stage {ms:.2f} ms; us/step mean/max: " + " ".join(f"{n}={st[:, i].mean():.2f}/{st[:, i].max():.2f}" for i, n in enumerate(names)))
