"""Fused ResBlock-pair kernel: single-CTA form (smem2) vs CTA-pair form (smem2c), generator-sized inputs, CUDA events."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from viettts_b200.engine import Engine
eng = Engine(0)
dev = torch.device("cuda", 0)


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


for C, k, dil, rows in [(64, 3, 1, 1_277_952), (64, 7, 3, 1_277_952), (64, 11, 5, 1_277_952), (32, 3, 1, 2_555_904), (32, 7, 3, 2_555_904), (32, 11, 5, 2_555_904)]:
    B = 32
    T = rows // B
    x = torch.randn(B, T, C, device=dev)
    w1 = torch.randn(k, C, C, device=dev) / (k * C) ** 0.5
    w2 = torch.randn(k, C, C, device=dev) / (k * C) ** 0.5
    b1 = torch.randn(C, device=dev) * 0.1
    b2 = torch.randn(C, device=dev) * 0.1
    flops = 2 * 2.0 * rows * C * C * k
    res, out = {}, {}
    for name in ("smem2", "smem2c"):
        eng.set_fused_pairs(False, kind=name)
        out[name] = eng.debug_pair(x, w1, b1, w2, b2, k, dil)
        res[name] = timed(lambda: eng.debug_pair(x, w1, b1, w2, b2, k, dil))
    diff = float((out["smem2"] - out["smem2c"]).abs().max())
    print(f"pair C={C} k={k:2d} d={dil}: single-CTA {res['smem2']:.3f} ms ({flops/res['smem2']/1e9:.0f} TFLOP/s alg) | CTA pairs {res['smem2c']:.3f} ms "
          f"({flops/res['smem2c']/1e9:.0f})   max |diff| = {diff}", flush=True)
