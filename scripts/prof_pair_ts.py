"""Times the fused ResBlock-pair kernels (tc_pair_ts.cu: A operand in TMEM; tc_pair.cu: A operand in smem) and the
two-launch path on generator-sized inputs with CUDA events (5 runs each, median)."""
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
    res = {}
    for name in ("smem2", "tmem", "smem"):
        eng.set_fused_pairs(False, kind=name)
        # debug_pair packs the weights on every call (cudaMalloc + pack kernel): subtract a k=1-sized call? no -- report it as is,
        # the pack kernels are ~20 us
        res[name] = timed(lambda: eng.debug_pair(x, w1, b1, w2, b2, k, dil))
    t2 = timed(lambda: (eng.debug_conv1d("bf16x3", x, w1, b1, k, dil, 0.1), eng.debug_conv1d("bf16x3", x, w2, b2, k, 1, 0.1, resid_t=x)))
    print(f"pair C={C} k={k:2d} d={dil}: pair2 {res['smem2']:.3f} ms ({flops/res['smem2']/1e9:.0f} TFLOP/s alg) | tmem-operand {res['tmem']:.3f} ({flops/res['tmem']/1e9:.0f}) | pair1 {res['smem']:.3f} "
          f"({flops/res['smem']/1e9:.0f}) | two launches {t2:.3f} ms ({flops/t2/1e9:.0f})", flush=True)
