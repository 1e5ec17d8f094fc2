"""CTA-pair form (tc_variant 3) vs the single-CTA form (1) of the tensor-core conv: device time and per-role stall counters."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from viettts_b200.engine import Engine
eng = Engine(0)
dev = torch.device("cuda", 0)
names = ["mma_total", "mma_wait_tmem", "mma_wait_A", "mma_wait_W", "prod_wait_slot", "conv0_wait_slot", "conv0_fill", "epi_wait_tmem", "epi_drain", "conv1_wait_slot", "conv1_fill"]
for C, k, rows in [(128, 7, 640_000), (128, 3, 640_000), (256, 7, 160_000)]:
    B = 32
    T = rows // B
    x = torch.randn(B, T, C, device=dev)
    res = torch.randn(B, T, C, device=dev)
    w = (torch.randn(k, C, C, device=dev) / (k * C) ** 0.5)
    b = torch.randn(C, device=dev) * 0.1
    outs = {}
    for var in (1, 3):
        eng.tc_stats(True, var)
        outs[var] = eng.debug_conv1d("bf16x3", x, w, b, k, 1, 0.1, res, None)
        eng.tc_stats(True, var)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            eng.debug_conv1d("bf16x3", x, w, b, k, 1, 0.1, res, None)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        st = eng.tc_stats(True, var)[:148, :11].astype(np.float64) / 5
        lead = st[st[:, 0] > 0]
        tot = lead[:, 0].mean()
        flops = 2.0 * rows * C * C * k
        print(f"C={C} k={k} variant {var}: {ms*1e3:.0f} us per call (incl. debug copies), issuing CTAs {len(lead)}, mma-role total {tot/1.9e3:.1f} us -> {flops/(tot/1.9e9)/1e12:.0f} TFLOP/s alg")
        print("   issuers: " + "  ".join(f"{n}={lead[:, i].mean()/tot*100:.0f}%" for i, n in enumerate(names) if i))
        if var == 3:
            peer = st[1::2]
            print("   rank 1 : " + "  ".join(f"{n}={peer[:, i].mean()/tot*100:.0f}%" for i, n in enumerate(names) if i))
    print("   max |pair - single| =", float((outs[1] - outs[3]).abs().max()))
