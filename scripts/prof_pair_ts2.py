"""Role counters of tc_pair_ts_kernel (vtts_debug_tc_stats rows of 32 counters per CTA, first 128 CTAs)."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from viettts_b200.engine import Engine
eng = Engine(0)
dev = torch.device("cuda", 0)
import os
KIND = os.environ.get("VTTS_PAIR", "smem2")
eng.set_fused_pairs(False, kind=KIND)
names = {0: "issA total", 1: "issA wait acc", 2: "issA wait slots", 3: "issA wait W", 4: "issB total", 5: "issB wait acc", 6: "issB wait slots", 7: "issB wait W",
         8: "conv total", 9: "conv wait stage", 12: "rep1 total", 13: "rep1 wait operand", 14: "rep1 wait slot free", 15: "rep1 lds+st issue", 16: "rep1 wait::st",
         17: "rep2 total", 18: "rep2 wait operand", 19: "rep2 wait slot free", 20: "rep2 lds+st issue", 21: "rep2 wait::st",
         22: "E1 total", 23: "E1 wait D1", 24: "E1 wait A2 free", 25: "E2 total", 26: "E2 wait D2"}
for C, k, dil, rows in [(64, 7, 3, 1_277_952), (32, 7, 3, 2_555_904)]:
    B = 32
    T = rows // B
    x = torch.randn(B, T, C, device=dev)
    w1 = torch.randn(k, C, C, device=dev) / (k * C) ** 0.5
    w2 = torch.randn(k, C, C, device=dev) / (k * C) ** 0.5
    b1 = torch.randn(C, device=dev) * 0.1
    b2 = torch.randn(C, device=dev) * 0.1
    eng.debug_pair(x, w1, b1, w2, b2, k, dil)
    eng.tc_stats(True)
    eng.debug_pair(x, w1, b1, w2, b2, k, dil)
    st = eng.tc_stats(True).reshape(-1)[: 128 * 32].reshape(128, 32).astype(np.float64)
    tot = st[:, 0].mean()
    V = (128 if C == 64 else 256) - (k - 1)
    tiles = (T + V - 1) // V * B / 148
    print(f"pair C={C} k={k} d={dil}: {tot/1.9e3:.0f} us per CTA, {tot/tiles:.0f} clk per tile ({tiles:.0f} tiles per CTA)")
    print("   " + " | ".join(f"{names[i]} {st[:, i].mean()/tot*100:.0f}%" for i in names if i % 4 or i < 8 or True))
