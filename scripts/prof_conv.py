"""Single tensor-core conv launches with the per-role stall counters (and for ncu captures)."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from viettts_b200.engine import Engine
eng = Engine(0)
dev = torch.device("cuda", 0)
names = ["mma_total", "mma_wait_tmem", "mma_wait_A", "mma_wait_W", "prod_wait_slot", "conv_wait_slot", "conv_fill", "epi_wait_tmem", "epi_drain"]
import os
VAR = int(os.environ.get("TCV", "0"))
CASES = [(32, 7, 1, 2_560_000, True), (32, 7, 1, 2_560_000, False), (64, 7, 1, 1_280_000, True), (128, 7, 1, 640_000, True), (128, 7, 1, 640_000, False), (256, 7, 1, 80_000, True)]
if VAR == 3:
    CASES = [(32, 3, 1, 2_560_000, True), (32, 7, 1, 2_560_000, True), (32, 11, 1, 2_560_000, True), (64, 3, 1, 1_280_000, True), (64, 7, 1, 1_280_000, True), (64, 11, 1, 1_280_000, True)]
elif VAR:
    CASES = [(128, 3, 1, 640_000, True), (128, 7, 1, 640_000, True), (128, 7, 1, 640_000, False), (128, 11, 1, 640_000, True), (256, 3, 1, 160_000, True), (256, 7, 1, 160_000, True), (256, 7, 1, 160_000, False)]
for C, k, dil, rows, resid in CASES:
    B = 32
    T = rows // B
    x = torch.randn(B, T, C, device=dev)
    res = torch.randn(B, T, C, device=dev) if resid else None
    w = (torch.randn(k, C, C, device=dev) / (k * C) ** 0.5)
    b = torch.randn(C, device=dev) * 0.1
    eng.debug_conv1d("bf16x3", x, w, b, k, dil, 0.1, res, None)
    eng.tc_stats(True, VAR - 1 if VAR else None)
    eng.debug_conv1d("bf16x3", x, w, b, k, dil, 0.1, res, None)
    st = eng.tc_stats(True, VAR - 1 if VAR else None)[:148, :9].astype(np.float64)
    tot = st[:, 0].mean()
    flops = 2.0 * rows * C * C * k
    print(f"C={C} k={k} rows={rows} resid={resid}: mma-role total {tot/1.9e3:.1f} us/CTA -> {flops/ (tot/1.9e9)/1e12:.0f} TFLOP/s alg")
    print("   " + "  ".join(f"{n}={st[:, i].mean()/tot*100:.0f}%" for i, n in enumerate(names) if i))
