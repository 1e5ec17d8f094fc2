import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from viettts_b200.engine import Engine
eng = Engine(0)
dev = torch.device("cuda", 0)
for C, k, dil, rows in [(64, 3, 1, 1_280_000), (64, 7, 3, 1_280_000), (64, 11, 5, 1_280_000), (32, 3, 1, 2_560_000), (32, 7, 3, 2_560_000), (32, 11, 5, 2_560_000)]:
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
    st = eng.tc_stats(True)[:148].astype(np.float64)
    tot = st[:, 0].mean()
    flops = 2 * 2.0 * rows * C * C * k
    gb = rows * C * 4 * 2 / 1e9
    n = {1: "mma_wait_acc", 2: "mma_wait_A1", 3: "mma_wait_W", 11: "mma_wait_A2", 4: "prod_wait", 5: "convA_wait", 6: "convA_fill", 9: "convB_wait", 10: "convB_fill", 7: "epi_wait", 8: "epi_busy"}
    print(f"pair C={C} k={k} d={dil}: {tot/1.9e3:.0f} us/CTA -> {flops/(tot/1.9e9)/1e12:.0f} TFLOP/s alg, {gb/(tot/1.9e9)/1e3:.2f} TB/s min-traffic")
    print("   " + " ".join(f"{v}={st[:, i].mean()/tot*100:.0f}%" for i, v in n.items()))
