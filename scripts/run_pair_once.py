"""One fused-pair launch per configuration (for ncu captures)."""
import os
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from viettts_b200.engine import Engine
eng = Engine(0)
dev = torch.device("cuda", 0)
eng.set_fused_pairs(False, kind=os.environ.get("VTTS_PAIR", "smem2"))
C, k, dil = int(os.environ.get("C", 64)), int(os.environ.get("K", 7)), int(os.environ.get("DIL", 3))
rows = 1_277_952 if C == 64 else 2_555_904
B = 32
T = rows // B
x = torch.randn(B, T, C, device=dev)
w1 = torch.randn(k, C, C, device=dev) / (k * C) ** 0.5
w2 = torch.randn(k, C, C, device=dev) / (k * C) ** 0.5
b1 = torch.randn(C, device=dev) * 0.1
b2 = torch.randn(C, device=dev) * 0.1
for _ in range(2):
    eng.debug_pair(x, w1, b1, w2, b2, k, dil)
torch.cuda.synchronize()
print("done")
