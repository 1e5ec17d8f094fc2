"""Diagnostic probe of the tcgen05 conv kernel (run on the GPU box under `timeout`).
Prints where one-hot inputs land, so a descriptor / layout misinterpretation can be read off the
output instead of guessed."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from viettts_b200.engine import Engine  # noqa: E402

eng = Engine(0)
dev = torch.device("cuda", 0)


def run(x, w, b, k, dil, slope=1.0, resid=None, lens=None, prec="bf16x3"):
    xt = torch.from_numpy(x).to(dev)
    wt = torch.from_numpy(w).to(dev)
    bt = torch.from_numpy(b).to(dev)
    rt = None if resid is None else torch.from_numpy(resid).to(dev)
    lt = None if lens is None else torch.from_numpy(lens.astype(np.int32)).to(dev)
    out = eng.debug_conv1d(prec, xt, wt, bt, k, dil, slope, rt, lt)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def ref_conv(x, w, b, k, dil, slope=1.0, resid=None):
    xt = torch.from_numpy(x).double()
    if slope != 1.0:
        xt = torch.nn.functional.leaky_relu(xt, slope)
    wt = torch.from_numpy(w).double().permute(2, 1, 0).contiguous()
    y = torch.nn.functional.conv1d(xt.transpose(1, 2), wt, torch.from_numpy(b).double(), padding=(k - 1) * dil // 2, dilation=dil).transpose(1, 2)
    if resid is not None:
        y = y + torch.from_numpy(resid).double()
    return y.numpy()


print("== one-hot probes (k=1, identity weights): expect out[t0,c0] = value, everything else 0")
for C in (32, 256):
    T = 256
    w = np.eye(C, dtype=np.float32)[None]
    b = np.zeros(C, np.float32)
    for (t0, c0, val) in [(0, 0, 1.0), (1, 0, 1.0), (9, 0, 1.0), (0, 1, 1.0), (0, 8, 1.0), (5, 13, 1.0), (130, 17, 1.0), (3, 2, 1.0 + 2.0**-10)]:
        x = np.zeros((1, T, C), np.float32)
        x[0, t0, c0] = val
        y = run(x, w, b, 1, 1)
        nz = np.argwhere(np.abs(y[0]) > 1e-6)
        items = [(int(t), int(o), float(y[0, t, o])) for t, o in nz[:6]]
        ok = len(nz) == 1 and nz[0][0] == t0 and nz[0][1] == c0 and abs(y[0, t0, c0] - val) < 1e-7
        print(f"C={C} x[{t0},{c0}]={val}: {'OK' if ok else 'MISMATCH'} nonzeros={len(nz)} {items}")

print("== weight one-hot probes: x = ones at row t0 (all channels), w[0,i0,o0]=1 -> out[t0,o0]=1")
for C in (32, 128):
    T = 128
    for (i0, o0) in [(0, 0), (1, 0), (8, 0), (0, 1), (0, 9), (17, 30)]:
        w = np.zeros((1, C, C), np.float32)
        w[0, i0, o0] = 1.0
        x = np.zeros((1, T, C), np.float32)
        x[0, 7, :] = np.arange(1, C + 1, dtype=np.float32)
        y = run(x, w, np.zeros(C, np.float32), 1, 1)
        nz = np.argwhere(np.abs(y[0]) > 1e-6)
        items = [(int(t), int(o), float(y[0, t, o])) for t, o in nz[:6]]
        ok = len(nz) == 1 and nz[0][0] == 7 and nz[0][1] == o0 and abs(y[0, 7, o0] - (i0 + 1)) < 1e-6
        print(f"C={C} w[{i0},{o0}]: {'OK' if ok else 'MISMATCH'} nonzeros={len(nz)} {items}")

print("== tap probes: k=3/7/11 dil=1/3/5, x one-hot at t0, w[j,c,c] = j+1 -> out[t0 - (j*dil - pad), c0] = j+1")
for (k, dil) in [(3, 1), (3, 5), (7, 3), (11, 5)]:
    C, T = 64, 300
    w = np.zeros((k, C, C), np.float32)
    for j in range(k):
        w[j] = np.eye(C) * (j + 1)
    x = np.zeros((1, T, C), np.float32)
    x[0, 150, 5] = 1.0
    y = run(x, w, np.zeros(C, np.float32), k, dil)
    r = ref_conv(x, w, np.zeros(C, np.float32), k, dil)
    print(f"k={k} dil={dil}: max err {np.abs(y - r).max():.3e}; nonzero rows got {sorted(set(np.argwhere(np.abs(y[0]) > 1e-6)[:, 0].tolist()))[:12]} want {sorted(set(np.argwhere(np.abs(r[0]) > 1e-6)[:, 0].tolist()))[:12]}")

print("== random parity per (C,k,dil), lrelu + bias + residual, ragged lengths")
rng = np.random.default_rng(0)
worst = 0.0
for C in (32, 64, 128, 256):
    for (k, dil) in [(3, 1), (3, 3), (7, 5), (11, 1), (11, 5)]:
        B, T = 2, 700
        x = rng.standard_normal((B, T, C)).astype(np.float32)
        w = (rng.standard_normal((k, C, C)) / np.sqrt(k * C)).astype(np.float32)
        b = rng.standard_normal(C).astype(np.float32) * 0.1
        res = rng.standard_normal((B, T, C)).astype(np.float32)
        lens = np.array([T, 333])
        y = run(x, w, b, k, dil, 0.1, res, lens)
        y32 = run(x, w, b, k, dil, 0.1, res, lens, prec="fp32")
        e_t, e_s = 0.0, 0.0
        for bb in range(B):
            n = lens[bb]
            r = ref_conv(x[bb : bb + 1, :n], w, b, k, dil, 0.1, res[bb : bb + 1, :n])
            e_t = max(e_t, np.abs(y[bb, :n] - r[0]).max())
            e_s = max(e_s, np.abs(y32[bb, :n] - r[0]).max())
        worst = max(worst, e_t)
        print(f"C={C} k={k} dil={dil}: bf16x3 err {e_t:.3e}   fp32-simt err {e_s:.3e}")
print("WORST bf16x3 layer error", worst)
print("PROBE_DONE")
