"""Independent cross-checks of the dm-haiku semantics the NAT restatement relies on (oracle/nat_oracle.py).

The NAT oracle's WIRING is pinned by executing the reference's own source on numpy stand-ins for jax / dm-haiku
(tests/test_reference_goldens.py); the stand-ins and the oracle both restate the third-party primitives, so these
tests pin the primitives themselves: they do the next best
thing: every haiku building block it restates is compared with torch's own, independently written implementation
of the same operator, after mapping haiku's documented parameter layout onto torch's:
  hk.LSTM      gates [i, g, f, o] along the 4H axis, forget-gate bias +1 added at run time, z = [x, h] W + b
  hk.BatchNorm eval mode: (x - mean) * scale * rsqrt(var + 1e-5) + offset
  hk.Conv1D    NWC, w[K, Cin, Cout], SAME padding (odd K: (K-1)/2 zeros each side)
  hk.Embed     table lookup
  jax.nn.softmax / relu / gelu(approximate=True) / softplus
"""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import nat_oracle as no


def test_lstm_step_equals_torch_lstmcell():
    rng = np.random.default_rng(0)
    X, H, B = 24, 16, 5
    w = rng.standard_normal((X + H, 4 * H)) * 0.3
    b = rng.standard_normal(4 * H) * 0.1
    cell = torch.nn.LSTMCell(X, H).double()
    # haiku column blocks [i, g, f, o] -> torch row blocks [i, f, g, o]; haiku adds 1 to f at run time
    order = [0, 2, 1, 3]
    wi = np.concatenate([w[:X, k * H:(k + 1) * H] for k in order], axis=1).T
    wh = np.concatenate([w[X:, k * H:(k + 1) * H] for k in order], axis=1).T
    bb = np.concatenate([b[k * H:(k + 1) * H] + (1.0 if k == 2 else 0.0) for k in order])
    with torch.no_grad():
        cell.weight_ih.copy_(torch.from_numpy(wi))
        cell.weight_hh.copy_(torch.from_numpy(wh))
        cell.bias_ih.copy_(torch.from_numpy(bb))
        cell.bias_hh.zero_()
        x = torch.from_numpy(rng.standard_normal((B, X)))
        h = torch.from_numpy(rng.standard_normal((B, H)))
        c = torch.from_numpy(rng.standard_normal((B, H)))
        h_ref, c_ref = cell(x, (h, c))
        h_or, c_or = no.lstm_step(x, h, c, torch.from_numpy(w), torch.from_numpy(b))
    np.testing.assert_allclose(h_or.numpy(), h_ref.numpy(), atol=1e-12)
    np.testing.assert_allclose(c_or.numpy(), c_ref.numpy(), atol=1e-12)


def test_batchnorm_eval_equals_torch():
    rng = np.random.default_rng(1)
    C = 12
    P = {"bn": dict(scale=rng.standard_normal((1, 1, C)) + 1.0, offset=rng.standard_normal((1, 1, C)))}
    S = {"bn/~/mean_ema": dict(average=rng.standard_normal((1, 1, C))), "bn/~/var_ema": dict(average=rng.uniform(0.5, 2.0, (1, 1, C)))}
    x = torch.from_numpy(rng.standard_normal((3, 7, C)))
    got = no.batchnorm_eval(x, P, S, "bn", torch.float64)
    ref = F.batch_norm(x.transpose(1, 2), torch.from_numpy(S["bn/~/mean_ema"]["average"].ravel()),
                       torch.from_numpy(S["bn/~/var_ema"]["average"].ravel()), torch.from_numpy(P["bn"]["scale"].ravel()),
                       torch.from_numpy(P["bn"]["offset"].ravel()), training=False, eps=1e-5).transpose(1, 2)
    np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=1e-12)


def test_conv1d_same_is_a_plain_correlation():
    """hk.Conv1D does NOT flip the kernel: out[t, o] = sum_{k, i} x[t + k - (K-1)/2, i] w[k, i, o]."""
    rng = np.random.default_rng(2)
    B, T, Ci, Co, K = 2, 9, 3, 4, 5
    x = rng.standard_normal((B, T, Ci))
    w = rng.standard_normal((K, Ci, Co))
    b = rng.standard_normal(Co)
    got = no.conv1d_same(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b)).numpy()
    ref = np.zeros((B, T, Co))
    for t in range(T):
        for k in range(K):
            s = t + k - (K - 1) // 2
            if 0 <= s < T:
                ref[:, t] += x[:, s] @ w[k]
    ref += b
    np.testing.assert_allclose(got, ref, atol=1e-12)


def test_upsample_is_a_row_stochastic_gaussian_attention():
    rng = np.random.default_rng(3)
    x = torch.from_numpy(rng.standard_normal((2, 6, 4)))
    dur = torch.from_numpy(rng.uniform(0.0, 4.0, (2, 6)))
    out, w = no.upsample(x, dur, 15)
    assert out.shape == (2, 15, 4) and w.shape == (2, 15, 6)
    np.testing.assert_allclose(w.sum(-1).numpy(), 1.0, atol=1e-12)
    # explicit formula for one entry (model.py:103-108)
    end = np.cumsum(dur.numpy(), axis=1)
    mid = end - dur.numpy() / 2
    logits = -((mid[1] - 7.0) ** 2) / 10.0
    ref = np.exp(logits - logits.max())
    ref /= ref.sum()
    np.testing.assert_allclose(w[1, 7].numpy(), ref, atol=1e-12)
    # a zero-duration token still receives weight
    dur0 = dur.clone()
    dur0[0, 2] = 0.0
    _, w0 = no.upsample(x, dur0, 15)
    assert float(w0[0, :, 2].max()) > 0


def test_reset_core_semantics_of_the_backward_lstm(acoustic_ckpt):
    """hk.ResetCore zeroes the state BEFORE the step where the flag is set: with lengths = L only the first backward
    step is flagged (a no-op on the zero initial state), so the backward half must equal a plain reversed LSTM."""
    P, S = acoustic_ckpt["params"], acoustic_ckpt["aux"]
    rng = np.random.default_rng(4)
    tok = rng.integers(4, 90, size=(1, 9))
    out = no.token_encoder(P, S, tok, np.array([9]), torch.float64)
    # run the conv stack again and a hand-rolled reversed LSTM
    x = no._t(P[no.T + "embed"]["embeddings"], torch.float64)[torch.as_tensor(tok).long()]
    for i in range(3):
        sfx = "" if i == 0 else f"_{i}"
        cw = P[no.T + "conv1_d" + sfx]
        x = torch.relu(no.batchnorm_eval(no.conv1d_same(x, no._t(cw["w"], torch.float64), no._t(cw["b"], torch.float64)), P, S,
                                         no.T + "batch_norm" + sfx, torch.float64))
    wb, bb = no._t(P[no.T + "lstm_1/linear"]["w"], torch.float64), no._t(P[no.T + "lstm_1/linear"]["b"], torch.float64)
    h = x.new_zeros(1, 256)
    c = x.new_zeros(1, 256)
    for t in range(8, -1, -1):
        h, c = no.lstm_step(x[:, t], h, c, wb, bb)
        np.testing.assert_allclose(out[:, t, 256:].numpy(), h.numpy(), atol=1e-12)
    # shorter length: positions >= len-1 restart from zero state, i.e. equal a run that starts there
    out_s = no.token_encoder(P, S, tok, np.array([5]), torch.float64)
    h = x.new_zeros(1, 256)
    c = x.new_zeros(1, 256)
    h4, _ = no.lstm_step(x[:, 4], h, c, wb, bb)
    np.testing.assert_allclose(out_s[:, 4, 256:].numpy(), h4.numpy(), atol=1e-12)
    h8, _ = no.lstm_step(x[:, 8], h, c, wb, bb)      # every padded position is reset too
    np.testing.assert_allclose(out_s[:, 8, 256:].numpy(), h8.numpy(), atol=1e-12)
