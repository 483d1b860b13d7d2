"""First-principles loop implementations (numpy, float64) used ONLY to pin `vqvae_oracle.py`.
TEST INFRASTRUCTURE.  Small shapes only.  Each function is written straight from the documented
TF-2.7 / Keras op definition, independently of the torch calls the oracle uses."""
from __future__ import annotations

import math

import numpy as np


def conv1d_naive(x, w, b, stride=1, dilation=1):
    """SAME cross-correlation: y[b,t,o] = bias[o] + sum_{j,i} x[b, t*s + j*d - left, i] w[j,i,o]."""
    B, L, Cin = x.shape
    k, _, Cout = w.shape
    out = -(-L // stride)
    pad = max((out - 1) * stride + (k - 1) * dilation + 1 - L, 0)
    left = pad // 2
    y = np.zeros((B, out, Cout), np.float64)
    for bb in range(B):
        for t in range(out):
            for j in range(k):
                u = t * stride + j * dilation - left
                if 0 <= u < L:
                    y[bb, t] += x[bb, u].astype(np.float64) @ w[j].astype(np.float64)
    return y + (0 if b is None else b.astype(np.float64))


def conv1d_transpose_naive(x, w, b, stride=2):
    """Gradient-of-SAME-conv definition: output length L*s, y[b, m*s + j - left, o] += x[b,m,i] w[j,o,i],
    left = max(k - s, 0)//2."""
    B, L, Cin = x.shape
    k, Cout, _ = w.shape
    left = max(k - stride, 0) // 2
    y = np.zeros((B, L * stride, Cout), np.float64)
    for bb in range(B):
        for m in range(L):
            for j in range(k):
                n = m * stride + j - left
                if 0 <= n < L * stride:
                    y[bb, n] += w[j].astype(np.float64) @ x[bb, m].astype(np.float64)
    return y + (0 if b is None else b.astype(np.float64))


def stft_mag_naive(x, n_fft, hop, win):
    """|tf.signal.stft|: frames of `win` samples every `hop`, periodic Hann, zero-padded at the END to
    n_fft, direct DFT."""
    B, T = x.shape
    n_frames = 1 + (T - win) // hop
    wdw = 0.5 - 0.5 * np.cos(2 * math.pi * np.arange(win) / win)
    nbin = n_fft // 2 + 1
    kk = np.arange(nbin)[:, None] * np.arange(win)[None, :]
    dft = np.exp(-2j * math.pi * kk / n_fft)
    out = np.zeros((B, n_frames, nbin))
    for bb in range(B):
        for f in range(n_frames):
            seg = x[bb, f * hop: f * hop + win].astype(np.float64) * wdw
            out[bb, f] = np.abs(dft @ seg)
    return out


def vq_indices_naive(flat, E):
    """argmin_k sum_d (x_d - E[d,k])^2 in float64, first minimum; also returns sorted top-2 distances."""
    x = flat.astype(np.float64)
    e = E.astype(np.float64)
    d = ((x[:, None, :] - e.T[None, :, :]) ** 2).sum(-1)
    idx = d.argmin(1)
    part = np.sort(d, axis=1)[:, :2]
    return idx, part


def adam_naive(p, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7):
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    return p - lr_t * m / (np.sqrt(v) + eps), m, v
