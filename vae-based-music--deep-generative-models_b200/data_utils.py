"""data_utils — the loss-head part of the reference's `data_utils.py` (:19-40): `STFT_ARGS`, `spectral`, `norm`,
plus the multi-scale spectral loss of `vqvae.py:309-326`.

ROUND-1 STATUS (a "next" row): the STFT is evaluated with torch.fft (cuFFT) on the device and
differentiated by torch autograd; it is not yet a libvqvae_b200 kernel.  Semantics follow tf.signal.stft: frames of
`window_length` samples every `hop_length`, periodic Hann window, zero padding at the END up to `n_fft`,
pad_end=False (torch.stft centres/pads differently and is NOT used)."""
from __future__ import annotations

import contextlib
import math

import torch

from .keras_compat import GradientTape, Scalar, convert_to_tensor, record

STFT_ARGS = [(2048, 1024, 512),  # n_fft
             (240, 120, 50),     # hop_length
             (1200, 600, 240)]   # window_size

_windows = {}


def _hann(n, device):
    key = (n, str(device))
    if key not in _windows:
        i = torch.arange(n, dtype=torch.float64)
        _windows[key] = (0.5 - 0.5 * torch.cos(2.0 * math.pi * i / n)).to(torch.float32).to(device)
    return _windows[key]


def spectral(x, n_fft, hop_length, window_length):
    """|STFT|: x [..., T] -> [..., frames, n_fft//2 + 1]   (data_utils.py:25-30)"""
    frames = x.unfold(-1, window_length, hop_length) * _hann(window_length, x.device)
    return torch.fft.rfft(frames, n=n_fft, dim=-1).abs()


def norm(x):
    """Frobenius norm over the last two axes (data_utils.py:33-40)."""
    return torch.sqrt((x * x).sum(dim=(-2, -1)))


_target_cache = {"key": None, "val": None}


def _target_specs(t):
    key = (t.data_ptr(), t._version, tuple(t.shape))
    if _target_cache["key"] != key:
        with torch.no_grad():
            specs = []
            for n_fft, hop, win in zip(*STFT_ARGS):
                s = spectral(t, n_fft, hop, win)
                specs.append((s, norm(s)))
        _target_cache["key"], _target_cache["val"] = key, specs
    return _target_cache["val"]


def clear_cache():
    _target_cache["key"] = _target_cache["val"] = None


class MultiSpectralLoss:
    """Lazy per-example multi-scale spectral convergence loss (vqvae.py:309-326); `reduce_mean` evaluates it."""

    def __init__(self, target, recon):
        self.target, self.recon = convert_to_tensor(target), recon

    def _per_example(self, t, r):
        losses = []
        for (st, nt), (n_fft, hop, win) in zip(_target_specs(t), zip(*STFT_ARGS)):
            losses.append(norm(st - spectral(r, n_fft, hop, win)) / nt)
        return torch.stack(losses, dim=-1).mean(dim=-1)

    def _reduce_mean(self):
        r = self.recon
        t = self.target.squeeze(-1)
        taped = GradientTape.current() is not None
        with (torch.enable_grad() if taped else contextlib.nullcontext()):
            rl = r.detach().requires_grad_(taped)
            val = self._per_example(t, rl.squeeze(-1)).mean()
        loss = val.detach().reshape(1)

        def bwd(g, needs):
            (dr,) = torch.autograd.grad(val, rl, grad_outputs=torch.full_like(val, float(g[0])))
            return [dr]

        record([r], [loss], bwd)
        return Scalar.leaf(loss)
