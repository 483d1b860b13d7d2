"""data_utils — the loss-head part of the reference's `data_utils.py` (:19-40): `STFT_ARGS`, `spectral`, `norm`,
plus the multi-scale spectral loss of `vqvae.py:309-326`.

SURVEY.md section 8f-1: the FFT itself is cuFFT (torch.fft.rfft / irfft on the device); everything around it — framing +
periodic Hann window + zero padding, magnitudes, Frobenius sums, the loss, its gradient with respect to the spectrum and
the overlap-add back onto the waveform — is libvqvae_b200 (csrc/spectral.cu), forward AND backward, with no autograd.
Semantics follow tf.signal.stft: frames of `window_length` samples every `hop_length`, periodic Hann window, zero padding
at the END up to `n_fft`, pad_end=False (torch.stft centres/pads differently and is NOT used)."""
from __future__ import annotations

import contextlib
import math

import torch

from . import ops
from .keras_compat import GradientTape, Scalar, convert_to_tensor, record

STFT_ARGS = [(2048, 1024, 512),  # n_fft
             (240, 120, 50),     # hop_length
             (1200, 600, 240)]   # window_size

def spectral(x, n_fft, hop_length, window_length):
    """|STFT|: x [..., T] -> [..., frames, n_fft//2 + 1]   (data_utils.py:25-30)"""
    x = convert_to_tensor(x)
    lead, T = x.shape[:-1], x.shape[-1]
    S = _rfft_frames(x.reshape(-1, T).contiguous(), n_fft, hop_length, window_length)
    return S.abs().reshape(*lead, S.shape[-2], S.shape[-1])


def norm(x):
    """Frobenius norm over the last two axes (data_utils.py:33-40)."""
    return torch.sqrt((x * x).sum(dim=(-2, -1)))


# |STFT| of the target, shared by the levels of ONE step.  The entry holds a strong reference to the tensor it was computed
# from and is valid only for that very object at that very version: a new batch is a new tensor object (its address may well be
# a recycled allocator block, which is why the address must not be the key), and the graph path's static input buffer changes
# version with every `copy_`.
_target_cache = {"ref": None, "version": None, "val": None}


def _rfft_frames(x2d, n_fft, hop, win):
    """complex64 [B, F, n_fft // 2 + 1]: cuFFT over the windowed frames produced by vqb_stft_frames."""
    return torch.fft.rfft(ops.stft_frames(x2d, n_fft, hop, win), dim=-1)


def _target_specs(src, t2):
    """per scale: (|S(target)| [B, F, bins]); and tsum [nscales, B] = their squared Frobenius norms.  `src` is the tensor object
    the caller was handed (identity + version are the cache key), `t2` its [B, T] view.  Both levels of the model compare
    against the same target (vqvae.py:119-127), so the second level of a step hits the cache."""
    c = _target_cache
    if c["ref"] is not src or c["version"] != src._version:
        mags, sums = [], []
        for n_fft, hop, win in zip(*STFT_ARGS):
            m, s_ = ops.spec_mag(_rfft_frames(t2, n_fft, hop, win))
            mags.append(m); sums.append(s_)
        c["ref"], c["version"], c["val"] = src, src._version, (mags, torch.stack(sums))
    return c["val"]


def clear_cache():
    _target_cache["ref"] = _target_cache["version"] = _target_cache["val"] = None


class MultiSpectralLoss:
    """Lazy per-example multi-scale spectral convergence loss (vqvae.py:309-326); `reduce_mean` evaluates it:
    mean over the batch of the mean over the three scales of ||S(x) - S(x_hat)||_F / ||S(x)||_F."""

    def __init__(self, target, recon):
        self.target, self.recon = convert_to_tensor(target), recon

    def _reduce_mean(self):
        r = self.recon
        B, T = r.shape[0], r.shape[1]
        t2, r2 = self.target.reshape(B, T).contiguous(), r.reshape(B, T).contiguous()
        mags, tsum = _target_specs(self.target, t2)
        scales = list(zip(*STFT_ARGS))
        taped = GradientTape.current() is not None
        dsum = ops.empty(len(scales), B)
        specs = []
        for s, (n_fft, hop, win) in enumerate(scales):
            S = _rfft_frames(r2, n_fft, hop, win)
            ops.spec_diff(S, mags[s], dsum[s])
            specs.append(S if taped else None)
        loss, coef = ops.spec_loss(dsum, tsum, taped)

        def bwd(g, needs):
            up = torch.full((1,), float(g[0]), dtype=torch.float32, device=r.device)
            dr = ops.empty(B, T)
            for s, (n_fft, hop, win) in enumerate(scales):
                G = ops.spec_grad(specs[s], mags[s], coef[s], up, n_fft)
                dfr = torch.fft.irfft(G, n=n_fft, dim=-1, norm="forward").contiguous()  # unnormalised C2R: no 1/n pass
                ops.stft_frames_bwd(dfr, T, hop, win, dr, s > 0)
            return [dr.reshape(r.shape)]

        record([r], [loss], bwd)
        return Scalar.leaf(loss)
