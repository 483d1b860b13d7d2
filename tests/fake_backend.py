"""TEST DOUBLE for libvqvae_b200.so: the same `vqb_*` entry points (include/vqb.h) evaluated on CPU memory with the
oracle, so that the package's HOST logic (tape, variable packing, Keras `training` resolution, metric plumbing,
data-parallel sharding arithmetic over gloo) can be unit-tested in a container without a GPU.

It is installed only by tests (`vqvae_b200._lib.set_backend(FakeBackend(), "cpu")`); the product never imports it
and raises without the CUDA library.  No parity claim is made from tests that use it."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from oracle import vqvae_oracle as O


def _t(ptr, shape, dtype=np.float32):
    if ptr is None:
        return None
    n = int(np.prod(shape)) if len(shape) else 1
    ct = {np.float32: C.c_float, np.int64: C.c_int64, np.uint8: C.c_uint8, np.int32: C.c_int32}[dtype]
    a = np.frombuffer((ct * max(n, 1)).from_address(int(ptr)), dtype=dtype)[:n].reshape(shape)
    return torch.from_numpy(a)


def _d(ref):
    return ref._obj


def _mix(z):
    m = (1 << 64) - 1
    z = (z + 0x9E3779B97F4A7C15) & m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    return z ^ (z >> 31)


class FakeBackend:
    def __init__(self):
        self.calls = []
        self._err = b""

    # ---- misc
    def vqb_version(self):
        return 100

    def vqb_kernel_launch_count(self):
        return 0

    def vqb_last_error(self):
        return self._err

    def vqb_device_check(self, dev):
        return 0

    # ---- conv
    @staticmethod
    def _act(x, d):
        return torch.relu(x) if d.relu_in else x

    def vqb_reduce_begin(self):
        return 0

    def vqb_reduce_flush(self, stream):
        return 0

    def vqb_conv1d_supports(self, dref, op):
        return 1

    def vqb_conv1d_transpose_supports(self, dref, op):
        return 1

    def vqb_conv1d_fwd(self, dref, x, w, b, res, y, stream):
        d = _d(dref)
        Lo = -(-d.L // d.stride)
        X = _t(x, (d.B, d.L, d.C_in)); W = _t(w, (d.k, d.C_in, d.C_out)); Bv = _t(b, (d.C_out,))
        out = O.conv1d(self._act(X, d), W, Bv, d.stride, d.dilation)
        if res is not None:
            out = out + _t(res, (d.B, Lo, d.C_out))
        _t(y, (d.B, Lo, d.C_out)).copy_(out)
        return 0

    def vqb_conv1d_dgrad(self, dref, dy, w, x, dx_add, dx, stream):
        d = _d(dref)
        Lo = -(-d.L // d.stride)
        W = _t(w, (d.k, d.C_in, d.C_out)); DY = _t(dy, (d.B, Lo, d.C_out))
        xr = torch.zeros(d.B, d.L, d.C_in, requires_grad=True)
        (g,) = torch.autograd.grad(O.conv1d(xr, W, None, d.stride, d.dilation), xr, DY)
        if d.relu_in:
            g = g * (_t(x, (d.B, d.L, d.C_in)) > 0)
        if dx_add is not None:
            g = g + _t(dx_add, (d.B, d.L, d.C_in))
        _t(dx, (d.B, d.L, d.C_in)).copy_(g)
        return 0

    def vqb_conv1d_wgrad_workspace_bytes(self, dref):
        return 64

    def vqb_conv1d_wgrad(self, dref, x, dy, dw, db, ws, wsn, stream):
        d = _d(dref)
        Lo = -(-d.L // d.stride)
        X = self._act(_t(x, (d.B, d.L, d.C_in)), d); DY = _t(dy, (d.B, Lo, d.C_out))
        wr = torch.zeros(d.k, d.C_in, d.C_out, requires_grad=True)
        (g,) = torch.autograd.grad(O.conv1d(X, wr, None, d.stride, d.dilation), wr, DY)
        _t(dw, (d.k, d.C_in, d.C_out)).copy_(g)
        if db is not None:
            _t(db, (d.C_out,)).copy_(DY.sum((0, 1)))
        return 0

    def vqb_conv1d_transpose_fwd(self, dref, x, w, b, y, stream):
        d = _d(dref)
        X = _t(x, (d.B, d.L, d.C_in)); W = _t(w, (d.k, d.C_out, d.C_in)); Bv = _t(b, (d.C_out,))
        _t(y, (d.B, d.L * d.stride, d.C_out)).copy_(O.conv1d_transpose(X, W, Bv, d.stride))
        return 0

    def vqb_conv1d_transpose_dgrad(self, dref, dy, w, dx, stream):
        d = _d(dref)
        W = _t(w, (d.k, d.C_out, d.C_in)); DY = _t(dy, (d.B, d.L * d.stride, d.C_out))
        xr = torch.zeros(d.B, d.L, d.C_in, requires_grad=True)
        (g,) = torch.autograd.grad(O.conv1d_transpose(xr, W, None, d.stride), xr, DY)
        _t(dx, (d.B, d.L, d.C_in)).copy_(g)
        return 0

    def vqb_conv1d_transpose_wgrad_workspace_bytes(self, dref):
        return 64

    def vqb_conv1d_transpose_wgrad(self, dref, x, dy, dw, db, ws, wsn, stream):
        d = _d(dref)
        X = _t(x, (d.B, d.L, d.C_in)); DY = _t(dy, (d.B, d.L * d.stride, d.C_out))
        wr = torch.zeros(d.k, d.C_out, d.C_in, requires_grad=True)
        (g,) = torch.autograd.grad(O.conv1d_transpose(X, wr, None, d.stride), wr, DY)
        _t(dw, (d.k, d.C_out, d.C_in)).copy_(g)
        if db is not None:
            _t(db, (d.C_out,)).copy_(DY.sum((0, 1)))
        return 0

    # ---- decoder tail (the two layers evaluated one after the other with the oracle)
    def vqb_dec_tail_supports(self, dref):
        d = _d(dref)
        return int(4 <= d.C_in <= 32 and d.C_in % 4 == 0)

    def vqb_dec_tail_fwd(self, dref, x, wt, bt, wf, bf, gbuf, recon, stream):
        d = _d(dref)
        X = _t(x, (d.B, d.L, d.C_in)); Wt = _t(wt, (4, d.C_mid, d.C_in)); Wf = _t(wf, (3, d.C_mid, 1))
        y = O.conv1d_transpose(X, Wt, _t(bt, (d.C_mid,)), 2)
        _t(recon, (d.B, 2 * d.L, 1)).copy_(O.conv1d(y, Wf, _t(bf, (1,)), 1, 1))
        return 0

    def vqb_dec_tail_bwd_workspace_bytes(self, dref):
        return 64

    def vqb_dec_tail_bwd(self, dref, x, dr, wt, bt, wf, gbuf, dx, dwt, dbt, dwf, dbf, ws, wsn, stream):
        d = _d(dref)
        X = _t(x, (d.B, d.L, d.C_in)).clone().requires_grad_(True)
        Wt = _t(wt, (4, d.C_mid, d.C_in)).clone().requires_grad_(True)
        Wf = _t(wf, (3, d.C_mid, 1)).clone().requires_grad_(True)
        Bt = (_t(bt, (d.C_mid,)).clone() if bt is not None else torch.zeros(d.C_mid)).requires_grad_(True)
        Bf = torch.zeros(1, requires_grad=True)  # the bias value does not enter any gradient
        r = O.conv1d(O.conv1d_transpose(X, Wt, Bt, 2), Wf, Bf, 1, 1)
        gx, gwt, gwf, gbt, gbf = torch.autograd.grad(r, (X, Wt, Wf, Bt, Bf), _t(dr, (d.B, 2 * d.L, 1)))
        if dx is not None:
            _t(dx, (d.B, d.L, d.C_in)).copy_(gx)
        _t(dwt, (4, d.C_mid, d.C_in)).copy_(gwt); _t(dwf, (3, d.C_mid, 1)).copy_(gwf)
        if dbt is not None:
            _t(dbt, (d.C_mid,)).copy_(gbt)
        if dbf is not None:
            _t(dbf, (1,)).copy_(gbf)
        return 0

    # ---- spectral loss pieces (torch restatement of csrc/spectral.cu)
    @staticmethod
    def _hann(win):
        i = torch.arange(win, dtype=torch.float64)
        return (0.5 - 0.5 * torch.cos(2.0 * np.pi * i / win)).float()

    def vqb_stft_frames(self, x, B, T, n_fft, hop, win, frames, stream):
        F = 1 + (T - win) // hop
        fr = _t(x, (B, T)).unfold(-1, win, hop) * self._hann(win)
        out = _t(frames, (B, F, n_fft))
        out.zero_()
        out[..., :win] = fr
        return 0

    def vqb_spec_workspace_bytes(self, B, per):
        return 64

    def vqb_spec_mag(self, S, B, per, mag, sums, ws, wsn, stream):
        s = _t(S, (B, per, 2))
        m = torch.sqrt(s[..., 0] ** 2 + s[..., 1] ** 2)
        _t(mag, (B, per)).copy_(m); _t(sums, (B,)).copy_((m * m).sum(1))
        return 0

    def vqb_spec_diff(self, S, mag_t, B, per, sums, ws, wsn, stream):
        s = _t(S, (B, per, 2))
        d = _t(mag_t, (B, per)) - torch.sqrt(s[..., 0] ** 2 + s[..., 1] ** 2)
        _t(sums, (B,)).copy_((d * d).sum(1))
        return 0

    def vqb_spec_loss(self, dsum, tsum, nscales, B, loss, coef, stream):
        nd, nt = torch.sqrt(_t(dsum, (nscales, B))), torch.sqrt(_t(tsum, (nscales, B)))
        _t(loss, (1,)).copy_((nd / nt).mean().reshape(1))
        if coef is not None:
            _t(coef, (nscales, B)).copy_(torch.where(nd > 0, 1.0 / (nscales * B * nt * nd), torch.zeros_like(nd)))
        return 0

    def vqb_spec_grad(self, S, mag_t, coef, upstream, B, per, bins, n_fft, G, stream):
        s = _t(S, (B, per, 2)); mt = _t(mag_t, (B, per))
        m = torch.sqrt(s[..., 0] ** 2 + s[..., 1] ** 2)
        k = torch.arange(per) % bins
        w = torch.where((k == 0) | (k == bins - 1), 1.0, 0.5)  # for the unnormalised inverse FFT (include/vqb.h)
        c = _t(upstream, (1,))[0] * _t(coef, (B,))[:, None] * (m - mt) / torch.where(m > 0, m, torch.ones_like(m)) * w
        c = torch.where(m > 0, c, torch.zeros_like(c))
        _t(G, (B, per, 2)).copy_(c[..., None] * s)
        return 0

    def vqb_stft_frames_bwd(self, dframes, B, T, n_fft, hop, win, accumulate, dx, stream):
        F = 1 + (T - win) // hop
        d = _t(dframes, (B, F, n_fft))[..., :win] * self._hann(win)
        out = torch.zeros(B, T)
        for f in range(F):
            out[:, f * hop:f * hop + win] += d[:, f]
        tgt = _t(dx, (B, T))
        tgt.copy_(tgt + out if accumulate else out)
        return 0

    # ---- resblock
    def vqb_resblock_supports(self, dref):
        return 1

    def vqb_resblock_fwd(self, dref, x, w1, b1, w2, b2, h, y, stream):
        d = _d(dref)
        X = _t(x, (d.B, d.L, d.C))
        H = O.conv1d(torch.relu(X), _t(w1, (3, d.C, d.F)), _t(b1, (d.F,)), 1, d.dilation)
        if h is not None:
            _t(h, (d.B, d.L, d.F)).copy_(H)
        _t(y, (d.B, d.L, d.C)).copy_(X + O.conv1d(torch.relu(H), _t(w2, (3, d.F, d.C)), _t(b2, (d.C,)), 1, 1))
        return 0

    def vqb_resblock_fwd_masks(self, dref, x, w1, b1, w2, b2, h, y, xbits, hbits, stream):
        d = _d(dref)
        self.vqb_resblock_fwd(dref, x, w1, b1, w2, b2, h, y, stream)
        w = (1 << torch.arange(32, dtype=torch.int64))
        for src, C_, dst in ((x, d.C, xbits), (h, d.F, hbits)):
            bits = ((_t(src, (d.B, d.L, C_)) > 0).to(torch.int64) * w[:C_]).sum(-1)
            bits = torch.where(bits >= 2 ** 31, bits - 2 ** 32, bits).to(torch.int32)
            _t(dst, (d.B, d.L), np.int32).copy_(bits)
        return 0

    def vqb_resblock_bwd_data_masks(self, dref, xbits, hbits, dy, w1, w2, dh, dx, stream):
        d = _d(dref)
        sh = torch.arange(32, dtype=torch.int64)
        XM = ((_t(xbits, (d.B, d.L), np.int32).to(torch.int64).unsqueeze(-1) >> sh[:d.C]) & 1).bool()
        HM = ((_t(hbits, (d.B, d.L), np.int32).to(torch.int64).unsqueeze(-1) >> sh[:d.F]) & 1).bool()
        DY = _t(dy, (d.B, d.L, d.C)); W1 = _t(w1, (3, d.C, d.F)); W2 = _t(w2, (3, d.F, d.C))
        hr = torch.zeros(d.B, d.L, d.F, requires_grad=True)
        (g2,) = torch.autograd.grad(O.conv1d(hr, W2, None, 1, 1), hr, DY)
        DH = g2 * HM
        xr = torch.zeros(d.B, d.L, d.C, requires_grad=True)
        (g1,) = torch.autograd.grad(O.conv1d(xr, W1, None, 1, d.dilation), xr, DH)
        _t(dh, (d.B, d.L, d.F)).copy_(DH)
        _t(dx, (d.B, d.L, d.C)).copy_(g1 * XM + DY)
        return 0

    def vqb_resblock_wgrad_workspace_bytes(self, dref):
        return 16

    def vqb_resblock_wgrad(self, dref, x, h, dy, dh, dw1, db1, dw2, db2, ws, wsn, stream):
        d = _d(dref)
        X = torch.relu(_t(x, (d.B, d.L, d.C))); H = torch.relu(_t(h, (d.B, d.L, d.F)))
        DY = _t(dy, (d.B, d.L, d.C)); DH = _t(dh, (d.B, d.L, d.F))
        w1 = torch.zeros(3, d.C, d.F, requires_grad=True)
        (g1,) = torch.autograd.grad(O.conv1d(X, w1, None, 1, d.dilation), w1, DH)
        w2 = torch.zeros(3, d.F, d.C, requires_grad=True)
        (g2,) = torch.autograd.grad(O.conv1d(H, w2, None, 1, 1), w2, DY)
        _t(dw1, (3, d.C, d.F)).copy_(g1); _t(dw2, (3, d.F, d.C)).copy_(g2)
        if db1 is not None:
            _t(db1, (d.F,)).copy_(DH.sum((0, 1)))
        if db2 is not None:
            _t(db2, (d.C,)).copy_(DY.sum((0, 1)))
        return 0

    def vqb_resblock_wgrad_batch_workspace_bytes(self, dref, n):
        return 16

    def vqb_resblock_wgrad_batch(self, dref, n, dils, x, h, dy, dh, dw1, db1, dw2, db2, ws, wsn, stream):
        import copy
        P = C.POINTER(C.c_void_p)
        arr = lambda a: C.cast(a, P)
        dl = C.cast(dils, C.POINTER(C.c_int32))
        for i in range(n):
            di = copy.copy(dref._obj)
            di.dilation = dl[i]
            rc = self.vqb_resblock_wgrad(C.byref(di), *[arr(a)[i] for a in (x, h, dy, dh, dw1, db1, dw2, db2)], ws, wsn, stream)
            if rc:
                return rc
        return 0

    def vqb_resblock_bwd_data(self, dref, x, h, dy, w1, w2, dh, dx, stream):
        d = _d(dref)
        X = _t(x, (d.B, d.L, d.C)); H = _t(h, (d.B, d.L, d.F)); DY = _t(dy, (d.B, d.L, d.C))
        W1 = _t(w1, (3, d.C, d.F)); W2 = _t(w2, (3, d.F, d.C))
        hr = torch.zeros(d.B, d.L, d.F, requires_grad=True)
        (g2,) = torch.autograd.grad(O.conv1d(hr, W2, None, 1, 1), hr, DY)
        DH = g2 * (H > 0)
        xr = torch.zeros(d.B, d.L, d.C, requires_grad=True)
        (g1,) = torch.autograd.grad(O.conv1d(xr, W1, None, 1, d.dilation), xr, DH)
        _t(dh, (d.B, d.L, d.F)).copy_(DH)
        _t(dx, (d.B, d.L, d.C)).copy_(g1 * (X > 0) + DY)
        return 0

    # ---- VQ
    def vqb_vq_fwd_workspace_bytes(self, dref):
        return 64

    def vqb_vq_fwd(self, dref, x, E, idx, q_st, q, loss, m_batch, n_batch, ws, wsn, stream):
        d = _d(dref)
        X = _t(x, (d.N, d.D)); Ev = _t(E, (d.D, d.K))
        I = O.vq_code_indices(X, Ev)
        _t(idx, (d.N,), np.int64).copy_(I)
        Q = Ev.t()[I]
        if q is not None:
            _t(q, (d.N, d.D)).copy_(Q)
        if q_st is not None:
            _t(q_st, (d.N, d.D)).copy_(X + (Q - X))
        if loss is not None:
            _t(loss, (1,)).copy_((d.beta * ((Q - X) ** 2).mean()).reshape(1))
        if m_batch is not None:
            mb, nb = O.vq_batch_stats(X, I, d.K)
            _t(m_batch, (d.D, d.K)).copy_(mb)
            _t(n_batch, (d.K,)).copy_(nb)
        return 0

    def vqb_vq_bwd(self, dref, dq, x, q, scale, dx, stream):
        d = _d(dref)
        X = _t(x, (d.N, d.D)); Q = _t(q, (d.N, d.D))
        g = scale * 2.0 * d.beta / (d.N * d.D) * (X - Q)
        if dq is not None:
            g = g + _t(dq, (d.N, d.D))
        _t(dx, (d.N, d.D)).copy_(g)
        return 0

    def vqb_vq_ema_update(self, D, K, gamma, thr, m_batch, n_batch, rows, E, m_t, N_t, metrics, stream):
        st = O.VQState(_t(E, (D, K)), _t(m_t, (D, K)), _t(N_t, (K,)))
        new, met = O.vq_ema_update(st, _t(m_batch, (D, K)), _t(n_batch, (K,)), _t(rows, (K, D)), gamma, thr)
        st.E.copy_(new.E); st.m_t.copy_(new.m_t); st.N_t.copy_(new.N_t)
        if metrics is not None:
            _t(metrics, (3,)).copy_(torch.stack([met["batch_usage"], met["usage"], met["entropy"]]))
        return 0

    def vqb_gather_rows(self, x, N, D, ids, n_ids, n_total, off, rows, stream):
        X = _t(x, (N, D)); I = _t(ids, (n_ids,), np.int64)
        r = (I % n_total) - off
        ok = (r >= 0) & (r < N)
        out = torch.zeros(n_ids, D)
        out[ok] = X[r[ok]]
        _t(rows, (n_ids, D)).copy_(out)
        return 0

    def vqb_restart_ids(self, N, K, seed, step, ids, stream):
        s = int(_t(step, (1,), np.int64)[0]) if step is not None else 0
        Nt = N if N >= K else N * (-(-K // N))
        key = _mix(seed ^ _mix(s))
        hb = 1
        while (1 << (2 * hb)) < Nt:
            hb += 1
        mask = (1 << hb) - 1

        def perm(v):  # the kernel's 4-round Feistel permutation, cycle-walked into [0, Nt)
            while True:
                L, R = v >> hb, v & mask
                for r in range(4):
                    L, R = R, L ^ (_mix(R ^ key ^ (((r + 1) * 0x9E3779B97F4A7C15) & ((1 << 64) - 1))) & mask)
                v = (L << hb) | R
                if v < Nt:
                    return v

        _t(ids, (K,), np.int64).copy_(torch.tensor([perm(i) for i in range(K)], dtype=torch.int64))
        return 0

    def vqb_gather_codes(self, E, D, K, idx, n, out, stream):
        _t(out, (n, D)).copy_(_t(E, (D, K)).t()[_t(idx, (n,), np.int64).clamp(0, K - 1)])
        return 0

    # ---- loss / optimiser
    def vqb_reduce_workspace_bytes(self, n):
        return 64

    def vqb_mse(self, x, r, n, scale, dr_add, loss, dr, ws, wsn, stream):
        X = _t(x, (n,)); R = _t(r, (n,))
        _t(loss, (1,)).copy_(((X - R) ** 2).mean().reshape(1))
        if dr is not None:
            g = scale * 2.0 / n * (R - X)
            if dr_add is not None:
                g = g + _t(dr_add, (n,))
            _t(dr, (n,)).copy_(g)
        return 0

    def vqb_adam_step(self, p, g, m, v, n, lr, b1, b2, eps, gs, step, stream):
        t = int(_t(step, (1,), np.int64)[0]) + 1
        P, G, M, V = _t(p, (n,)), _t(g, (n,)), _t(m, (n,)), _t(v, (n,))
        O.adam_step([P], [G * gs], [M], [V], t, lr, b1, b2, eps)
        return 0

    def vqb_resstack_supports(self, d):
        return 0  # the CPU double composes residual blocks one by one

    def vqb_adam_step_dev(self, p, g, m, v, n, lr_dev, b1, b2, eps, gs, step, stream):
        return self.vqb_adam_step(p, g, m, v, n, float(_t(lr_dev, (1,))[0]), b1, b2, eps, gs, step, stream)

    def vqb_lincomb(self, n_out, term_start, term_ptr, term_coef, out, stream):
        st = C.cast(term_start, C.POINTER(C.c_int32))
        ptrs = C.cast(term_ptr, C.POINTER(C.c_void_p))
        cf = C.cast(term_coef, C.POINTER(C.c_float))
        o = _t(out, (n_out,))
        for i in range(n_out):
            s_ = torch.zeros((), dtype=torch.float32)
            for j in range(st[i], st[i + 1]):
                s_ = s_ + torch.tensor(cf[j], dtype=torch.float32) * _t(ptrs[j], (1,))[0]
            o[i] = s_
        return 0

    def vqb_increment(self, c, stream):
        _t(c, (1,), np.int64).add_(1)
        return 0
