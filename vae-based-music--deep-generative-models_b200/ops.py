"""Thin tensor-level wrappers over the C ABI (include/vqb.h).  torch is used only to own device memory and to
name the stream; every computation below is one or more `vqb_*` calls into libvqvae_b200.so."""
from __future__ import annotations

import contextlib
import ctypes as C

import torch

from . import _lib
from ._lib import ConvDesc, ResblockDesc, ResstackDesc, TailDesc, VQDesc, call, ptr

F32 = torch.float32


def empty(*shape, dtype=F32):
    return torch.empty(*shape, dtype=dtype, device=_lib.device())


def zeros(*shape, dtype=F32):
    return torch.zeros(*shape, dtype=dtype, device=_lib.device())


_deferred = None  # list of workspaces kept alive while weight-gradient reductions are queued (reduce_begin/flush)


def _ws(nbytes: int):
    t = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=_lib.device())
    if _deferred is not None:
        _deferred.append(t)
    return t


def reduce_begin():
    """Queue the fixed-order reductions of the following *_wgrad calls (vqb_reduce_begin)."""
    global _deferred
    _deferred = []
    call("vqb_reduce_begin")


def reduce_flush():
    """Launch all queued reductions as a few batched kernels (vqb_reduce_flush)."""
    global _deferred
    wg_flush()
    call("vqb_reduce_flush", _lib.stream())
    _deferred = None


def _chk(t, name):
    if t is None:
        return
    if t.dtype != F32 or not t.is_contiguous() or t.device != _lib.device():
        raise ValueError(f"{name}: expected a contiguous float32 tensor on {_lib.device()}, got {t.dtype} "
                         f"contiguous={t.is_contiguous()} on {t.device}")


def out_len(L, stride):
    return -(-L // stride)


# ----------------------------------------------------------------------------------------------- conv
def _cdesc(B, L, cin, cout, k, stride, dilation, relu_in, precision=0):
    return ConvDesc(B, L, cin, cout, k, stride, dilation, int(bool(relu_in)), precision)


def _pick(d, op, transpose=False, plain=True):
    """Keeps d.precision if libvqvae_b200 has a kernel for this (shape, op) in it and the call needs no fused extras
    the tensor-core kernels lack (`plain`), else falls back to the exact fp32 path."""
    if d.precision not in _lib.PRECISIONS.values():
        raise _lib.VQBError(f"precision code {d.precision} is not available (include/vqb.h: VQB_PREC_*)")
    if d.precision:
        fn = _lib.lib().vqb_conv1d_transpose_supports if transpose else _lib.lib().vqb_conv1d_supports
        if not plain or not fn(C.byref(d), op):
            d.precision = 0
    return d


def conv1d_fwd(x, w, b, stride=1, dilation=1, relu_in=False, residual=None, precision=0):
    _chk(x, "x"); _chk(w, "w"); _chk(b, "b"); _chk(residual, "residual")
    B, L, cin = x.shape
    k, wcin, cout = w.shape
    if wcin != cin:
        raise ValueError(f"Conv1D: input has {cin} channels, kernel expects {wcin}")
    y = empty(B, out_len(L, stride), cout)
    d = _pick(_cdesc(B, L, cin, cout, k, stride, dilation, relu_in, precision), 0, plain=residual is None)
    call("vqb_conv1d_fwd", C.byref(d), ptr(x), ptr(w), ptr(b), ptr(residual), ptr(y), _lib.stream())
    return y


def conv1d_dgrad(dy, w, x_shape, x=None, stride=1, dilation=1, relu_in=False, dx_add=None, precision=0):
    _chk(dy, "dy"); _chk(w, "w"); _chk(x, "x"); _chk(dx_add, "dx_add")
    B, L, cin = x_shape
    k, _, cout = w.shape
    dx = empty(B, L, cin)
    d = _pick(_cdesc(B, L, cin, cout, k, stride, dilation, relu_in, precision), 1, plain=dx_add is None)
    call("vqb_conv1d_dgrad", C.byref(d), ptr(dy), ptr(w), ptr(x), ptr(dx_add), ptr(dx), _lib.stream())
    return dx


def conv1d_wgrad(x, dy, dw, db, stride=1, dilation=1, relu_in=False, precision=0):
    _chk(x, "x"); _chk(dy, "dy"); _chk(dw, "dw"); _chk(db, "db")
    B, L, cin = x.shape
    k, _, cout = dw.shape
    d = _pick(_cdesc(B, L, cin, cout, k, stride, dilation, relu_in, precision), 2)  # no tensor-core kernel: exact fp32
    n = _lib.lib().vqb_conv1d_wgrad_workspace_bytes(C.byref(d))
    ws = _ws(n)
    call("vqb_conv1d_wgrad", C.byref(d), ptr(x), ptr(dy), ptr(dw), ptr(db), ptr(ws), ws.numel(), _lib.stream())


def conv1d_transpose_fwd(x, w, b, stride=2, precision=0):
    _chk(x, "x"); _chk(w, "w"); _chk(b, "b")
    B, L, cin = x.shape
    k, cout, wcin = w.shape
    if wcin != cin:
        raise ValueError(f"Conv1DTranspose: input has {cin} channels, kernel expects {wcin}")
    y = empty(B, L * stride, cout)
    d = _pick(_cdesc(B, L, cin, cout, k, stride, 1, 0, precision), 0, transpose=True)
    call("vqb_conv1d_transpose_fwd", C.byref(d), ptr(x), ptr(w), ptr(b), ptr(y), _lib.stream())
    return y


def conv1d_transpose_dgrad(dy, w, x_shape, stride=2, precision=0):
    _chk(dy, "dy"); _chk(w, "w")
    B, L, cin = x_shape
    k, cout, _ = w.shape
    dx = empty(B, L, cin)
    d = _pick(_cdesc(B, L, cin, cout, k, stride, 1, 0, precision), 1, transpose=True)
    call("vqb_conv1d_transpose_dgrad", C.byref(d), ptr(dy), ptr(w), ptr(dx), _lib.stream())
    return dx


def conv1d_transpose_wgrad(x, dy, dw, db, stride=2, precision=0):
    _chk(x, "x"); _chk(dy, "dy"); _chk(dw, "dw"); _chk(db, "db")
    B, L, cin = x.shape
    k, cout, _ = dw.shape
    d = _pick(_cdesc(B, L, cin, cout, k, stride, 1, 0, precision), 2, transpose=True)
    n = _lib.lib().vqb_conv1d_transpose_wgrad_workspace_bytes(C.byref(d))
    ws = _ws(n)
    call("vqb_conv1d_transpose_wgrad", C.byref(d), ptr(x), ptr(dy), ptr(dw), ptr(db), ptr(ws), ws.numel(),
         _lib.stream())


# --------------------------------------------------------------------------------------- decoder tail
def dec_tail_supported(cin, cmid):
    """1 if libvqvae_b200 fuses Conv1DTranspose(cmid, 4, strides=2) -> Conv1D(1, 3) for this input width."""
    return bool(_lib.lib().vqb_dec_tail_supports(C.byref(TailDesc(1, 1, int(cin), int(cmid)))))


def dec_tail_fwd(x, wt, bt, wf, bf):
    """recon [B, 2L, 1] = Conv1D(1,3)(Conv1DTranspose(k=4,s=2)(x)) as one composed operator; returns (recon, gbuf)."""
    _chk(x, "x"); _chk(wt, "wt"); _chk(bt, "bt"); _chk(wf, "wf"); _chk(bf, "bf")
    B, L, cin = x.shape
    k, cmid, wcin = wt.shape
    if k != 4 or wcin != cin or tuple(wf.shape) != (3, cmid, 1):
        raise ValueError(f"decoder tail: kernels {tuple(wt.shape)} / {tuple(wf.shape)} do not match input {tuple(x.shape)}")
    recon, gbuf = empty(B, 2 * L, 1), empty(_lib.TAIL_GBUF)
    d = TailDesc(B, L, cin, cmid)
    call("vqb_dec_tail_fwd", C.byref(d), ptr(x), ptr(wt), ptr(bt), ptr(wf), ptr(bf), ptr(gbuf), ptr(recon), _lib.stream())
    return recon, gbuf


def dec_tail_bwd(x, dr, wt, bt, wf, gbuf, dwt, dbt, dwf, dbf, want_dx=True):
    _chk(x, "x"); _chk(dr, "dr"); _chk(dwt, "dwt"); _chk(dbt, "dbt"); _chk(dwf, "dwf"); _chk(dbf, "dbf")
    B, L, cin = x.shape
    d = TailDesc(B, L, cin, wt.shape[1])
    dx = empty(B, L, cin) if want_dx else None
    ws = _ws(_lib.lib().vqb_dec_tail_bwd_workspace_bytes(C.byref(d)))
    call("vqb_dec_tail_bwd", C.byref(d), ptr(x), ptr(dr), ptr(wt), ptr(bt), ptr(wf), ptr(gbuf), ptr(dx), ptr(dwt),
         ptr(dbt), ptr(dwf), ptr(dbf), ptr(ws), ws.numel(), _lib.stream())
    return dx


# ------------------------------------------------------------------------------------------- resblock
def resblock_precision(C_, F_, dilation, precision):
    """The requested precision if libvqvae_b200 has a kernel for this block shape in it, else fp32."""
    if precision == 0:
        return 0
    d = ResblockDesc(1, 1, C_, F_, dilation, precision)
    return precision if _lib.lib().vqb_resblock_supports(C.byref(d)) else 0


def resblock_fwd(x, w1, b1, w2, b2, dilation, precision=0, want_h=True):
    """(y, h).  want_h=False (inference, tensor-core precisions only): the intermediate is not stored, h is None."""
    for t, n in ((x, "x"), (w1, "w1"), (b1, "b1"), (w2, "w2"), (b2, "b2")):
        _chk(t, n)
    B, L, Cc = x.shape
    Fc = w1.shape[2]
    if w1.shape[1] != Cc or w2.shape[1] != Fc or w2.shape[2] != Cc:
        raise ValueError(f"ResnetConv1DBlock: kernel shapes {tuple(w1.shape)}, {tuple(w2.shape)} do not fit input {tuple(x.shape)}")
    h = empty(B, L, Fc) if (want_h or precision == 0) else None
    y = empty(B, L, Cc)
    d = ResblockDesc(B, L, Cc, Fc, dilation, precision)
    call("vqb_resblock_fwd", C.byref(d), ptr(x), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(h), ptr(y), _lib.stream())
    return y, h


def resblock_fwd_masks(x, w1, b1, w2, b2, dilation, precision):
    """resblock_fwd that also returns the sign masks of x and h (int32 [B, L], bit c = channel c > 0) for
    resblock_bwd_data_masks.  Tensor-core precisions / shapes only (resblock_precision(...) != 0)."""
    for t, n in ((x, "x"), (w1, "w1"), (b1, "b1"), (w2, "w2"), (b2, "b2")):
        _chk(t, n)
    B, L, Cc = x.shape
    Fc = w1.shape[2]
    if w1.shape[1] != Cc or w2.shape[1] != Fc or w2.shape[2] != Cc:
        raise ValueError(f"ResnetConv1DBlock: kernel shapes {tuple(w1.shape)}, {tuple(w2.shape)} do not fit input {tuple(x.shape)}")
    h, y = empty(B, L, Fc), empty(B, L, Cc)
    xbits = torch.empty(B, L, dtype=torch.int32, device=_lib.device())
    hbits = torch.empty(B, L, dtype=torch.int32, device=_lib.device())
    d = ResblockDesc(B, L, Cc, Fc, dilation, precision)
    call("vqb_resblock_fwd_masks", C.byref(d), ptr(x), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(h), ptr(y), ptr(xbits),
         ptr(hbits), _lib.stream())
    return y, h, xbits, hbits


def resblock_bwd_data_masks(xbits, hbits, dy, w1, w2, dilation, precision):
    _chk(dy, "dy")
    B, L, Cc = dy.shape
    Fc = w1.shape[2]
    dh, dx = empty(B, L, Fc), empty(B, L, Cc)
    d = ResblockDesc(B, L, Cc, Fc, dilation, precision)
    call("vqb_resblock_bwd_data_masks", C.byref(d), ptr(xbits), ptr(hbits), ptr(dy), ptr(w1), ptr(w2), ptr(dh), ptr(dx),
         _lib.stream())
    return dx, dh


def resblock_bwd_data(x, h, dy, w1, w2, dilation, precision=0):
    _chk(dy, "dy")
    B, L, Cc = x.shape
    Fc = h.shape[2]
    dh, dx = empty(B, L, Fc), empty(B, L, Cc)
    d = ResblockDesc(B, L, Cc, Fc, dilation, precision)
    call("vqb_resblock_bwd_data", C.byref(d), ptr(x), ptr(h), ptr(dy), ptr(w1), ptr(w2), ptr(dh), ptr(dx),
         _lib.stream())
    return dx, dh


# ------------------------------------------------------------------------------------------- resstack
RESSTACK_MAX = 4  # VQB_RESSTACK_MAX_BLOCKS


def _sdesc(B, L, Cc, dilations, precision):
    d = ResstackDesc()
    d.B, d.L, d.C, d.n_blocks, d.precision = B, L, Cc, len(dilations), precision
    for i, v in enumerate(dilations[:RESSTACK_MAX]):
        d.dilations[i] = int(v)
    return d


def resstack_supported(Cc, dilations, precision):
    """True if libvqvae_b200 runs these residual blocks (first-convolution dilations, in execution order) as ONE fused launch."""
    if precision == 0 or not 1 <= len(dilations) <= RESSTACK_MAX:
        return False
    return bool(_lib.lib().vqb_resstack_supports(C.byref(_sdesc(1, 1, Cc, list(dilations), precision))))


def _parr(ts):
    return C.cast((C.c_void_p * len(ts))(*[ptr(t) for t in ts]), C.c_void_p)


def resstack_workspace(Cc, dilations, precision):
    """A workspace for resstack_fwd(..., ws=...) that the caller keeps for itself (one per DilatedResnet1D)."""
    d = _sdesc(1, 1, Cc, list(dilations), precision)
    return torch.empty(max(int(_lib.lib().vqb_resstack_workspace_bytes(C.byref(d))), 16), dtype=torch.uint8, device=_lib.device())


def resstack_fwd(x, w1s, b1s, w2s, b2s, dilations, precision, train, ws=None):
    """A chain of len(dilations) residual blocks in one launch (vqb_resstack_fwd).  Returns (ys, hs, xbits, hbits, ws): under a
    tape (`train`) every block's h, output and sign masks (lists of n tensors) and the workspace with the packed operand
    images of both directions (hand it to resstack_bwd_data), else ys = [None, ..., y_last] and hs, xbits, hbits, ws = None
    (nothing but the stack's output is written)."""
    _chk(x, "x")
    for t in (*w1s, *b1s, *w2s, *b2s):
        _chk(t, "weight")
    B, L, Cc = x.shape
    n = len(dilations)
    d = _sdesc(B, L, Cc, list(dilations), precision)
    # ws given: the caller's own buffer (nothing else ever uses it), which lets the packing launch overlap the previous kernel
    entry = "vqb_resstack_fwd_private_ws" if ws is not None else "vqb_resstack_fwd"
    if ws is None:
        ws = _ws(_lib.lib().vqb_resstack_workspace_bytes(C.byref(d)))
    if train:
        ys = [empty(B, L, Cc) for _ in range(n)]
        hs = [empty(B, L, Cc) for _ in range(n)]
        xb = [torch.empty(B, L, dtype=torch.int32, device=_lib.device()) for _ in range(n)]
        hb = [torch.empty(B, L, dtype=torch.int32, device=_lib.device()) for _ in range(n)]
        call(entry, C.byref(d), ptr(x), _parr(w1s), _parr(b1s), _parr(w2s), _parr(b2s), _parr(hs), _parr(ys),
             _parr(xb), _parr(hb), ptr(ws), ws.numel(), _lib.stream())
        return ys, hs, xb, hb, ws
    ys = [None] * (n - 1) + [empty(B, L, Cc)]
    call(entry, C.byref(d), ptr(x), _parr(w1s), _parr(b1s), _parr(w2s), _parr(b2s), None, _parr(ys), None, None,
         ptr(ws), ws.numel(), _lib.stream())
    return ys, None, None, None, None


def resstack_bwd_data(dy, w1s, w2s, xbits, hbits, dilations, precision, fwd_ws=None):
    """Data gradients of the chain (vqb_resstack_bwd_data): returns (dxs, dhs), dxs[i] = gradient at the input of block i,
    dhs[i] = gradient at the output of its first convolution.  fwd_ws: the workspace returned by the resstack_fwd(train=True)
    call of the same blocks and (unchanged) weights — the operand images packed there are reused (no packing launch)."""
    _chk(dy, "dy")
    B, L, Cc = dy.shape
    n = len(dilations)
    d = _sdesc(B, L, Cc, list(dilations), precision)
    dhs = [empty(B, L, Cc) for _ in range(n)]
    dxs = [empty(B, L, Cc) for _ in range(n)]
    if fwd_ws is not None:
        call("vqb_resstack_bwd_data_packed", C.byref(d), ptr(dy), _parr(xbits), _parr(hbits), _parr(dhs), _parr(dxs),
             ptr(fwd_ws), fwd_ws.numel(), _lib.stream())
        return dxs, dhs
    ws = _ws(_lib.lib().vqb_resstack_workspace_bytes(C.byref(d)))
    call("vqb_resstack_bwd_data", C.byref(d), ptr(dy), _parr(w1s), _parr(w2s), _parr(xbits), _parr(hbits), _parr(dhs),
         _parr(dxs), ptr(ws), ws.numel(), _lib.stream())
    return dxs, dhs


# ------------------------------------------------------------------------------------------------- VQ
_wg_queue = []      # residual-block weight-gradient problems waiting for their stack mates (same shape, same stream)
WG_BATCH = 4        # blocks per launch: a DilatedResnet1D of the reference models (vqb_resblock_wgrad_batch takes <= 4)


def _wg_key(x, h, precision):
    dev = _lib.device()
    st = torch.cuda.current_stream(dev) if dev.type == "cuda" else None
    return (tuple(x.shape), h.shape[2], precision, st.cuda_stream if st is not None else 0), st


def wg_flush():
    """Launch the queued residual-block weight gradients: one vqb_resblock_wgrad_batch call for up to WG_BATCH blocks."""
    global _wg_queue
    q, _wg_queue = _wg_queue, []
    if not q:
        return
    key, stream, items = q[0][0], q[0][1], [it[2] for it in q]
    (B, L, Cc), Fc, precision, _ = key
    n = len(items)
    # on the stream the problems were queued on (their operands were produced there), whatever the current one is
    with (torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()):
        if n == 1:
            x, h, dy, dh, dw1, db1, dw2, db2, dil = items[0]
            _resblock_wgrad_now(x, h, dy, dh, dw1, db1, dw2, db2, dil, precision)
            return
        d = ResblockDesc(B, L, Cc, Fc, 1, precision)
        arr = lambda k: (C.c_void_p * n)(*[ptr(it[k]) for it in items])
        dils = (C.c_int32 * n)(*[it[8] for it in items])
        ws = _ws(_lib.lib().vqb_resblock_wgrad_batch_workspace_bytes(C.byref(d), n))
        call("vqb_resblock_wgrad_batch", C.byref(d), n, C.cast(dils, C.c_void_p),
             *[C.cast(arr(k), C.c_void_p) for k in range(8)], ptr(ws), ws.numel(), _lib.stream())


def resblock_wgrad(x, h, dy, dh, dw1, db1, dw2, db2, dilation, precision=0):
    """Both weight (+ bias) gradients of a residual block (vqb_resblock_wgrad): conv1 from (ReLU(x), dh, dilation), conv2 from
    (ReLU(h), dy, 1).  One kernel launch in the tensor-core precisions; inside a reduce_begin() / reduce_flush() window (a
    backward pass) consecutive blocks of one shape on one stream are collected and launched together, WG_BATCH at a time
    (vqb_resblock_wgrad_batch: the per-launch fixed cost of the persistent kernel is paid once per DilatedResnet1D)."""
    for t, n in ((x, "x"), (h, "h"), (dy, "dy"), (dh, "dh"), (dw1, "dw1"), (db1, "db1"), (dw2, "dw2"), (db2, "db2")):
        _chk(t, n)
    if _deferred is None or precision == 0 or WG_BATCH <= 1:
        return _resblock_wgrad_now(x, h, dy, dh, dw1, db1, dw2, db2, dilation, precision)
    key, stream = _wg_key(x, h, precision)
    if _wg_queue and _wg_queue[0][0] != key:
        wg_flush()
    _wg_queue.append((key, stream, (x, h, dy, dh, dw1, db1, dw2, db2, dilation)))
    if len(_wg_queue) >= WG_BATCH:
        wg_flush()


def _resblock_wgrad_now(x, h, dy, dh, dw1, db1, dw2, db2, dilation, precision=0):
    B, L, Cc = x.shape
    d = ResblockDesc(B, L, Cc, h.shape[2], dilation, precision)
    ws = _ws(_lib.lib().vqb_resblock_wgrad_workspace_bytes(C.byref(d)))
    call("vqb_resblock_wgrad", C.byref(d), ptr(x), ptr(h), ptr(dy), ptr(dh), ptr(dw1), ptr(db1), ptr(dw2), ptr(db2), ptr(ws),
         ws.numel(), _lib.stream())


VQ_BF16_IO = True  # vq_fwd takes bfloat16 latents (vqb_vq_fwd_bf16)


def vq_fwd(flat, E, beta, want_q_st=True, want_q=True, m_batch=None, n_batch=None, precision=0):
    """flat [N,D], E [D,K] -> idx int64 [N], q_st, q, loss[1].  A bfloat16 `flat` selects vqb_vq_fwd_bf16: q_st / q come back
    in bfloat16, codebook, loss and statistics stay fp32 (include/vqb.h)."""
    bf = flat.dtype == torch.bfloat16
    if bf:
        if not flat.is_contiguous() or flat.device != _lib.device():
            raise ValueError(f"x: expected a contiguous tensor on {_lib.device()}")
    else:
        _chk(flat, "x")
    _chk(E, "embeddings")
    N, D = flat.shape
    if E.shape[0] != D:
        raise ValueError(f"VectorQuantizer: input depth {D} != embedding_dim {E.shape[0]}")
    K = E.shape[1]
    idx = empty(N, dtype=torch.int64)
    q_st = empty(N, D, dtype=flat.dtype) if want_q_st else None
    q = empty(N, D, dtype=flat.dtype) if want_q else None
    loss = empty(1)
    d = VQDesc(N, D, K, beta, precision)
    ws = _ws(_lib.lib().vqb_vq_fwd_workspace_bytes(C.byref(d)))
    call("vqb_vq_fwd_bf16" if bf else "vqb_vq_fwd", C.byref(d), ptr(flat), ptr(E), ptr(idx), ptr(q_st), ptr(q), ptr(loss),
         ptr(m_batch), ptr(n_batch), ptr(ws), ws.numel(), _lib.stream())
    return idx, q_st, q, loss


def vq_bwd(dq, flat, q, beta, loss_scale):
    N, D = flat.shape
    dx = empty(N, D)
    d = VQDesc(N, D, 0, beta, 0)
    call("vqb_vq_bwd", C.byref(d), ptr(dq), ptr(flat), ptr(q), float(loss_scale), ptr(dx), _lib.stream())
    return dx


def vq_ema_update(E, m_t, N_t, m_batch, n_batch, restart_rows, gamma, threshold, metrics=None):
    D, K = E.shape
    call("vqb_vq_ema_update", D, K, float(gamma), float(threshold), ptr(m_batch), ptr(n_batch), ptr(restart_rows),
         ptr(E), ptr(m_t), ptr(N_t), ptr(metrics), _lib.stream())


def gather_rows(flat, ids, n_total=None, row_offset=0, out=None):
    N, D = flat.shape
    rows = out if out is not None else empty(ids.numel(), D)
    call("vqb_gather_rows", ptr(flat), N, D, ptr(ids), ids.numel(), int(n_total if n_total is not None else N),
         int(row_offset), ptr(rows), _lib.stream())
    return rows


def restart_ids(N, K, seed, step_counter):
    ids = empty(K, dtype=torch.int64)
    call("vqb_restart_ids", int(N), int(K), int(seed), ptr(step_counter), ptr(ids), _lib.stream())
    return ids


def gather_codes(E, idx):
    D, K = E.shape
    idx = idx.contiguous()
    out = empty(*idx.shape, D)
    call("vqb_gather_codes", ptr(E), D, K, ptr(idx), idx.numel(), ptr(out), _lib.stream())
    return out


# ---------------------------------------------------------------------------------------- loss / optimiser
def mse(x, r, loss_scale=None, dr_add=None):
    """returns (loss[1], dr or None).  dr = loss_scale * 2 (r - x)/n (+ dr_add) when loss_scale is given."""
    _chk(x, "x"); _chk(r, "r")
    if x.shape != r.shape:
        raise ValueError(f"mse: shapes differ {tuple(x.shape)} vs {tuple(r.shape)}")
    n = x.numel()
    loss = empty(1)
    dr = empty(r.shape) if loss_scale is not None else None
    ws = _ws(_lib.lib().vqb_reduce_workspace_bytes(n))
    call("vqb_mse", ptr(x), ptr(r), n, float(loss_scale or 0.0), ptr(dr_add), ptr(loss), ptr(dr), ptr(ws),
         ws.numel(), _lib.stream())
    return loss, dr


def adam_step(p, g, m, v, lr, b1, b2, eps, grad_scale, step_counter):
    call("vqb_adam_step", ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), float(lr), float(b1), float(b2), float(eps),
         float(grad_scale), ptr(step_counter), _lib.stream())


def adam_step_dev(p, g, m, v, lr_dev, b1, b2, eps, grad_scale, step_counter):
    """adam_step with the learning rate in a one-element device tensor (read when the kernel runs)."""
    call("vqb_adam_step_dev", ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), ptr(lr_dev), float(b1), float(b2), float(eps),
         float(grad_scale), ptr(step_counter), _lib.stream())


def lincomb(outputs):
    """outputs: list of term lists [(tensor with >= 1 element, coefficient), ...] -> float32 vector [len(outputs)] of the
    weighted sums of the tensors' first elements, ONE launch (vqb_lincomb)."""
    starts, ptrs, coefs, keep = [0], [], [], []
    for terms in outputs:
        for t, c in terms:
            _chk(t, "term")
            ptrs.append(ptr(t)); coefs.append(float(c)); keep.append(t)
        starts.append(len(ptrs))
    out = empty(len(outputs))
    n = len(ptrs)
    call("vqb_lincomb", len(outputs), C.cast((C.c_int32 * len(starts))(*starts), C.c_void_p),
         C.cast((C.c_void_p * max(n, 1))(*ptrs), C.c_void_p), C.cast((C.c_float * max(n, 1))(*coefs), C.c_void_p), ptr(out),
         _lib.stream())
    return out


def increment(counter):
    call("vqb_increment", ptr(counter), _lib.stream())


# ------------------------------------------------------------------------------------- spectral loss pieces
def stft_frames(x2d, n_fft, hop, win):
    """x [B, T] -> windowed, zero-padded frames [B, F, n_fft] (tf.signal.stft framing, data_utils.py:25-30)."""
    _chk(x2d, "x")
    B, T = x2d.shape
    F = 1 + (T - win) // hop
    fr = empty(B, F, n_fft)
    call("vqb_stft_frames", ptr(x2d), B, T, n_fft, hop, win, ptr(fr), _lib.stream())
    return fr


def _spec_ws(B, per):
    return _ws(_lib.lib().vqb_spec_workspace_bytes(B, per))


def spec_mag(S):
    """S complex64 [B, F, bins] -> (|S| [B, F, bins], sum |S|^2 per example [B])."""
    Sr = torch.view_as_real(S)
    B, per = S.shape[0], S.shape[1] * S.shape[2]
    mag, sums, ws = empty(*S.shape), empty(B), _spec_ws(B, per)
    call("vqb_spec_mag", ptr(Sr), B, per, ptr(mag), ptr(sums), ptr(ws), ws.numel(), _lib.stream())
    return mag, sums


def spec_diff(S, mag_t, sums):
    """sums[b] = sum (mag_t - |S|)^2 over example b (the squared norm of data_utils.norm(S_t - S_r))."""
    Sr = torch.view_as_real(S)
    B, per = S.shape[0], S.shape[1] * S.shape[2]
    ws = _spec_ws(B, per)
    call("vqb_spec_diff", ptr(Sr), ptr(mag_t), B, per, ptr(sums), ptr(ws), ws.numel(), _lib.stream())


def spec_loss(dsum, tsum, want_coef):
    nscales, B = dsum.shape
    loss, coef = empty(1), (empty(nscales, B) if want_coef else None)
    call("vqb_spec_loss", ptr(dsum), ptr(tsum), nscales, B, ptr(loss), ptr(coef), _lib.stream())
    return loss, coef


def spec_grad(S, mag_t, coef_row, upstream, n_fft):
    Sr = torch.view_as_real(S)
    B, F, bins = S.shape
    G = torch.empty_like(S)
    call("vqb_spec_grad", ptr(Sr), ptr(mag_t), ptr(coef_row), ptr(upstream), B, F * bins, bins, n_fft,
         ptr(torch.view_as_real(G)), _lib.stream())
    return G


def stft_frames_bwd(dframes, T, hop, win, dx, accumulate):
    B, F, n_fft = dframes.shape
    call("vqb_stft_frames_bwd", ptr(dframes), B, T, n_fft, hop, win, int(bool(accumulate)), ptr(dx), _lib.stream())
