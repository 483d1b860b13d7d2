"""tcgen05 (tensor-core) kernels against the oracle.  Operands are rounded to bf16 / tf32 before the MMA (fp32
accumulate), so the tolerance is that of the operand format, stated per test: bf16 2^-8 ~ 4e-3 per operand,
tf32 2^-11 ~ 5e-4 per operand; errors are measured relative to the output's max magnitude."""
import numpy as np
import pytest
import torch

from oracle import vqvae_oracle as O

pytestmark = pytest.mark.gpu
TOL = {"bf16": 1.5e-2, "tf32": 2e-3, "bf16x2": 1e-4, "bf16x3": 2e-5, "fp16x2": 2e-5}  # bf16x3 (24 mantissa bits) and fp16x2 (22) are fp32-grade


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).float().cuda().contiguous()


def rel(got, want):
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    return float((got - want).abs().max() / want.abs().max().clamp_min(1e-12))


@pytest.mark.parametrize("prec", ["tf32", "bf16", "bf16x2", "bf16x3", "fp16x2"])
@pytest.mark.parametrize("B,L,d", [(2, 1000, 1), (3, 881, 27), (1, 254, 3), (2, 20, 27), (2, 255, 9), (1, 1, 1), (4, 3520, 9)])
def test_resblock_tc_fwd_bwd(gpu, prec, B, L, d):
    ops, P = gpu.ops, gpu._lib.PRECISIONS[prec]
    rng = np.random.default_rng(L + d)
    C = 32
    x = rng.normal(size=(B, L, C)).astype(np.float32)
    w1 = (rng.normal(size=(3, C, C)) / np.sqrt(3 * C)).astype(np.float32); b1 = (rng.normal(size=C) * 0.1).astype(np.float32)
    w2 = (rng.normal(size=(3, C, C)) / np.sqrt(3 * C)).astype(np.float32); b2 = (rng.normal(size=C) * 0.1).astype(np.float32)
    dy = rng.normal(size=(B, L, C)).astype(np.float32)
    assert ops.resblock_precision(C, C, d, P) == P
    ts = [torch.tensor(a, requires_grad=True) for a in (x, w1, b1, w2, b2)]
    h_ref = O.conv1d(torch.relu(ts[0]), ts[1], ts[2], 1, d)
    y_ref = ts[0] + O.conv1d(torch.relu(h_ref), ts[3], ts[4], 1, 1)
    gx, gh = torch.autograd.grad(y_ref, (ts[0], h_ref), torch.tensor(dy))
    y, h = ops.resblock_fwd(dev(x), dev(w1), dev(b1), dev(w2), dev(b2), d, P)
    torch.cuda.synchronize()
    assert rel(h, h_ref) < TOL[prec], ("h", rel(h, h_ref))
    assert rel(y, y_ref) < TOL[prec], ("y", rel(y, y_ref))
    # the oracle's h keeps the ReLU masks identical, so the comparison isolates the arithmetic of the backward kernel
    dx, dh = ops.resblock_bwd_data(dev(x), dev(h_ref.detach().numpy()), dev(dy), dev(w1), dev(w2), d, P)
    torch.cuda.synchronize()
    assert rel(dh, gh) < TOL[prec], ("dh", rel(dh, gh))
    assert rel(dx, gx) < TOL[prec], ("dx", rel(dx, gx))


@pytest.mark.parametrize("env", [{"VQB_RB_MB": "1"}, {"VQB_RB_MB": "2"}, {"VQB_RB_TMA": "1"}])
@pytest.mark.parametrize("B,L,d", [(2, 1000, 1), (3, 881, 27), (1, 254, 3), (2, 20, 27), (1, 1, 1), (4, 3520, 9)])
def test_resblock_fp16x2_kernel_variants(gpu, monkeypatch, env, B, L, d):
    """The fp16x2 residual block has three kernel variants picked per call (resblock_tc.cu: dispatch_rb): 128-row tiles with
    two CTAs per SM, 256-row tiles, and 256-row tiles whose outputs leave through TMA bulk-tensor stores (tma.cuh: swizzled
    shared-memory images, rows beyond L clipped by the hardware).  Each must meet the same fp32-grade tolerance."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    ops, P = gpu.ops, gpu._lib.PRECISIONS["fp16x2"]
    rng = np.random.default_rng(L + d)
    C = 32
    x = rng.normal(size=(B, L, C)).astype(np.float32)
    w1 = (rng.normal(size=(3, C, C)) / np.sqrt(3 * C)).astype(np.float32); b1 = (rng.normal(size=C) * 0.1).astype(np.float32)
    w2 = (rng.normal(size=(3, C, C)) / np.sqrt(3 * C)).astype(np.float32); b2 = (rng.normal(size=C) * 0.1).astype(np.float32)
    dy = rng.normal(size=(B, L, C)).astype(np.float32)
    ts = [torch.tensor(a, requires_grad=True) for a in (x, w1, b1, w2, b2)]
    h_ref = O.conv1d(torch.relu(ts[0]), ts[1], ts[2], 1, d)
    y_ref = ts[0] + O.conv1d(torch.relu(h_ref), ts[3], ts[4], 1, 1)
    gx, gh = torch.autograd.grad(y_ref, (ts[0], h_ref), torch.tensor(dy))
    y, h = ops.resblock_fwd(dev(x), dev(w1), dev(b1), dev(w2), dev(b2), d, P)
    dx, dh = ops.resblock_bwd_data(dev(x), dev(h_ref.detach().numpy()), dev(dy), dev(w1), dev(w2), d, P)
    torch.cuda.synchronize()
    for name, got, want in (("h", h, h_ref), ("y", y, y_ref), ("dh", dh, gh), ("dx", dx, gx)):
        assert rel(got, want) < TOL["fp16x2"], (env, name, rel(got, want))


@pytest.mark.parametrize("prec", ["tf32", "bf16", "bf16x2", "bf16x3", "fp16x2"])
def test_resblock_tc_full_size_matches_fp32_kernel(gpu, prec):
    """[32, 14080, 32] (the largest stage of SMALL_VQ_VAE at batch 32): tensor-core path vs the exact-fp32 CUDA path."""
    ops, P = gpu.ops, gpu._lib.PRECISIONS[prec]
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(32, 14080, 32, device="cuda", generator=g)
    w1 = torch.randn(3, 32, 32, device="cuda", generator=g) / 96 ** 0.5
    w2 = torch.randn(3, 32, 32, device="cuda", generator=g) / 96 ** 0.5
    b1 = torch.randn(32, device="cuda", generator=g) * 0.1
    b2 = torch.randn(32, device="cuda", generator=g) * 0.1
    dy = torch.randn(32, 14080, 32, device="cuda", generator=g)
    for d in (1, 27):
        y0, h0 = ops.resblock_fwd(x, w1, b1, w2, b2, d, 0)
        y1, h1 = ops.resblock_fwd(x, w1, b1, w2, b2, d, P)
        assert rel(h1, h0) < TOL[prec] and rel(y1, y0) < TOL[prec]
        dx0, dh0 = ops.resblock_bwd_data(x, h0, dy, w1, w2, d, 0)
        dx1, dh1 = ops.resblock_bwd_data(x, h0, dy, w1, w2, d, P)
        assert rel(dh1, dh0) < TOL[prec] and rel(dx1, dx0) < TOL[prec]


@pytest.mark.parametrize("xs,ws,bs", [(1e-6, 1.0, 0.0), (3e4, 1.0, 0.1), (1.0, 1e-5, 1e-7), (1e-3, 50.0, 10.0), (1e-12, 1e-12, 0.0),
                                      (1e12, 1e3, 1.0), (0.0, 1.0, 0.3)])
def test_resblock_fp16x2_is_scale_free(gpu, xs, ws, bs):
    """fp16x2 multiplies every operand tile / weight tensor by a power of two that fits it into fp16's range before the
    split (tc.cuh) and undoes it in the epilogue: the result must stay fp32-grade for magnitudes far outside fp16's range,
    for tiles of very different magnitude inside one launch, and for an all-zero input (scale 1)."""
    ops, P = gpu.ops, gpu._lib.PRECISIONS["fp16x2"]
    rng = np.random.default_rng(5)
    B, L, C, d = 3, 2100, 32, 3
    x = (rng.normal(size=(B, L, C)) * xs).astype(np.float32)
    x[1, 700:1400] *= 1e-4   # whole tiles far below their neighbours
    x[2, 10] *= 300.0        # one dominant row inside a tile
    w1 = (rng.normal(size=(3, C, C)) / np.sqrt(3 * C) * ws).astype(np.float32); b1 = (rng.normal(size=C) * bs).astype(np.float32)
    w2 = (rng.normal(size=(3, C, C)) / np.sqrt(3 * C) * ws).astype(np.float32); b2 = (rng.normal(size=C) * bs).astype(np.float32)
    dy = (rng.normal(size=(B, L, C)) * xs).astype(np.float32)
    ts = [torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (x, w1, b1, w2, b2)]
    h_ref = O.conv1d(torch.relu(ts[0]), ts[1], ts[2], 1, d)
    y_ref = ts[0] + O.conv1d(torch.relu(h_ref), ts[3], ts[4], 1, 1)
    gx, gh = torch.autograd.grad(y_ref, (ts[0], h_ref), torch.tensor(dy, dtype=torch.float64))
    y, h = ops.resblock_fwd(dev(x), dev(w1), dev(b1), dev(w2), dev(b2), d, P)
    dx, dh = ops.resblock_bwd_data(dev(x), dev(h_ref.detach().float().numpy()), dev(dy), dev(w1), dev(w2), d, P)
    torch.cuda.synchronize()
    for name, got, want in (("h", h, h_ref), ("y", y, y_ref), ("dh", dh, gh), ("dx", dx, gx)):
        assert bool(torch.isfinite(got).all()), name
        if float(want.abs().max()) > 1e-30:  # below that the fp32 result itself underflows
            assert rel(got, want) < TOL["fp16x2"], (name, rel(got, want))
    # a quiet stretch next to a loud one keeps its own relative accuracy (per-tile scales)
    quiet = slice(800, 1300)
    if xs > 0:
        assert rel(h[1, quiet], h_ref[1, quiet]) < 1e-4


@pytest.mark.parametrize("prec", ["bf16", "bf16x2", "bf16x3"])
@pytest.mark.parametrize("B,L,d,relu", [(2, 1000, 1, 1), (3, 881, 27, 1), (1, 256, 3, 0), (2, 20, 27, 1), (5, 3520, 9, 1), (1, 1, 1, 1),
                                        (32, 14080, 3, 1)])
def test_conv_wgrad_tc(gpu, prec, B, L, d, relu):
    """tcgen05 weight gradient (time = MMA K dimension, persistent CTAs) vs autograd of the oracle conv."""
    ops, P = gpu.ops, gpu._lib.PRECISIONS[prec]
    g = torch.Generator().manual_seed(L + d)
    x = torch.randn(B, L, 32, generator=g)
    dy = torch.randn(B, L, 32, generator=g)
    w = torch.zeros(3, 32, 32, requires_grad=True)
    bz = torch.zeros(32, requires_grad=True)
    y = O.conv1d(torch.relu(x) if relu else x, w, bz, 1, d)
    gw, gb = torch.autograd.grad(y, (w, bz), dy)
    dw, db = ops.empty(3, 32, 32), ops.empty(32)
    ops.conv1d_wgrad(x.cuda(), dy.cuda(), dw, db, 1, d, bool(relu), P)
    torch.cuda.synchronize()
    # operands are rounded to bf16/tf32 but the sum runs over B*L terms: the error averages down, the tolerance holds
    assert rel(dw, gw) < TOL[prec], rel(dw, gw)
    assert rel(db, gb) < TOL[prec], rel(db, gb)
    # determinism: two runs are bit-identical (fixed-order reduction of per-CTA partials, no atomics)
    dw2, db2 = ops.empty(3, 32, 32), ops.empty(32)
    ops.conv1d_wgrad(x.cuda(), dy.cuda(), dw2, db2, 1, d, bool(relu), P)
    assert torch.equal(dw, dw2) and torch.equal(db, db2)


@pytest.mark.parametrize("prec", ["bf16", "bf16x3", "fp16x2"])
@pytest.mark.parametrize("B,L,d", [(2, 1000, 1), (3, 881, 27), (1, 254, 3), (2, 20, 27), (2, 255, 9), (1, 1, 1), (4, 3520, 9), (5, 126, 1),
                                   (3, 127, 3)])
def test_resblock_sign_mask_variants(gpu, prec, B, L, d):
    """vqb_resblock_fwd_masks writes, next to h and y, one uint32 per position for x and for h (bit c = channel c > 0: one half-word per row and 16-channel half from the two
    epilogues, every in-range row exactly once across overlapping tiles);
    vqb_resblock_bwd_data_masks must give bit for bit what vqb_resblock_bwd_data gives from the fp32 tensors."""
    ops, P = gpu.ops, gpu._lib.PRECISIONS[prec]
    rng = np.random.default_rng(3 * L + d)
    C = 32
    x = dev(rng.normal(size=(B, L, C)).astype(np.float32))
    w1 = dev((rng.normal(size=(3, C, C)) / np.sqrt(3 * C)).astype(np.float32)); b1 = dev((rng.normal(size=C) * 0.1).astype(np.float32))
    w2 = dev((rng.normal(size=(3, C, C)) / np.sqrt(3 * C)).astype(np.float32)); b2 = dev((rng.normal(size=C) * 0.1).astype(np.float32))
    dy = dev(rng.normal(size=(B, L, C)).astype(np.float32))
    y0, h0 = ops.resblock_fwd(x, w1, b1, w2, b2, d, P)
    xb = hb = None
    y1, h1, xb, hb = ops.resblock_fwd_masks(x, w1, b1, w2, b2, d, P)
    y2, h2 = ops.resblock_fwd(x, w1, b1, w2, b2, d, P, want_h=False)  # inference form: h is not stored
    torch.cuda.synchronize()
    assert torch.equal(y0, y1) and torch.equal(h0, h1) and h2 is None and torch.equal(y0, y2)
    sh = torch.arange(32, device="cuda")
    assert torch.equal(((xb.long().unsqueeze(-1) >> sh) & 1).bool(), x > 0)
    assert torch.equal(((hb.long().unsqueeze(-1) >> sh) & 1).bool(), h1 > 0)
    dx0, dh0 = ops.resblock_bwd_data(x, h1, dy, w1, w2, d, P)
    dx1, dh1 = ops.resblock_bwd_data_masks(xb, hb, dy, w1, w2, d, P)
    torch.cuda.synchronize()
    assert torch.equal(dx0, dx1) and torch.equal(dh0, dh1)


@pytest.mark.parametrize("prec", ["fp32", "bf16", "bf16x3", "fp16x2"])
@pytest.mark.parametrize("B,L,d", [(2, 1000, 1), (3, 881, 27), (1, 128, 3), (2, 20, 9), (1, 1, 1), (32, 440, 9), (4, 3520, 3)])
def test_resblock_wgrad_matches_two_conv_wgrads(gpu, prec, B, L, d):
    """vqb_resblock_wgrad (one launch for both convolutions of a block on the tensor-core paths: the two problems share the
    grid) gives exactly what two vqb_conv1d_wgrad calls give — same kernel, same tile order per problem only when the CTA
    split is the same, so the comparison is to the oracle's autograd within the mode's tolerance and to a second run bit for
    bit (deterministic fixed-order reduction)."""
    ops, P = gpu.ops, gpu._lib.PRECISIONS[prec]
    g = torch.Generator().manual_seed(7 * L + d)
    x, h = torch.randn(B, L, 32, generator=g), torch.randn(B, L, 32, generator=g)
    dy, dh = torch.randn(B, L, 32, generator=g), torch.randn(B, L, 32, generator=g)
    w1 = torch.zeros(3, 32, 32, requires_grad=True); b1 = torch.zeros(32, requires_grad=True)
    w2 = torch.zeros(3, 32, 32, requires_grad=True); b2 = torch.zeros(32, requires_grad=True)
    g1 = torch.autograd.grad(O.conv1d(torch.relu(x), w1, b1, 1, d), (w1, b1), dh)
    g2 = torch.autograd.grad(O.conv1d(torch.relu(h), w2, b2, 1, 1), (w2, b2), dy)
    outs = [[ops.empty(3, 32, 32), ops.empty(32), ops.empty(3, 32, 32), ops.empty(32)] for _ in range(2)]
    for o in outs:
        ops.resblock_wgrad(x.cuda(), h.cuda(), dy.cuda(), dh.cuda(), o[0], o[1], o[2], o[3], d, P)
    torch.cuda.synchronize()
    tol = 2e-5 if prec == "fp32" else TOL["bf16x3" if prec == "fp16x2" else prec]
    for got, want in zip(outs[0], (*g1, *g2)):
        assert rel(got, want) < tol, (prec, rel(got, want))
    assert all(torch.equal(a, b) for a, b in zip(outs[0], outs[1]))


@pytest.mark.parametrize("prec", ["fp32", "tf32", "bf16x3", "fp16x2"])  # tf32 has no tensor-core wgrad: exact fp32 inside
@pytest.mark.parametrize("B,L,nblk", [(2, 1000, 4), (3, 881, 3), (1, 1, 2), (32, 110, 4), (4, 3520, 4), (2, 300, 5)])
def test_resblock_wgrad_batching(gpu, prec, B, L, nblk):
    """Inside a reduce_begin() / reduce_flush() window ops.resblock_wgrad collects the blocks of a stack and launches them
    together (vqb_resblock_wgrad_batch: up to 4 blocks = 8 problems share the grid of one tcgen05 wgrad launch; a 5th block
    starts the next batch).  Every block's four gradients must match the oracle's autograd, and a second pass must be bit
    identical (fixed CTA split, fixed-order reductions)."""
    ops, P = gpu.ops, gpu._lib.PRECISIONS[prec]
    g = torch.Generator().manual_seed(11 * L + nblk)
    dil = [27, 9, 3, 1, 27][:nblk]
    blocks = []
    for i in range(nblk):
        x, h = torch.randn(B, L, 32, generator=g), torch.randn(B, L, 32, generator=g)
        dy, dh = torch.randn(B, L, 32, generator=g), torch.randn(B, L, 32, generator=g)
        w1 = torch.zeros(3, 32, 32, requires_grad=True); b1 = torch.zeros(32, requires_grad=True)
        w2 = torch.zeros(3, 32, 32, requires_grad=True); b2 = torch.zeros(32, requires_grad=True)
        g1 = torch.autograd.grad(O.conv1d(torch.relu(x), w1, b1, 1, dil[i]), (w1, b1), dh)
        g2 = torch.autograd.grad(O.conv1d(torch.relu(h), w2, b2, 1, 1), (w2, b2), dy)
        blocks.append(((x.cuda(), h.cuda(), dy.cuda(), dh.cuda()), (*g1, *g2)))
    runs = []
    for _ in range(2):
        outs = [[ops.empty(3, 32, 32), ops.empty(32), ops.empty(3, 32, 32), ops.empty(32)] for _ in range(nblk)]
        ops.reduce_begin()
        for i, (t, _) in enumerate(blocks):
            ops.resblock_wgrad(*t, *outs[i], dil[i], P)
        ops.reduce_flush()
        torch.cuda.synchronize()
        runs.append(outs)
    tol = 2e-5 if prec in ("fp32", "tf32") else TOL["bf16x3"]
    for i, (_, want) in enumerate(blocks):
        for got, w in zip(runs[0][i], want):
            assert rel(got, w) < tol, (prec, i, rel(got, w))
        assert all(torch.equal(a, b) for a, b in zip(runs[0][i], runs[1][i]))


@pytest.mark.parametrize("N,K,kind", [(28160, 512, "normal"), (3520, 512, "init"), (1 << 16, 2048, "normal"), (4096, 512, "nearties"),
                                      (440, 512, "normal"), (1000, 256, "normal"), (129, 1024, "normal")])
def test_vq_search_tc_is_exact(gpu, N, K, kind):
    """tcgen05 nearest-code search (bf16 MMA scores, exact fp32 re-evaluation of every code within the rounding margin):
    same indices as the exact-fp32 kernel, and the fp64 index-parity rule holds."""
    from tests.test_gpu_kernels import check_indices
    ops = gpu.ops
    rng = np.random.default_rng(N + K)
    D = 64
    x = rng.normal(size=(N, D)).astype(np.float32)
    if kind == "init":
        E = rng.uniform(-0.05, 0.05, size=(D, K)).astype(np.float32)
    elif kind == "nearties":
        E = (x[rng.integers(0, N, K)] + 1e-3 * rng.normal(size=(K, D))).T.astype(np.float32).copy()
    else:
        E = rng.normal(size=(D, K)).astype(np.float32)
    idx0, _, _, loss0 = ops.vq_fwd(dev(x), dev(E), 0.25, True, True, None, None, 0)
    m_batch, n_batch = ops.empty(D, K), ops.empty(K)
    idx1, q_st, q, loss1 = ops.vq_fwd(dev(x), dev(E), 0.25, True, True, m_batch, n_batch, gpu._lib.PREC_BF16)
    torch.cuda.synchronize()
    check_indices(idx1, x, E)
    assert int((idx0 != idx1).sum()) == 0
    assert float(n_batch.sum()) == N and torch.equal(q.cpu(), torch.tensor(E).t()[idx1.cpu()])
    assert float(loss0) == float(loss1)


@pytest.mark.parametrize("prec", ["bf16", "bf16x2", "bf16x3"])
@pytest.mark.parametrize("B,L", [(2, 512), (3, 1000), (2, 333), (1, 2), (2, 14080)])
def test_strided_conv_tc(gpu, prec, B, L):
    """Tensor-core Conv1D(32, 4, strides=2) / Conv1DTranspose(32, 4, strides=2) forward and data gradient (conv_tc.cu)
    against the exact fp32 kernels of the same library on the same inputs; tolerance = that of the operand format."""
    ops, P = gpu.ops, gpu._lib.PRECISIONS[prec]
    g = torch.Generator(device="cuda").manual_seed(L + B)
    x = torch.randn(B, L, 32, device="cuda", generator=g)
    w = torch.randn(4, 32, 32, device="cuda", generator=g) * 0.1
    b = torch.randn(32, device="cuda", generator=g)
    Lo = (L + 1) // 2
    dy = torch.randn(B, Lo, 32, device="cuda", generator=g)
    tol = TOL[prec]
    for got, want, what in (
            (ops.conv1d_fwd(x, w, b, 2, 1, False, None, P), ops.conv1d_fwd(x, w, b, 2, 1, False, None, 0), "conv fwd"),
            (ops.conv1d_dgrad(dy, w, x.shape, None, 2, 1, False, None, P), ops.conv1d_dgrad(dy, w, x.shape, None, 2, 1, False, None, 0), "conv dgrad"),
            (ops.conv1d_transpose_fwd(x, w, b, 2, P), ops.conv1d_transpose_fwd(x, w, b, 2, 0), "convT fwd"),
            (ops.conv1d_transpose_dgrad(torch.randn(B, 2 * L, 32, device="cuda", generator=torch.Generator(device="cuda").manual_seed(7)), w, x.shape, 2, P),
             ops.conv1d_transpose_dgrad(torch.randn(B, 2 * L, 32, device="cuda", generator=torch.Generator(device="cuda").manual_seed(7)), w, x.shape, 2, 0), "convT dgrad")):
        assert got.shape == want.shape, what
        err = float((got - want).abs().max() / want.abs().max())
        assert err < tol, (what, err)


@pytest.mark.parametrize("prec", ["bf16", "bf16x2", "bf16x3"])
@pytest.mark.parametrize("B,L", [(2, 512), (3, 1000), (2, 333), (1, 2), (4, 7040)])
def test_strided_conv_wgrad_tc(gpu, prec, B, L):
    """Tensor-core weight / bias gradients of Conv1D(32, 4, strides=2) and Conv1DTranspose(32, 4, strides=2) (wgrad4_tc.cu)
    against the exact fp32 kernels of the same library."""
    ops, P = gpu.ops, gpu._lib.PRECISIONS[prec]
    g = torch.Generator(device="cuda").manual_seed(3 * L + B)
    x = torch.randn(B, L, 32, device="cuda", generator=g)
    dy = torch.randn(B, (L + 1) // 2, 32, device="cuda", generator=g)
    dyT = torch.randn(B, 2 * L, 32, device="cuda", generator=g)
    tol = TOL[prec] * (4 if prec == "bf16" else 1)
    for fn, a, b_ in ((ops.conv1d_wgrad, x, dy), (ops.conv1d_transpose_wgrad, x, dyT)):
        got_w, got_b, want_w, want_b = ops.empty(4, 32, 32), ops.empty(32), ops.empty(4, 32, 32), ops.empty(32)
        if fn is ops.conv1d_wgrad:
            fn(a, b_, got_w, got_b, 2, 1, False, P); fn(a, b_, want_w, want_b, 2, 1, False, 0)
        else:
            fn(a, b_, got_w, got_b, 2, P); fn(a, b_, want_w, want_b, 2, 0)
        assert float((got_w - want_w).abs().max() / want_w.abs().max()) < tol, fn.__name__
        assert float((got_b - want_b).abs().max() / want_b.abs().max().clamp_min(1e-6)) < tol, fn.__name__ + " bias"


@pytest.mark.parametrize("prec", ["bf16", "bf16x2", "bf16x3", "fp16x2"])
@pytest.mark.parametrize("cin,cout", [(32, 64), (64, 32)])
@pytest.mark.parametrize("B,L,dil", [(2, 512, 1), (3, 1000, 1), (2, 333, 2), (1, 1, 1), (1, 127, 8), (32, 3520, 1), (2, 129, 3)])
def test_latent_conv3_tc(gpu, prec, cin, cout, B, L, dil):
    """Tensor-core Conv1D(64, 3, 1) on 32 channels / Conv1D(32, 3, 1) on 64 channels (encdec.py:38,60; conv3_tc.cu), forward and
    data gradient, against the exact fp32 kernels of the same library and (small cases) the oracle's conv1d; tolerance = that of
    the operand format (fp16x2 runs the bf16x3 kernel here)."""
    ops, P = gpu.ops, gpu._lib.PRECISIONS[prec]
    g = torch.Generator(device="cuda").manual_seed(L + B + cin)
    x = torch.randn(B, L, cin, device="cuda", generator=g)
    w = torch.randn(3, cin, cout, device="cuda", generator=g) * 0.1
    b = torch.randn(cout, device="cuda", generator=g)
    dy = torch.randn(B, L, cout, device="cuda", generator=g)
    tol = TOL["bf16x3" if prec == "fp16x2" else prec]
    d = gpu.ops._cdesc(B, L, cin, cout, 3, 1, dil, False, P)
    assert gpu._lib.lib().vqb_conv1d_supports(gpu.ops.C.byref(d), 0) == 1 and gpu._lib.lib().vqb_conv1d_supports(gpu.ops.C.byref(d), 1) == 1
    got_y, want_y = ops.conv1d_fwd(x, w, b, 1, dil, False, None, P), ops.conv1d_fwd(x, w, b, 1, dil, False, None, 0)
    got_dx, want_dx = ops.conv1d_dgrad(dy, w, x.shape, None, 1, dil, False, None, P), ops.conv1d_dgrad(dy, w, x.shape, None, 1, dil, False, None, 0)
    for got, want, what in ((got_y, want_y, "fwd"), (got_dx, want_dx, "dgrad")):
        assert got.shape == want.shape, what
        err = float((got - want).abs().max() / want.abs().max())
        assert err < tol, (what, err)
    if B * L <= 4000:
        xo = x.cpu().requires_grad_(True)
        yo = O.conv1d(xo, w.cpu(), b.cpu(), 1, dil)
        (dxo,) = torch.autograd.grad(yo, xo, dy.cpu())
        assert float((got_y.cpu() - yo.detach()).abs().max() / yo.abs().max()) < tol
        assert float((got_dx.cpu() - dxo).abs().max() / dxo.abs().max()) < tol
