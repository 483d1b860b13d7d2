"""The fused DilatedResnet1D kernel (csrc/resstack_tc.cu: vqb_resstack_fwd / vqb_resstack_bwd_data, one launch per stack of up to
4 residual blocks, TMA-fed, activations on chip) against the oracle's per-block restatement of resnet.py:7-59 and against the
library's own per-block kernels, through the C ABI.  Arithmetic is the fp16x2 mode (fp32-grade: 2e-5 of the output's largest
magnitude, as for the per-block kernel in test_gpu_tc.py)."""
import numpy as np
import pytest
import torch

from oracle import vqvae_oracle as O

pytestmark = pytest.mark.gpu
TOL = 2e-5


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).float().cuda().contiguous()


def rel(got, want):
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    return float((got - want).abs().max() / want.abs().max().clamp_min(1e-12))


def make(rng, B, L, n, wscale=1.0, bscale=0.1):
    C = 32
    x = rng.normal(size=(B, L, C)).astype(np.float32)
    ws = [[(wscale * rng.normal(size=(3, C, C)) / np.sqrt(3 * C)).astype(np.float32), (bscale * rng.normal(size=C)).astype(np.float32),
           (wscale * rng.normal(size=(3, C, C)) / np.sqrt(3 * C)).astype(np.float32), (bscale * rng.normal(size=C)).astype(np.float32)]
          for _ in range(n)]
    return x, ws


def oracle_stack(x, ws, dils, dy=None):
    """per block: h, y (resnet.py:11-18,29); with dy also the gradients at every block input and at every h"""
    xt = torch.tensor(x, requires_grad=dy is not None)
    cur, hs, ys = xt, [], []
    for (w1, b1, w2, b2), d in zip(ws, dils):
        h = O.conv1d(torch.relu(cur), torch.tensor(w1), torch.tensor(b1), 1, d)
        cur = cur + O.conv1d(torch.relu(h), torch.tensor(w2), torch.tensor(b2), 1, 1)
        hs.append(h); ys.append(cur)
    if dy is None:
        return hs, ys, None, None
    ins = [xt] + ys[:-1]
    g = torch.autograd.grad(ys[-1], ins + hs, torch.tensor(dy))
    return hs, ys, list(g[:len(ins)]), list(g[len(ins):])


CASES = [(2, 1000, (1, 3, 9, 27)), (2, 1000, (27, 9, 3, 1)), (3, 881, (1, 3, 9, 27)), (1, 110, (27, 9, 3, 1)), (2, 296, (1, 3, 9, 27)),
         (2, 297, (1, 3)), (1, 1, (1,)), (4, 3520, (9, 27)), (1, 600, (3, 3, 3, 3)), (2, 440, (1, 1, 1))]


@pytest.fixture(params=["auto", "3", "4"])
def tile_rows(request, monkeypatch):
    """the stack kernel has two tile heights picked per call (resstack_tc.cu: use_rs4): 384 rows (rs_kernel) and 512 rows with an
    in-place operand buffer (rs4_kernel); VQB_RS_MB forces one, so every case below runs on both"""
    if request.param != "auto":
        monkeypatch.setenv("VQB_RS_MB", request.param)
    return request.param


@pytest.mark.parametrize("B,L,dils", CASES)
def test_resstack_inference_forward(gpu, tile_rows, B, L, dils):
    ops, P = gpu.ops, gpu._lib.PRECISIONS["fp16x2"]
    assert ops.resstack_supported(32, dils, P)
    rng = np.random.default_rng(L + sum(dils))
    x, ws = make(rng, B, L, len(dils))
    _, ys_ref, _, _ = oracle_stack(x, ws, dils)
    W = [[dev(a) for a in blk] for blk in ws]
    ys, hs, xb, hb, fws = ops.resstack_fwd(dev(x), [w[0] for w in W], [w[1] for w in W], [w[2] for w in W], [w[3] for w in W],
                                      dils, P, train=False)
    torch.cuda.synchronize()
    assert hs is None and all(y is None for y in ys[:-1])
    assert rel(ys[-1], ys_ref[-1]) < TOL, rel(ys[-1], ys_ref[-1])


@pytest.mark.parametrize("B,L,dils", CASES)
def test_resstack_training_forward_and_data_gradient(gpu, tile_rows, B, L, dils):
    """Under a tape the forward writes every block's h, output and sign masks; the data-gradient chain reads the masks and
    writes every block's dh / dx (the operands of vqb_resblock_wgrad_batch)."""
    ops, P = gpu.ops, gpu._lib.PRECISIONS["fp16x2"]
    rng = np.random.default_rng(L + sum(dils) + 1)
    n = len(dils)
    x, ws = make(rng, B, L, n)
    dy = rng.normal(size=(B, L, 32)).astype(np.float32)
    hs_ref, ys_ref, gx_ref, gh_ref = oracle_stack(x, ws, dils, dy)
    W = [[dev(a) for a in blk] for blk in ws]
    ys, hs, xb, hb, fws = ops.resstack_fwd(dev(x), [w[0] for w in W], [w[1] for w in W], [w[2] for w in W], [w[3] for w in W],
                                      dils, P, train=True)
    torch.cuda.synchronize()
    ins_ref = [torch.tensor(x)] + [y.detach() for y in ys_ref[:-1]]
    for i in range(n):
        assert rel(hs[i], hs_ref[i]) < TOL, ("h", i, rel(hs[i], hs_ref[i]))
        assert rel(ys[i], ys_ref[i]) < TOL, ("y", i, rel(ys[i], ys_ref[i]))
        # sign masks: bit c of word t = channel c > 0; the kernel's own h / x decide (values within rounding of 0 may differ
        # from the oracle's sign, so compare against the kernel's tensors — exactly)
        xin = dev(x) if i == 0 else ys[i - 1]
        for bits, src in ((xb[i], xin), (hb[i], hs[i])):
            want = ((src > 0).to(torch.int64) << torch.arange(32, device="cuda")).sum(-1)
            got = bits.to(torch.int64) & 0xffffffff
            assert torch.equal(got, want), ("bits", i)
    # masks of the ORACLE's tensors, so that the comparison isolates the arithmetic of the gradient chain
    def pack(t):
        w = ((t > 0).to(torch.int64) << torch.arange(32)).sum(-1)
        return (w - ((w >> 31) << 32)).to(torch.int32).cuda()
    xb_ref = [pack(t) for t in ins_ref]
    hb_ref = [pack(h.detach()) for h in hs_ref]
    dxs, dhs = ops.resstack_bwd_data(dev(dy), [w[0] for w in W], [w[2] for w in W], xb_ref, hb_ref, dils, P)
    torch.cuda.synchronize()
    for i in range(n):
        assert rel(dhs[i], gh_ref[i]) < TOL, ("dh", i, rel(dhs[i], gh_ref[i]))
        assert rel(dxs[i], gx_ref[i]) < TOL, ("dx", i, rel(dxs[i], gx_ref[i]))


@pytest.mark.parametrize("dils", [(1, 3, 9, 27), (27, 9, 3, 1)])
def test_resstack_full_size_matches_block_kernels(gpu, tile_rows, dils):
    """[32, 14080, 32] (the largest stage of SMALL_VQ_VAE at batch 32): the fused stack against the chain of per-block fp16x2
    launches and the exact-fp32 CUDA-core kernels."""
    ops, P = gpu.ops, gpu._lib.PRECISIONS["fp16x2"]
    g = torch.Generator(device="cuda").manual_seed(1)
    B, L, C = 32, 14080, 32
    x = torch.randn(B, L, C, device="cuda", generator=g)
    W = [[torch.randn(3, C, C, device="cuda", generator=g) / 96 ** 0.5, torch.randn(C, device="cuda", generator=g) * 0.1,
          torch.randn(3, C, C, device="cuda", generator=g) / 96 ** 0.5, torch.randn(C, device="cuda", generator=g) * 0.1]
         for _ in dils]
    dy = torch.randn(B, L, C, device="cuda", generator=g)
    ys, hs, xb, hb, fws = ops.resstack_fwd(x, [w[0] for w in W], [w[1] for w in W], [w[2] for w in W], [w[3] for w in W], dils, P, True)
    yi, _, _, _, _ = ops.resstack_fwd(x, [w[0] for w in W], [w[1] for w in W], [w[2] for w in W], [w[3] for w in W], dils, P, False)
    cur = x
    for i, d in enumerate(dils):
        y0, h0 = ops.resblock_fwd(cur, *W[i], d, 0)
        assert rel(hs[i], h0) < TOL and rel(ys[i], y0) < TOL, (i, rel(hs[i], h0), rel(ys[i], y0))
        cur = y0
    assert rel(yi[-1], cur) < TOL
    assert torch.equal(yi[-1], ys[-1])  # the inference and the training instantiation compute the same numbers
    dxs, dhs = ops.resstack_bwd_data(dy, [w[0] for w in W], [w[2] for w in W], xb, hb, dils, P)
    # ... and from the operand images the training forward packed (vqb_resstack_bwd_data_packed: no packing launch): same bits
    dxp, dhp = ops.resstack_bwd_data(dy, None, None, xb, hb, dils, P, fwd_ws=fws)
    assert all(torch.equal(a, b) for a, b in zip(dxs + dhs, dxp + dhp))
    gcur = dy
    for i in reversed(range(len(dils))):
        xin = x if i == 0 else ys[i - 1]
        dx0, dh0 = ops.resblock_bwd_data(xin, hs[i], gcur, W[i][0], W[i][2], dils[i], 0)
        assert rel(dhs[i], dh0) < TOL and rel(dxs[i], dx0) < TOL, (i, rel(dhs[i], dh0), rel(dxs[i], dx0))
        gcur = dx0


@pytest.mark.parametrize("xs,ws,bs", [(1e-6, 1.0, 0.0), (3e4, 1.0, 0.1), (1.0, 1e-3, 1e-5), (1e-3, 8.0, 10.0), (0.0, 1.0, 0.3)])
def test_resstack_is_scale_free(gpu, tile_rows, xs, ws, bs):
    """operand scales are chosen per tile and convolution from exact maxima and L1 bounds: magnitudes far outside fp16's range,
    growing or shrinking through the stack, and an all-zero input must all stay fp32-grade"""
    ops, P = gpu.ops, gpu._lib.PRECISIONS["fp16x2"]
    rng = np.random.default_rng(3)
    dils = (1, 3, 9, 27)
    x, wts = make(rng, 2, 700, 4, wscale=ws, bscale=bs)
    x = (x * xs).astype(np.float32)
    x[1, 300:] *= 1e-3  # tiles of very different magnitude inside one launch
    _, ys_ref, _, _ = oracle_stack(x, wts, dils)
    W = [[dev(a) for a in blk] for blk in wts]
    ys, *_ = ops.resstack_fwd(dev(x), [w[0] for w in W], [w[1] for w in W], [w[2] for w in W], [w[3] for w in W], dils, P, False)
    torch.cuda.synchronize()
    assert torch.isfinite(ys[-1]).all()
    assert rel(ys[-1], ys_ref[-1]) < TOL, rel(ys[-1], ys_ref[-1])


def test_resstack_unsupported_shapes_are_refused(gpu):
    ops, L = gpu.ops, gpu._lib
    P = L.PRECISIONS["fp16x2"]
    assert not ops.resstack_supported(64, (1, 3), P)            # width
    assert not ops.resstack_supported(32, (1, 3, 9, 27, 1), P)  # more than 4 blocks per launch
    assert not ops.resstack_supported(32, (81,), P)             # dilation beyond the guard rows
    assert not ops.resstack_supported(32, (1, 3), L.PRECISIONS["fp32"])
    x = torch.zeros(1, 64, 64, device="cuda")
    w = torch.zeros(3, 64, 64, device="cuda"); b = torch.zeros(64, device="cuda")
    with pytest.raises(L.VQBError):
        ops.resstack_fwd(x, [w], [b], [w], [b], (1,), P, False)


@pytest.mark.parametrize("B,L", [(2, 1000), (32, 3520)])
def test_resstack_is_independent_of_what_ran_before(gpu, tile_rows, B, L):
    """The kernel keeps state across tiles in shared / tensor memory (operand buffer rewritten in place, guard rows, per-tile
    maxima, accumulators): the result of a call must not depend on what the SMs processed before it — bit for bit, also after a
    launch full of huge values and after a launch of another shape."""
    ops, P = gpu.ops, gpu._lib.PRECISIONS["fp16x2"]
    rng = np.random.default_rng(11)
    dils = (1, 3, 9, 27)
    x, wts = make(rng, B, L, 4)
    W = [[dev(a) for a in blk] for blk in wts]
    args = ([w[0] for w in W], [w[1] for w in W], [w[2] for w in W], [w[3] for w in W], dils, P)
    xd = dev(x)
    first = ops.resstack_fwd(xd, *args, True)
    junk = torch.full((B, L + 777, 32), 3.0e30, device="cuda")
    junk[:, ::3] = -1.0e-30
    ops.resstack_fwd(junk, *args, False)
    again = ops.resstack_fwd(xd, *args, True)
    for a, b in zip(first[0] + first[1] + first[2] + first[3], again[0] + again[1] + again[2] + again[3]):
        assert torch.equal(a, b)
    dy = dev(rng.normal(size=(B, L, 32)).astype(np.float32))
    g1 = ops.resstack_bwd_data(dy, None, None, first[2], first[3], dils, P, fwd_ws=first[4])
    ops.resstack_fwd(junk, *args, False)
    g2 = ops.resstack_bwd_data(dy, None, None, first[2], first[3], dils, P, fwd_ws=first[4])
    for a, b in zip(g1[0] + g1[1], g2[0] + g2[1]):
        assert torch.equal(a, b)
