"""Parity of the CUDA kernels (through the C ABI, via vqvae_b200.ops) against the CPU oracle on identical inputs.
fp32 contractions: |err| <= 1e-4 * scale (different summation order only); integer results bit-exact;
EMA update bit-exact (separately rounded ops)."""
import os

import numpy as np
import pytest
import torch

from oracle import vqvae_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a)).to(dtype).cuda().contiguous()


def close(got, want, tol=1e-4, what=""):
    got = got.detach().cpu().double().numpy() if torch.is_tensor(got) else np.asarray(got, np.float64)
    want = want.detach().cpu().double().numpy() if torch.is_tensor(want) else np.asarray(want, np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    scale = max(float(np.abs(want).max()), 1e-6)
    err = float(np.abs(got - want).max())
    assert err <= tol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


CONV_CASES = [  # B, L, cin, cout, k, s, d, relu
    (2, 28160 // 16, 1, 32, 4, 2, 1, 0),   # first stage (Cin = 1)
    (1, 7, 1, 32, 4, 2, 1, 0), (3, 1001, 1, 32, 4, 2, 1, 0), (2, 333, 1, 16, 6, 2, 1, 0), (2, 64, 1, 8, 3, 1, 1, 0),  # ... ragged / other taps
    (3, 500, 32, 32, 3, 1, 1, 1), (3, 500, 32, 32, 3, 1, 3, 1), (2, 500, 32, 32, 3, 1, 9, 1),
    (2, 501, 32, 32, 3, 1, 27, 1),         # res-block convs, odd length
    (2, 40, 32, 32, 3, 1, 27, 1),          # window shorter than the receptive field
    (2, 440, 32, 32, 4, 2, 1, 0), (2, 441, 64, 32, 4, 2, 1, 0),  # down-sampling, odd length pads (1,2)
    (2, 300, 32, 64, 3, 1, 1, 0), (2, 300, 64, 32, 3, 1, 1, 0),  # proj / pre
    (2, 1000, 64, 1, 3, 1, 1, 0),          # final 64 -> 1
    (1, 257, 64, 64, 3, 1, 3, 1), (2, 100, 128, 32, 3, 1, 1, 1), (1, 64, 8, 24, 6, 3, 1, 0),  # other widths / stride 3
    (1, 1, 32, 32, 3, 1, 1, 0),
]


@pytest.mark.parametrize("B,L,cin,cout,k,s,d,relu", CONV_CASES)
def test_conv1d_fwd_dgrad_wgrad(gpu, B, L, cin, cout, k, s, d, relu):
    ops = gpu.ops
    rng = np.random.default_rng(B * 1000 + L + cin + 7 * d)
    x = rng.normal(size=(B, L, cin)).astype(np.float32)
    w = (rng.normal(size=(k, cin, cout)) / np.sqrt(k * cin)).astype(np.float32)
    b = rng.normal(size=cout).astype(np.float32)
    Lo = -(-L // s)
    res = rng.normal(size=(B, Lo, cout)).astype(np.float32)
    dy = rng.normal(size=(B, Lo, cout)).astype(np.float32)
    add = rng.normal(size=(B, L, cin)).astype(np.float32)
    xt = torch.tensor(x, requires_grad=True); wt = torch.tensor(w, requires_grad=True); bt = torch.tensor(b, requires_grad=True)
    act = torch.relu(xt) if relu else xt
    y_ref = O.conv1d(act, wt, bt, s, d)
    gx, gw, gb = torch.autograd.grad(y_ref, (xt, wt, bt), torch.tensor(dy))
    y = ops.conv1d_fwd(dev(x), dev(w), dev(b), s, d, relu, dev(res))
    close(y, y_ref.detach() + torch.tensor(res), what="fwd")
    dx = ops.conv1d_dgrad(dev(dy), dev(w), x.shape, dev(x) if relu else None, s, d, relu, dev(add))
    close(dx, gx + torch.tensor(add), what="dgrad")
    dw, db = torch.empty_like(dev(w)), torch.empty_like(dev(b))
    ops.conv1d_wgrad(dev(x), dev(dy), dw, db, s, d, relu)
    close(dw, gw, 2e-4, "wgrad"); close(db, gb, 2e-4, "bgrad")


@pytest.mark.parametrize("B,L,cin,cout,k,s", [(2, 440, 32, 32, 4, 2), (2, 441, 32, 64, 4, 2), (1, 1, 32, 32, 4, 2),
                                              (2, 50, 16, 8, 6, 3), (2, 64, 128, 32, 4, 2), (1, 30, 8, 8, 2, 1)])
def test_conv1d_transpose_fwd_dgrad_wgrad(gpu, B, L, cin, cout, k, s):
    ops = gpu.ops
    rng = np.random.default_rng(L + cin)
    x = rng.normal(size=(B, L, cin)).astype(np.float32)
    w = (rng.normal(size=(k, cout, cin)) / np.sqrt(k * cin)).astype(np.float32)
    b = rng.normal(size=cout).astype(np.float32)
    dy = rng.normal(size=(B, L * s, cout)).astype(np.float32)
    xt = torch.tensor(x, requires_grad=True); wt = torch.tensor(w, requires_grad=True); bt = torch.tensor(b, requires_grad=True)
    y_ref = O.conv1d_transpose(xt, wt, bt, s)
    gx, gw, gb = torch.autograd.grad(y_ref, (xt, wt, bt), torch.tensor(dy))
    close(ops.conv1d_transpose_fwd(dev(x), dev(w), dev(b), s), y_ref.detach(), what="fwd")
    close(ops.conv1d_transpose_dgrad(dev(dy), dev(w), x.shape, s), gx, what="dgrad")
    dw, db = torch.empty_like(dev(w)), torch.empty_like(dev(b))
    ops.conv1d_transpose_wgrad(dev(x), dev(dy), dw, db, s)
    close(dw, gw, 2e-4, "wgrad"); close(db, gb, 2e-4, "bgrad")


def test_golden_primitives(gpu):
    ops = gpu.ops
    g = np.load(os.path.join(GOLD, "primitives.npz"))
    for name in ("c_k3d1", "c_k3d9", "c_k4s2", "c_k4s2_odd", "c_k3_out1"):
        k, s, d = (int(v) for v in g[f"{name}.cfg"])
        close(ops.conv1d_fwd(dev(g[f"{name}.x"]), dev(g[f"{name}.w"]), dev(g[f"{name}.b"]), s, d), g[f"{name}.y"], what=name)
    for name in ("t_k4s2", "t_k6s3"):
        k, s = (int(v) for v in g[f"{name}.cfg"])
        close(ops.conv1d_transpose_fwd(dev(g[f"{name}.x"]), dev(g[f"{name}.w"]), dev(g[f"{name}.b"]), s), g[f"{name}.y"], what=name)


@pytest.mark.parametrize("B,L,C,F,d", [(3, 880, 32, 32, 1), (2, 881, 32, 32, 27), (2, 300, 32, 32, 9), (1, 200, 64, 32, 3),
                                       (2, 20, 32, 32, 27)])
def test_resblock_fwd_bwd(gpu, B, L, C, F, d):
    ops = gpu.ops
    rng = np.random.default_rng(L + d)
    x = rng.normal(size=(B, L, C)).astype(np.float32)
    w1 = (rng.normal(size=(3, C, F)) / np.sqrt(3 * C)).astype(np.float32); b1 = rng.normal(size=F).astype(np.float32) * 0.1
    w2 = (rng.normal(size=(3, F, C)) / np.sqrt(3 * F)).astype(np.float32); b2 = rng.normal(size=C).astype(np.float32) * 0.1
    dy = rng.normal(size=(B, L, C)).astype(np.float32)
    ts = [torch.tensor(a, requires_grad=True) for a in (x, w1, b1, w2, b2)]
    y_ref = O.resblock(*ts, d)
    grads = torch.autograd.grad(y_ref, ts, torch.tensor(dy))
    y, h = ops.resblock_fwd(dev(x), dev(w1), dev(b1), dev(w2), dev(b2), d)
    close(y, y_ref.detach(), what="y")
    dx, dh = ops.resblock_bwd_data(dev(x), h, dev(dy), dev(w1), dev(w2), d)
    close(dx, grads[0], what="dx")
    dw2, db2, dw1, db1 = (torch.empty_like(dev(a)) for a in (w2, b2, w1, b1))
    ops.conv1d_wgrad(h, dev(dy), dw2, db2, 1, 1, True)
    ops.conv1d_wgrad(dev(x), dh, dw1, db1, 1, d, True)
    close(dw1, grads[1], 2e-4, "dw1"); close(db1, grads[2], 2e-4, "db1"); close(dw2, grads[3], 2e-4, "dw2"); close(db2, grads[4], 2e-4, "db2")


def check_indices(idx, x, E):
    """Index-parity rule (north star / SURVEY 8c): equal to the exact (fp64) argmin except where the top-2 distance gap
    is below 1e-5 relative.  "Relative" is taken against the magnitude the reference's own fp32 expression carries,
    ||x||^2 + ||e||^2 (VectorQuantizer.py:175-182 forms (xx + ee) - 2 x.e, so its rounding error scales with that sum,
    not with the possibly tiny distance itself)."""
    xt, Et = torch.as_tensor(x).double(), torch.as_tensor(E).double()
    d64 = O.vq_distances(xt, Et)
    srt = torch.sort(d64, dim=1)
    ref = d64.argmin(1)
    scale = (xt ** 2).sum(1) + (Et ** 2).sum(0)[ref]
    gap_ok = (srt.values[:, 1] - srt.values[:, 0]) > 1e-5 * scale
    got = idx.cpu()
    bad = (got != ref) & gap_ok
    assert int(bad.sum()) == 0, f"{int(bad.sum())} indices differ outside the near-tie allowance"
    # and a differing index inside the allowance must still be (near-)optimal
    sel = d64.gather(1, got[:, None])[:, 0]
    assert float(((sel - srt.values[:, 0]) / scale).max()) < 1e-5
    return ref


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("N,D,K,kind", [(28160, 64, 512, "normal"), (1 << 16, 64, 2048, "normal"), (4096, 64, 512, "nearties"),
                                        (440, 64, 512, "init"), (1000, 48, 100, "normal"), (33, 200, 9, "normal"), (5, 64, 512, "normal")])
def test_vq_forward_bf16_activations(gpu, prec, N, D, K, kind):
    """vqb_vq_fwd_bf16 (BASELINE configs[3] "bf16"): x, q_st, q in bfloat16, everything in between the fp32 arithmetic of the
    fp32 entry applied to float(x).  Oracle = VectorQuantizer.py:86-124,170-186 on float(x): indices by the same near-tie rule,
    q = bf16_rn(E[:, idx]) and q_st = bf16_rn(fl(x + fl(q - x))) bit for bit, loss / statistics as in the fp32 test.  prec = the
    search arithmetic (exact fp32 CUDA cores / tensor cores + exact re-ranking)."""
    ops = gpu.ops
    rng = np.random.default_rng(N + K + 1)
    xb = torch.tensor(rng.normal(size=(N, D)).astype(np.float32)).bfloat16()
    x = xb.float().numpy()
    if kind == "init":
        E = rng.uniform(-0.05, 0.05, size=(D, K)).astype(np.float32)
    elif kind == "nearties":
        E = (x[rng.integers(0, N, K)] + 1e-3 * rng.normal(size=(K, D))).T.astype(np.float32).copy()
    else:
        E = rng.normal(size=(D, K)).astype(np.float32)
    m_batch, n_batch = ops.empty(D, K), ops.empty(K)
    P = gpu._lib.PRECISIONS[prec]
    idx, q_st, q, loss = ops.vq_fwd(xb.cuda().contiguous(), dev(E), 0.25, True, True, m_batch, n_batch, P)
    assert idx.dtype == torch.int64 and q_st.dtype == torch.bfloat16 and q.dtype == torch.bfloat16
    check_indices(idx, x, E)
    ic = idx.cpu()
    qr = torch.tensor(E).t()[ic]
    xt = torch.tensor(x)
    assert torch.equal(q.cpu(), qr.bfloat16())
    assert torch.equal(q_st.cpu(), (xt + (qr - xt)).bfloat16())
    want_loss = 0.25 * ((qr.double() - xt.double()) ** 2).mean()
    assert abs(float(loss) - float(want_loss)) <= 1e-5 * float(want_loss)
    mb, nb = O.vq_batch_stats(xt.double(), ic, K)
    assert torch.equal(n_batch.cpu().double(), nb)
    close(m_batch, mb, 1e-5, "m_batch")
    # the fp32 entry on float(x) picks the same codes
    idx32, _, _, loss32 = ops.vq_fwd(dev(x), dev(E), 0.25, False, False, None, None, P)
    assert torch.equal(idx32, idx) and float(loss32) == float(loss)


@pytest.mark.parametrize("N,D,K,kind", [(28160, 64, 512, "normal"), (3520, 64, 512, "init"), (1 << 16, 64, 2048, "normal"),
                                        (4096, 64, 512, "nearties"), (3200, 2, 6, "normal"), (440, 64, 512, "normal"),
                                        (1000, 48, 100, "normal"), (33, 200, 9, "normal")])
def test_vq_forward(gpu, N, D, K, kind):
    ops = gpu.ops
    rng = np.random.default_rng(N + K)
    x = rng.normal(size=(N, D)).astype(np.float32)
    if kind == "init":
        E = rng.uniform(-0.05, 0.05, size=(D, K)).astype(np.float32)
    elif kind == "nearties":
        E = (x[rng.integers(0, N, K)] + 1e-3 * rng.normal(size=(K, D))).T.astype(np.float32).copy()
    else:
        E = rng.normal(size=(D, K)).astype(np.float32)
    m_batch, n_batch = ops.empty(D, K), ops.empty(K)
    idx, q_st, q, loss = ops.vq_fwd(dev(x), dev(E), 0.25, True, True, m_batch, n_batch)
    assert idx.dtype == torch.int64
    check_indices(idx, x, E)
    ic = idx.cpu()
    qr = torch.tensor(E).t()[ic]
    assert torch.equal(q.cpu(), qr)                                   # gather is exact
    assert torch.equal(q_st.cpu(), torch.tensor(x) + (qr - torch.tensor(x)))  # fl(x + fl(q - x)), bit for bit
    want_loss = 0.25 * ((qr.double() - torch.tensor(x).double()) ** 2).mean()
    assert abs(float(loss) - float(want_loss)) <= 1e-5 * float(want_loss)
    mb, nb = O.vq_batch_stats(torch.tensor(x).double(), ic, K)
    assert torch.equal(n_batch.cpu().double(), nb)                    # counts are exact
    close(m_batch, mb, 1e-5, "m_batch")
    # inference form: no statistics, same indices
    idx2, _, _, _ = ops.vq_fwd(dev(x), dev(E), 0.25, False, False)
    assert torch.equal(idx2, idx)


def test_vq_golden_and_bwd(gpu):
    ops = gpu.ops
    v = np.load(os.path.join(GOLD, "vq.npz"))
    idx, q_st, q, loss = ops.vq_fwd(dev(v["x"]), dev(v["E"]), 0.25)
    safe = (v["top2"][:, 1] - v["top2"][:, 0]) > 1e-5 * v["top2"][:, 0]
    assert np.array_equal(idx.cpu().numpy()[safe], v["idx"][safe])
    if np.array_equal(idx.cpu().numpy(), v["idx"]):
        assert np.array_equal(q_st.cpu().numpy(), v["q_st"])
        assert abs(float(loss) - float(v["commit"])) < 1e-6 * float(v["commit"])
    dq = np.random.default_rng(0).normal(size=v["x"].shape).astype(np.float32)
    dx = ops.vq_bwd(dev(dq), dev(v["x"]), q, 0.25, 1.0)
    want = dq + 2 * 0.25 / v["x"].size * (v["x"] - q.cpu().numpy())
    close(dx, want, 1e-6, "vq_bwd")


def test_vq_empty_batch(gpu):
    ops = gpu.ops
    m_batch, n_batch = ops.empty(4, 8), ops.empty(8)
    idx, q_st, q, loss = ops.vq_fwd(ops.empty(0, 4), dev(np.ones((4, 8))), 0.25, True, True, m_batch, n_batch)
    assert idx.numel() == 0 and float(loss) == 0.0 and float(n_batch.sum()) == 0.0


def test_ema_update_bit_exact(gpu):
    ops = gpu.ops
    v = np.load(os.path.join(GOLD, "vq.npz"))
    E, m_t, N_t = dev(v["E"]), dev(v["m_t_in"]), dev(v["N_t_in"])
    met = ops.zeros(3)
    ops.vq_ema_update(E, m_t, N_t, dev(v["m_batch"]), dev(v["n_batch"]), dev(v["rows"]), 0.99, 1.0, met)
    assert np.array_equal(N_t.cpu().numpy(), v["N_t_out"])
    assert np.array_equal(m_t.cpu().numpy(), v["m_t_out"])
    assert np.array_equal(E.cpu().numpy(), v["E_out"])
    assert (v["N_t_out"] < 1.0).any() and (v["N_t_out"] >= 1.0).any()  # both the alive and the restart branch ran
    np.testing.assert_allclose(met.cpu().numpy(), v["metrics"], rtol=1e-5)


def test_ema_knife_edge(gpu):
    """a code hit exactly once on step 1 has N_t = fl(fl(.99*1)+fl(.01*1)); FMA contraction would flip
    alive/dead.  Compare against the oracle's separately rounded arithmetic for hit counts 0..5."""
    ops = gpu.ops
    K, D = 6, 4
    E = np.arange(D * K, dtype=np.float32).reshape(D, K) / 7
    n_batch = np.arange(K, dtype=np.float32)
    m_batch = (np.arange(D * K, dtype=np.float32).reshape(D, K) * 0.37).astype(np.float32)
    rows = np.full((K, D), -1.0, np.float32)
    st = O.VQState(torch.tensor(E), torch.tensor(E), torch.ones(K))
    new, _ = O.vq_ema_update(st, torch.tensor(m_batch), torch.tensor(n_batch), torch.tensor(rows))
    Ed, md, Nd = dev(E), dev(E), dev(np.ones(K))
    ops.vq_ema_update(Ed, md, Nd, dev(m_batch), dev(n_batch), dev(rows), 0.99, 1.0)
    assert np.array_equal(Nd.cpu().numpy(), new.N_t.numpy())
    assert np.array_equal(Ed.cpu().numpy(), new.E.numpy())
    assert np.array_equal((Nd.cpu().numpy() >= 1.0), (new.N_t.numpy() >= 1.0))


def test_restart_rows(gpu):
    ops = gpu.ops
    step = ops.zeros(1, dtype=torch.int64)
    for N, K in ((28160, 512), (440, 512), (7, 512), (512, 512)):
        ids = ops.restart_ids(N, K, 1234, step).cpu().numpy()
        Nt = N if N >= K else N * (-(-K // N))
        assert ids.min() >= 0 and ids.max() < Nt and len(set(ids.tolist())) == K  # K distinct rows of the tiled batch
    a = ops.restart_ids(28160, 512, 1234, step).cpu().numpy()
    # a pseudo-random sample, not an arithmetic progression: the gaps between consecutive ids vary (tf.random.shuffle stand-in)
    assert len(set(np.diff(a).tolist())) > 400
    # the CPU double of the ABI (tests/fake_backend.py) implements the same permutation: host-logic tests see the device's picks
    from tests.fake_backend import FakeBackend
    import ctypes as C
    buf = np.zeros(512, np.int64); st0 = np.zeros(1, np.int64)
    FakeBackend().vqb_restart_ids(28160, 512, 1234, st0.ctypes.data, buf.ctypes.data, None)
    assert np.array_equal(buf, a)
    ops.increment(step)
    b = ops.restart_ids(28160, 512, 1234, step).cpu().numpy()
    assert not np.array_equal(a, b) and np.array_equal(a, ops.restart_ids(28160, 512, 1234, ops.zeros(1, dtype=torch.int64)).cpu().numpy())
    x = np.random.default_rng(0).normal(size=(440, 8)).astype(np.float32)
    ids = ops.restart_ids(440, 512, 5, step)
    rows = ops.gather_rows(dev(x), ids)
    assert np.array_equal(rows.cpu().numpy(), np.tile(x, (2, 1))[ids.cpu().numpy()])  # == shuffle(_tile(x))[:K] for that permutation
    # two-rank ownership split sums to the single-rank pick
    r0 = ops.gather_rows(dev(x[:220]), ids, 440, 0); r1 = ops.gather_rows(dev(x[220:]), ids, 440, 220)
    assert torch.equal(r0 + r1, rows)


def test_gather_codes_mse_adam(gpu):
    ops = gpu.ops
    rng = np.random.default_rng(0)
    E = rng.normal(size=(16, 40)).astype(np.float32)
    idx = rng.integers(0, 40, size=(3, 50))
    assert np.array_equal(ops.gather_codes(dev(E), dev(idx, torch.int64)).cpu().numpy(), E.T[idx])
    x = rng.normal(size=(5, 1000, 1)).astype(np.float32); r = rng.normal(size=(5, 1000, 1)).astype(np.float32)
    loss, dr = ops.mse(dev(x), dev(r), loss_scale=1.0)
    assert abs(float(loss) - float(((x.astype(np.float64) - r) ** 2).mean())) < 1e-6
    close(dr, 2 * (r - x) / x.size, 1e-6, "dmse")
    n = 100003
    p = rng.normal(size=n).astype(np.float32); g = rng.normal(size=n).astype(np.float32)
    pd, md, vd = dev(p), ops.zeros(n), ops.zeros(n)
    step = ops.zeros(1, dtype=torch.int64)
    pt, mt, vt = [torch.tensor(p.copy())], [torch.zeros(n)], [torch.zeros(n)]
    for t in (1, 2, 3):
        ops.adam_step(pd, dev(g), md, vd, 1e-3, 0.9, 0.999, 1e-7, 1.0, step); ops.increment(step)
        O.adam_step(pt, [torch.tensor(g)], mt, vt, t)
    close(pd, pt[0], 1e-6, "adam")


def test_errors_are_reported_not_swallowed(gpu):
    ops = gpu.ops
    with pytest.raises(ValueError):
        ops.conv1d_fwd(dev(np.zeros((1, 8, 4))), dev(np.zeros((3, 5, 4))), None)
    with pytest.raises(gpu._lib.VQBError, match="k="):
        ops.conv1d_fwd(dev(np.zeros((1, 8, 4))), dev(np.zeros((17, 4, 4))), None)
    with pytest.raises(gpu._lib.VQBError, match="not built|precision|UNIMPLEMENTED|available"):
        ops.conv1d_fwd(dev(np.zeros((1, 8, 4))), dev(np.zeros((3, 4, 4))), None, precision=7)


TAIL_CASES = [  # B, L, cin, cmid, bias
    (3, 37, 8, 16, True), (2, 300, 32, 64, True), (2, 513, 32, 64, False), (1, 1, 32, 64, True), (4, 2, 4, 8, True),
    (2, 28160 // 8, 32, 64, True),
]


@pytest.mark.parametrize("B,L,cin,cmid,bias", TAIL_CASES)
def test_decoder_tail_matches_two_layer_oracle(gpu, B, L, cin, cmid, bias):
    """vqb_dec_tail_*: Conv1DTranspose(cmid, 4, 2) followed by Conv1D(1, 3) (encdec.py:67-68,148) run as ONE composed
    operator equals the two layers evaluated one after the other — forward, input gradient and all four parameter
    gradients (fp32; only the association of the sums differs)."""
    ops = gpu.ops
    rng = np.random.default_rng(B * 100 + L + cin)
    x = rng.normal(size=(B, L, cin)).astype(np.float32)
    wt = (rng.normal(size=(4, cmid, cin)) / np.sqrt(2 * cin)).astype(np.float32)
    wf = (rng.normal(size=(3, cmid, 1)) / np.sqrt(3 * cmid)).astype(np.float32)
    bt = rng.normal(size=cmid).astype(np.float32) if bias else None
    bf = rng.normal(size=1).astype(np.float32) if bias else None
    dr = rng.normal(size=(B, 2 * L, 1)).astype(np.float32)
    T = lambda a: None if a is None else torch.tensor(a, dtype=torch.float64, requires_grad=True)
    xt, wtt, wft, btt, bft = T(x), T(wt), T(wf), T(bt), T(bf)
    r_ref = O.conv1d(O.conv1d_transpose(xt, wtt, btt, 2), wft, bft, 1, 1)
    leaves = [t for t in (xt, wtt, wft, btt, bft) if t is not None]
    g = torch.autograd.grad(r_ref, leaves, torch.tensor(dr, dtype=torch.float64))
    gx, gwt, gwf = g[0], g[1], g[2]
    D = lambda a: None if a is None else dev(a)
    recon, gbuf = ops.dec_tail_fwd(dev(x), dev(wt), D(bt), dev(wf), D(bf))
    close(recon, r_ref.detach(), tol=2e-5, what="recon")
    dwt, dwf = ops.empty(4, cmid, cin), ops.empty(3, cmid, 1)
    dbt, dbf = (ops.empty(cmid), ops.empty(1)) if bias else (None, None)
    dx = ops.dec_tail_bwd(dev(x), dev(dr), dev(wt), D(bt), dev(wf), gbuf, dwt, dbt, dwf, dbf)
    close(dx, gx, tol=2e-5, what="dx")
    close(dwt, gwt, tol=5e-5, what="dwt")
    close(dwf, gwf, tol=5e-5, what="dwf")
    if bias:
        close(dbt, g[3], tol=5e-5, what="dbt")
        close(dbf, g[4], tol=5e-5, what="dbf")
    # and against the library's own two-layer path
    y = ops.conv1d_transpose_fwd(dev(x), dev(wt), D(bt) if bias else ops.zeros(cmid), 2)
    r2 = ops.conv1d_fwd(y, dev(wf), D(bf) if bias else ops.zeros(1), 1, 1, False, None)
    close(recon, r2, tol=2e-5, what="recon vs two kernels")


@pytest.mark.parametrize("B,T", [(3, 2048), (2, 28160), (1, 4096)])
def test_multispectral_loss_kernels_match_oracle(gpu, B, T):
    """csrc/spectral.cu around cuFFT (framing + periodic Hann + end padding, magnitudes, Frobenius sums, loss, gradient w.r.t.
    the spectrum, overlap-add) against the oracle's tf.signal.stft restatement and its autograd gradient
    (vqvae.py:309-326, data_utils.py:25-40)."""
    V = gpu
    rng = np.random.default_rng(T + B)
    x = rng.uniform(0, 1, size=(B, T, 1)).astype(np.float32)
    r = (x + 0.1 * rng.normal(size=(B, T, 1))).astype(np.float32)
    rt = torch.tensor(r, requires_grad=True)
    want = O.multispectral_loss(torch.tensor(x).squeeze(-1), rt.squeeze(-1)).mean()
    (gwant,) = torch.autograd.grad(want, rt)
    V.data_utils.clear_cache()
    rd = dev(r)
    with V.GradientTape() as tape:
        loss = V.data_utils.MultiSpectralLoss(dev(x), rd)._reduce_mean()
    node = tape.nodes[-1]
    (dr,) = node.bwd([1.0], [True])
    assert abs(float(loss) - float(want)) <= 2e-5 * abs(float(want))
    close(dr, gwant, tol=2e-4, what="d loss / d recon")
    # the public helpers keep the reference's semantics (data_utils.py:25-40)
    s = V.data_utils.spectral(dev(x).squeeze(-1), 512, 50, 240)
    close(s, O.spectral(torch.tensor(x).squeeze(-1), 512, 50, 240), tol=2e-5, what="spectral")


def test_lincomb_one_launch(gpu):
    """vqb_lincomb: the loss bookkeeping of train_step (vqvae.py:127-146) as one launch; terms are added in order, in fp32."""
    ops = gpu.ops
    g = torch.Generator(device="cuda").manual_seed(11)
    leaves = [torch.randn(3, device="cuda", generator=g) for _ in range(7)]
    outs = [[(leaves[0][0:1], 1.0), (leaves[1][1:2], 1.0), (leaves[2][2:3], 1.0)], [], [(leaves[3][0:1], 0.25)],
            [(leaves[i][0:1], float(i)) for i in range(7)]]
    got = ops.lincomb(outs).cpu()
    want = []
    for terms in outs:
        acc = torch.zeros((), dtype=torch.float32)
        for t, c in terms:
            acc = acc + torch.tensor(c, dtype=torch.float32) * t.cpu()[0]
        want.append(acc)
    assert torch.equal(got, torch.stack(want))
