"""Pins the oracle (oracle/vqvae_oracle.py) against first-principles loop implementations (oracle/naive.py), its own
fp64 twin, autograd consistency and the committed golden fixtures.  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import naive as Nv
from oracle import vqvae_oracle as O
from oracle.make_golden import TINY, tiny_case

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("L,cin,cout,k,s,d", [(17, 3, 5, 3, 1, 1), (16, 3, 5, 3, 1, 3), (20, 2, 4, 4, 2, 1),
                                              (21, 2, 4, 4, 2, 1), (30, 4, 3, 3, 1, 9), (18, 3, 2, 6, 3, 1),
                                              (5, 2, 2, 3, 1, 27), (1, 2, 2, 3, 1, 1)])
def test_conv1d_matches_loops(L, cin, cout, k, s, d):
    rng = np.random.default_rng(L * 31 + k)
    x = rng.normal(size=(2, L, cin)).astype(np.float32)
    w = rng.normal(size=(k, cin, cout)).astype(np.float32)
    b = rng.normal(size=cout).astype(np.float32)
    y = O.conv1d(torch.tensor(x).double(), torch.tensor(w).double(), torch.tensor(b).double(), s, d).numpy()
    np.testing.assert_allclose(y, Nv.conv1d_naive(x, w, b, s, d), atol=1e-12)


@pytest.mark.parametrize("L,cin,cout,k,s", [(9, 3, 5, 4, 2), (8, 2, 4, 6, 3), (7, 2, 3, 2, 1), (5, 2, 2, 3, 2), (1, 2, 2, 4, 2)])
def test_conv1d_transpose_matches_loops(L, cin, cout, k, s):
    rng = np.random.default_rng(L * 17 + k)
    x = rng.normal(size=(2, L, cin)).astype(np.float32)
    w = rng.normal(size=(k, cout, cin)).astype(np.float32)
    b = rng.normal(size=cout).astype(np.float32)
    y = O.conv1d_transpose(torch.tensor(x).double(), torch.tensor(w).double(), torch.tensor(b).double(), s).numpy()
    assert y.shape == (2, L * s, cout)
    np.testing.assert_allclose(y, Nv.conv1d_transpose_naive(x, w, b, s), atol=1e-12)


def test_conv_transpose_k4s2_two_phase_formula():
    """y[2m] = x[m]W1 + x[m-1]W3, y[2m+1] = x[m+1]W0 + x[m]W2 (+bias)."""
    rng = np.random.default_rng(0)
    x = rng.normal(size=(1, 6, 2)); w = rng.normal(size=(4, 3, 2)); b = rng.normal(size=3)
    y = O.conv1d_transpose(torch.tensor(x), torch.tensor(w), torch.tensor(b), 2).numpy()[0]
    xp = np.concatenate([np.zeros((1, 2)), x[0], np.zeros((1, 2))])
    for m in range(6):
        np.testing.assert_allclose(y[2 * m], w[1] @ xp[m + 1] + w[3] @ xp[m] + b, atol=1e-12)
        np.testing.assert_allclose(y[2 * m + 1], w[0] @ xp[m + 2] + w[2] @ xp[m + 1] + b, atol=1e-12)


def test_stft_matches_direct_dft():
    rng = np.random.default_rng(3)
    x = rng.normal(size=(2, 700)).astype(np.float32)
    got = O.spectral(torch.tensor(x).double(), 512, 50, 240).numpy()
    np.testing.assert_allclose(got, Nv.stft_mag_naive(x, 512, 50, 240), atol=1e-9)
    assert got.shape == (2, 1 + (700 - 240) // 50, 257)


def test_stft_frame_counts_at_reference_window():
    """frames x bins at T=28160: [113,1025], [230,513], [559,257]."""
    x = torch.zeros(1, 28160)
    shapes = [tuple(O.spectral(x, n, h, w).shape[1:]) for n, h, w in zip(*O.STFT_ARGS)]
    assert shapes == [(113, 1025), (230, 513), (559, 257)]


def test_vq_indices_and_tie_break():
    rng = np.random.default_rng(5)
    x = rng.normal(size=(200, 8)).astype(np.float32)
    E = rng.normal(size=(8, 20)).astype(np.float32)
    E[:, 7] = E[:, 3]  # exact duplicate code: first minimum must win
    idx = O.vq_code_indices(torch.tensor(x), torch.tensor(E)).numpy()
    ref, top2 = Nv.vq_indices_naive(x, E)
    gap_ok = (top2[:, 1] - top2[:, 0]) > 1e-5 * top2[:, 0]
    dup = np.isin(ref, [3, 7])
    assert np.all(idx[gap_ok & ~dup] == ref[gap_ok & ~dup])
    assert not np.any(idx == 7)
    assert idx.dtype == np.int64


def test_vq_forward_values_and_gradient():
    rng = np.random.default_rng(6)
    x = torch.tensor(rng.normal(size=(2, 10, 4)).astype(np.float32), requires_grad=True)
    E = torch.tensor(rng.normal(size=(4, 6)).astype(np.float32))
    q_st, idx, commit, q = O.vq_forward(x, E, 0.25)
    assert torch.equal(q, E.t()[idx].reshape(x.shape))
    assert torch.equal(q_st, x + (q - x))
    up = torch.tensor(rng.normal(size=(2, 10, 4)).astype(np.float32))
    (g,) = torch.autograd.grad((q_st * up).sum() + commit, x)
    expect = up + 2 * 0.25 / x.numel() * (x.detach() - q)
    np.testing.assert_allclose(g.numpy(), expect.numpy(), rtol=1e-6, atol=1e-7)


def test_ema_knife_edge_and_restart():
    """A code hit exactly once on step 1 ends with N_t = fl(fl(.99*1)+fl(.01*1)) and stays alive iff that is >= 1."""
    K, D = 4, 2
    E = torch.arange(D * K, dtype=torch.float32).reshape(D, K) / 10
    st = O.VQState(E.clone(), E.clone(), torch.ones(K))
    flat = torch.tensor([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]])
    idx = torch.tensor([0, 0, 1])
    mb, nb = O.vq_batch_stats(flat, idx, K)
    rows = O.restart_rows_from_perm(flat, K, np.array([2, 0, 1, 3, 4, 5]))  # N=3 < K=4 -> tiled to 6 rows
    assert rows.shape == (K, D) and torch.equal(rows[3], flat[0])
    new, met = O.vq_ema_update(st, mb, nb, rows)
    g, om = np.float32(0.99), np.float32(1.0 - 0.99)
    N_expect = np.array([g * 1 + om * 2, g * 1 + om * 1, g, g], np.float32)
    np.testing.assert_array_equal(new.N_t.numpy(), N_expect)
    alive = N_expect >= 1.0
    assert list(alive) == [True, bool(np.float32(g + om) >= 1.0), False, False]
    for k in range(K):
        if alive[k]:
            np.testing.assert_array_equal(new.E[:, k].numpy(), (new.m_t[:, k] / new.N_t[k]).numpy())
        else:
            np.testing.assert_array_equal(new.E[:, k].numpy(), rows[k].numpy())
    assert float(met["batch_usage"]) == 2.0


def test_adam_matches_textbook():
    rng = np.random.default_rng(8)
    p = rng.normal(size=7); g = rng.normal(size=7)
    pt, m, v = [torch.tensor(p.copy())], [torch.zeros(7, dtype=torch.float64)], [torch.zeros(7, dtype=torch.float64)]
    pn, mn, vn = p.copy(), np.zeros(7), np.zeros(7)
    for t in (1, 2, 3):
        O.adam_step(pt, [torch.tensor(g)], m, v, t)
        pn, mn, vn = Nv.adam_naive(pn, g, mn, vn, t)
    np.testing.assert_allclose(pt[0].numpy(), pn, rtol=1e-12)


def test_model_structure_matches_survey():
    spec = O.ModelSpec(T=28160, **O.SMALL_VQ_VAE)
    counts = [sum(int(np.prod(s)) for op in spec.level_ops(l) for s in op.param_shapes()) for l in range(2)]
    tensors = [sum(len(op.param_shapes()) for op in spec.level_ops(l)) for l in range(2)]
    assert counts == [302337, 496705] and sum(tensors) == 484
    # MACs per input sample, forward: 115 280 including the VQ distance
    macs = 0.0
    for l in range(2):
        rate = 1.0
        for op in spec.level_ops(l):
            if op.kind == "conv":
                rate /= op.stride
                macs += rate * op.k * op.cin * op.cout
            elif op.kind == "convT":
                macs += rate * op.k * op.cin * op.cout  # k/stride real taps per output, stride outputs per input
                rate *= op.stride
            else:
                macs += rate * 2 * 3 * op.cin * op.cout
        hop = int(np.prod([s ** d for s, d in zip(spec.strides[: l + 1], spec.down_depth[: l + 1])]))
        macs += spec.num_embeddings * spec.latent_dim / hop
    assert round(macs) == 115280


def test_fp32_oracle_close_to_fp64_twin():
    spec, weights, vq, x = tiny_case()
    r32 = O.forward_losses(spec, weights, vq, torch.tensor(x), torch.float32)
    r64 = O.forward_losses(spec, weights, vq, torch.tensor(x), torch.float64)
    for a, b in zip(r32, r64):
        np.testing.assert_allclose(a["recon"].numpy(), b["recon"].numpy(), rtol=2e-4, atol=2e-5)
        assert abs(float(a["level_loss"]) - float(b["level_loss"])) < 1e-4 * abs(float(b["level_loss"]))


def test_autograd_matches_finite_differences():
    spec = O.ModelSpec(T=2048, levels=1, latent_dim=4, num_embeddings=8, down_depth=(2,), strides=(2,),
                       dilation_factor=3, residual_width=4, residual_depth=1)
    weights, vq = O.init_model(spec, 1, bias_scale=0.1)
    x = torch.tensor(np.random.default_rng(2).uniform(0, 1, size=(1, 2048, 1)))
    res, grads = O.loss_and_grads(spec, weights, vq, x, torch.float64)

    def total(ws):
        r = O.forward_losses(spec, [ws], vq, x, torch.float64)[0]
        return float(r["level_loss"]), r["idx"]

    base_idx = res[0]["idx"]
    checked = 0
    # decoder parameters only: through the encoder the straight-through estimator (VectorQuantizer.py:114) is by design
    # not the derivative of the piecewise-constant quantiser, so finite differences do not apply there
    ne = spec.n_enc_params(0)
    for pi in (ne, ne + 2, len(weights[0]) - 2):
        w = weights[0][pi].astype(np.float64)
        flat_i = int(np.argmax(np.abs(grads[0][pi].numpy()).ravel()))
        for eps in (1e-6,):
            wp, wm = [a.astype(np.float64) for a in weights[0]], [a.astype(np.float64) for a in weights[0]]
            wp[pi] = w.copy(); wp[pi].ravel()[flat_i] += eps
            wm[pi] = w.copy(); wm[pi].ravel()[flat_i] -= eps
            (lp, ip), (lm, im) = total(wp), total(wm)
            if not (torch.equal(ip, base_idx) and torch.equal(im, base_idx)):
                continue  # a code flipped: the loss is not differentiable there
            fd = (lp - lm) / (2 * eps)
            an = float(grads[0][pi].numpy().ravel()[flat_i])
            assert abs(fd - an) < 1e-4 * max(1.0, abs(an)), (pi, fd, an)
            checked += 1
    assert checked >= 2


def test_golden_fixtures_reproduce():
    """The committed vectors are what the oracle produces today (guards against silent oracle drift)."""
    g = np.load(os.path.join(GOLD, "primitives.npz"))
    for name in ("c_k3d1", "c_k3d9", "c_k4s2", "c_k4s2_odd", "c_k3_out1"):
        k, s, d = g[f"{name}.cfg"]
        y = O.conv1d(torch.tensor(g[f"{name}.x"]), torch.tensor(g[f"{name}.w"]), torch.tensor(g[f"{name}.b"]), int(s), int(d))
        np.testing.assert_allclose(y.numpy(), g[f"{name}.y"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(g[f"{name}.y"], Nv.conv1d_naive(g[f"{name}.x"], g[f"{name}.w"], g[f"{name}.b"], int(s), int(d)),
                                   rtol=1e-4, atol=1e-4)
    for name in ("t_k4s2", "t_k6s3"):
        k, s = g[f"{name}.cfg"]
        np.testing.assert_allclose(g[f"{name}.y"], Nv.conv1d_transpose_naive(g[f"{name}.x"], g[f"{name}.w"], g[f"{name}.b"], int(s)),
                                   rtol=1e-4, atol=1e-4)
    v = np.load(os.path.join(GOLD, "vq.npz"))
    ref, _ = Nv.vq_indices_naive(v["x"], v["E"])
    safe = (v["top2"][:, 1] - v["top2"][:, 0]) > 1e-5 * v["top2"][:, 0]
    assert np.all(v["idx"][safe] == ref[safe])
    spec, weights, vq, x = tiny_case()
    t = np.load(os.path.join(GOLD, "tiny_model.npz"))
    np.testing.assert_array_equal(t["x"], x)
    res = O.forward_losses(spec, weights, vq, torch.tensor(x))
    for l in range(spec.levels):
        np.testing.assert_allclose(res[l]["recon"].numpy(), t[f"recon{l}"], rtol=1e-4, atol=1e-5)


def test_relu_flip_bounds_cover_a_perturbed_block():
    """oracle.relu_flip_bounds: a residual block's w1 / b1 gradients are discontinuous where h crosses 0.  Shifting h by less than
    tol * max|h| (here: a tiny change of b1) may flip masks; the change of the gradients must stay inside the bound."""
    rng = np.random.default_rng(0)
    C, L, d = 8, 4000, 3
    x = torch.tensor(rng.normal(size=(1, L, C)).astype(np.float32))
    w1 = torch.tensor((rng.normal(size=(3, C, C)) / np.sqrt(3 * C)).astype(np.float32))
    w2 = torch.tensor((rng.normal(size=(3, C, C)) / np.sqrt(3 * C)).astype(np.float32))
    b2 = torch.zeros(C)
    dy = torch.tensor(rng.normal(size=(1, L, C)).astype(np.float32))
    tol = 1e-3

    def grads(b1v):
        p = [w1.clone().requires_grad_(True), b1v.clone().requires_grad_(True), w2.clone().requires_grad_(True), b2.clone().requires_grad_(True)]
        O.TRACE = []
        y = O.resblock(x, *p, d)
        rec, O.TRACE = O.TRACE, None
        allg = torch.autograd.grad((y * dy).sum(), p + [rec[0]["ah"]])
        return p, rec, allg

    b1 = torch.tensor((0.1 * rng.normal(size=C)).astype(np.float32))
    p, rec, g0 = grads(b1)
    bounds = O.relu_flip_bounds(rec, [g0[4]], 4, p, tol)
    assert bounds[2] == 0.0 and bounds[3] == 0.0 and bounds[0] > 0 and bounds[1] > 0
    hmax = float(rec[0]["h"].abs().max())
    flips_seen = 0
    for sign in (+1, -1):
        p2, rec2, g1 = grads(b1 + sign * 0.5 * tol * hmax)
        flips_seen += int(((rec2[0]["h"] > 0) != (rec[0]["h"] > 0)).sum())
        assert float((g1[1] - g0[1]).abs().max()) <= bounds[1] * 1.01 + 1e-5     # b1
        assert float((g1[0] - g0[0]).abs().max()) <= bounds[0] * 1.01 + 1e-4     # w1
    assert flips_seen > 0  # the perturbation really crossed zeros
