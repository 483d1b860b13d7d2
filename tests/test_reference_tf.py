"""Pins the oracle against the REAL reference (TensorFlow + /root/reference sources) wherever TensorFlow exists.

Here (no TensorFlow, Python 3.12, no network) every test in this file is collected and SKIPPED; DESIGN.md section 2 says
"parity unpinned" for exactly that reason.  On a box with TensorFlow >= 2.4 and the reference sources:
  * test_oracle_matches_live_reference runs vqvae.VQVAE itself (oracle/reference_tf.py) on the oracle's tiny synthetic
    case and compares reconstructions, the three losses, code indices, every gradient and one EMA step;
  * test_oracle_matches_reference_fixture does the same against tests/golden/reference_tf.npz once
    `python -m oracle.make_reference_fixtures` has been run on such a box and the file committed (it needs no TensorFlow).
Tolerances are the north star's: 1e-3 relative for reconstructions / losses / gradients, identical indices except where the
fp64 top-2 distance gap is below 1e-5 relative."""
import os

import numpy as np
import pytest
import torch

from oracle import reference_tf as RT
from oracle import vqvae_oracle as O
from oracle.make_golden import tiny_case

FIXTURE = os.path.join(os.path.dirname(__file__), "golden", "reference_tf.npz")
REL = 1e-3


def _rel(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-12))


def _check_against(ref):
    """oracle (fp32) vs a dict of reference outputs (live or fixture)"""
    spec, weights, vq, x = tiny_case()
    res, grads = O.loss_and_grads(spec, weights, vq, torch.tensor(x))
    for l in range(spec.levels):
        assert _rel(res[l]["recon"].numpy(), ref[f"recon{l}"]) < REL
        got = [float(res[l][k]) for k in ("recon_loss", "commit_loss", "spec_loss")]
        np.testing.assert_allclose(got, ref[f"losses{l}"], rtol=REL)
        # indices: equal except where the fp64 gap between best and second-best distance is below 1e-5 relative
        idx_o, idx_r = res[l]["idx"].numpy(), np.asarray(ref[f"idx{l}"]).reshape(-1)
        z64 = torch.tensor(np.asarray(ref[f"z{l}"], np.float64)).reshape(-1, spec.latent_dim)
        d = O.vq_distances(z64, torch.tensor(vq[l]["E"], dtype=torch.float64))
        top2 = torch.sort(d, dim=1).values[:, :2].numpy()
        scale = (z64 ** 2).sum(1).numpy() + float((torch.tensor(vq[l]["E"]) ** 2).sum(0).max())
        tie = (top2[:, 1] - top2[:, 0]) < 1e-5 * scale
        assert np.all((idx_o == idx_r) | tie), f"level {l}: {int(((idx_o != idx_r) & ~tie).sum())} index mismatches"
        for i, g in enumerate(grads[l]):
            want = ref[f"g{l}_{i:03d}"]
            assert _rel(g.numpy(), want) < REL or np.abs(want).max() < 1e-7, (l, i)
        # one EMA step (VectorQuantizer.py:118-159) on the reference's own encoder output
        z = torch.tensor(np.asarray(ref[f"z{l}"], np.float32)).reshape(-1, spec.latent_dim)
        E = torch.tensor(vq[l]["E"])
        idx = O.vq_code_indices(z, E)
        mb, nb = O.vq_batch_stats(z, idx, spec.num_embeddings)
        st = O.VQState(E.clone(), torch.tensor(vq[l]["m_t"]).clone(), torch.tensor(vq[l]["N_t"]).clone())
        rows = O.restart_rows_from_perm(z, spec.num_embeddings, np.arange(z.shape[0]))
        new, _ = O.vq_ema_update(st, mb, nb, rows, spec.gamma, spec.threshold)
        np.testing.assert_allclose(new.N_t.numpy(), ref[f"ema{l}_N_t"], rtol=1e-6)
        np.testing.assert_allclose(new.m_t.numpy(), ref[f"ema{l}_m_t"], rtol=1e-5, atol=1e-7)
        alive = np.asarray(ref[f"ema{l}_alive"]).astype(bool)  # restarted codes come from tf.random.shuffle: not comparable
        assert np.array_equal(alive, (new.N_t.numpy() >= spec.threshold))
        np.testing.assert_allclose(new.E.numpy()[:, alive], np.asarray(ref[f"ema{l}_E"])[:, alive], rtol=1e-5, atol=1e-7)


def test_oracle_matches_live_reference():
    ok, why = RT.available()
    if not ok:
        pytest.skip(why)
    spec, weights, vq, x = tiny_case()
    _check_against(RT.run_case(spec, weights, vq, x))


def test_oracle_matches_reference_fixture():
    if not os.path.exists(FIXTURE):
        pytest.skip("tests/golden/reference_tf.npz not generated yet (needs a TensorFlow box: python -m oracle.make_reference_fixtures)")
    _check_against(dict(np.load(FIXTURE)))


@pytest.mark.gpu
def test_cuda_path_matches_reference_fixture():
    """the CUDA path against the reference's own outputs (not the oracle's) once the fixture exists"""
    if not os.path.exists(FIXTURE):
        pytest.skip("tests/golden/reference_tf.npz not generated yet")
    import vqvae_b200 as V
    from tests.test_gpu_model import tiny_model
    ref = dict(np.load(FIXTURE))
    V._lib._BACKEND = None
    V._lib.lib()
    m, spec, weights, vq, x = tiny_model(V)
    recons, losses = m(x, training=False)
    for l in range(spec.levels):
        assert _rel(recons[l].cpu().numpy(), ref[f"recon{l}"]) < REL
        got = [float(losses[k][l]) for k in ("recon_losses", "commit_losses", "spec_losses")]
        np.testing.assert_allclose(got, ref[f"losses{l}"], rtol=REL)


def test_reference_loader_reports_why_it_is_unavailable():
    ok, why = RT.available()
    assert ok or why  # never silently "available"; bench.py --impl reference prints `why` next to kind="port"


def test_checker_plumbing_on_the_fp64_twin():
    """The comparison code above cannot run against TensorFlow here; exercise every line of it with the oracle's fp64 twin
    standing in for the reference (same dict layout as oracle.reference_tf.run_case)."""
    spec, weights, vq, x = tiny_case()
    res, grads = O.loss_and_grads(spec, weights, vq, torch.tensor(x), dtype=torch.float64)
    ref = {}
    for l in range(spec.levels):
        ref[f"recon{l}"] = res[l]["recon"].numpy()
        ref[f"losses{l}"] = np.array([float(res[l][k]) for k in ("recon_loss", "commit_loss", "spec_loss")])
        ref[f"idx{l}"] = res[l]["idx"].numpy()
        ref[f"z{l}"] = res[l]["z"].numpy()
        for i, g in enumerate(grads[l]):
            ref[f"g{l}_{i:03d}"] = g.numpy()
        z = res[l]["z"].float().reshape(-1, spec.latent_dim)
        E = torch.tensor(vq[l]["E"])
        idx = O.vq_code_indices(z, E)
        mb, nb = O.vq_batch_stats(z, idx, spec.num_embeddings)
        st = O.VQState(E.clone(), torch.tensor(vq[l]["m_t"]).clone(), torch.tensor(vq[l]["N_t"]).clone())
        new, _ = O.vq_ema_update(st, mb, nb, O.restart_rows_from_perm(z, spec.num_embeddings, np.arange(z.shape[0])[::-1].copy()),
                                 spec.gamma, spec.threshold)
        ref[f"ema{l}_m_t"], ref[f"ema{l}_N_t"], ref[f"ema{l}_E"] = new.m_t.numpy(), new.N_t.numpy(), new.E.numpy()
        ref[f"ema{l}_alive"] = new.N_t.numpy() >= spec.threshold
    _check_against(ref)
