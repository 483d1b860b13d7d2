"""Generates the committed golden fixtures under tests/golden/ from the oracle (fp32, seeded).
Run from the repo root:  python -m oracle.make_golden
The reference itself cannot be executed here (TensorFlow 2.7 is not installable offline), so these vectors pin the
ORACLE's outputs — any later change of the restatement shows up as a golden mismatch — and give the GPU tests a
fixed, file-based target that does not depend on /root/reference or on torch's CPU kernels at test time."""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import vqvae_oracle as O

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TINY = dict(T=2048, levels=2, latent_dim=16, num_embeddings=32, down_depth=(3, 2), strides=(2, 2),
            dilation_factor=3, residual_width=8, residual_depth=2)


def tiny_case(seed=0, B=3):
    spec = O.ModelSpec(**TINY)
    weights, vq = O.init_model(spec, seed, bias_scale=0.1)
    rng = np.random.Generator(np.random.PCG64(seed + 100))
    x = rng.uniform(0, 1, size=(B, spec.T, 1)).astype(np.float32)
    return spec, weights, vq, x


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    # 1. primitive ops
    rng = np.random.Generator(np.random.PCG64(7))
    prim = {}
    for name, (L, cin, cout, k, s, d) in dict(c_k3d1=(50, 8, 8, 3, 1, 1), c_k3d9=(64, 8, 16, 3, 1, 9),
                                               c_k4s2=(64, 1, 8, 4, 2, 1), c_k4s2_odd=(63, 16, 8, 4, 2, 1),
                                               c_k3_out1=(40, 16, 1, 3, 1, 1)).items():
        x = rng.normal(size=(2, L, cin)).astype(np.float32)
        w = rng.normal(size=(k, cin, cout)).astype(np.float32) * 0.3
        b = rng.normal(size=(cout,)).astype(np.float32)
        y = O.conv1d(torch.tensor(x), torch.tensor(w), torch.tensor(b), s, d).numpy()
        prim.update({f"{name}.x": x, f"{name}.w": w, f"{name}.b": b, f"{name}.y": y,
                     f"{name}.cfg": np.array([k, s, d])})
    for name, (L, cin, cout, k, s) in dict(t_k4s2=(33, 8, 16, 4, 2), t_k6s3=(20, 8, 4, 6, 3)).items():
        x = rng.normal(size=(2, L, cin)).astype(np.float32)
        w = rng.normal(size=(k, cout, cin)).astype(np.float32) * 0.3
        b = rng.normal(size=(cout,)).astype(np.float32)
        y = O.conv1d_transpose(torch.tensor(x), torch.tensor(w), torch.tensor(b), s).numpy()
        prim.update({f"{name}.x": x, f"{name}.w": w, f"{name}.b": b, f"{name}.y": y, f"{name}.cfg": np.array([k, s])})
    np.savez_compressed(os.path.join(OUT, "primitives.npz"), **prim)

    # 2. VQ forward + EMA
    N, D, K = 300, 16, 40
    x = rng.normal(size=(N, D)).astype(np.float32)
    E = rng.normal(size=(D, K)).astype(np.float32)
    xt, Et = torch.tensor(x), torch.tensor(E)
    q_st, idx, commit, q = O.vq_forward(xt, Et, 0.25)
    d64 = O.vq_distances(xt.double(), Et.double())
    top2 = torch.sort(d64, dim=1).values[:, :2].numpy()
    mb, nb = O.vq_batch_stats(xt, idx, K)
    perm = rng.permutation(N)
    rows = O.restart_rows_from_perm(xt, K, perm)
    st = O.VQState(Et.clone(), Et.clone() * 0.5, torch.ones(K) * torch.tensor(rng.uniform(0.5, 3, K)).float())
    st_in = dict(E=st.E.numpy().copy(), m_t=st.m_t.numpy().copy(), N_t=st.N_t.numpy().copy())
    new, met = O.vq_ema_update(st, mb, nb, rows)
    np.savez_compressed(os.path.join(OUT, "vq.npz"), x=x, E=E, idx=idx.numpy(), q_st=q_st.numpy(), q=q.numpy(),
                        commit=commit.numpy(), top2=top2, m_batch=mb.numpy(), n_batch=nb.numpy(),
                        perm=perm, rows=rows.numpy(), m_t_in=st_in["m_t"], N_t_in=st_in["N_t"],
                        E_out=new.E.numpy(), m_t_out=new.m_t.numpy(), N_t_out=new.N_t.numpy(),
                        metrics=np.array([float(met["batch_usage"]), float(met["usage"]), float(met["entropy"])]))

    # 3. tiny two-level model: forward, losses, gradients, two training steps
    spec, weights, vq, x = tiny_case()
    res, grads = O.loss_and_grads(spec, weights, vq, torch.tensor(x))
    out = {"x": x}
    for l in range(spec.levels):
        for i, w in enumerate(weights[l]):
            out[f"w{l}_{i:03d}"] = w
        for i, g in enumerate(grads[l]):
            out[f"g{l}_{i:03d}"] = g.numpy()
        for k in ("E", "m_t", "N_t"):
            out[f"vq{l}_{k}"] = vq[l][k]
        out[f"recon{l}"] = res[l]["recon"].numpy()
        out[f"idx{l}"] = res[l]["idx"].numpy()
        out[f"losses{l}"] = np.array([float(res[l][k]) for k in ("recon_loss", "commit_loss", "spec_loss")])
    tr = O.OracleTrainer(spec, weights, vq)
    for _ in range(2):
        tr.train_step(torch.tensor(x))
    for l in range(spec.levels):
        for i, w in enumerate(tr.w[l]):
            out[f"w2_{l}_{i:03d}"] = w.numpy()
        out[f"vq2_{l}_E"] = tr.vq[l].E.numpy()
        out[f"vq2_{l}_N_t"] = tr.vq[l].N_t.numpy()
    np.savez_compressed(os.path.join(OUT, "tiny_model.npz"), **out)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
