"""Measures, on the GPU, how far each arithmetic mode (fp32, fp16x2, bf16x3, bf16x2, tf32, bf16) is from the fp32 oracle on the
SMALL_VQ_VAE forward + gradients at batch 2 (the quantities the north star bounds at 1e-3 relative in fp32).
Writes gpurun_out/precision_report.json.  Test infrastructure (uses the oracle)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vqvae_b200 as V  # noqa: E402
from oracle import vqvae_oracle as O  # noqa: E402


def main():
    spec = O.ModelSpec(T=28160, **O.SMALL_VQ_VAE)
    weights, vq = O.init_model(spec, 0, bias_scale=0.02)
    rng = np.random.Generator(np.random.PCG64(int(os.environ.get("SEED", "0"))))
    x = rng.uniform(0, 1, size=(2, 28160, 1)).astype(np.float32)
    res, grads = O.loss_and_grads(spec, weights, vq, torch.tensor(x))
    res64, grads64 = O.loss_and_grads(spec, weights, vq, torch.tensor(x), torch.float64)
    out = {"oracle_fp32_vs_fp64": {}}
    for l in range(2):
        gm = max(float(g.abs().max()) for g in grads64[l])
        out["oracle_fp32_vs_fp64"][f"level{l}"] = {
            "recon": float((res[l]["recon"].double() - res64[l]["recon"]).abs().max() / res64[l]["recon"].abs().max()),
            "grad_max_rel_to_largest": max(float((a.double() - b).abs().max()) for a, b in zip(grads[l], grads64[l])) / gm,
            "idx_mismatch": int((res[l]["idx"] != res64[l]["idx"]).sum())}
    for prec in (sys.argv[1:] or ("fp32", "fp16x2", "bf16x3", "bf16x2", "tf32", "bf16")):
        V.keras_compat.reset_name_counters(); V.set_seed(0)
        m = V.VQVAE((28160, 1), **V.SMALL_VQ_VAE)
        m.use_cuda_graph = False
        m.set_precision(prec)
        if os.environ.get("VQB_NO_FUSED") == "1":  # per-block kernels instead of the fused DilatedResnet1D launches
            from vqvae_b200.resnet import DilatedResnet1D
            for mm in m.vqvaes:
                for lay in mm._flatten_layers():
                    if isinstance(lay, DilatedResnet1D):
                        lay.use_fused_stack = False
        for l in range(2):
            for v, w in zip(m.vqvaes[l].trainable_variables, weights[l]):
                v.assign(w)
            m.vqs[l].embeddings.assign(vq[l]["E"]); m.vqs[l].m_t.assign(vq[l]["m_t"]); m.vqs[l].N_t.assign(vq[l]["N_t"])
        with V.GradientTape() as tape:
            total = V.keras.Scalar(); outs = []
            for l in range(2):
                rec, r, s, c = m._level_losses(l, V.keras.convert_to_tensor(x), False)
                outs.append((rec, r, c, s)); total += r + c + s
        g = tape.gradient(total, m.trainable_variables)
        rep, i = {}, 0
        for l in range(2):
            rec, r, c, s = outs[l]
            idx = m.encode(x)[l].reshape(-1).cpu()
            gm = max(float(t.abs().max()) for t in grads[l])
            gerr, gerr_own, worst = 0.0, 0.0, []
            names = [v.name for v in m.vqvaes[l].trainable_variables]
            gerr_own64 = 0.0
            for j, want in enumerate(grads[l]):
                e = float((g[i].cpu() - want).abs().max())
                gerr = max(gerr, e / gm)
                unit = max(float(want.abs().max()), 1e-3 * gm)
                own = e / unit
                gerr_own = max(gerr_own, own)
                e64 = float((g[i].cpu().double() - grads64[l][j]).abs().max())      # this mode against the fp64 twin ("truth")
                o64 = float((want.double() - grads64[l][j]).abs().max())            # the fp32 oracle against the fp64 twin
                gerr_own64 = max(gerr_own64, e64 / unit)
                worst.append((own, j, names[j], list(want.shape), float(want.abs().max()), e, e64 / unit, o64 / unit))
                i += 1
            worst.sort(reverse=True)
            rep[f"level{l}"] = {
                "recon_rel": float((rec.cpu() - res[l]["recon"]).abs().max() / res[l]["recon"].abs().max()),
                "recon_loss_rel": abs(float(r) - float(res[l]["recon_loss"])) / float(res[l]["recon_loss"]),
                "commit_loss_rel": abs(float(c) - float(res[l]["commit_loss"])) / float(res[l]["commit_loss"]),
                "spec_loss_rel": abs(float(s) - float(res[l]["spec_loss"])) / float(res[l]["spec_loss"]),
                "grad_err_rel_to_largest_grad": gerr, "grad_err_rel_to_own_max": gerr_own, "grad_err_rel_to_own_max_vs_fp64": gerr_own64,
                "idx_mismatch": int((idx != res[l]["idx"]).sum()), "n_idx": int(idx.numel()),
                "worst_tensors": [dict(rel=w[0], index=w[1], name=w[2], shape=w[3], own_max=w[4], abs_err=w[5], rel_vs_fp64=w[6],
                                       oracle_fp32_rel_vs_fp64=w[7]) for w in worst[:6]]}
        out[prec] = rep
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "precision_report.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
