"""BASELINE.json configs[3]: VectorQuantizer-only sweep — 2^20 latents x K in {512, 2048} x D = 64, forward (indices,
gather, straight-through output, commitment loss, batch statistics) and forward + EMA update.  Prints one JSON line per case:
latents/s and the fraction of the HBM roofline (algorithmic 520 B per latent: x 256 + q_st 256 + idx 8)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vqvae_b200 as V  # noqa: E402

ops = V.ops
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
N, D = 1 << 20, 64
g = torch.Generator(device="cuda").manual_seed(0)
xs = [torch.randn(N, D, device="cuda", generator=g) for _ in range(2)]   # 2 x 268 MB: alternate so inputs are not L2 resident
for K in (512, 2048):
    E = torch.randn(D, K, device="cuda", generator=g)
    m_t, N_t = E.clone(), torch.ones(K, device="cuda")
    mb, nb, rows, met = ops.empty(D, K), ops.empty(K), ops.zeros(K, D), ops.zeros(3)
    for prec, io in (("fp32", "fp32"), ("bf16", "fp32"), ("bf16", "bf16")):
        P = V._lib.PRECISIONS[prec]
        xin = [x.to(torch.bfloat16) for x in xs] if io == "bf16" else xs   # bfloat16 latents: vqb_vq_fwd_bf16
        for ema in (False, True):
            def one(i):
                ops.vq_fwd(xin[i % 2], E, 0.25, True, False, mb, nb, P)
                if ema:
                    ops.vq_ema_update(E.clone(), m_t, N_t, mb, nb, rows, 0.99, 1.0, met)
            for i in range(3):
                one(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 10
            e0.record()
            for i in range(n):
                one(i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            lat = N / (ms * 1e-3)
            t_hbm = N * (264 if io == "bf16" else 520) / (peaks["hbm_gbs"] * 1e9)
            t_mma = N * 2.0 * K * D / (peaks["bf16_tflops_sustained"] * 1e12)
            print(json.dumps({"workload": f"VQ-only N=2^20 K={K} D=64 {'fwd+EMA' if ema else 'fwd'}", "search": "tcgen05+exact-rerank" if prec != "fp32" else "fp32 CUDA cores", "io": io,
                              "ms": ms, "latents_per_s": lat, "roofline_ms": max(t_hbm, t_mma) * 1e3,
                              "bound": "hbm" if t_hbm > t_mma else "tensor", "frac_of_roofline": max(t_hbm, t_mma) * 1e3 / ms}))
