"""One VectorQuantizer forward (2^20 latents x K codes x 64 dims, tensor-core search) between cudaProfilerStart/Stop.
Usage: python tools/profile_vq.py [K]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vqvae_b200 as V  # noqa: E402

ops = V.ops
K = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N, D = 1 << 20, 64
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(N, D, device="cuda", generator=g)
E = torch.randn(D, K, device="cuda", generator=g)
mb, nb = ops.empty(D, K), ops.empty(K)
P = V._lib.PRECISIONS["bf16"]
for _ in range(2):
    ops.vq_fwd(x, E, 0.25, True, False, mb, nb, P)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.vq_fwd(x, E, 0.25, True, False, mb, nb, P)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
