"""One eager SMALL_VQ_VAE train_step (batch 32) bracketed by cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum ...   (per-launch time list)
Usage: python tools/profile_step.py [precision]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vqvae_b200 as V  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
V.set_seed(0)
m = V.VQVAE((28160, 1), **V.SMALL_VQ_VAE)
m.use_cuda_graph = False
m.set_precision(prec)
m.compile(optimizer=V.keras.optimizers.Adam())
x = torch.from_numpy(np.random.default_rng(0).uniform(0, 1, size=(32, 28160, 1)).astype(np.float32)).cuda()
for _ in range(2):
    m.train_step((x, None))
torch.cuda.synchronize()
torch.cuda.profiler.start()
m.train_step((x, None))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
