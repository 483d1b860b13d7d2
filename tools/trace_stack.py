"""Pipeline time stamps of the fused residual-stack kernel (csrc/resstack_tc.cu, TRACE instantiation): clock64 at the hand-offs of
CTA 0's second tile — issuer (weights there, commit of every M block) and the first epilogue warp of every M block (accumulator
ready, all-epilogues barrier passed, row work done incl. operand rows, proxy fence done).
Usage: python tools/trace_stack.py   (prints clocks relative to the issuer's first stamp)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
buf = torch.zeros(5 * 9 * 8, dtype=torch.int64, device="cuda")
os.environ["VQB_RS_TRACE"] = hex(buf.data_ptr())
import vqvae_b200 as V  # noqa: E402

ops, P = V.ops, V._lib.PRECISIONS["fp16x2"]
g = torch.Generator(device="cuda").manual_seed(0)
B, L, C = 32, 14080, 32
x = torch.randn(B, L, C, device="cuda", generator=g)
dils = (1, 3, 9, 27)
W1 = [torch.randn(3, C, C, device="cuda", generator=g) * 0.1 for _ in dils]
W2 = [torch.randn(3, C, C, device="cuda", generator=g) * 0.1 for _ in dils]
Bz = [torch.zeros(C, device="cuda") for _ in dils]
for _ in range(3):
    ops.resstack_fwd(x, W1, Bz, W2, Bz, dils, P, False)
torch.cuda.synchronize()
t = buf.cpu().view(5, 9, 8)
t0 = int(t[0, 0, 0])
names = {0: ["w_full", "commit0", "commit1", "commit2", "commit3", "-", "-", "-"],
         1: ["acc_ready", "allepi", "rows_done", "fenced", "-", "-", "-", "-"]}
for k in range(8):
    print(f"conv {k}")
    for role in range(5):
        row = [int(v) - t0 if int(v) else None for v in t[role, k]]
        nm = names[0 if role == 0 else 1]
        print("   ", "issuer " if role == 0 else f"epi mb{role - 1}", " ".join(f"{n}={v}" for n, v in zip(nm, row) if v is not None))
