"""Pipeline time stamps of the residual-block weight-gradient kernel (csrc/wgrad_tc.cu, TRACE instantiation): clock64 at the
hand-offs of CTA 0's tiles 6..9 — the issuing thread (wait for the operand buffer, MMAs issued) and converter warp 0 (loads of
the next tile issued, buffer free, rows converted, fenced + arrived).  Usage: python tools/trace_wgrad.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
buf = torch.zeros(32, dtype=torch.int64, device="cuda")
os.environ["VQB_WG_TRACE"] = hex(buf.data_ptr())
import vqvae_b200 as V  # noqa: E402

ops, P = V.ops, V._lib.PRECISIONS["fp16x2"]
g = torch.Generator(device="cuda").manual_seed(0)
B, L, C = 32, 14080, 32
xs = [torch.randn(B, L, C, device="cuda", generator=g) for _ in range(4)]
dw, db, dw2, db2 = ops.empty(3, C, C), ops.empty(C), ops.empty(3, C, C), ops.empty(C)
for _ in range(2):
    ops.reduce_begin()
    for j, dl in enumerate((1, 3, 9, 27)):
        ops.resblock_wgrad(xs[j % 4], xs[(j + 1) % 4], xs[(j + 2) % 4], xs[(j + 3) % 4], dw, db, dw2, db2, dl, P)
    ops.reduce_flush()
torch.cuda.synchronize()
t = buf.cpu().view(4, 8)
names = ["iss: waits for the buffer", "iss: buffer full", "iss: MMAs issued", "conv0: waits for the buffer", "conv0: buffer free",
         "conv0: rows converted", "conv0: fenced + arrived", "conv0: next loads issued"]
t0 = min(int(v) for v in t.flatten() if int(v))
for c, ti, n in sorted((int(t[i, e]) - t0, f"tile {6 + i}", names[e]) for i in range(4) for e in range(8) if int(t[i, e])):
    print(f"{c:8d}  {ti}  {n}")
