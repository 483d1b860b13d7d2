"""Pipeline time stamps of the resident-codebook VQ search (csrc/vq_tc.cu: vq2_kernel<TRACE>): clock64 at the hand-offs of CTA 0's
tiles 3 and 4 — the issuing thread (x tile staged, accumulator drained, chunk committed) and scan warps 0 / 15 (accumulator
complete, chunk scanned, next tile staged, rows merged).  Usage: python tools/trace_vq.py   (clocks relative to the first stamp)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
buf = torch.zeros(48, dtype=torch.int64, device="cuda")
os.environ["VQB_VQ_TRACE"] = hex(buf.data_ptr())
import vqvae_b200 as V  # noqa: E402

ops = V.ops
N, D, K = 1 << 20, 64, 512
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(N, D, device="cuda", generator=g)
E = torch.randn(D, K, device="cuda", generator=g)
for _ in range(2):
    ops.vq_fwd(x, E, 0.25, True, False, None, None, V._lib.PRECISIONS["bf16"])
torch.cuda.synchronize()
t = buf.cpu().view(2, 24)
names = ["iss: x staged", "iss: acc0 drained", "iss: chunk0 committed", "iss: acc1 drained", "iss: chunk1 committed",
         "scan0: iteration start", "scan0: acc0 complete", "scan0: chunk0 scanned", "scan0: acc1 complete", "scan0: chunk1 scanned",
         "scan0: next tile staged", "scan0: rows merged", "iss: sees acc0 complete", "scan15: chunk0 scanned", "scan15: chunk1 scanned", "iss: sees acc1 complete", "scan0: iteration start (2nd stamp)", "scan0: bar acc0 passed", "scan0: bar acc1 passed", "iss: SPIN sees acc0 complete", "iss: SPIN sees acc1 complete", "scan0: partials merged", "scan0: index stored", "scan0: iteration end"]
t0 = min(int(v) for v in t.flatten() if int(v))
ev = sorted((int(t[i, e]) - t0, f"tile {3 + i}", names[e]) for i in range(2) for e in range(24) if int(t[i, e]))
for c, ti, n in ev:
    print(f"{c:8d}  {ti}  {n}")
