import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vqvae_b200 as V
from oracle import vqvae_oracle as O
ops = V.ops
for prec in ("bf16", "tf32"):
    P = V._lib.PRECISIONS[prec]
    for (B, L, d) in ((1, 256, 1), (2, 1000, 3)):
        g = torch.Generator().manual_seed(1)
        x = torch.randn(B, L, 32, generator=g); dy = torch.randn(B, L, 32, generator=g)
        w = torch.zeros(3, 32, 32, requires_grad=True); bz = torch.zeros(32, requires_grad=True)
        y = O.conv1d(x, w, bz, 1, d)
        gw, gb = torch.autograd.grad(y, (w, bz), dy)
        dw, db = ops.zeros(3, 32, 32) + 7.0, ops.zeros(32) + 7.0
        ops.conv1d_wgrad(x.cuda(), dy.cuda(), dw, db, 1, d, False, P)
        torch.cuda.synchronize()
        print(prec, B, L, d, "dw[1,:2,:4]", dw[1, :2, :4].cpu().numpy(), "ref", gw[1, :2, :4].numpy())
        print("   db[:4]", db[:4].cpu().numpy(), "ref", gb[:4].numpy())
        print("   transposed? ", float((dw.cpu() - gw.transpose(1, 2)).abs().max()), "direct", float((dw.cpu() - gw).abs().max()), "max|gw|", float(gw.abs().max()))
