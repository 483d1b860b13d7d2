"""Runs one hot kernel a few times (for `ncu --set full -k regex:...`).  Usage: python tools/bench_kernel.py resblock_fwd[_masks]|resblock_bwd[_masks]|wgrad|resblock_wgrad|vq|stack_infer|stack_train|stack_bwd|stack_wgrad [precision]
stack_*: the fused DilatedResnet1D kernel (4 blocks, dilations 1,3,9,27; DILS=27,9,3,1 to reverse) — one launch per call."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vqvae_b200 as V  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "resblock_fwd"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
P = V._lib.PRECISIONS[prec]
ops = V.ops
g = torch.Generator(device="cuda").manual_seed(0)
B, L, C = int(os.environ.get("B", "32")), int(os.environ.get("L", "14080")), 32
xs = [torch.randn(B, L, C, device="cuda", generator=g) for _ in range(2)]
dy = torch.randn(B, L, C, device="cuda", generator=g)
w1 = torch.randn(3, C, C, device="cuda", generator=g) * 0.1
w2 = torch.randn(3, C, C, device="cuda", generator=g) * 0.1
b1 = torch.zeros(C, device="cuda"); b2 = torch.zeros(C, device="cuda")
y, h = ops.resblock_fwd(xs[0], w1, b1, w2, b2, 1, P)
dw, db = ops.empty(3, C, C), ops.empty(C)
dw2, db2 = ops.empty(3, C, C), ops.empty(C)
if what.endswith("_masks"):
    _, _, xb, hb = ops.resblock_fwd_masks(xs[0], w1, b1, w2, b2, 1, P)
if what.startswith("stack"):
    dils = tuple(int(v) for v in os.environ.get("DILS", "1,3,9,27").split(","))
    W1 = [torch.randn(3, C, C, device="cuda", generator=g) * 0.1 for _ in dils]
    W2 = [torch.randn(3, C, C, device="cuda", generator=g) * 0.1 for _ in dils]
    Bz = [torch.zeros(C, device="cuda") for _ in dils]
    _, _, sxb, shb, sws = ops.resstack_fwd(xs[0], W1, Bz, W2, Bz, dils, P, True)
w4 = torch.randn(4, C, C, device="cuda", generator=g) * 0.1
if what == "conv_in1_wgrad":
    xa = torch.randn(B, 2 * L, 1, device="cuda", generator=g)
    dw1c = ops.empty(4, 1, C)
if what.startswith("conv3"):
    w3u, b3u = torch.randn(3, C, 64, device="cuda", generator=g) * 0.1, torch.zeros(64, device="cuda")
    w3d = torch.randn(3, 64, C, device="cuda", generator=g) * 0.1
    x64 = torch.randn(B, L, 64, device="cuda", generator=g)
dw4 = ops.empty(4, C, C)
dyh = dy[:, :L // 2].contiguous()
n = int(os.environ.get("N", "6"))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def one(i):
    if what == "resblock_fwd":
        ops.resblock_fwd(xs[i % 2], w1, b1, w2, b2, 1, P)
    elif what == "resblock_fwd_masks":
        ops.resblock_fwd_masks(xs[i % 2], w1, b1, w2, b2, 1, P)
    elif what == "resblock_bwd_masks":
        ops.resblock_bwd_data_masks(xb, hb, dy, w1, w2, int(os.environ.get("DIL", "1")), P)
    elif what == "stack_infer":
        ops.resstack_fwd(xs[i % 2], W1, Bz, W2, Bz, dils, P, False)
    elif what == "stack_train":
        ops.resstack_fwd(xs[i % 2], W1, Bz, W2, Bz, dils, P, True)
    elif what == "stack_bwd":
        ops.resstack_bwd_data(dy, W1, W2, sxb, shb, dils, P, fwd_ws=sws if os.environ.get("PACKED", "1") == "1" else None)
    elif what in ("conv3_up", "conv3_down"):   # latent-rate Conv1D(64, 3, 1) on 32 channels / Conv1D(32, 3, 1) on 64 (conv3_tc_kernel)
        if what == "conv3_up":
            ops.conv1d_fwd(xs[i % 2], w3u, b3u, 1, 1, False, None, P)
        else:
            ops.conv1d_fwd(x64, w3d, b1, 1, 1, False, None, P)
    elif what == "conv_in1_wgrad":   # weight gradient of the C_in = 1 first convolution (exact fp32)
        ops.conv1d_wgrad(xa, dy, dw1c, db, 2, 1, False, 0)
    elif what == "conv_down":      # Conv1D(32, 4, strides=2) 32 -> 32 (conv_tc_kernel): the encoder's down-sampling convolution
        ops.conv1d_fwd(xs[i % 2], w4, b1, 2, 1, False, None, P)
    elif what == "conv_down_wgrad":
        ops.conv1d_wgrad(xs[i % 2], dy[:, :L // 2].contiguous() if False else dyh, dw4, db, 2, 1, False, P)
    elif what == "resblock_bwd":
        ops.resblock_bwd_data(xs[i % 2], h, dy, w1, w2, int(os.environ.get("DIL", "1")), P)
    elif what == "resblock_wgrad":
        ops.resblock_wgrad(xs[i % 2], h, dy, xs[(i + 1) % 2], dw, db, dw2, db2, int(os.environ.get("DIL", "1")), P)
    elif what == "stack_wgrad":    # the weight gradients of a whole DilatedResnet1D: 4 blocks = 8 problems, one launch
        ops.reduce_begin()
        for j, dl in enumerate((1, 3, 9, 27)):
            ops.resblock_wgrad(xs[(i + j) % 2], h, dy, xs[(i + j + 1) % 2], dw, db, dw2, db2, dl, P)
        ops.reduce_flush()
    elif what == "wgrad":
        ops.conv1d_wgrad(xs[i % 2], dy, dw, db, 1, 1, True, P)
    elif what == "vq":
        x = xs[0].view(-1, 64)
        E = torch.randn(64, 512, device="cuda", generator=g)
        ops.vq_fwd(x, E, 0.25, True, True, ops.empty(64, 512), ops.empty(512), P if prec in ("bf16", "tf32") else 0)


for i in range(4):
    one(i)
torch.cuda.synchronize()
if os.environ.get("GRAPH", "0") == "1":  # device time without the host side of the calls: n calls captured, replayed
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(n):
            one(i)
    gr.replay()
    torch.cuda.synchronize()
    e0.record()
    gr.replay()
    e1.record()
else:
    e0.record()
    for i in range(n):
        one(i)
    e1.record()
torch.cuda.synchronize()
print(what, prec, f"B={B} L={L}", "ms per call", e0.elapsed_time(e1) / n)
