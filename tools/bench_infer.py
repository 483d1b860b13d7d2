"""BASELINE.json configs[4]: SMALL_VQ_VAE encode -> quantize -> decode inference over long synthetic audio
(2^20 samples per window, 8 windows per job).  One process per GPU (torchrun), windows are independent: replicas only,
no collective on the data path; rank r takes windows r, r + world, ...  Prints one JSON line (rank 0): audio samples/s
of the whole job (CUDA events, max over ranks)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vqvae_b200 as V  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=8)
    ap.add_argument("--log2-window", type=int, default=20)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--precision", default="fp16x2")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        V.dist.init_from_env("nccl")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    T = 1 << args.log2_window
    V.set_seed(0)
    m = V.VQVAE((T, 1), **V.SMALL_VQ_VAE)
    m.set_precision(args.precision)
    mine = list(range(rank, args.windows, world))
    rng = np.random.Generator(np.random.PCG64(5))
    x = torch.from_numpy(rng.uniform(0, 1, size=(args.windows, T, 1)).astype(np.float32))[mine].cuda()

    def run():
        codes = m.encode(x)
        return codes, [m.decode(c, level=l) for l, c in enumerate(codes)]

    for _ in range(2):
        run()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        codes, recons = run()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.iters], device="cuda", dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"metric": "VQ-VAE encode->quantize->decode audio samples/sec (both levels)",
                          "value": args.windows * T / (float(ms) * 1e-3), "unit": "samples/s", "n_gpus": world,
                          "ms_per_pass": float(ms), "config": {"workload": f"{args.windows} windows x 2^{args.log2_window} samples, "
                          f"SMALL_VQ_VAE encode+decode of levels 0 and 1, precision {args.precision}", "parallelism": "replicas"},
                          "codes_shape": [list(c.shape) for c in codes], "recon_shape": [list(r.shape) for r in recons]}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
