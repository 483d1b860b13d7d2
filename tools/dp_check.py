"""Data-parallel check on real GPUs (run under torchrun, one rank per GPU, NCCL):
batch-sharded SMALL_VQ_VAE training (ONE flat all-reduce of gradients + EMA statistics + restart rows + loss scalars per
step, CUDA-graph path) must (1) leave every rank with bit-identical weights / codebooks and (2) match the un-sharded
run of the same global batch on one GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vqvae_b200 as V  # noqa: E402


def build(seed=0):
    V.keras_compat.reset_name_counters()
    V.set_seed(seed)
    m = V.VQVAE((28160, 1), **V.SMALL_VQ_VAE)
    m.compile(optimizer=V.keras.optimizers.Adam())
    return m


def main():
    V.dist.init_from_env("nccl")
    rank, world = V.dist.rank(), V.dist.world_size()
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    per = 4
    rng = np.random.Generator(np.random.PCG64(7))
    x = rng.uniform(0, 1, size=(per * world, 28160, 1)).astype(np.float32)
    m = build()
    steps = 4
    for _ in range(steps):
        logs = m.train_step((x[rank * per:(rank + 1) * per], None))
    torch.cuda.synchronize()
    flat = torch.cat([m._packed.params] + [t.value.reshape(-1) for vq in m.vqs for t in (vq.embeddings, vq.m_t, vq.N_t)])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    td.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    loss = float(logs["loss"])
    if rank == 0:
        # un-sharded reference on this GPU (no process group involvement: force world size 1)
        orig_ws, orig_rank = V.dist.world_size, V.dist.rank
        V.dist.world_size, V.dist.rank = (lambda: 1), (lambda: 0)
        V.vqvae.vdist.world_size, V.vqvae.vdist.rank = V.dist.world_size, V.dist.rank
        ref = build()
        for _ in range(steps):
            rlogs = ref.train_step((x, None))
        torch.cuda.synchronize()
        V.dist.world_size, V.dist.rank = orig_ws, orig_rank
        rflat = torch.cat([ref._packed.params] + [t.value.reshape(-1) for vq in ref.vqs for t in (vq.embeddings, vq.m_t, vq.N_t)])
        n = m._packed.params.numel()
        upd = (flat[:n] - build()._packed.params).abs().max()
        werr = float((flat[:n] - rflat[:n]).abs().max())
        verr = float((flat[n:] - rflat[n:]).abs().max() / rflat[n:].abs().max())
        print(f"ranks identical: {same}; loss dp={loss:.6f} single={float(rlogs['loss']):.6f}; "
              f"max |w_dp - w_single| = {werr:.3e} (largest update {float(upd):.3e}); codebook/EMA state rel err = {verr:.3e}")
        ok = same and abs(loss - float(rlogs["loss"])) < 2e-2 * abs(float(rlogs["loss"])) and werr < 0.3 * float(upd) + 1e-6
        print("DP_CHECK", "OK" if ok else "FAIL")
    td.barrier()
    td.destroy_process_group()


if __name__ == "__main__":
    main()
