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
    logs = m.train_step((x[rank * per:(rank + 1) * per], None))
    torch.cuda.synchronize()
    g1 = m._packed.grads.clone() / world                       # all-reduced (summed) gradient of step 1
    s1 = torch.cat([t.value.reshape(-1) for vq in m.vqs for t in (vq.embeddings, vq.m_t, vq.N_t)]).clone()
    l1 = float(logs["loss"])
    for _ in range(steps - 1):
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
        ref = build()
        rlogs = ref.train_step((x, None))
        torch.cuda.synchronize()
        rg1 = ref._packed.grads.clone()
        rs1 = torch.cat([t.value.reshape(-1) for vq in ref.vqs for t in (vq.embeddings, vq.m_t, vq.N_t)]).clone()
        rl1 = float(rlogs["loss"])
        for _ in range(steps - 1):
            rlogs = ref.train_step((x, None))
        torch.cuda.synchronize()
        V.dist.world_size, V.dist.rank = orig_ws, orig_rank
        gerr = float((g1 - rg1).abs().max() / rg1.abs().max())
        serr = float((s1 - rs1).abs().max() / rs1.abs().max())
        print(f"ranks identical after {steps} steps: {same}")
        print(f"step 1: loss dp={l1:.6f} single={rl1:.6f}; gradient rel err = {gerr:.3e}; codebook/EMA state rel err = {serr:.3e}")
        print(f"step {steps}: loss dp={loss:.6f} single={float(rlogs['loss']):.6f}")
        ok = same and abs(l1 - rl1) < 1e-4 * abs(rl1) and gerr < 1e-3 and serr < 1e-4 and \
            abs(loss - float(rlogs["loss"])) < 5e-2 * abs(float(rlogs["loss"]))
        print("DP_CHECK", "OK" if ok else "FAIL")
    td.barrier()
    td.destroy_process_group()


if __name__ == "__main__":
    main()
