#!/usr/bin/env python
"""bench.py — the reference's headline workload on B200: SMALL_VQ_VAE training step (forward + backward of both levels,
codebook EMA, Adam), batch 32 windows of 28160 samples per GPU, fp32 (BASELINE.json configs[1]; weak scaling under
torchrun).

  python bench.py [--gpus N] [--steps K] [--warmup W]          this repo's CUDA path
  python bench.py --impl reference [...]                       the CPU restatement of the reference (oracle port)

Prints ONE JSON line (rank 0).  `value` = audio samples/s of the whole job with the batch resident in HBM (CUDA-graph
replays, CUDA-event timing, max over ranks); `e2e` = the same metric through the public API (model.train_step on a
pinned HOST batch: H2D copy in, loss scalar D2H out, every step); `roofline` = the dominant kernel timed alone with
CUDA events; `cpu_baseline` = the oracle timed on the host cores on a bounded sample of the same workload."""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_WINDOW = 28160
FLOP_PER_SAMPLE_FWD_BWD = 691680.0  # SURVEY.md section 8d (both levels; conv + VQ distance, bwd = 2x fwd)
METRIC = "VQ-VAE fwd+bwd audio samples/sec"
DTYPE = {"fp32": "fp32", "bf16x3": "fp32", "fp16x2": "fp32", "bf16x2": "bf16x2", "bf16": "bf16", "tf32": "tf32"}
PRECISION_NOTE = {
    "fp32": "exact fp32 FMA on CUDA cores for every contraction",
    "bf16x3": "fp32-grade on tensor cores: each fp32 operand of the residual-block convolutions (208 of 242 convs) is split "
              "into 3 bf16 pieces (8+8+8 = 24 mantissa bits), all piece products on tcgen05 with fp32 TMEM accumulation; the "
              "remaining convolutions and the VQ re-ranking are exact fp32",
    "fp16x2": "fp32-grade on tensor cores: in the residual-block convolutions (208 of 242 convs) every fp32 operand is scaled by "
              "a power of two (per tile for activations, per convolution for weights) and split into 2 fp16 pieces (11+11 "
              "mantissa bits), 3 piece products on tcgen05 with fp32 TMEM accumulation; the strided convolutions and weight "
              "gradients use 3 bf16 pieces (24 bits); the remaining convolutions and the VQ re-ranking are exact fp32",
    "bf16x2": "2 bf16 pieces per operand (~2^-16 products)", "bf16": "bf16 operands, fp32 accumulate",
    "tf32": "tf32 operands, fp32 accumulate"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """Samples SM clock / throttle reasons every 100 ms during the timed region (pynvml; nvidia-smi semantics)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop, self._th = [], set(), None, threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, repr(e)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._th = threading.Thread(target=self._run, daemon=True)
            self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._th:
            self._th.join()

    def summary(self):
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_run(steps, warmup, batch):
    """The reference's arithmetic on the host cores: oracle train_step (fwd, bwd, Adam, EMA), all threads."""
    import numpy as np
    import torch
    from oracle import vqvae_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = O.ModelSpec(T=T_WINDOW, **O.SMALL_VQ_VAE)
    weights, vq = O.init_model(spec, 0)
    rng = np.random.Generator(np.random.PCG64(0))
    x = torch.tensor(rng.uniform(0, 1, size=(batch, T_WINDOW, 1)).astype(np.float32))
    tr = O.OracleTrainer(spec, weights, vq)
    for _ in range(warmup):
        tr.train_step(x)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.train_step(x)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dict(value=batch * T_WINDOW / dt, ms_per_step=1e3 * dt, cores=cores,
                sample=f"{steps} oracle train_step(s) (fwd+bwd+Adam+EMA, both levels) on {batch} windows of {T_WINDOW} "
                       f"samples, torch CPU fp32, {cores} threads")


_JSON_OUT = None


def emit(line):
    """The ONE JSON line of this run, on the real stdout."""
    print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)


def main():
    # stdout carries exactly one JSON line: file descriptor 1 is pointed at stderr for everything else (NCCL prints its
    # version banner to stdout when NCCL_DEBUG is set, libraries print warnings), the line itself goes to the saved descriptor
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="windows per GPU")
    ap.add_argument("--cpu-batch", type=int, default=0,
                    help="windows per step of the CPU arms; 0 = the same as --batch for --impl reference, 4 for the bounded "
                         "cpu_baseline sample inside the GPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run train_step eagerly (profiling)")
    ap.add_argument("--precision", default="fp16x2", choices=["fp32", "tf32", "bf16", "bf16x2", "bf16x3", "fp16x2"],
                    help="fp16x2 (default): fp32-grade two-piece fp16 split in the residual blocks (see PRECISION_NOTE); bf16x3: fp32 operands split into 3 bf16 pieces = all 24 mantissa bits, piece products on "
                         "tcgen05, fp32 accumulation (meets the fp32 parity contract, tests/test_gpu_model.py); fp32: exact "
                         "CUDA-core FMA path; bf16 / tf32 / bf16x2: reduced-precision tensor-core modes")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": f"SMALL_VQ_VAE train_step (fwd+bwd both levels, codebook EMA, Adam), {args.batch} windows x "
                          f"{T_WINDOW} samples per GPU, fp32 (BASELINE.json configs[1]; configs[2] when gpus > 1)",
              "levels": 2, "latent_dim": 64, "num_embeddings": 512, "window": T_WINDOW,
              "global_batch": args.batch * world, "parallelism": f"dp{world}",
              "l2_policy": "working set per step (~5 GB of activations) exceeds the 126 MB L2; no explicit flush"}

    if args.impl == "reference":
        if rank != 0:
            return
        # The reference's own CPU path on this box's host cores, on the SAME workload (args.batch windows per step; a step of
        # 32 windows takes ~2.5 s on 16 threads), bounded to 3 timed steps.  The real thing first: the unmodified reference
        # under TensorFlow (oracle/reference_tf.py) when TensorFlow is importable; otherwise the oracle port of it.
        steps, warm = min(args.steps, 3), min(args.warmup, 1)
        batch = args.cpu_batch if args.cpu_batch > 0 else args.batch
        from oracle import reference_tf as RT
        ok, why = RT.available()
        if ok:
            r, kind, note = RT.time_train_step(batch, steps, warm, T_WINDOW), "tensorflow", ""
        else:
            r, kind = cpu_reference_run(steps, warm, batch), "port"
            note = f" (oracle port of the reference's CPU path; the reference itself was not runnable: {why})"
        config["workload"] = (f"SMALL_VQ_VAE train_step (fwd+bwd both levels, codebook EMA, Adam), {batch} windows x {T_WINDOW} "
                              f"samples per step on the host CPU, fp32 (BASELINE.json configs[1])")
        config["global_batch"] = batch
        config["parallelism"] = "cpu"
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "samples/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": kind,
                                 "sample": r["sample"] + note},
                "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    import numpy as np
    import torch
    import vqvae_b200 as V
    V.dist.init_from_env("nccl")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    lib = V._lib.lib()
    assert V._lib.is_native()

    V.set_seed(0)
    model = V.VQVAE((T_WINDOW, 1), **V.SMALL_VQ_VAE)
    if world > 1:  # identical initial weights / codebooks on every rank
        V.dist.broadcast(model._packed.params, 0)
        for vq in model.vqs:
            for v in (vq.embeddings, vq.m_t, vq.N_t):
                V.dist.broadcast(v.value, 0)
    model.compile(optimizer=V.keras.optimizers.Adam())
    model.use_cuda_graph = not args.no_graph
    model.set_precision(args.precision)
    config["precision"] = args.precision
    config["precision_note"] = PRECISION_NOTE[args.precision]
    rng = np.random.Generator(np.random.PCG64(1000 + rank))
    x_host = torch.from_numpy(rng.uniform(0, 1, size=(args.batch, T_WINDOW, 1)).astype(np.float32)).pin_memory()
    x_dev = x_host.cuda(non_blocking=True)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # eager step (counts kernels) + graph capture + warm-up replays
    n0 = lib.vqb_kernel_launch_count()
    model.train_step((x_dev, None))
    torch.cuda.synchronize()
    launches_per_step = lib.vqb_kernel_launch_count() - n0
    for _ in range(max(args.warmup, 3)):
        logs = model.train_step((x_dev, None))
    barrier()

    # ---- device-resident timing
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            logs = model.train_step((x_dev, None))
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1) / args.steps
    # ---- end to end: pinned host batch in, loss scalar out, every step (its own warm-up: the host-side path differs)
    for _ in range(3):
        float(model.train_step((x_host, None))["loss"])
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss_host = 0.0
    for _ in range(args.steps):
        logs = model.train_step((x_host, None))
        loss_host = float(logs["loss"])  # D2H read of the step's result
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)) / args.steps
    t = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    samples = args.batch * T_WINDOW * world
    pk = peaks()
    line = {"metric": METRIC, "value": samples / (ms * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DTYPE[args.precision], "data": "synthetic", "config": config,
            "e2e": {"value": samples / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches_per_step * args.steps), "gpu_launches_per_step": int(launches_per_step),
            "clocks": clk.summary(), "final_loss": loss_host,
            "step_tflops": FLOP_PER_SAMPLE_FWD_BWD * samples / (ms * 1e-3) / 1e12,
            "step_frac_of_bf16_sustained": FLOP_PER_SAMPLE_FWD_BWD * samples / (ms * 1e-3) / 1e12 / (pk["tf_sust"] * world)}
    if not args.no_roofline:
        line.update(roofline_section(V, pk, args.precision))
        if world == 1:
            line.update(vq_section(V, pk))
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(1, 1, args.cpu_batch if args.cpu_batch > 0 else 4)
        line["cpu_baseline"] = {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
                                "sample": r["sample"]}
    emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()


def ncu_traffic(name):
    """DRAM bytes (read + write) of one launch from the committed `ncu --set full` capture (profiles/r1_ncu.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r1_ncu.json")))[name]["traffic_bytes"]
    except Exception:
        return None


def roofline_section(V, pk, precision="fp32"):
    """The dominant kernel timed ALONE with CUDA events on its launch stream: the fused residual-block forward at the
    largest stage of the model ([32, 14080, 32], dilation 1) — 208 of the model's 242 forward convolutions are inside
    such blocks, and the block kernels (forward + data gradient) are the largest share of the step (profiles/).
    After fusion the block is HBM-bound: 384 algorithmic bytes per time position (read x, write h for the backward pass,
    write y) against 12 288 FLOP, i.e. 26 us of HBM time vs 3.4 us of bf16 tensor time at this shape (20 us with the 6
    piece products of bf16x3).  Two input buffers are alternated; inputs + outputs per launch (173 MB) exceed the L2."""
    import torch
    ops = V.ops
    P = V._lib.PRECISIONS[precision]
    B, L, C, d = 32, 14080, 32, 1
    g = torch.Generator(device="cuda").manual_seed(0)
    xs = [torch.randn(B, L, C, device="cuda", generator=g) for _ in range(2)]
    w1 = torch.randn(3, C, C, device="cuda", generator=g) * 0.1
    w2 = torch.randn(3, C, C, device="cuda", generator=g) * 0.1
    b1 = torch.zeros(C, device="cuda"); b2 = torch.zeros(C, device="cuda")
    # what train_step launches under a tape: on the tensor-core paths the forward also writes the two sign-mask words per
    # position for the data gradient (+8 B on 384)
    fwd = ops.resblock_fwd_masks if P else ops.resblock_fwd
    for i in range(4):
        fwd(xs[i % 2], w1, b1, w2, b2, d, P)
    torch.cuda.synchronize()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fwd(xs[i % 2], w1, b1, w2, b2, d, P)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = 2.0 * B * L * (3 * C * C) * 2          # two k=3 C->C convolutions
    bytes_alg = B * L * (C * 4 * 3.0 + (8.0 if P else 0.0))  # read x, write h (kept for backward), write y (+ 2 mask words)
    tf = flops / (ms * 1e-3) / 1e12
    gbs = bytes_alg / (ms * 1e-3) / 1e9
    kname = "vqb_resblock_fwd [32,14080,32] dil 1 (" + ("fp32 path: 2 x tgc_kernel" if precision == "fp32" else "rb_tc_kernel, tcgen05 " + precision) + ")"
    return {"roofline": {"kernel": kname, "bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s",
                         "frac": gbs / pk["hbm"], "traffic": ncu_traffic("rb_fwd_" + precision),
                         "peak_source": pk["src"] + " HBM copy bandwidth", "ms_per_launch": ms,
                         "algorithmic_bytes_per_launch": bytes_alg,
                         "algorithmic_bytes_per_unit": "384 B per time position (SURVEY 8d / DESIGN 4) + 8 B of sign masks on the tensor-core paths"},
            "roofline_tensor": {"bound": "tensor", "achieved": tf, "peak": pk["tf_burst"], "unit": "TFLOP/s",
                                "frac": tf / pk["tf_burst"], "flop_per_launch": flops,
                                "note": "algorithmic FLOP of the two convolutions; the tensor pipe is not the limiter of this block"}}


def vq_section(V, pk):
    """BASELINE.json's second metric: VectorQuantizer latents/s (configs[3]: 2^20 latents x 512 codes x 64 dims, forward =
    indices + gather + straight-through output + commitment loss + batch statistics), against its HBM roofline
    (520 algorithmic bytes per latent: x 256 + q_st 256 + idx 8)."""
    import torch
    ops = V.ops
    N, D, K = 1 << 20, 64, 512
    g = torch.Generator(device="cuda").manual_seed(0)
    xs = [torch.randn(N, D, device="cuda", generator=g) for _ in range(2)]  # 2 x 268 MB alternate: not L2 resident
    E = torch.randn(D, K, device="cuda", generator=g)
    mb, nb = ops.empty(D, K), ops.empty(K)
    P = V._lib.PRECISIONS["bf16"]  # tensor-core search with exact fp32 re-ranking: same indices as the fp32 search
    for i in range(3):
        ops.vq_fwd(xs[i % 2], E, 0.25, True, False, mb, nb, P)
    torch.cuda.synchronize()
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        ops.vq_fwd(xs[i % 2], E, 0.25, True, False, mb, nb, P)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    t_hbm = N * 520 / (pk["hbm"] * 1e9)
    return {"vq": {"metric": "VQ latents/sec", "workload": "VectorQuantizer forward, 2^20 latents x 512 codes x 64 dims, fp32 I/O, exact fp32 indices",
                   "value": N / (ms * 1e-3), "unit": "latents/s", "ms": ms,
                   "roofline": {"bound": "hbm", "achieved": N * 520 / (ms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                "frac": t_hbm * 1e3 / ms, "algorithmic_bytes_per_unit": 520}}}


if __name__ == "__main__":
    main()
