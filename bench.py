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
    ap.add_argument("--workload", default="train", choices=["train", "infer"],
                    help="train (default): BASELINE.json configs[1] / [2]; infer: configs[4] (8 windows x 2^20 samples, encode -> "
                         "quantize -> decode, windows spread over the GPUs) as the headline line")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): --batch windows per GPU; strong: the global batch stays --batch windows, split over the GPUs")
    ap.add_argument("--no-extra", action="store_true", help="skip the `infer` and `strong` sections of the default line")
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

    def make_model():
        V.set_seed(0)
        m = V.VQVAE((T_WINDOW, 1), **V.SMALL_VQ_VAE)
        if world > 1:  # identical initial weights / codebooks on every rank
            V.dist.broadcast(m._packed.params, 0)
            for vq in m.vqs:
                for v in (vq.embeddings, vq.m_t, vq.N_t):
                    V.dist.broadcast(v.value, 0)
        m.compile(optimizer=V.keras.optimizers.Adam())
        m.use_cuda_graph = not args.no_graph
        m.set_precision(args.precision)
        return m

    if args.workload == "infer":  # configs[4] as the headline
        with ClockSampler(local) as clk:
            sec = infer_section(V, world, rank, args.precision, iters=max(args.steps // 4, 3))["infer"]
        if rank == 0:
            config["workload"] = sec["workload"]
            config["parallelism"] = f"replicas x{world}"
            emit({"metric": sec["metric"], "value": sec["value"], "unit": sec["unit"], "n_gpus": world, "steps": max(args.steps // 4, 3),
                  "warmup": 2, "ms_per_step": sec["ms_per_pass"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                  "dtype": DTYPE[args.precision], "data": "synthetic", "config": config, "clocks": clk.summary()})
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    if args.scaling == "strong":
        if args.batch % world:
            raise SystemExit(f"--scaling strong: the global batch {args.batch} is not divisible by {world} GPUs")
        config["global_batch"] = args.batch
        args.batch //= world
        config["workload"] = config["workload"].replace("32 windows x", f"{args.batch} windows x") + " [strong scaling: global batch fixed]"
    model = make_model()
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
    extra = {}
    if not args.no_extra and args.scaling == "weak":
        del model  # its graphs hold ~5 GB of activations
        torch.cuda.empty_cache()
        if world > 1 and 32 % world == 0:   # the same job with the global batch fixed at 32 windows
            extra.update(strong_section(V, make_model, world, rank, 32, min(args.steps, 10)))
            torch.cuda.empty_cache()
        extra.update(infer_section(V, world, rank, args.precision))  # configs[4]
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    samples = args.batch * T_WINDOW * world
    pk = peaks()
    line = {"metric": METRIC, "value": samples / (ms * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": DTYPE[args.precision], "data": "synthetic", "config": config, **extra,
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
    """DRAM bytes (read + write) of one launch from the committed `ncu --set full` captures (profiles/r2_ncu.json, r1_ncu.json)."""
    for f in ("r2_ncu.json", "r1_ncu.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", f)))[name]["traffic_bytes"]
        except Exception:
            continue
    return None


def device_ms(fn, n=20, warm=3):
    """Average device time of one call of `fn(i)`: n calls captured into a CUDA graph and replayed between two CUDA events on
    the launching stream (the host side of the calls — ctypes, tensor-map encoding, allocations — is not in the number, as it
    is not in a captured train_step either).  Inputs alternate (i % 2) and exceed the L2 per call."""
    import torch
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def roofline_section(V, pk, precision="fp32"):
    """The dominant kernels timed ALONE (CUDA events on their launch stream, graph replays) at the largest stage of the model,
    [32, 14080, 32]: after round 2 the residual stacks run as ONE launch per DilatedResnet1D (vqb_resstack_fwd under a tape,
    vqb_resstack_bwd_data; 40 % of the step) next to the batched weight-gradient kernel (22 %).
    Algorithmic work of a stack (4 blocks = 8 k=3 32->32 convolutions): 49 152 FLOP and 256 B (read x, write y) per time
    position — with the activations on chip the stack is TENSOR-bound by SURVEY 8d, so `roofline` is quoted on the tensor
    roofline; the HBM view (algorithmic 256 B and the bytes the training variants really write: `design_bytes`) rides along.
    `roofline` = the lower of the two training kernels."""
    import torch
    ops = V.ops
    P = V._lib.PRECISIONS[precision]
    B, L, C = 32, 14080, 32
    dils = (1, 3, 9, 27)
    g = torch.Generator(device="cuda").manual_seed(0)
    xs = [torch.randn(B, L, C, device="cuda", generator=g) for _ in range(2)]
    dy = torch.randn(B, L, C, device="cuda", generator=g)
    W1 = [torch.randn(3, C, C, device="cuda", generator=g) * 0.1 for _ in dils]
    W2 = [torch.randn(3, C, C, device="cuda", generator=g) * 0.1 for _ in dils]
    Bz = [torch.zeros(C, device="cuda") for _ in dils]
    pos = B * L
    flop_stack = 2.0 * pos * (3 * C * C) * 2 * len(dils)
    kernels = []

    def entry(name, ms, flop, alg_b, design_b, traffic_key=None):
        tf, gbs = flop / (ms * 1e-3) / 1e12, alg_b / (ms * 1e-3) / 1e9
        e = {"kernel": name, "ms_per_launch": ms, "tflops": tf, "frac_tensor": tf / pk["tf_sust"], "gbs_algorithmic": gbs,
             "frac_hbm_algorithmic": gbs / pk["hbm"], "gbs_design": design_b / (ms * 1e-3) / 1e9,
             "frac_hbm_design": design_b / (ms * 1e-3) / 1e9 / pk["hbm"], "flop_per_launch": flop,
             "algorithmic_bytes_per_launch": alg_b, "design_bytes_per_launch": design_b,
             "traffic": ncu_traffic(traffic_key) if traffic_key else None}
        kernels.append(e)
        return e

    fused = bool(P) and ops.resstack_supported(C, dils, P)
    if fused:
        _, hs, xb, hb, fws = ops.resstack_fwd(xs[0], W1, Bz, W2, Bz, dils, P, True)
        n = len(dils)
        t_inf = device_ms(lambda i: ops.resstack_fwd(xs[i % 2], W1, Bz, W2, Bz, dils, P, False))
        t_trn = device_ms(lambda i: ops.resstack_fwd(xs[i % 2], W1, Bz, W2, Bz, dils, P, True))
        t_bwd = device_ms(lambda i: ops.resstack_bwd_data(dy, W1, W2, xb, hb, dils, P, fwd_ws=fws))  # as train_step runs it
        e_inf = entry("rs_kernel<0>: vqb_resstack_fwd, inference (4 blocks, dilations 1,3,9,27) [32,14080,32]", t_inf, flop_stack,
                      pos * 256.0, pos * 256.0, "rs_infer")
        e_trn = entry("rs_kernel<1>: vqb_resstack_fwd under a tape (stores h_i, y_i, sign masks)", t_trn, flop_stack, pos * 256.0,
                      pos * (128.0 + n * (256.0 + 8.0)), "rs_train")
        e_bwd = entry("rs_kernel<2>: vqb_resstack_bwd_data (stores dh_i, dx_i)", t_bwd, flop_stack, pos * 256.0,
                      pos * (128.0 + n * (256.0 + 8.0)), "rs_bwd")
        dom = min((e_trn, e_bwd), key=lambda e: e["frac_tensor"])
    # the per-block kernels (round 1's hot kernels; still the path of the other precisions)
    w1, w2, b0 = W1[0], W2[0], Bz[0]
    fwd = ops.resblock_fwd_masks if P else ops.resblock_fwd
    t_bf = device_ms(lambda i: fwd(xs[i % 2], w1, b0, w2, b0, 1, P))
    flop_blk = 2.0 * pos * (3 * C * C) * 2
    e_bf = entry("vqb_resblock_fwd" + ("_masks" if P else "") + " (one block, dilation 1)", t_bf, flop_blk, pos * 256.0,
                 pos * (384.0 + (8.0 if P else 0.0)), "rb_fwd_" + precision)
    if P:
        _, hh, xb1, hb1 = ops.resblock_fwd_masks(xs[0], w1, b0, w2, b0, 1, P)
        t_bb = device_ms(lambda i: ops.resblock_bwd_data_masks(xb1, hb1, dy, w1, w2, 1, P))
        entry("vqb_resblock_bwd_data_masks (one block, dilation 1)", t_bb, flop_blk, pos * 256.0, pos * 392.0, "rb_bwd_" + precision)
    if not fused:
        dom = min(kernels, key=lambda e: e["frac_hbm_algorithmic"])
        return {"roofline": {"kernel": dom["kernel"], "bound": "hbm", "achieved": dom["gbs_algorithmic"], "peak": pk["hbm"],
                             "unit": "GB/s", "frac": dom["frac_hbm_algorithmic"], "traffic": dom["traffic"],
                             "peak_source": pk["src"] + " HBM copy bandwidth", "ms_per_launch": dom["ms_per_launch"],
                             "algorithmic_bytes_per_unit": "256 B per time position (read x, write y)",
                             "design_bytes_per_launch": dom["design_bytes_per_launch"]},
                "roofline_kernels": kernels}
    return {"roofline": {"kernel": dom["kernel"], "bound": "tensor", "achieved": dom["tflops"], "peak": pk["tf_sust"],
                         "unit": "TFLOP/s", "frac": dom["frac_tensor"], "traffic": dom["traffic"],
                         "peak_source": pk["src"] + " bf16 matmul, sustained", "ms_per_launch": dom["ms_per_launch"],
                         "algorithmic_flop_per_unit": "49 152 FLOP per time position (8 convolutions x 2*3*32*32), fp32-equivalent: "
                                                      "the fp16x2 arithmetic issues 3 piece products per FLOP counted here",
                         "hbm": {"algorithmic_bytes_per_unit": "256 B per time position for the whole stack (read x, write y)",
                                 "achieved_algorithmic": dom["gbs_algorithmic"], "frac_algorithmic": dom["frac_hbm_algorithmic"],
                                 "design_bytes_per_launch": dom["design_bytes_per_launch"],
                                 "achieved_design": dom["gbs_design"], "frac_design": dom["frac_hbm_design"], "peak": pk["hbm"]}},
            "roofline_kernels": kernels}


def vq_section(V, pk):
    """BASELINE.json's second metric: VectorQuantizer latents/s (configs[3]: 2^20 latents x K in {512, 2048} x 64 dims, forward =
    indices + gather + straight-through output + commitment loss + batch statistics).  Roofline per SURVEY 8d:
    max(520 B [264 B with bf16 I/O] per latent / HBM, 2 K D FLOP / bf16 tensor peak)."""
    import torch
    ops = V.ops
    N, D = 1 << 20, 64
    g = torch.Generator(device="cuda").manual_seed(0)
    xs = [torch.randn(N, D, device="cuda", generator=g) for _ in range(2)]  # 2 x 268 MB alternate: not L2 resident
    P = V._lib.PRECISIONS["bf16"]  # tensor-core search with exact fp32 re-ranking: same indices as the fp32 search
    sweep = []
    for K in (512, 2048):
        E = torch.randn(D, K, device="cuda", generator=g)
        mb, nb = ops.empty(D, K), ops.empty(K)
        for io in ("fp32", "bf16"):
            if io == "bf16" and not getattr(ops, "VQ_BF16_IO", False):
                sweep.append({"K": K, "io": io, "unavailable": "this build of libvqvae_b200 has no vqb_vq_fwd_bf16"})
                continue
            xin = [x.to(torch.bfloat16) for x in xs] if io == "bf16" else xs   # a bfloat16 input selects vqb_vq_fwd_bf16
            ms = device_ms(lambda i: ops.vq_fwd(xin[i % 2], E, 0.25, True, False, mb, nb, P), n=10)
            per = 264 if io == "bf16" else 520
            t_hbm, t_mma = N * per / (pk["hbm"] * 1e9), N * 2.0 * K * D / (pk["tf_sust"] * 1e12)
            bound = "hbm" if t_hbm >= t_mma else "tensor"
            sweep.append({"K": K, "io": io, "ms": ms, "latents_per_s": N / (ms * 1e-3), "bound": bound,
                          "roofline_ms": max(t_hbm, t_mma) * 1e3, "frac": max(t_hbm, t_mma) * 1e3 / ms,
                          "algorithmic_bytes_per_unit": per, "flop_per_unit": 2 * K * D})
    head = sweep[0]
    return {"vq": {"metric": "VQ latents/sec", "workload": "VectorQuantizer forward, 2^20 latents x 512 codes x 64 dims, fp32 I/O, exact fp32 indices",
                   "value": head["latents_per_s"], "unit": "latents/s", "ms": head["ms"],
                   "roofline": {"bound": head["bound"], "achieved": N * 520 / (head["ms"] * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                "frac": head["frac"], "algorithmic_bytes_per_unit": 520}},
            "vq_sweep": sweep}


def infer_section(V, world, rank, precision, windows=8, log2_window=20, iters=3):
    """BASELINE.json configs[4]: SMALL_VQ_VAE encode -> quantize -> decode of both levels over `windows` windows of 2^20 samples.
    Windows are independent: rank r takes windows r, r + world, ... (replicas only, no collective on the data path; strong
    scaling: the job is the same 8 windows at every N).  Device time, max over ranks."""
    import numpy as np
    import torch
    T = 1 << log2_window
    V.set_seed(0)
    m = V.VQVAE((T, 1), **V.SMALL_VQ_VAE)
    m.set_precision(precision)
    mine = list(range(rank, windows, world))
    rng = np.random.Generator(np.random.PCG64(5))
    x = torch.from_numpy(rng.uniform(0, 1, size=(windows, T, 1)).astype(np.float32))[mine].cuda()

    def run():
        codes = m.encode(x)
        return codes, [m.decode(c, level=l) for l, c in enumerate(codes)]

    ms = 0.0
    if mine:
        for _ in range(2):
            run()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    if mine:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            codes, recons = run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t[0])
    del m, x
    torch.cuda.empty_cache()
    return {"infer": {"metric": "VQ-VAE encode->quantize->decode audio samples/sec (both levels)", "value": windows * T / (ms * 1e-3),
                      "unit": "samples/s", "ms_per_pass": ms, "n_gpus": world, "scaling": "strong",
                      "workload": f"{windows} windows x 2^{log2_window} samples (BASELINE.json configs[4]), encode + decode of levels 0 "
                                  f"and 1, {len(list(range(0, windows, world)))} window(s) per GPU, replicas only"}}


def strong_section(V, model_factory, world, rank, global_batch, steps):
    """SURVEY 8d C3, strong scaling: the global batch stays 32 windows, every rank trains on 32 / N of them (same model, one
    all-reduce per step).  Returns samples/s of the whole job (device time, max over ranks)."""
    import numpy as np
    import torch
    per = global_batch // world
    model = model_factory()
    rng = np.random.Generator(np.random.PCG64(2000 + rank))
    x = torch.from_numpy(rng.uniform(0, 1, size=(per, T_WINDOW, 1)).astype(np.float32)).cuda()
    for _ in range(5):
        model.train_step((x, None))
    torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        model.train_step((x, None))
    e1.record()
    torch.distributed.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t[0])
    return {"strong": {"metric": METRIC, "value": per * world * T_WINDOW / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms,
                       "scaling": "strong", "global_batch": per * world, "windows_per_gpu": per, "n_gpus": world}}


if __name__ == "__main__":
    main()
