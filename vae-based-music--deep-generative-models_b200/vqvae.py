"""vqvae — drop-in for the reference's `vqvae.py`: `get_vqvae` (:15-21) and `VQVAE` (:24-326) with the same
constructor, attributes (`vqvaes`, `encoders`, `decoders`, `vqs`, `levels`, ...) and methods (`train_step`,
`test_step`, `call`, `encode`, `decode`, `update_metrics`, `get_quantizer`, `_multispectral_loss`, `compile`, `fit`,
`evaluate`).  Every level is an independent VQ-VAE on the same input (Jukebox style).

What is different underneath:
  * all trainable variables of all levels live in ONE packed parameter buffer / ONE gradient buffer, so the data-
    parallel gradient all-reduce and the Adam step are single launches;
  * `train_step` is captured into CUDA graphs after its first eager execution for a given batch shape (the reference
    runs ~1500 eager TF kernels per step, vqvae.py:110 has @tf.function commented out) — graph A = forward + backward
    of every level, [NCCL all-reduce of gradients + EMA statistics + loss scalars under data parallelism],
    graph B = EMA codebook update + Adam;
  * `train_step_training` controls the `training` flag the VectorQuantizer sees inside train_step.  Strict Keras-2.7
    resolution of the reference's `self.vqvaes[level](x)` gives training=False (the functional model's `call` default is
    None -> False, and the nested layer inherits it), i.e. the reference as written never runs its EMA update during
    `fit`; the benchmark configurations ask for "training with codebook EMA", so the default here is True.  Set it to
    None for the literal Keras resolution.
"""
from __future__ import annotations

import contextlib
from itertools import chain

import numpy as np
import torch

from . import _lib, dist as vdist, ops
from .VectorQuantizer import VectorQuantizer
from .data_utils import STFT_ARGS, MultiSpectralLoss, norm, spectral  # noqa: F401
from .encdec import Decoder, Encoder, print_dec_layer  # noqa: F401
from .keras_compat import (GradientTape, Input, MeanSquaredError, Model, Packed, Scalar, convert_to_tensor, metrics,
                           on_stream, reduce_mean)


def get_vqvae(input_shape, encoder, decoder, vq, level=0):
    inputs = Input(shape=input_shape)
    encoder_outputs = encoder(inputs)
    quantized_latents, _ = vq(encoder_outputs)
    reconstructions = decoder(quantized_latents)
    return Model(inputs, reconstructions, name="vq_vae_{}".format(level))


class VQVAE(Model):
    """@levels: number of independent VQ-VAEs, bottom (finest) to top."""

    def __init__(self, input_shape, levels, latent_dim, down_depth, strides, num_embeddings=128, residual_width=64,
                 residual_depth=4, dilation_factor=1, train_variance=1.0, **kwargs):
        super(VQVAE, self).__init__(**kwargs)
        self.levels = levels
        self._arch = (list(down_depth), list(strides), residual_depth, dilation_factor)  # for the halo of time-tiled inference
        self.train_variance = train_variance
        self.latent_dim = latent_dim
        self.num_embeddings = num_embeddings
        self.vqs = [VectorQuantizer(num_embeddings, latent_dim, level=level, name="vector_quantizer_{}".format(level))
                    for level in range(levels)]
        self.encoders = [Encoder(output_dim=latent_dim, residual_width=residual_width, residual_depth=residual_depth,
                                 depth=level + 1, down_depth=down_depth[:level + 1], strides=strides[:level + 1],
                                 dilation_factor=dilation_factor, name="encoder_{}".format(level))
                         for level in range(levels)]
        self.decoders = [Decoder(output_dim=input_shape[-1], embed_width=latent_dim, residual_width=residual_width,
                                 residual_depth=residual_depth, depth=level + 1, down_depth=down_depth[:level + 1],
                                 strides=strides[:level + 1], dilation_factor=dilation_factor,
                                 name="decoder_{}".format(level)) for level in range(levels)]
        self.vqvaes = [get_vqvae(input_shape, self.encoders[level], self.decoders[level], self.vqs[level], level)
                       for level in range(levels)]

        self.total_loss_tracker = metrics.Mean(name="total_loss")
        self.reconstruction_loss_tracker = metrics.Mean(name="reconstruction_loss")
        self.vq_loss_tracker = metrics.Mean(name="vq_loss")
        self.spectral_loss_tracker = metrics.Mean(name="spectral_loss")
        self.level_loss_trackers = [metrics.Mean(name="[{}]level_loss".format(level)) for level in range(levels)]
        self.recon_loss_trackers = [metrics.Mean(name="[{}]recon_loss".format(level)) for level in range(levels)]
        self.vq_loss_trackers = [metrics.Mean(name="[{}]vq_loss".format(level)) for level in range(levels)]
        self.spectral_loss_trackers = [metrics.Mean(name="[{}]spectral_loss".format(level)) for level in range(levels)]
        self.loss_fn = MeanSquaredError(reduction="none")
        self.built = True

        # ---- implementation state --------------------------------------------------------------------------
        self.train_step_training = True   # see module docstring
        self.test_step_training = None    # what test_step passes as `training`: None = the literal call `self.vqvaes[level](x)`
                                          # (vqvae.py:158), resolved by keras_compat like Keras 2.7 (-> False: no EMA in evaluate);
                                          # True if your Keras lets the VectorQuantizer's own default win
        self.use_cuda_graph = _lib.is_native() or _lib._BACKEND is None
        self.use_level_streams = True     # levels >= 1 on side CUDA streams inside train_step (see _forward_backward)
        self._level_streams = []
        self.spectral_weight = 1.0        # parity tests may switch the torch-side loss head off (0.0)
        self._graphs = {}
        self._pack_variables()

    # ------------------------------------------------------------------------------------------------------
    def _pack_variables(self):
        tv = list(chain.from_iterable(m.trainable_variables for m in self.vqvaes))
        D, K = self.latent_dim, self.num_embeddings
        per_level = D * K + K + K * D
        n_scalars = 3 * self.levels
        self._packed = Packed(tv, extra_floats=self.levels * per_level + n_scalars)
        ex = self._packed.extra
        for l, vq in enumerate(self.vqs):
            o = l * per_level
            vq.bind_stats(ex[o:o + D * K].view(D, K), ex[o + D * K:o + D * K + K],
                          ex[o + D * K + K:o + per_level].view(K, D))
        self._comm_scalars = ex[self.levels * per_level:]
        # metric totals of every tracker in one vector (one add per step)
        trackers = self._all_trackers()
        self._metric_totals = ops.zeros(len(trackers))
        for i, m in enumerate(trackers):
            m._bind(self._metric_totals[i:i + 1])

    def set_precision(self, precision="fp32"):
        """Arithmetic of the contraction kernels: "fp32" (exact CUDA-core FMA), "tf32" or "bf16" (tcgen05 tensor cores,
        fp32 accumulate).  Layers whose shape has no tensor-core kernel keep fp32 (vqb_resblock_supports)."""
        from .keras_compat import Conv1D, Conv1DTranspose
        from .resnet import ResnetConv1DBlock
        code = _lib.PRECISIONS[precision]
        for m in self.vqvaes:
            for l in m._flatten_layers():
                if isinstance(l, (ResnetConv1DBlock, Conv1D, Conv1DTranspose)):
                    l.precision = code  # shapes without a tensor-core kernel fall back to fp32 per call (ops._pick)
        for vq in self.vqs:
            vq.precision = code  # tensor-core search + exact fp32 re-ranking (same indices as the fp32 search)
        self.precision = precision
        self._graphs = {}
        return self

    def compile(self, optimizer=None, **kw):
        """vqvae.py:362.  A captured train_step has the optimizer's slot / counter / learning-rate buffers baked in by address:
        a new optimizer invalidates the captures (the learning rate itself lives in device memory and may change freely)."""
        super(VQVAE, self).compile(optimizer=optimizer, **kw)
        self._graphs = {}

    def _extra_state(self):
        return [np.asarray(0 if vq._step is None else int(vq._step.item()), np.int64) for vq in self.vqs]

    def _set_extra_state(self, arrays):
        for vq, a in zip(self.vqs, arrays):
            vq._buffers()
            vq._step.fill_(int(a))

    def _all_trackers(self):
        return self.metrics + list(chain.from_iterable(vq.metrics for vq in self.vqs))

    @property
    def metrics(self):
        return [
            self.total_loss_tracker,
            self.reconstruction_loss_tracker,
            self.vq_loss_tracker,
            self.spectral_loss_tracker,
            *self.level_loss_trackers,
            *self.recon_loss_trackers,
            *self.vq_loss_trackers,
            *self.spectral_loss_trackers,
        ]

    def reset_metrics(self):
        for m in self._all_trackers():
            m.reset_state()

    @property
    def trainable_variables(self):
        return list(chain.from_iterable(m.trainable_variables for m in self.vqvaes))

    @property
    def variables(self):
        out = []
        for l in range(self.levels):
            out += self.vqvaes[l].variables
        return out

    # ------------------------------------------------------------------------------------------------------
    def _level_losses(self, level, x, training, stream=None):
        """One level of the loss computation shared by train_step / test_step / call (vqvae.py:121-131).  `stream`: side CUDA
        stream for the level's encoder -> VQ -> decoder (already forked by the caller); the loss head runs on the caller's."""
        with on_stream(stream):
            reconstructions = self.vqvaes[level](x) if training is None else self.vqvaes[level](x, training=training)
        if stream is not None:
            torch.cuda.current_stream().wait_stream(stream)
            reconstructions.record_stream(torch.cuda.current_stream())  # allocated on `stream`, read by the loss head here
        reconstruction_loss = reduce_mean(self.loss_fn(x, reconstructions))
        spectral_loss = reduce_mean(self._multispectral_loss(x, reconstructions)) if self.spectral_weight else Scalar()
        commit_loss = sum(self.vqvaes[level].losses)
        return reconstructions, reconstruction_loss, spectral_loss, commit_loss

    def _forward_backward(self, x):
        """The tape part of train_step (vqvae.py:113-143)."""
        commit_losses, recon_losses, spectral_losses, level_losses = [], [], [], []
        total_loss = Scalar()
        # The levels are independent models on the same batch: levels >= 1 run their encoder -> VQ -> decoder (forward and,
        # through the tape's stream tags, backward) on side streams, so that their many small-grid kernels (the deep stages:
        # 14 to 110 tiles for 148 SMs) fill in next to level 0's.  Forked here, before level 0 is enqueued.
        streams = [None] * self.levels
        if self.use_level_streams and self.levels > 1 and _lib.device().type == "cuda":
            if len(self._level_streams) < self.levels - 1:
                self._level_streams = [torch.cuda.Stream(device=_lib.device()) for _ in range(self.levels - 1)]
            for l in range(1, self.levels):
                streams[l] = self._level_streams[l - 1]
                streams[l].wait_stream(torch.cuda.current_stream())
        with GradientTape() as tape:
            for level in range(self.levels):  # bottom to top
                _, reconstruction_loss, spectral_loss, commit_loss = self._level_losses(
                    level, x, self.train_step_training, streams[level])
                level_loss = reconstruction_loss + commit_loss + spectral_loss
                commit_losses.append(commit_loss)
                recon_losses.append(reconstruction_loss)
                spectral_losses.append(spectral_loss)
                level_losses.append(level_loss)
                total_loss += level_loss
        trainable_vars = self.trainable_variables
        grads = tape.gradient(total_loss, trainable_vars)
        return grads, trainable_vars, (level_losses, recon_losses, commit_losses, spectral_losses)

    def train_step(self, data):
        raw = data[0] if isinstance(data, (tuple, list)) else data
        if self.optimizer is None:
            raise RuntimeError("VQVAE.train_step: call compile(optimizer=...) first")
        world = vdist.world_size()
        for vq in self.vqs:
            vq.defer_ema = world > 1
            vq.shard = (world, vdist.rank())
        self.optimizer.grad_scale = 1.0 / world
        try:
            if self.use_cuda_graph and _lib.device().type == "cuda":
                return self._graph_train_step(raw)
            x = convert_to_tensor(raw)
            grads, tvars, losses = self._forward_backward(x)
            losses = self._exchange(losses, world)
            for vq in self.vqs:
                vq.apply_ema()
            self.optimizer.apply_gradients(zip(grads, tvars))
            return self.update_metrics(*losses)
        finally:
            for vq in self.vqs:  # a later direct call (`model(x, training=True)`, `vq(x)`) applies its own EMA again
                vq.defer_ema = False

    @contextlib.contextmanager
    def _vq_metrics_in_step_vector(self):
        """Inside the graph path the VQ usage / entropy metrics travel in the step vector (`_step_vector`), so the layers must
        not also update their trackers one by one — for the duration of that step only."""
        for vq in self.vqs:
            vq.skip_metric_update = True
        try:
            yield
        finally:
            for vq in self.vqs:
                vq.skip_metric_update = False

    def _exchange(self, losses, world):
        """Data parallelism: ONE all-reduce of [gradients | EMA statistics + restart rows | loss scalars]."""
        if world == 1:
            return losses
        level_losses, recon_losses, commit_losses, spectral_losses = losses
        L = self.levels
        vec = torch.cat([s.tensor().reshape(1) for s in (*recon_losses, *commit_losses, *spectral_losses)])
        self._comm_scalars.copy_(vec)
        vdist.all_reduce_sum(self._packed.comm)
        sc = self._comm_scalars / world
        rec = [Scalar.leaf(sc[i:i + 1]) for i in range(L)]
        com = [Scalar.leaf(sc[L + i:L + i + 1]) for i in range(L)]
        spe = [Scalar.leaf(sc[2 * L + i:2 * L + i + 1]) for i in range(L)]
        lev = [rec[i] + com[i] + spe[i] for i in range(L)]
        return lev, rec, com, spe

    # ---- CUDA-graph path ------------------------------------------------------------------------------------
    def _step_vector(self, losses):
        """All per-step metric increments, in `_all_trackers()` order, as one device vector."""
        level_losses, recon_losses, commit_losses, spectral_losses = losses
        parts = [sum(level_losses), sum(recon_losses), sum(commit_losses), sum(spectral_losses),
                 *level_losses, *recon_losses, *commit_losses, *spectral_losses]
        outs = [list(p.terms) for p in parts]   # every entry is a linear combination of the loss kernels' scalars ...
        for vq in self.vqs:                     # ... or one of a VQ layer's three usage / entropy metrics
            buf = vq._metrics_buf if (vq._metrics_buf is not None and self.train_step_training) else None
            outs.extend([[(buf[i:i + 1], 1.0)] if buf is not None else [] for i in range(3)])
        return ops.lincomb(outs)                # one launch (vqb_lincomb) instead of a library kernel per `+` and a concatenation

    def _graph_train_step(self, raw):
        key = (tuple(raw.shape), vdist.world_size(), self.train_step_training, self.spectral_weight,
               getattr(self, "precision", "fp32"), self.use_level_streams, id(self.optimizer))
        st = self._graphs.get(key)
        world = vdist.world_size()
        if st is not None:
            # straight into the graph's static input buffer (pinned host memory -> one asynchronous H2D copy)
            src = raw if isinstance(raw, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(raw, dtype=np.float32))
            st["x"].copy_(src.reshape(st["x"].shape), non_blocking=True)
        if st is None:
            x = convert_to_tensor(raw)
            # first step for this shape: run it eagerly (it is a real training step and doubles as warm-up)
            self._graphs[key] = {"graph": None, "x": x.clone()}
            with self._vq_metrics_in_step_vector():
                grads, tvars, losses = self._forward_backward(x)
                losses = self._exchange(losses, world)
                for vq in self.vqs:
                    vq.apply_ema()
                self.optimizer.apply_gradients(zip(grads, tvars))
                vec = self._step_vector(losses)
            return self._accumulate(vec)
        if st["graph"] is None:
            with self._vq_metrics_in_step_vector():
                self._capture(st, world)
        self.optimizer.refresh_lr()  # the captured Adam kernel reads the learning rate from device memory
        if st.get("gone") is not None:   # single process: the whole step is one graph launch
            st["gone"].replay()
        else:
            st["ga"].replay()
            if world > 1:
                vdist.all_reduce_sum(self._packed.comm)
            st["gb"].replay()
        self.optimizer.step_done()
        return self._accumulate(st["vec"])

    def _tail(self, grads, tvars, losses, world):
        """what follows the exchange: reduced loss scalars, EMA codebook update, Adam, the metric vector"""
        if world > 1:
            L = self.levels
            sc = self._comm_scalars / world
            rec = [Scalar.leaf(sc[i:i + 1]) for i in range(L)]
            com = [Scalar.leaf(sc[L + i:L + i + 1]) for i in range(L)]
            spe = [Scalar.leaf(sc[2 * L + i:2 * L + i + 1]) for i in range(L)]
            losses = ([rec[i] + com[i] + spe[i] for i in range(L)], rec, com, spe)
        for vq in self.vqs:
            vq.apply_ema()
        self.optimizer.apply_gradients(zip(grads, tvars))
        return self._step_vector(losses)

    def _head(self, x, world):
        grads, tvars, losses = self._forward_backward(x)
        if world > 1:
            level_losses, recon_losses, commit_losses, spectral_losses = losses
            self._comm_scalars.copy_(torch.cat(
                [s.tensor().reshape(1) for s in (*recon_losses, *commit_losses, *spectral_losses)]))
        return grads, tvars, losses

    def _capture(self, st, world):
        """Captures train_step for st["x"].  One process: ONE graph (forward + backward of every level, EMA + Adam + metric
        vector).  Data parallel, default: graph A (forward + backward), torch.distributed.all_reduce of [gradients | EMA
        statistics | loss scalars] between the replays, graph B (EMA + Adam + metrics).  With VQB_DP_INGRAPH=1 the collective is
        captured too — ONE graph per step — through dist.GraphComm (a communicator of our own on the loaded NCCL library:
        ncclAllReduce on the capture stream; capturing torch.distributed.all_reduce itself hung at the first replay).  Verified on
        2 and 8 B200 (bit-identical ranks), but measured 0.4-0.8 % slower than the two-graph form, hence opt-in."""
        torch.cuda.synchronize()
        self.optimizer.refresh_lr()
        pool = torch.cuda.graph_pool_handle()
        st["gone"] = None
        gc = vdist.graph_comm() if world > 1 else None   # a communicator whose all-reduce can be captured (dist.GraphComm)
        if world == 1 or gc is not None:
            kw = {}
            if gc is not None:
                # NCCL sets up channels / protocols lazily per message size: once at full size outside the capture; its proxy thread
                # may touch the CUDA API while we capture, hence the thread-local capture mode
                tmp = torch.zeros_like(self._packed.comm)
                gc.all_reduce_sum(tmp)
                torch.cuda.synchronize()
                kw["capture_error_mode"] = "thread_local"
            g1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1, pool=pool, **kw):
                grads, tvars, losses = self._head(st["x"], world)
                if gc is not None:
                    gc.all_reduce_sum(self._packed.comm)   # [gradients | EMA statistics + restart rows | loss scalars], inside the graph
                st["vec"] = self._tail(grads, tvars, losses, world)
            st["gone"] = g1
            st["graph"] = True
            return
        ga = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ga, pool=pool):
            grads, tvars, losses = self._head(st["x"], world)
        st["ga"] = ga
        gb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gb, pool=pool):
            st["vec"] = self._tail(grads, tvars, losses, world)
        st["gb"] = gb
        st["graph"] = True

    def _accumulate(self, vec):
        self._metric_totals += vec
        ema_ran = bool(self.train_step_training)  # the VQ usage / entropy entries of `vec` are fresh only after an EMA update
        for m in (self._all_trackers() if ema_ran else self.metrics):
            m._count += 1
        return self._metric_dict()

    # ------------------------------------------------------------------------------------------------------
    def test_step(self, data):
        x = convert_to_tensor(data[0] if isinstance(data, (tuple, list)) else data)
        commit_losses, recon_losses, spectral_losses, level_losses = [], [], [], []
        for level in range(self.levels):  # bottom to top
            # the reference calls self.vqvaes[level](x) (vqvae.py:158); Keras resolves that to training=False
            _, reconstruction_loss, spectral_loss, commit_loss = self._level_losses(level, x, self.test_step_training)
            level_loss = reconstruction_loss + commit_loss + spectral_loss
            commit_losses.append(commit_loss)
            recon_losses.append(reconstruction_loss)
            spectral_losses.append(spectral_loss)
            level_losses.append(level_loss)
        return self.update_metrics(level_losses, recon_losses, commit_losses, spectral_losses)

    def call(self, x, training=False):
        """For callback model calls (vqvae.py:178-206): returns (recons, dict of per-level loss lists)."""
        if isinstance(x, tuple):
            x, _ = x
        x = convert_to_tensor(x)
        commit_losses, recon_losses, spectral_losses, level_losses, recons = [], [], [], [], []
        for level in range(self.levels):
            reconstructions, reconstruction_loss, spectral_loss, commit_loss = self._level_losses(level, x, training)
            recons.append(reconstructions)
            level_loss = reconstruction_loss + commit_loss + spectral_loss
            commit_losses.append(commit_loss)
            recon_losses.append(reconstruction_loss)
            spectral_losses.append(spectral_loss)
            level_losses.append(level_loss)
        return recons, {"level_losses": level_losses, "recon_losses": recon_losses, "commit_losses": commit_losses,
                        "spec_losses": spectral_losses}

    def _halo(self, level):
        """(halo in input samples, hop) of level `level`: an upper bound of the receptive-field radius of its encoder (and of its
        decoder, in output samples) — the summed extents (k - 1) * dilation * cumulative stride of all its convolutions —
        rounded up to a multiple of the level's hop, so that tile boundaries keep the phase of every strided convolution."""
        down_depth, strides, depth, factor = self._arch
        per_stack = sum(2 * factor ** i + 2 for i in range(depth))  # one DilatedResnet1D: k=3 dilated + k=3 per block
        j, ext = 1, 0
        for n, st in zip(down_depth[:level + 1], strides[:level + 1]):
            for _ in range(n):
                ext += (2 * st - 1) * j
                j *= st
                ext += per_stack * j
            ext += 2 * j
        return -(-ext // j) * j, j

    def _encode_once(self, x, level):
        enc_outputs = self.encoders[level](x, training=False)
        latent_output, latent_codes = self.vqs[level](enc_outputs, training=False)
        return latent_codes.view(enc_outputs.shape[:-1])  # (N, T_downsampled) int64

    def encode_level(self, x, level, chunk=1):
        """codes (N, T / hop) of one level (vqvae.py:208-219).  chunk > 1 (a TODO in the reference, which ignores the argument)
        time-tiles a long window: the T samples are encoded in `chunk` pieces, each with `_halo(level)` samples of real context
        on either side, and only the codes of the piece itself are kept — the same codes as the one-pass call (every code
        depends on inputs inside the halo only), with the activation memory of a piece instead of the whole window."""
        x = convert_to_tensor(x)
        T = x.shape[1]
        halo, hop = self._halo(level)
        if chunk <= 1 or T % hop or T <= hop * chunk:
            return self._encode_once(x, level)
        seg = -(-(T // hop) // chunk) * hop
        parts = []
        for s0 in range(0, T, seg):
            a, b, e = max(s0 - halo, 0), min(s0 + seg + halo, T), min(s0 + seg, T)
            codes = self._encode_once(x[:, a:b].contiguous(), level)
            parts.append(codes[:, (s0 - a) // hop:(e - a) // hop])
        return torch.cat(parts, dim=1)

    def encode(self, x, start_level=0, end_level=None, chunk=1):
        """list of code tensors for levels [start_level, end_level)  (vqvae.py:221-236)"""
        if end_level is None:
            end_level = self.levels
        return [self.encode_level(x, i, chunk) for i in range(start_level, end_level)]

    def _decode_once(self, zq, level):
        level_vq = self.vqs[level]
        quantized = ops.gather_codes(level_vq.embeddings.value, zq)
        return self.decoders[level](quantized, training=False)

    def decode_level(self, zq, level, chunk=1):
        """zq (N, T) int64 codes -> (N, T * hop, 1)  (vqvae.py:238-251): codebook gather + decoder; chunk > 1 time-tiles the code
        sequence the same way as encode_level (halo in latent positions)."""
        zq = convert_to_tensor(zq, torch.int64)
        Tl = zq.shape[1]
        halo, hop = self._halo(level)
        hl = halo // hop
        if chunk <= 1 or Tl <= chunk:
            return self._decode_once(zq, level)
        seg = -(-Tl // chunk)
        parts = []
        for s0 in range(0, Tl, seg):
            a, b, e = max(s0 - hl, 0), min(s0 + seg + hl, Tl), min(s0 + seg, Tl)
            y = self._decode_once(zq[:, a:b].contiguous(), level)
            parts.append(y[:, (s0 - a) * hop:(e - a) * hop])
        return torch.cat(parts, dim=1)

    def decode(self, zq, level=0, chunk=1):
        return self.decode_level(zq, level, chunk)

    def update_metrics(self, level_losses, recon_losses, commit_losses, spectral_losses):
        self.total_loss_tracker.update_state(sum(level_losses))
        self.reconstruction_loss_tracker.update_state(sum(recon_losses))
        self.vq_loss_tracker.update_state(sum(commit_losses))
        self.spectral_loss_tracker.update_state(sum(spectral_losses))
        for level in range(self.levels):
            self.level_loss_trackers[level].update_state(level_losses[level])
            self.recon_loss_trackers[level].update_state(recon_losses[level])
            self.vq_loss_trackers[level].update_state(commit_losses[level])
            self.spectral_loss_trackers[level].update_state(spectral_losses[level])
        return self._metric_dict()

    def _metric_dict(self):
        """The flat dict `update_metrics` returns (vqvae.py:276-304), same key order."""
        ret_metrics = dict(loss=self.total_loss_tracker.result(),
                           recon_loss=self.reconstruction_loss_tracker.result(),
                           vqvae_loss=self.vq_loss_tracker.result(),
                           spectral_loss=self.spectral_loss_tracker.result())
        for level in range(self.levels):
            for t in (self.level_loss_trackers, self.recon_loss_trackers, self.vq_loss_trackers,
                      self.spectral_loss_trackers):
                ret_metrics[t[level].name] = t[level].result()
            ret_metrics.update({m.name: m.result() for m in self.vqs[level].metrics})
        return ret_metrics

    def get_quantizer(self):
        return self.vqs[0]

    def _multispectral_loss(self, target, recon, **kwargs):
        return MultiSpectralLoss(target, recon)
