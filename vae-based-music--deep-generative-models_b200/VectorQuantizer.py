"""VectorQuantizer — drop-in for the reference's `VectorQuantizer.py` (class `VectorQuantizer`, :7-199).

Same constructor, attributes (`embeddings [D,K]`, `m_t`, `N_t`, `num_embeddings`, `metrics`, `losses`) and methods
(`call(x, training=True, debug=False) -> (quantized, indices)`, `get_code_indices`, `get_usage_count`, `_tile`);
the work is done by `vqb_vq_fwd` / `vqb_vq_bwd` / `vqb_vq_ema_update` (include/vqb.h).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, ops
from .keras_compat import KerasTensor, Mean, Scalar, Variable, _is_symbolic, convert_to_tensor, layers, record, rng


class VectorQuantizer(layers.Layer):
    def __init__(self, num_embeddings, embedding_dim, beta=0.25, codebook_usage_threshold=1.0, decay_rate=0.99,
                 level=0, **kwargs):
        super().__init__(**kwargs)
        self.embedding_dim = embedding_dim
        self.num_embeddings = num_embeddings
        self.beta = beta  # VectorQuantizer.py:19-21
        self.codebook_usage_threshold = codebook_usage_threshold
        self.gamma = decay_rate
        # tf.random_uniform_initializer() default range (-0.05, 0.05); non-trainable (VectorQuantizer.py:25,38-44)
        w = rng().uniform(-0.05, 0.05, size=(embedding_dim, num_embeddings)).astype(np.float32)
        self.embeddings = Variable(w, trainable=False, name="embeddings_vqvae")
        self.m_t = Variable(self.embeddings, trainable=False, name="m_t")                       # :48-51
        self.N_t = Variable(np.ones((num_embeddings,), np.float32), trainable=False, name="N_t")  # :57-60
        self.batch_usage_tracker = Mean(name="[{}]batch_codebook_usage".format(level))
        self.usage_tracker = Mean(name="[{}]codebook_usage".format(level))
        self.entropy_tracker = Mean(name="[{}]codebook_entropy".format(level))
        self.built = True
        # --- implementation state (not part of the reference surface)
        self.precision = _lib.PREC_FP32
        self.restart_seed = 0x5EED + level
        self.restart_ids = None        # injected int64 [K] row numbers (parity tests); None -> device RNG
        self._step = None              # device int64 counter feeding the restart RNG
        self._stats = None             # (m_batch [D,K], n_batch [K], restart_rows [K,D]) — may be views of the comm buffer
        self._metrics_buf = None       # [3] batch usage, running usage, entropy of the last EMA update
        self.defer_ema = False         # data parallel: statistics are all-reduced before apply_ema()
        self.shard = (1, 0)            # (world size, rank) for the restart-row pick
        self.skip_metric_update = False
        self._pending = False

    @property
    def metrics(self):
        return [self.batch_usage_tracker, self.usage_tracker, self.entropy_tracker]

    # ------------------------------------------------------------------------------------------------
    def _buffers(self):
        if self._stats is None:
            D, K = self.embedding_dim, self.num_embeddings
            self._stats = (ops.empty(D, K), ops.empty(K), ops.empty(K, D))
        if self._metrics_buf is None:
            self._metrics_buf = ops.zeros(3)
        if self._step is None:
            self._step = ops.zeros(1, dtype=torch.int64)
        return self._stats

    def bind_stats(self, m_batch, n_batch, rows):
        """Place the batch statistics in caller-owned memory (the all-reduce buffer under data parallelism)."""
        self._stats = (m_batch, n_batch, rows)

    def call(self, x, training=True, debug=False):
        if _is_symbolic(x):
            return KerasTensor(x.shape), KerasTensor((None,))
        D = self.embedding_dim
        if x.shape[-1] != D:
            raise ValueError(f"{self.name}: last dimension {x.shape[-1]} != embedding_dim {D}")
        input_shape = x.shape
        xc = x if x.is_contiguous() else x.contiguous()
        flat = xc.view(-1, D)                                          # VectorQuantizer.py:79
        m_batch = n_batch = rows = None
        if training:
            m_batch, n_batch, rows = self._buffers()
        # nearest code, gather, commitment loss, straight-through output and batch statistics: one call (:83-124)
        idx, q_st, q, loss = ops.vq_fwd(flat, self.embeddings.value, self.beta, True, True, m_batch, n_batch,
                                        self.precision)
        commitment_loss = Scalar.leaf(loss)
        self.add_loss(commitment_loss)                                  # :107
        quantized = q_st.view(input_shape)
        beta = self.beta

        def bwd(g, needs):
            dq, c = g
            if not needs[0]:
                return [None]
            dqf = None if dq is None else dq.contiguous().view(-1, D)
            return [ops.vq_bwd(dqf, flat, q, beta, float(c or 0.0)).view(input_shape)]

        record([x], [quantized, loss], bwd)

        if training:
            N = flat.shape[0]
            world, rank = self.shard
            ids = self.restart_ids if self.restart_ids is not None else \
                ops.restart_ids(N * world, self.num_embeddings, self.restart_seed, self._step)
            n_tot = N * world
            if n_tot < self.num_embeddings:                              # _tile, :191-199
                n_tot *= -(-self.num_embeddings // n_tot)
            # rows of the (tiled, virtually shuffled) encoder outputs owned by this rank; tiling = ids mod N_total
            ops.gather_rows(flat, convert_to_tensor(ids, torch.int64), N * world, rank * N, out=rows)
            self._pending = True
            if not self.defer_ema:
                self.apply_ema()
        if debug:
            print("VQ input (Encoder Output): ", x)
            print("VQ output: ", quantized)
        return quantized, idx

    def apply_ema(self):
        """EMA + dead-code restart + usage metrics (VectorQuantizer.py:128-159) from the current batch statistics."""
        if not self._pending:
            return
        m_batch, n_batch, rows = self._stats
        ops.vq_ema_update(self.embeddings.value, self.m_t.value, self.N_t.value, m_batch, n_batch, rows, self.gamma,
                          self.codebook_usage_threshold, self._metrics_buf)
        ops.increment(self._step)
        self._pending = False
        if not self.skip_metric_update:
            self.batch_usage_tracker.update_state(self._metrics_buf[0:1])
            self.usage_tracker.update_state(self._metrics_buf[1:2])
            self.entropy_tracker.update_state(self._metrics_buf[2:3])

    def get_code_indices(self, flattened_inputs):
        """(N, D) -> (N,) int64 (VectorQuantizer.py:170-186)."""
        flat = convert_to_tensor(flattened_inputs)
        idx, _, _, _ = ops.vq_fwd(flat, self.embeddings.value, self.beta, False, False, None, None, self.precision)
        return idx

    def get_usage_count(self):
        return self.N_t

    def _tile(self, x):
        x = convert_to_tensor(x)
        nt = x.shape[0]
        if nt < self.num_embeddings:
            x = x.repeat(((self.num_embeddings + nt - 1) // nt, 1))
        return x
