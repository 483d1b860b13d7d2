"""reference_tf — runs the UNMODIFIED reference (vqvae.py / encdec.py / resnet.py / VectorQuantizer.py / data_utils.py) under
TensorFlow when TensorFlow exists, so that the oracle can be pinned against the real thing.  TEST INFRASTRUCTURE: imported
only by tests/, oracle/make_reference_fixtures.py and bench.py's `--impl reference` arm.

TensorFlow 2.7 is not installable in the build container (no network, Python 3.12), so nothing here has ever run in
it: `available()` is False and everything that depends on it is skipped.  On a box that has TensorFlow >= 2.4 and a
checkout of the reference (REFERENCE_DIR, default /root/reference, or $VQB_REFERENCE_DIR, or baseline/_ref):

    python -m oracle.make_reference_fixtures        # writes tests/golden/reference_tf.npz
    python -m pytest tests/test_reference_tf.py     # oracle vs the live reference, and vs the fixture

The reference files import `tensorflow_addons` (never used: `grep tfa\\.` -> 0 hits) and, in data_utils.py, librosa /
matplotlib / sklearn (file-loading helpers outside the path); those are stubbed in sys.modules when absent.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_MODS = None


def reference_dir():
    for d in (os.environ.get("VQB_REFERENCE_DIR"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if d and os.path.exists(os.path.join(d, "vqvae.py")):
            return d
    return None


def available():
    """(ok, why): TensorFlow importable AND the reference sources present."""
    if reference_dir() is None:
        return False, "reference sources not found (VQB_REFERENCE_DIR, /root/reference, baseline/_ref)"
    try:
        importlib.import_module("tensorflow")
    except Exception as e:  # ImportError, or a broken wheel
        return False, f"tensorflow not importable: {e!r}"
    return True, ""


def _stub(name, **attrs):
    if name in sys.modules:
        return
    try:
        importlib.import_module(name)
    except Exception:
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        parent, _, child = name.rpartition(".")
        if parent and parent in sys.modules:
            setattr(sys.modules[parent], child, m)


def load():
    """Imports the reference modules (once) and returns them as a namespace: .tf .vqvae .encdec .resnet .VectorQuantizer
    .data_utils"""
    global _MODS
    if _MODS is not None:
        return _MODS
    ok, why = available()
    if not ok:
        raise ImportError(why)
    import tensorflow as tf
    _stub("tensorflow_addons")
    _stub("librosa")
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    _stub("sklearn")
    _stub("sklearn.model_selection", train_test_split=None)
    d = reference_dir()
    if d not in sys.path:
        sys.path.insert(0, d)
    ns = types.SimpleNamespace(tf=tf)
    for name in ("data_utils", "resnet", "encdec", "VectorQuantizer", "vqvae"):
        setattr(ns, name, importlib.import_module(name))
    _MODS = ns
    return ns


def build_model(spec, weights, vq):
    """vqvae.VQVAE built with the oracle's ModelSpec and loaded with its synthetic weights: trainable variables are in layer
    creation order (kernel, bias per conv; encoder then decoder) — the order of oracle.init_model; VQ state by attribute."""
    R = load()
    m = R.vqvae.VQVAE(input_shape=(spec.T, spec.channels), levels=spec.levels, latent_dim=spec.latent_dim,
                      num_embeddings=spec.num_embeddings, down_depth=list(spec.down_depth), strides=list(spec.strides),
                      dilation_factor=spec.dilation_factor, residual_width=spec.residual_width,
                      residual_depth=spec.residual_depth)
    for l in range(spec.levels):
        tv = m.vqvaes[l].trainable_variables
        assert len(tv) == len(weights[l]), (len(tv), len(weights[l]))
        for v, w in zip(tv, weights[l]):
            assert tuple(v.shape) == tuple(w.shape), (v.name, v.shape, w.shape)
            v.assign(w)
        q = m.vqs[l]
        q.embeddings.assign(vq[l]["E"]); q.m_t.assign(vq[l]["m_t"]); q.N_t.assign(vq[l]["N_t"])
    return m


def run_case(spec, weights, vq, x):
    """Everything the parity contract names, from the live reference, as numpy arrays:
    per level: recon, the three losses, code indices, gradients of (recon + commit + spectral) w.r.t. the level's trainable
    variables (vqvae.py:119-143), and the VQ state after ONE training-mode call of the quantizer on the encoder output
    (VectorQuantizer.py:118-159; `E_alive` marks the codes whose new embedding does not come from tf.random.shuffle)."""
    R = load()
    tf = R.tf
    m = build_model(spec, weights, vq)
    xt = tf.constant(x)
    out = {}
    recons, losses = m(xt, training=False)
    codes = m.encode(xt)
    for l in range(spec.levels):
        out[f"recon{l}"] = recons[l].numpy()
        out[f"losses{l}"] = np.array([float(losses[k][l]) for k in ("recon_losses", "commit_losses", "spec_losses")])
        out[f"idx{l}"] = codes[l].numpy().reshape(-1)
        tv = m.vqvaes[l].trainable_variables
        with tf.GradientTape() as tape:
            r = m.vqvaes[l](xt, training=False)
            total = (tf.reduce_mean(m.loss_fn(xt, r)) + tf.reduce_mean(m._multispectral_loss(xt, r))
                     + sum(m.vqvaes[l].losses))
        for i, g in enumerate(tape.gradient(total, tv)):
            out[f"g{l}_{i:03d}"] = g.numpy()
        # one EMA step of this level's quantizer (training=True is the layer's own default, VectorQuantizer.py:75)
        z = m.encoders[l](xt, training=False)
        q = m.vqs[l]
        q(z, training=True)
        out[f"ema{l}_m_t"] = q.m_t.numpy(); out[f"ema{l}_N_t"] = q.N_t.numpy(); out[f"ema{l}_E"] = q.embeddings.numpy()
        out[f"ema{l}_alive"] = (q.N_t.numpy() >= q.codebook_usage_threshold)
        out[f"z{l}"] = z.numpy()
    return out


def time_train_step(batch, steps, warmup, T=28160):
    """bench.py --impl reference: the reference's own train_step (vqvae.py:111-146) on SMALL_VQ_VAE, TF CPU, all host threads.
    Returns dict(value samples/s, ms_per_step, cores, sample)."""
    import time
    R = load()
    tf = R.tf
    try:
        tf.config.set_visible_devices([], "GPU")  # the arm is the reference's CPU path
    except Exception:
        pass
    m = R.vqvae.VQVAE(input_shape=(T, 1), levels=2, latent_dim=64, num_embeddings=512, down_depth=[5, 3], strides=[2, 2],
                      dilation_factor=3, residual_width=32)  # vqvae.py:352-353
    m.compile(optimizer=tf.keras.optimizers.Adam())
    rng = np.random.Generator(np.random.PCG64(0))
    x = tf.constant(rng.uniform(0, 1, size=(batch, T, 1)).astype(np.float32))
    for _ in range(warmup):
        m.train_step((x, None))
    t0 = time.perf_counter()
    for _ in range(steps):
        logs = m.train_step((x, None))
    float(logs["loss"])
    dt = (time.perf_counter() - t0) / max(steps, 1)
    cores = os.cpu_count() or 1
    return dict(value=batch * T / dt, ms_per_step=1e3 * dt, cores=cores,
                sample=f"{steps} train_step(s) of the unmodified reference (vqvae.VQVAE, TensorFlow {tf.__version__}, CPU, "
                       f"{cores} host threads) on {batch} windows of {T} samples")
