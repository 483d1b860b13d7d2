"""Host-side logic of the package (Keras-style API, tape, packed buffers, training-flag resolution, metric plumbing,
data-parallel sharding) exercised on CPU through the test double of the C ABI (tests/fake_backend.py).
No numerical-parity claim is made here — the kernels are checked by the gpu-marked tests."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from oracle import vqvae_oracle as O
from oracle.make_golden import TINY, tiny_case


def build_tiny(V, seed=0):
    spec, weights, vq, x = tiny_case(seed)
    kw = {k: (list(v) if isinstance(v, tuple) else v) for k, v in TINY.items() if k != "T"}
    m = V.VQVAE((TINY["T"], 1), **kw)
    m.use_cuda_graph = False
    for l in range(spec.levels):
        m.vqvaes[l].set_weights_trainable = None
        for v, w in zip(m.vqvaes[l].trainable_variables, weights[l]):
            v.assign(w)
        m.vqs[l].embeddings.assign(vq[l]["E"]); m.vqs[l].m_t.assign(vq[l]["m_t"]); m.vqs[l].N_t.assign(vq[l]["N_t"])
        m.vqs[l].restart_ids = np.arange(max(TINY["num_embeddings"], 1))
    return m, spec, weights, vq, x


def test_structure_names_and_variable_order(cpu_backend):
    V = cpu_backend
    m = V.VQVAE((28160, 1), **V.SMALL_VQ_VAE)
    spec = O.ModelSpec(T=28160, **O.SMALL_VQ_VAE)
    for l in range(2):
        want = [s for op in spec.level_ops(l) for s in op.param_shapes()]
        assert [v.shape for v in m.vqvaes[l].trainable_variables] == want
    assert len(m.trainable_variables) == 484
    assert sum(int(np.prod(v.shape)) for v in m.trainable_variables) == 799042
    assert m.vqvaes[0].name == "vq_vae_0" and m.encoders[1].name == "encoder_1" and m.vqs[0].name == "vector_quantizer_0"
    enc0 = m.encoders[0].model.layers[0]  # EncoderConvBlock
    assert type(enc0).__name__ == "EncoderConvBlock" and len(enc0.model.layers) == 11
    stack = enc0.model.layers[1]
    assert "dilated_resnet1d" in stack.name
    block = stack.model.layers[2]
    assert [type(l).__name__ for l in block.model.layers] == ["ReLU", "Conv1D", "ReLU", "Conv1D"]
    assert block.model.layers[1].name == "dilated_cov1d_dr-9" and block.model.layers[1].dilation_rate == 9
    dec_stack = m.decoders[0].model.layers[0].model.layers[1]
    assert [b.dilation for b in dec_stack.model.layers] == [27, 9, 3, 1]  # reversed for decoders (resnet.py:54-55)
    # every trainable variable is a view of ONE parameter buffer, in order
    base = m._packed.params.data_ptr()
    off = 0
    for v in m.trainable_variables:
        assert v.value.data_ptr() == base + 4 * off and v.grad.data_ptr() == m._packed.grads.data_ptr() + 4 * off
        off += (v.value.numel() + 3) & ~3  # 16-byte aligned slices
    # VQ state: non-trainable [D,K], m_t = E, N_t = ones (VectorQuantizer.py:38-60)
    vq = m.get_quantizer()
    assert vq.embeddings.shape == (64, 512) and not vq.embeddings.trainable
    assert torch.equal(vq.m_t.value, vq.embeddings.value) and float(vq.N_t.value.sum()) == 512
    assert float(vq.embeddings.value.abs().max()) <= 0.05
    assert [t.name for t in vq.metrics] == ["[0]batch_codebook_usage", "[0]codebook_usage", "[0]codebook_entropy"]
    V.print_dec_layer(m.decoders[0])


def test_bad_arguments_raise(cpu_backend):
    V = cpu_backend
    with pytest.raises(AssertionError, match="not Legit"):  # encdec.py:84-85
        V.Encoder(8, 8, 2, depth=2, down_depth=[2], strides=[2, 2])
    with pytest.raises(AssertionError, match="not Legit"):
        V.Decoder(1, 8, 8, 2, depth=1, down_depth=[2], strides=[2, 2])
    blk = V.ResnetConv1DBlock(8, 8, dilation=3)
    with pytest.raises(ValueError):
        blk(np.zeros((1, 16, 4), np.float32))
    vq = V.VectorQuantizer(6, 2)
    with pytest.raises(ValueError):
        vq(np.zeros((2, 5, 3), np.float32))
    with pytest.raises(NotImplementedError):
        V.keras.layers.Conv1D(4, 3, padding="valid")


def test_training_flag_resolution(cpu_backend):
    """Keras-2.7 rules: top-level VQ call defaults to training=True (signature default, VectorQuantizer.py:75);
    inside the functional model it inherits the outer value (False when nothing is passed)."""
    V = cpu_backend
    m, spec, weights, vq, x = build_tiny(V)
    E0 = m.vqs[0].embeddings.numpy().copy()
    z = m.encoders[0](x)
    m.vqs[0](z)  # top level: EMA runs
    assert not np.array_equal(m.vqs[0].embeddings.numpy(), E0)
    E1 = m.vqs[0].embeddings.numpy().copy()
    m.vqvaes[0](x)  # functional model, nothing passed -> False
    assert np.array_equal(m.vqs[0].embeddings.numpy(), E1)
    m.vqvaes[0](x, training=True)
    assert not np.array_equal(m.vqs[0].embeddings.numpy(), E1)
    E2 = m.vqs[0].embeddings.numpy().copy()
    m(x)  # VQVAE.call default training=False (vqvae.py:178)
    m.test_step((x, None))
    assert np.array_equal(m.vqs[0].embeddings.numpy(), E2)
    assert len(m.vqvaes[0].losses) == 1  # exactly the commitment loss of the last call (add_loss, :107)


def test_forward_and_train_steps_match_oracle(cpu_backend):
    V = cpu_backend
    m, spec, weights, vq, x = build_tiny(V)
    recons, losses = m(x, training=False)
    ref = O.forward_losses(spec, weights, vq, torch.tensor(x))
    for l in range(spec.levels):
        np.testing.assert_allclose(recons[l].numpy(), ref[l]["recon"].numpy(), rtol=1e-5, atol=1e-6)
        for key, rk in (("recon_losses", "recon_loss"), ("commit_losses", "commit_loss"), ("spec_losses", "spec_loss")):
            assert abs(float(losses[key][l]) - float(ref[l][rk])) < 1e-5
    m.compile(optimizer=V.keras.optimizers.Adam())
    tr = O.OracleTrainer(spec, weights, vq)
    for step in range(2):
        logs = m.train_step((x, np.zeros(len(x))))
        tr.train_step(torch.tensor(x))
    assert m.optimizer.iterations == 2
    for l in range(spec.levels):
        for a, b in zip(m.vqvaes[l].trainable_variables, tr.w[l]):
            np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(m.vqs[l].embeddings.numpy(), tr.vq[l].E.numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(m.vqs[l].N_t.numpy(), tr.vq[l].N_t.numpy(), rtol=1e-6)
    want_keys = ["loss", "recon_loss", "vqvae_loss", "spectral_loss"]
    for l in range(spec.levels):
        want_keys += [f"[{l}]level_loss", f"[{l}]recon_loss", f"[{l}]vq_loss", f"[{l}]spectral_loss",
                      f"[{l}]batch_codebook_usage", f"[{l}]codebook_usage", f"[{l}]codebook_entropy"]
    assert list(logs) == want_keys  # vqvae.py:276-304
    assert abs(float(logs["loss"]) - sum(float(logs[f"[{l}]level_loss"]) for l in range(spec.levels))) < 1e-5


def test_encode_decode_roundtrip_shapes(cpu_backend):
    V = cpu_backend
    m, spec, weights, vq, x = build_tiny(V)
    zs = m.encode(x)
    assert [tuple(z.shape) for z in zs] == [(3, 2048 // 8), (3, 2048 // 32)] and zs[0].dtype == torch.int64
    assert [tuple(z.shape) for z in m.encode(x, start_level=1)] == [(3, 64)]
    ref = O.forward_losses(spec, weights, vq, torch.tensor(x))
    for l in range(2):
        assert torch.equal(zs[l].reshape(-1), ref[l]["idx"])
        y = m.decode(zs[l], level=l)
        # decode(gather(E, idx)) equals the decoder applied to the raw (not straight-through) quantised latents
        want = O.run_ops(spec.dec_ops[l], [torch.tensor(w) for w in weights[l][spec.n_enc_params(l):]], ref[l]["q"])
        np.testing.assert_allclose(y.numpy(), want.numpy(), rtol=1e-5, atol=1e-6)


def test_fit_evaluate_and_checkpoint(cpu_backend, tmp_path):
    V = cpu_backend
    m, spec, weights, vq, x = build_tiny(V)
    m.compile(optimizer=V.keras.optimizers.Adam())
    xs = np.concatenate([x, x[:2]])  # 5 windows, batch 2 -> ragged last batch
    h = m.fit(xs, np.zeros(5), batch_size=2, epochs=2, verbose=0, shuffle=False)
    assert len(h.history["loss"]) == 2 and m.optimizer.iterations == 6
    ev = m.evaluate(xs, None, batch_size=2, verbose=0, return_dict=True)
    assert ev["loss"] > 0 and "[1]codebook_entropy" in ev
    m.save_weights(str(tmp_path / "ck"))
    m2, *_ = build_tiny(V, seed=3)
    m2.load_weights(str(tmp_path / "ck"))
    for a, b in zip(m.variables, m2.variables):
        assert np.array_equal(a.numpy(), b.numpy())


def test_generic_tape_with_plain_layers(cpu_backend):
    """Conv1D / ReLU / add used directly (the reference's un-fused formulation of the residual block, resnet.py:29)
    give the same result and gradients as the fused block."""
    V = cpu_backend
    K = V.keras
    rng = np.random.default_rng(0)
    x = torch.tensor(rng.normal(size=(2, 40, 8)).astype(np.float32))
    blk = V.ResnetConv1DBlock(8, 8, dilation=3)
    pre = K.layers.Conv1D(8, 3, padding="same")
    with V.GradientTape() as tape:
        h0 = pre(x)
        y_fused = blk(h0)
        loss = K.reduce_mean(K.losses.MeanSquaredError()(np.zeros((2, 40, 8), np.float32), y_fused))
    vars_ = pre.trainable_variables + blk.trainable_variables
    g_fused = [g.clone() for g in tape.gradient(loss, vars_)]
    with V.GradientTape() as tape:
        h0 = pre(x)
        y_plain = K.layers.add([h0, blk.model(h0)])
        loss2 = K.reduce_mean(K.losses.MeanSquaredError()(np.zeros((2, 40, 8), np.float32), y_plain))
    g_plain = tape.gradient(loss2, vars_)
    np.testing.assert_allclose(y_fused.numpy(), y_plain.numpy(), rtol=1e-5, atol=1e-6)
    for a, b in zip(g_fused, g_plain):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-4, atol=1e-7)


# ---------------------------------------------------------------------------------------------------------------
def _dp_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import vqvae_b200 as V
    from tests.fake_backend import FakeBackend
    V._lib.set_backend(FakeBackend(), "cpu")
    td.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    m, spec, weights, vq, x = build_tiny(V)
    xs = np.concatenate([x, x[::-1]])[:4]  # global batch 4
    m.compile(optimizer=V.keras.optimizers.Adam())
    per = len(xs) // world
    for _ in range(2):
        logs = m.train_step((xs[rank * per:(rank + 1) * per], None))
    out = {f"w{i}": v.numpy() for i, v in enumerate(m.trainable_variables)}
    for l in range(spec.levels):
        out[f"E{l}"] = m.vqs[l].embeddings.numpy(); out[f"N{l}"] = m.vqs[l].N_t.numpy()
    out["loss"] = np.float32(float(logs["loss"]))
    np.savez(os.path.join(tmp, f"rank{rank}.npz"), **out)
    td.destroy_process_group()


def test_data_parallel_two_ranks_equal_unsharded(cpu_backend, tmp_path):
    """world_size 2 over gloo: batch-sharded training (one flat all-reduce of gradients + EMA statistics + restart
    rows + loss scalars) must reproduce the un-sharded step, and leave both ranks bit-identical."""
    V = cpu_backend
    port = 29500 + os.getpid() % 2000
    mp.spawn(_dp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for k in r0.files:
        assert np.array_equal(r0[k], r1[k]), f"ranks differ in {k}"
    m, spec, weights, vq, x = build_tiny(V)
    xs = np.concatenate([x, x[::-1]])[:4]
    m.compile(optimizer=V.keras.optimizers.Adam())
    for _ in range(2):
        logs = m.train_step((xs, None))
    for i, v in enumerate(m.trainable_variables):
        np.testing.assert_allclose(r0[f"w{i}"], v.numpy(), rtol=2e-4, atol=2e-6)
    for l in range(spec.levels):
        np.testing.assert_allclose(r0[f"E{l}"], m.vqs[l].embeddings.numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(r0[f"N{l}"], m.vqs[l].N_t.numpy(), rtol=1e-6)
    assert abs(float(r0["loss"]) - float(logs["loss"])) < 1e-4


def test_conditioner_decoder_conv_block_host_logic(cpu_backend):
    """The ConditionerNet-shaped DecoderConvBlock (depth 8, cyclic dilation 4) through the layer classes on the
    CPU test double: variable count and order, dilation schedule, output shape, tape gradients (same body as the GPU test)."""
    from tests.test_gpu_model import test_conditioner_decoder_conv_block as body
    body(cpu_backend, "fp32", 1e-3)


def test_resblock_sign_mask_variants_on_the_test_double(cpu_backend):
    """vqb_resblock_fwd_masks / vqb_resblock_bwd_data_masks (include/vqb.h): bit c of xbits / hbits[b, t] is set iff channel c of
    x / h is > 0, and the masked data gradient equals the one computed from the fp32 tensors (ABI semantics, CPU double)."""
    V = cpu_backend
    ops = V.ops
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 37, 32, generator=g); dy = torch.randn(2, 37, 32, generator=g)
    w1 = torch.randn(3, 32, 32, generator=g) * 0.1; w2 = torch.randn(3, 32, 32, generator=g) * 0.1
    b1 = torch.randn(32, generator=g) * 0.1; b2 = torch.randn(32, generator=g) * 0.1
    y0, h0 = ops.resblock_fwd(x, w1, b1, w2, b2, 3, 0)
    y1, h1, xb, hb = ops.resblock_fwd_masks(x, w1, b1, w2, b2, 3, 0)
    assert torch.equal(y0, y1) and torch.equal(h0, h1) and xb.dtype == torch.int32 and tuple(xb.shape) == (2, 37)
    sh = torch.arange(32)
    assert torch.equal(((xb.long().unsqueeze(-1) >> sh) & 1).bool(), x > 0)
    assert torch.equal(((hb.long().unsqueeze(-1) >> sh) & 1).bool(), h0 > 0)
    dx0, dh0 = ops.resblock_bwd_data(x, h0, dy, w1, w2, 3, 0)
    dx1, dh1 = ops.resblock_bwd_data_masks(xb, hb, dy, w1, w2, 3, 0)
    assert torch.equal(dx0, dx1) and torch.equal(dh0, dh1)


def test_block_weight_gradient_queue(cpu_backend):
    """ops.resblock_wgrad inside a reduce_begin() / reduce_flush() window (a backward pass) holds the blocks of a stack back and
    hands them to vqb_resblock_wgrad_batch four at a time; a change of shape flushes the queue; outside the window and in
    fp32 every call is immediate.  Host logic on the CPU double: the batched results equal the immediate ones."""
    V = cpu_backend
    ops, P = V.ops, V._lib.PREC_FP16X2
    g = torch.Generator().manual_seed(3)

    def block(B, L):
        t = [torch.randn(B, L, 32, generator=g) for _ in range(4)]
        return t, [torch.empty(3, 32, 32), torch.empty(32), torch.empty(3, 32, 32), torch.empty(32)]

    blocks = [block(2, 50) for _ in range(5)] + [block(1, 20)]
    dil = [27, 9, 3, 1, 27, 3]
    want = []
    for (t, _), d in zip(blocks, dil):
        o = [torch.empty(3, 32, 32), torch.empty(32), torch.empty(3, 32, 32), torch.empty(32)]
        ops.resblock_wgrad(*t, *o, d, P)        # no window: immediate
        assert not ops._wg_queue
        want.append(o)
    calls = []
    real = V._lib._BACKEND.vqb_resblock_wgrad_batch
    V._lib._BACKEND.vqb_resblock_wgrad_batch = lambda dref, n, *a: (calls.append(n), real(dref, n, *a))[1]
    try:
        ops.reduce_begin()
        for i, ((t, o), d) in enumerate(zip(blocks, dil)):
            ops.resblock_wgrad(*t, *o, d, P)
            assert len(ops._wg_queue) == [1, 2, 3, 0, 1, 1][i]   # full batch after 4; the new shape flushes the single leftover
        ops.reduce_flush()
    finally:
        V._lib._BACKEND.vqb_resblock_wgrad_batch = real
    assert calls == [4] and not ops._wg_queue  # batches of one go through vqb_resblock_wgrad
    for (_, o), w in zip(blocks, want):
        for a, b in zip(o, w):
            assert torch.equal(a, b)
    ops.reduce_begin()
    t, o = blocks[0]
    ops.resblock_wgrad(*t, *o, 1, 0)            # fp32: never queued
    assert not ops._wg_queue
    ops.reduce_flush()


# ---------------------------------------------------------------------------------------------------------------
# regressions for the round-1 review findings
def test_evaluate_scores_every_batch_against_its_own_target(cpu_backend):
    """The |STFT| of the target is cached per step, keyed on the tensor OBJECT (+ version): two different batches that happen
    to live at the same address (a recycled allocator block on the GPU; one reused numpy buffer here) must not share it."""
    V = cpu_backend
    m, spec, weights, vq, x = build_tiny(V)
    rng = np.random.default_rng(5)
    x2 = rng.uniform(0, 1, size=x.shape).astype(np.float32)
    want = []
    for xb in (x, x2):
        ref = O.forward_losses(spec, weights, vq, torch.tensor(xb))
        want.append([float(ref[l]["spec_loss"]) for l in range(spec.levels)])
    buf = torch.empty(x.shape, dtype=torch.float32)  # ONE buffer, filled with two batches in turn
    got = []
    for xb in (x, x2):
        buf.copy_(torch.tensor(xb))
        _, losses = m(buf, training=False)
        got.append([float(s) for s in losses["spec_losses"]])
    np.testing.assert_allclose(got, want, rtol=2e-4)
    assert abs(got[0][0] - got[1][0]) > 1e-4  # the two batches really differ


def test_adam_learning_rate_is_live_and_accepts_schedules(cpu_backend):
    """`optimizer.learning_rate` is read at every update (it lives in a device scalar the kernel reads, so a captured
    train_step follows it too), and may be a callable step -> lr (Keras LearningRateSchedule)."""
    V = cpu_backend
    K = V.keras
    p0 = np.ones(8, np.float32)
    g = np.full(8, 0.5, np.float32)

    def run(lr, steps):
        v = K.Variable(p0.copy(), name="p")
        opt = K.optimizers.Adam(learning_rate=lr(0) if callable(lr) else lr)
        opt.learning_rate = lr
        out = []
        for _ in range(steps):
            opt.apply_gradients([(torch.tensor(g), v)])
            out.append(v.numpy().copy())
        return out, opt

    const, _ = run(1e-2, 3)
    sched, opt = run(lambda step: 1e-2 if step < 1 else 0.0, 3)
    np.testing.assert_allclose(sched[0], const[0])
    assert np.array_equal(sched[1], sched[0]) and np.array_equal(sched[2], sched[0])  # lr = 0 from the second update on
    assert opt.iterations == 3 and opt._host_iterations == 3
    # reassigning a plain number takes effect on the next update
    v = K.Variable(p0.copy(), name="q")
    opt = K.optimizers.Adam(learning_rate=1e-2)
    opt.apply_gradients([(torch.tensor(g), v)])
    a = v.numpy().copy()
    opt.learning_rate = 0.0
    opt.apply_gradients([(torch.tensor(g), v)])
    assert np.array_equal(v.numpy(), a)


def test_compile_drops_captured_graphs_and_step_flags_are_scoped(cpu_backend):
    V = cpu_backend
    m, spec, weights, vq, x = build_tiny(V)
    m._graphs["stale"] = object()
    m.compile(optimizer=V.keras.optimizers.Adam())
    assert m._graphs == {}
    m.train_step((x, None))
    assert all(not q.defer_ema and not q.skip_metric_update for q in m.vqs)


def test_time_tiled_encode_decode_equals_one_pass(cpu_backend):
    """encode_level / decode_level with chunk > 1 (vqvae.py:208,238: `chunk` is an unused TODO in the reference): pieces with
    `_halo(level)` samples of context on either side give the codes / audio of the one-pass call."""
    V = cpu_backend
    kw = dict(levels=2, latent_dim=8, num_embeddings=16, down_depth=[2, 1], strides=[2, 2], dilation_factor=2, residual_width=4,
              residual_depth=2)
    T = 4096
    m = V.VQVAE((T, 1), **kw)
    assert m._halo(0) == (80, 4) and m._halo(1) == (192, 8)   # summed (k-1) * dilation * cumulative stride, rounded up to the hop
    x = np.random.default_rng(0).uniform(0, 1, size=(2, T, 1)).astype(np.float32)
    for level in range(2):
        one = m.encode_level(x, level)
        for chunk in (2, 3, 7):
            assert torch.equal(m.encode_level(x, level, chunk=chunk), one), (level, chunk)
        y = m.decode_level(one, level)
        for chunk in (2, 5):
            np.testing.assert_allclose(m.decode_level(one, level, chunk=chunk).numpy(), y.numpy(), rtol=0, atol=1e-6)
    assert [tuple(c.shape) for c in m.encode(x, chunk=4)] == [(2, T // 4), (2, T // 8)]


def test_checkpoint_carries_optimizer_and_restart_state(cpu_backend, tmp_path):
    V = cpu_backend
    m, spec, weights, vq, x = build_tiny(V)
    m.compile(optimizer=V.keras.optimizers.Adam())
    for _ in range(2):
        m.train_step((x, None))
    m.save_weights(str(tmp_path / "ck"))
    m2, *_ = build_tiny(V, seed=5)
    m2.compile(optimizer=V.keras.optimizers.Adam())
    m2.load_weights(str(tmp_path / "ck"))
    assert m2.optimizer.iterations == 2
    # both continue identically: same weights after one more step (Adam moments and bias-correction step were restored)
    m.train_step((x, None)); m2.train_step((x, None))
    for a, b in zip(m.variables, m2.variables):
        assert np.array_equal(a.numpy(), b.numpy()), a.name
    # a file of another architecture is refused by NAME, not just by shape
    m3 = V.VQVAE((TINY["T"], 1), **{**{k: (list(v) if isinstance(v, tuple) else v) for k, v in TINY.items() if k != "T"}, "levels": 1,
                                     "down_depth": [3], "strides": [2]})
    with pytest.raises(ValueError):
        m3.load_weights(str(tmp_path / "ck"))


def test_step_vector_is_one_lincomb_call(cpu_backend):
    """The metric increments of a captured step (`VQVAE._step_vector`: total / per-kind / per-level losses in `_all_trackers()`
    order + the VQ layers' usage / entropy metrics) come from ONE vqb_lincomb call and equal the sums formed term by term."""
    V = cpu_backend
    m, spec, weights, vq, x = build_tiny(V)
    m.compile(optimizer=V.keras.optimizers.Adam())
    grads, tvars, losses = m._forward_backward(V.keras_compat.convert_to_tensor(x))
    n0 = V._lib.launches()
    vec = m._step_vector(losses)
    assert V._lib.launches() - n0 == 1
    level, recon, commit, spectral = losses
    L = len(level)
    want = [sum(float(v) for v in level), sum(float(v) for v in recon), sum(float(v) for v in commit), sum(float(v) for v in spectral)]
    want += [float(v) for v in (*level, *recon, *commit, *spectral)]
    got = vec.cpu().numpy()
    assert got.shape == (4 + 4 * L + 3 * L,)
    np.testing.assert_allclose(got[:4 + 4 * L], np.array(want, dtype=np.float32), rtol=1e-6)


def test_in_graph_collective_is_opt_in_and_nccl_only(cpu_backend, monkeypatch):
    """dist.graph_comm(): None unless VQB_DP_INGRAPH=1 AND torch.distributed runs on NCCL (the CPU tests' gloo group never gets one)."""
    V = cpu_backend
    monkeypatch.setattr(V.dist, "_graph_comm", None)
    monkeypatch.delenv("VQB_DP_INGRAPH", raising=False)
    assert V.dist.graph_comm() is None
    monkeypatch.setattr(V.dist, "_graph_comm", None)
    monkeypatch.setenv("VQB_DP_INGRAPH", "1")
    assert V.dist.graph_comm() is None   # no process group at all in this test
    monkeypatch.setattr(V.dist, "_graph_comm", None)
