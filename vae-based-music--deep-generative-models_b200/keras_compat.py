"""The slice of the Keras-2.7 surface the reference's VQ-VAE path is written against (`keras.layers.Layer`,
`Sequential`, functional `keras.Model`, `keras.Input`, `layers.Conv1D/Conv1DTranspose/ReLU/add`,
`keras.metrics.Mean`, `keras.optimizers.Adam`, `keras.losses.MeanSquaredError`, `tf.GradientTape`), re-hosted on
device buffers whose arithmetic is done by libvqvae_b200.so.  Tensors are `torch.Tensor`s on the GPU used purely as
typed device memory; no torch op computes anything on the hot path.

Keras behaviours reproduced on purpose (they decide what the reference computes):
  * `training` resolution of `Layer.__call__` (TF 2.7 `base_layer._set_training_mode`): explicit value, else the
    enclosing layer call's value, else the `call` signature default, else False.
  * `layer.losses` holds the `add_loss` values of the last top-level call and is cleared when it starts.
  * auto-generated layer names (`conv1d_3`, `resnet_conv1d_block_7`, `dilated_resnet1d`, `re_lu`, ...), layer
    creation order = `trainable_variables` order (kernel, bias per conv).
  * glorot-uniform kernels / zero biases.
"""
from __future__ import annotations

import inspect
import math
import re
import time
from collections import defaultdict
from typing import List, Optional, Sequence

import re as _re

import numpy as np
import torch

from . import _lib, ops

# ------------------------------------------------------------------------------------------------ utils
_rng = np.random.Generator(np.random.PCG64(0))


def set_seed(seed: int):
    """Seeds the initialiser stream (the analogue of tf.random.set_seed for this package)."""
    global _rng
    _rng = np.random.Generator(np.random.PCG64(seed))


def rng() -> np.random.Generator:
    return _rng


def convert_to_tensor(x, dtype=torch.float32):
    """numpy / list / torch (any device) -> contiguous tensor on the library's device."""
    if isinstance(x, Variable):
        return x.value
    dev = _lib.device()
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(x))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if t.device != dev:
        t = t.to(dev, non_blocking=True)
    return t.contiguous()


_name_counts = defaultdict(int)


def _snake(name: str) -> str:
    s = re.sub("(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    return re.sub("([a-z])([A-Z])", r"\1_\2", s).lower()


def _unique_name(base: str) -> str:
    n = _name_counts[base]
    _name_counts[base] += 1
    return base if n == 0 else f"{base}_{n}"


def reset_name_counters():
    _name_counts.clear()


# --------------------------------------------------------------------------------------------- Variable
class Variable:
    """tf.Variable stand-in: a named device buffer (+ a gradient buffer once packed or differentiated)."""

    def __init__(self, initial_value, trainable=True, name=None, dtype=torch.float32):
        if isinstance(initial_value, Variable):
            initial_value = initial_value.value.clone()
        self.value = convert_to_tensor(initial_value, dtype).clone() if not isinstance(initial_value, torch.Tensor) \
            else convert_to_tensor(initial_value, dtype)
        self.trainable = trainable
        self.name = name or "Variable"
        self.grad: Optional[torch.Tensor] = None

    @property
    def shape(self):
        return tuple(self.value.shape)

    @property
    def dtype(self):
        return self.value.dtype

    def numpy(self):
        return self.value.detach().cpu().numpy()

    def assign(self, v):
        self.value.copy_(convert_to_tensor(v, self.value.dtype).reshape(self.value.shape))
        return self

    def _rebind(self, view: torch.Tensor, grad_view: Optional[torch.Tensor] = None):
        view.copy_(self.value)
        self.value = view
        self.grad = grad_view

    def __repr__(self):
        return f"<Variable {self.name} shape={self.shape} trainable={self.trainable}>"


# ----------------------------------------------------------------------------------- scalars and the tape
class Scalar:
    """A scalar loss living on the device: a linear combination of leaf 0-d tensors produced by loss kernels.
    Supports the arithmetic train_step does on losses (`a + b + c`, `sum(list)`, `total += x`)."""

    __slots__ = ("terms", "_cache")

    def __init__(self, terms=None):
        self.terms = list(terms or [])  # [(tensor[1], coeff)]
        self._cache = None

    @staticmethod
    def leaf(t):
        return Scalar([(t, 1.0)])

    def tensor(self):
        if self._cache is None:
            if not self.terms:
                self._cache = ops.zeros(1)
            else:
                acc = None
                for t, c in self.terms:
                    v = t if c == 1.0 else t * c
                    acc = v if acc is None else acc + v
                self._cache = acc
        return self._cache

    def __add__(self, o):
        if isinstance(o, Scalar):
            return Scalar(self.terms + o.terms)
        if o == 0:
            return Scalar(self.terms)
        return NotImplemented

    __radd__ = __add__

    def __mul__(self, c):
        return Scalar([(t, k * float(c)) for t, k in self.terms])

    __rmul__ = __mul__

    def __float__(self):
        return float(self.tensor().item())

    def numpy(self):
        return np.float32(float(self))

    def __repr__(self):
        return f"Scalar({float(self):.6g})"


class _Node:
    __slots__ = ("inputs", "outputs", "bwd", "stream")

    def __init__(self, inputs, outputs, bwd, stream=None):
        self.inputs, self.outputs, self.bwd, self.stream = inputs, outputs, bwd, stream


_node_stream = None  # side CUDA stream the ops being recorded run on (None = the caller's stream)


class on_stream:
    """Runs a self-contained part of the computation (one level's encoder -> VQ -> decoder: nothing but the input batch and
    read-only parameters is shared with the rest) on a side CUDA stream, and tags the tape nodes recorded inside so that
    `GradientTape.gradient` replays their backward on the same stream.  The caller forks the stream (`stream.wait_stream`)
    before and joins it (`current.wait_stream(stream)`) after; works eagerly and under CUDA-graph capture alike.
    A `None` stream makes it a no-op."""

    def __init__(self, stream):
        self.stream = stream

    def __enter__(self):
        global _node_stream
        self._prev = _node_stream
        if self.stream is not None:
            _node_stream = self.stream
            self._ctx = torch.cuda.stream(self.stream)
            self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        global _node_stream
        if self.stream is not None:
            self._ctx.__exit__(*exc)
        _node_stream = self._prev
        return False


class GradientTape:
    """tf.GradientTape stand-in (vqvae.py:119,143): records the layer-level ops executed inside the context and
    replays their hand-written backward kernels in reverse order."""

    _stack: List["GradientTape"] = []

    def __init__(self, persistent=False):
        self.nodes: List[_Node] = []

    def __enter__(self):
        GradientTape._stack.append(self)
        return self

    def __exit__(self, *exc):
        GradientTape._stack.pop()
        return False

    @staticmethod
    def current() -> Optional["GradientTape"]:
        return GradientTape._stack[-1] if GradientTape._stack else None

    def gradient(self, target, sources):
        if not isinstance(target, Scalar):
            raise TypeError("GradientTape.gradient: target must be a scalar loss produced by this package")
        grads = {}
        for t, c in target.terms:
            grads[id(t)] = grads.get(id(t), 0.0) + c
        produced = {id(o) for n in self.nodes for o in n.outputs}
        for v in sources:
            v._grad_written = False
        ops.reduce_begin()  # the ~480 per-variable partial-sum reductions of one backward pass are flushed together
        try:
            self._backprop(grads, produced)
        finally:
            ops.reduce_flush()
        return [v.grad if getattr(v, "_grad_written", False) else None for v in sources]

    def _backprop(self, grads, produced):
        """Reverse sweep.  Nodes tagged with a side stream (`on_stream`) replay there.  A gradient tensor that crosses streams
        makes the consumer's stream wait for the producer's and is `record_stream`ed, so that the caching allocator does not
        hand its block to the producer's stream again while the consumer still reads it (also under graph capture, where the
        reuse would be baked into the graph); side streams are joined into the caller's stream at the end."""
        cuda = torch.cuda.is_available() and any(n.stream is not None for n in self.nodes)
        main = torch.cuda.current_stream() if cuda else None
        prod = {}     # id(gradient tensor) -> stream it was produced on
        used = []
        last_ns = main

        def sync_in(t, ns):
            if cuda and isinstance(t, torch.Tensor):
                ps = prod.get(id(t), main)
                if ps is not ns:
                    ns.wait_stream(ps)
                    t.record_stream(ns)

        try:
            for node in reversed(self.nodes):
                gouts = [grads.pop(id(o), None) for o in node.outputs]
                if all(g is None for g in gouts):
                    continue
                needs = [(not isinstance(i, Variable)) and id(i) in produced for i in node.inputs]
                ns = node.stream if node.stream is not None else main
                if ns is not last_ns:
                    ops.wg_flush()  # leftovers of the previous stream's queue (launched on that stream)
                    last_ns = ns
                if node.stream is not None and node.stream not in used:
                    node.stream.wait_stream(main)  # everything enqueued so far (the loss heads) precedes this level's sweep
                    used.append(node.stream)
                for g in gouts:
                    sync_in(g, ns)
                with on_stream(node.stream):
                    gins = node.bwd(gouts, needs)
                    for inp, g in zip(node.inputs, gins):
                        if g is None:
                            continue
                        k = id(inp)
                        if k in grads:
                            sync_in(grads[k], ns)
                            g = grads[k] + g
                        grads[k] = g
                        if cuda and isinstance(g, torch.Tensor):
                            prod[id(g)] = ns
        finally:
            ops.wg_flush()  # queued block weight gradients go out on their own stream BEFORE it is joined
            for s in used:
                main.wait_stream(s)


def record(inputs, outputs, bwd):
    t = GradientTape.current()
    if t is not None:
        t.nodes.append(_Node(list(inputs), list(outputs), bwd, _node_stream))


def grad_buffer(v: Variable) -> torch.Tensor:
    """Where a weight-gradient kernel writes: the variable's slice of the packed gradient buffer (or a private
    buffer for an unpacked variable)."""
    if v.grad is None:
        v.grad = torch.empty_like(v.value)
    return v.grad


def write_grad(v: Variable, fn):
    """fn(buffer) must overwrite `buffer` with the gradient; a second contribution in one pass is accumulated."""
    if getattr(v, "_grad_written", False):
        tmp = torch.empty_like(v.value)
        fn(tmp)
        if ops._deferred is not None:  # queued reductions must land before the sum is formed
            ops.reduce_flush()
            ops.reduce_begin()
        v.grad += tmp
    else:
        fn(grad_buffer(v))
        v._grad_written = True


# ---------------------------------------------------------------------------------------- symbolic input
class KerasTensor:
    """Shape-only placeholder produced by `keras.Input`; calling layers on it builds them."""

    def __init__(self, shape, source=None, index=0):
        self.shape = tuple(shape)
        self._source = source  # (layer, input KerasTensor)
        self._index = index


def Input(shape, batch_size=None, name=None):
    return KerasTensor((batch_size,) + tuple(shape))


def _is_symbolic(x):
    return isinstance(x, KerasTensor)


# ------------------------------------------------------------------------------------------------ Layer
class _CallContext:
    depth = 0
    training = None


class Layer:
    def __init__(self, name=None, trainable=True, dtype=None, **kwargs):
        if kwargs:
            raise TypeError(f"{type(self).__name__}: unexpected keyword arguments {sorted(kwargs)}")
        object.__setattr__(self, "_sublayers", [])
        object.__setattr__(self, "_own_vars", [])
        self.name = name if name is not None else _unique_name(_snake(type(self).__name__))
        self.trainable = trainable
        self.built = False
        self._losses = []
        sig = inspect.signature(self.call)
        self._expects_training = "training" in sig.parameters
        d = sig.parameters["training"].default if self._expects_training else None
        self._default_training = None if d is inspect.Parameter.empty else d

    # -- attribute tracking (creation order defines variable order, as in Keras)
    def __setattr__(self, k, v):
        if isinstance(v, Layer):
            if all(v is not s for s in self._sublayers):
                self._sublayers.append(v)
        elif isinstance(v, Variable):
            if all(v is not s for s in self._own_vars):
                self._own_vars.append(v)
        elif isinstance(v, (list, tuple)) and v and all(isinstance(e, Layer) for e in v):
            for e in v:
                if all(e is not s for s in self._sublayers):
                    self._sublayers.append(e)
        object.__setattr__(self, k, v)

    def add_weight(self, name, shape, initializer="glorot_uniform", trainable=True):
        shape = tuple(int(s) for s in shape)
        if initializer == "zeros":
            a = np.zeros(shape, np.float32)
        elif initializer == "ones":
            a = np.ones(shape, np.float32)
        elif initializer == "glorot_uniform":
            rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
            lim = math.sqrt(6.0 / (shape[-2] * rf + shape[-1] * rf))
            a = _rng.uniform(-lim, lim, size=shape).astype(np.float32)
        elif callable(initializer):
            a = np.asarray(initializer(shape), np.float32)
        else:
            raise ValueError(f"unknown initializer {initializer!r}")
        v = Variable(a, trainable=trainable, name=f"{self.name}/{name}:0")
        self._own_vars.append(v)
        return v

    # -- variable / loss collections
    def _flatten_layers(self):
        for s in self._sublayers:
            yield s
            yield from s._flatten_layers()

    @property
    def layers(self):
        return list(self._sublayers)

    @property
    def variables(self):
        out = list(self._own_vars)
        for s in self._sublayers:
            out += s.variables
        seen, uniq = set(), []
        for v in out:
            if id(v) not in seen:
                seen.add(id(v))
                uniq.append(v)
        return uniq

    weights = variables

    @property
    def trainable_variables(self):
        return [v for v in self.variables if v.trainable] if self.trainable else []

    trainable_weights = trainable_variables

    @property
    def non_trainable_variables(self):
        return [v for v in self.variables if not (v.trainable and self.trainable)]

    def get_weights(self):
        return [v.numpy() for v in self.variables]

    def set_weights(self, weights):
        vs = self.variables
        if len(weights) != len(vs):
            raise ValueError(f"{self.name}.set_weights: expected {len(vs)} arrays, got {len(weights)}")
        for v, w in zip(vs, weights):
            if tuple(np.shape(w)) != v.shape:
                raise ValueError(f"{self.name}.set_weights: {v.name} has shape {v.shape}, got {np.shape(w)}")
            v.assign(w)

    def count_params(self):
        return int(sum(np.prod(v.shape) for v in self.variables))

    def add_loss(self, loss):
        self._losses.append(loss)

    @property
    def losses(self):
        out = list(self._losses)
        for s in self._sublayers:
            out += s.losses
        return out

    def _clear_losses(self):
        self._losses = []
        for s in self._sublayers:
            s._clear_losses()

    @property
    def metrics(self):
        return []

    # -- building and calling
    def build(self, input_shape):
        self.built = True

    def call(self, inputs, **kwargs):
        return inputs

    def __call__(self, *args, **kwargs):
        inputs = args[0] if args else kwargs.get("inputs")
        # training-mode resolution, TF 2.7 base_layer._set_training_mode
        training = kwargs.get("training", None)
        if self._expects_training:
            if training is None:
                if _CallContext.training is not None:
                    training = _CallContext.training
                else:
                    training = self._default_training if self._default_training is not None else False
            kwargs["training"] = training
        else:
            kwargs.pop("training", None)
            training = _CallContext.training
        top = _CallContext.depth == 0
        if top:
            self._clear_losses()
        if not self.built:
            shp = tuple(inputs.shape) if hasattr(inputs, "shape") else None
            self.build(shp)
            self.built = True
        prev = _CallContext.training
        _CallContext.depth += 1
        _CallContext.training = training
        try:
            if not _is_symbolic(inputs) and args and not isinstance(inputs, (torch.Tensor, tuple, list)):
                args = (convert_to_tensor(inputs),) + tuple(args[1:])
            out = self.call(*args, **kwargs)
        finally:
            _CallContext.depth -= 1
            _CallContext.training = prev
        if top and _is_symbolic(inputs):
            outs = out if isinstance(out, (tuple, list)) else (out,)
            for i, o in enumerate(outs):
                if isinstance(o, KerasTensor):
                    o._source, o._index = (self, inputs), i
        return out

    def summary(self, print_fn=print):
        print_fn(f'Layer "{self.name}" ({type(self).__name__}): {self.count_params():,} params')
        for s in self._sublayers:
            print_fn(f"  {s.name:40s} {type(s).__name__:24s} {s.count_params():>10,}")


class Sequential(Layer):
    def __init__(self, layers=None, name=None):
        super().__init__(name=name)
        self._seq: List[Layer] = []
        for l in layers or []:
            self.add(l)

    def add(self, layer):
        self._seq.append(layer)
        if all(layer is not s for s in self._sublayers):
            self._sublayers.append(layer)

    @property
    def layers(self):
        return list(self._seq)

    def call(self, inputs, training=None):
        x = inputs
        for l in self._seq:
            x = l(x)
        return x


# ------------------------------------------------------------------------------------- primitive layers
class ReLU(Layer):
    """layers.ReLU (resnet.py:12,16).  Inside ResnetConv1DBlock it is fused into the convolution that follows; this
    standalone form exists for direct use of the layer."""

    def call(self, x):
        if _is_symbolic(x):
            return KerasTensor(x.shape)
        y = torch.relu(x)
        record([x], [y], lambda g, needs: [g[0] * (x > 0)])
        return y


class Add(Layer):
    def call(self, xs):
        if _is_symbolic(xs[0]):
            return KerasTensor(xs[0].shape)
        y = xs[0]
        for t in xs[1:]:
            y = y + t
        # the tape keys on tensor identity, so register every summand as an input of this node
        ins = list(xs)
        record(ins, [y], lambda g, needs: [g[0] for _ in ins])
        return y


def add(xs):
    """layers.add (resnet.py:29)."""
    return Add()(xs)


def _same_only(padding, who):
    if str(padding).lower() != "same":
        raise NotImplementedError(f"{who}: only padding='same' is implemented (the reference uses nothing else)")


class Conv1D(Layer):
    """layers.Conv1D(filters, kernel_size, strides, padding='same', dilation_rate) — resnet.py:13,17;
    encdec.py:33,38,60,148.  kernel [k, Cin, Cout] glorot-uniform, bias zeros."""

    def __init__(self, filters, kernel_size, strides=1, padding="valid", dilation_rate=1, use_bias=True,
                 name=None, **kw):
        super().__init__(name=name, **kw)
        _same_only(padding, "Conv1D")
        self.filters, self.kernel_size = int(filters), int(kernel_size)
        self.strides, self.dilation_rate, self.use_bias = int(strides), int(dilation_rate), use_bias
        self.kernel = self.bias = None
        self.precision = _lib.PREC_FP32

    def build(self, input_shape):
        cin = int(input_shape[-1])
        self.kernel = self.add_weight("kernel", (self.kernel_size, cin, self.filters))
        self.bias = self.add_weight("bias", (self.filters,), "zeros") if self.use_bias else None
        self.built = True

    def call(self, x):
        if _is_symbolic(x):
            L = x.shape[1]
            return KerasTensor((x.shape[0], None if L is None else -(-L // self.strides), self.filters))
        return conv1d_op(x, self.kernel, self.bias, self.strides, self.dilation_rate, False, self.precision)


def conv1d_op(x, kernel: Variable, bias: Optional[Variable], stride, dilation, relu_in, precision=0):
    y = ops.conv1d_fwd(x, kernel.value, None if bias is None else bias.value, stride, dilation, relu_in,
                       None, precision)

    def bwd(g, needs):
        dy = g[0].contiguous()
        write_grad(kernel, lambda buf: ops.conv1d_wgrad(
            x, dy, buf, None if bias is None else grad_buffer(bias), stride, dilation, relu_in, precision))
        if bias is not None:
            bias._grad_written = True
        dx = ops.conv1d_dgrad(dy, kernel.value, x.shape, x if relu_in else None, stride, dilation, relu_in, None,
                              precision) if needs[0] else None
        return [dx]

    record([x], [y], bwd)
    return y


class Conv1DTranspose(Layer):
    """layers.Conv1DTranspose(filters, kernel_size, strides, padding='same') — encdec.py:67-68.
    kernel [k, Cout, Cin]."""

    def __init__(self, filters, kernel_size, strides=1, padding="valid", use_bias=True, name=None, **kw):
        super().__init__(name=name, **kw)
        _same_only(padding, "Conv1DTranspose")
        self.filters, self.kernel_size, self.strides, self.use_bias = int(filters), int(kernel_size), int(strides), use_bias
        self.kernel = self.bias = None
        self.precision = _lib.PREC_FP32

    def build(self, input_shape):
        cin = int(input_shape[-1])
        self.kernel = self.add_weight("kernel", (self.kernel_size, self.filters, cin))
        self.bias = self.add_weight("bias", (self.filters,), "zeros") if self.use_bias else None
        self.built = True

    def call(self, x):
        if _is_symbolic(x):
            L = x.shape[1]
            return KerasTensor((x.shape[0], None if L is None else L * self.strides, self.filters))
        kernel, bias, s, prec = self.kernel, self.bias, self.strides, self.precision
        y = ops.conv1d_transpose_fwd(x, kernel.value, None if bias is None else bias.value, s, prec)

        def bwd(g, needs):
            dy = g[0].contiguous()
            write_grad(kernel, lambda buf: ops.conv1d_transpose_wgrad(
                x, dy, buf, None if bias is None else grad_buffer(bias), s, prec))
            if bias is not None:
                bias._grad_written = True
            return [ops.conv1d_transpose_dgrad(dy, kernel.value, x.shape, s, prec) if needs[0] else None]

        record([x], [y], bwd)
        return y


def decoder_tail_fusable(up: "Conv1DTranspose", out: "Conv1D") -> bool:
    """The decoder's last Conv1DTranspose(k=4, s=2) followed by Conv1D(1, 3) (encdec.py:67-68,148) in a shape
    libvqvae_b200 runs as one composed linear operator."""
    return (isinstance(up, Conv1DTranspose) and isinstance(out, Conv1D) and up.built and out.built
            and up.kernel_size == 4 and up.strides == 2 and out.kernel_size == 3 and out.strides == 1
            and out.dilation_rate == 1 and out.filters == 1
            and ops.dec_tail_supported(up.kernel.shape[2], up.filters))


def decoder_tail_op(x, up: "Conv1DTranspose", out: "Conv1D"):
    """out(up(x)) without materialising up(x); the backward writes the gradients of both layers' variables."""
    wt, bt, wf, bf = up.kernel, up.bias, out.kernel, out.bias
    val = lambda v: None if v is None else v.value
    recon, gbuf = ops.dec_tail_fwd(x, wt.value, val(bt), wf.value, val(bf))

    def bwd(g, needs):
        dr = g[0].contiguous()
        if any(getattr(v, "_grad_written", False) for v in (wt, wf)):
            # a layer applied twice in one pass: write temporaries and accumulate (what write_grad does per variable)
            tmp = [torch.empty_like(v.value) if v is not None else None for v in (wt, bt, wf, bf)]
            dx = ops.dec_tail_bwd(x, dr, wt.value, val(bt), wf.value, gbuf, tmp[0], tmp[1], tmp[2], tmp[3], needs[0])
            for v, t in zip((wt, bt, wf, bf), tmp):
                if v is None:
                    continue
                if getattr(v, "_grad_written", False):
                    v.grad += t
                else:
                    grad_buffer(v).copy_(t)
                    v._grad_written = True
            return [dx]
        dx = ops.dec_tail_bwd(x, dr, wt.value, val(bt), wf.value, gbuf, grad_buffer(wt),
                              None if bt is None else grad_buffer(bt), grad_buffer(wf),
                              None if bf is None else grad_buffer(bf), needs[0])
        for v in (wt, bt, wf, bf):
            if v is not None:
                v._grad_written = True
        return [dx]

    record([x], [recon], bwd)
    return recon


# ------------------------------------------------------------------------------------------------ Model
class History:
    def __init__(self):
        self.history = defaultdict(list)
        self.epoch = []


class Model(Layer):
    """keras.Model: functional form `Model(inputs, outputs, name=)` (vqvae.py:21) or subclassed."""

    def __init__(self, inputs=None, outputs=None, name=None, **kw):
        super().__init__(name=name, **kw)
        self._fn_inputs, self._fn_outputs = inputs, outputs
        self.optimizer = None
        self.stop_training = False
        if inputs is not None:
            if not isinstance(inputs, KerasTensor) or not isinstance(outputs, KerasTensor):
                raise NotImplementedError("functional Model: single input / single output graphs only")
            self.built = True
            for l in self._graph_layers():
                if all(l is not s for s in self._sublayers):
                    self._sublayers.append(l)

    def _graph_layers(self):
        order = []

        def walk(kt):
            if kt is self._fn_inputs or kt._source is None:
                return
            layer, src = kt._source
            walk(src)
            if all(layer is not l for l in order):
                order.append(layer)

        walk(self._fn_outputs)
        return order

    def call(self, inputs, training=None, mask=None):
        if self._fn_inputs is None:
            raise NotImplementedError("subclassed Model must implement call()")
        memo = {}

        def ev(kt):
            if kt is self._fn_inputs:
                return inputs
            layer, src = kt._source
            key = id(src), id(layer)
            if key not in memo:
                memo[key] = layer(ev(src))
            out = memo[key]
            return out[kt._index] if isinstance(out, (tuple, list)) else out

        return ev(self._fn_outputs)

    # -- training loop plumbing (vqvae.py:362-363: compile(optimizer=Adam()), fit(x, y, batch_size, epochs))
    def compile(self, optimizer=None, **kw):
        self.optimizer = optimizer

    def reset_metrics(self):
        for m in self.metrics:
            m.reset_state()

    def train_step(self, data):
        raise NotImplementedError

    def test_step(self, data):
        raise NotImplementedError

    @staticmethod
    def _batches(x, y, batch_size):
        n = len(x)
        for i in range(0, n, batch_size):
            yield (x[i:i + batch_size], None if y is None else y[i:i + batch_size])

    def fit(self, x=None, y=None, batch_size=32, epochs=1, verbose=1, callbacks=None, validation_data=None,
            shuffle=True, **kw):
        hist = History()
        callbacks = callbacks or []
        for cb in callbacks:
            if hasattr(cb, "set_model"):
                cb.set_model(self)
        n = len(x)
        for epoch in range(epochs):
            self.reset_metrics()
            t0 = time.time()
            order = _rng.permutation(n) if shuffle else np.arange(n)
            xs = x[order] if shuffle else x
            ys = None if y is None else (y[order] if shuffle else y)
            logs = {}
            for cb in callbacks:
                if hasattr(cb, "on_epoch_begin"):
                    cb.on_epoch_begin(epoch)
            steps = 0
            for bx, by in self._batches(xs, ys, batch_size):
                logs = self.train_step((bx, by))
                steps += 1
            logs = {k: float(v) for k, v in logs.items()}
            if validation_data is not None:
                val = self.evaluate(validation_data[0], validation_data[1] if len(validation_data) > 1 else None,
                                    batch_size=batch_size, verbose=0, return_dict=True)
                logs.update({"val_" + k: v for k, v in val.items()})
            hist.epoch.append(epoch)
            for k, v in logs.items():
                hist.history[k].append(v)
            if verbose:
                dt = time.time() - t0
                body = " - ".join(f"{k}: {v:.4f}" for k, v in logs.items())
                print(f"Epoch {epoch + 1}/{epochs}\n{steps}/{steps} - {dt:.1f}s {1e3 * dt / max(steps, 1):.0f}ms/step - {body}")
            for cb in callbacks:
                if hasattr(cb, "on_epoch_end"):
                    cb.on_epoch_end(epoch, logs)
            if self.stop_training:
                break
        return hist

    def evaluate(self, x=None, y=None, batch_size=32, verbose=1, return_dict=False, **kw):
        self.reset_metrics()
        logs = {}
        for bx, by in self._batches(x, y, batch_size):
            logs = self.test_step((bx, by))
        logs = {k: float(v) for k, v in logs.items()}
        if verbose:
            print(" - ".join(f"{k}: {v:.4f}" for k, v in logs.items()))
        return logs if return_dict else list(logs.values())

    def save_weights(self, path):
        """All variables (Keras `variables` order: conv kernels / biases, then the VQ state embeddings / m_t / N_t per level) plus
        what a resumed run needs to continue bit for bit: the optimizer's Adam moments and step counter and the VQ layers'
        restart-RNG step.  The file is a flat .npz keyed "index|variable name" (+ "opt|..." / "state|..." entries) — NOT the
        tf.train.Checkpoint format of the reference's CheckpointManager (src/callback/vae_monitor.py:56-58)."""
        out = {f"{i:04d}|{v.name}": v.numpy() for i, v in enumerate(self.variables)}
        opt = getattr(self, "optimizer", None)
        if opt is not None and getattr(opt, "_iterations", None) is not None:
            out["opt|iterations"] = np.asarray(opt.iterations, np.int64)
            packed = getattr(self, "_packed", None)
            for key, (m, v) in opt._slots.items():
                if packed is not None and key == ("flat", packed.params.data_ptr()):
                    out["opt|flat|m"], out["opt|flat|v"] = m.cpu().numpy(), v.cpu().numpy()
        for i, st in enumerate(self._extra_state()):
            out[f"state|{i:03d}"] = st
        np.savez(path, **out)

    def _extra_state(self):
        """numpy arrays of non-variable state to checkpoint (overridden where a model has any)"""
        return []

    def _set_extra_state(self, arrays):
        pass

    def load_weights(self, path):
        data = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
        wkeys = sorted(k for k in data.files if k[:4].isdigit())
        vs = self.variables
        if len(wkeys) != len(vs):
            raise ValueError(f"{self.name}.load_weights: the file holds {len(wkeys)} variables, the model has {len(vs)}")
        for k, v in zip(wkeys, vs):  # names are validated: equal shapes are not enough to accept a file
            name = k.split("|", 1)[1]
            # Keras uniquifies auto-generated layer names per process ("conv1d_48"): compare without that counter
            if _re.sub(r"_\d+(?=/)", "", name) != _re.sub(r"_\d+(?=/)", "", v.name):
                raise ValueError(f"{self.name}.load_weights: variable {k[:4]} is '{name}' in the file but '{v.name}' in the model")
        self.set_weights([data[k] for k in wkeys])
        opt = getattr(self, "optimizer", None)
        packed = getattr(self, "_packed", None)
        if opt is not None and "opt|iterations" in data.files:
            opt._counter().fill_(int(data["opt|iterations"]))
            opt._host_iterations = int(data["opt|iterations"])
            if packed is not None and "opt|flat|m" in data.files:
                m, v = opt._slot(("flat", packed.params.data_ptr()), packed.params)
                m.copy_(torch.from_numpy(data["opt|flat|m"])); v.copy_(torch.from_numpy(data["opt|flat|v"]))
        st = [data[k] for k in sorted(k for k in data.files if k.startswith("state|"))]
        if st:
            self._set_extra_state(st)
        if hasattr(self, "_graphs"):
            self._graphs = {}


# ---------------------------------------------------------------------------------------------- metrics
class Mean:
    """keras.metrics.Mean: running mean of scalars, kept on the device (no host sync until `result()` is read).
    The running total is a 1-element device tensor updated in place; a model may bind the totals of all its trackers
    to slices of one vector (`_bind`) so that a whole step's metrics are accumulated by one vector add."""

    def __init__(self, name="mean", dtype=None):
        self.name = name
        self._total = None
        self._count = 0

    def _bind(self, view):
        if self._total is not None:
            view.copy_(self._total)
        else:
            view.zero_()
        self._total = view

    def update_state(self, value, sample_weight=None):
        t = value.tensor() if isinstance(value, Scalar) else (
            value if isinstance(value, torch.Tensor) else torch.as_tensor(float(value), device=_lib.device()))
        t = t.reshape(-1)[:1].to(torch.float32)
        if self._total is None:
            self._total = torch.zeros(1, dtype=torch.float32, device=_lib.device())
        self._total += t
        self._count += 1

    def result(self):
        return MetricValue(self._total, self._count)

    def reset_state(self):
        if self._total is not None:
            self._total.zero_()
        self._count = 0

    reset_states = reset_state


class MetricValue:
    """Lazy scalar: float() / numpy() read it back from the device."""

    __slots__ = ("_t", "_n")

    def __init__(self, t, n):
        self._t, self._n = t, n

    def __float__(self):
        return 0.0 if (self._t is None or self._n == 0) else float(self._t.item()) / self._n

    def numpy(self):
        return np.float32(float(self))

    def __format__(self, spec):
        return format(float(self), spec)

    def __repr__(self):
        return f"{float(self):.6g}"


# -------------------------------------------------------------------------------------------- optimizer
class Adam:
    """keras.optimizers.Adam (Keras 2.7, non-amsgrad) — vqvae.py:362.  One fused multi-tensor kernel when the
    variables are slices of one packed buffer, otherwise one kernel per variable."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, name="Adam"):
        # learning_rate: a number, or a callable step -> number (keras LearningRateSchedule, e.g. the reference's CustomSchedule,
        # src/transformer/multi_head_attention.py:82); may be reassigned at any time, also after train_step was captured
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon
        self._iterations = None
        self._host_iterations = 0  # mirror of the device counter (no synchronisation needed to evaluate a schedule)
        self._lr_dev = None        # the learning rate the kernels read: one fp32 in device memory
        self._lr_host = None       # the value last written to it
        self._slots = {}
        self.grad_scale = 1.0  # 1/world_size under data parallelism (gradients arrive summed)

    @property
    def iterations(self):
        return 0 if self._iterations is None else int(self._iterations.item())

    def _counter(self):
        if self._iterations is None:
            self._iterations = ops.zeros(1, dtype=torch.int64)
        return self._iterations

    def current_lr(self):
        """learning rate of the NEXT update (schedules are evaluated at the number of updates applied so far, as Keras does)"""
        lr = self.learning_rate
        return float(lr(self._host_iterations)) if callable(lr) else float(lr)

    def refresh_lr(self):
        """Writes current_lr() into the device scalar the Adam kernel reads.  Called before every update — eager ones and
        replays of a captured train_step alike (the copy itself is never part of a capture)."""
        if self._lr_dev is None:
            self._lr_dev = ops.zeros(1)
        lr = self.current_lr()
        if self._lr_host != lr:
            self._lr_dev.fill_(lr)
            self._lr_host = lr
        return self._lr_dev

    def step_done(self):
        """bookkeeping of one applied update whose kernels ran from a CUDA-graph replay"""
        self._host_iterations += 1

    def _slot(self, key, like):
        if key not in self._slots:
            self._slots[key] = (torch.zeros_like(like), torch.zeros_like(like))
        return self._slots[key]

    def _lr_for_launch(self):
        capturing = torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
        if capturing:  # the replay path refreshes the scalar before every replay
            if self._lr_dev is None:
                raise RuntimeError("Adam: refresh_lr() must run once before the update is captured into a CUDA graph")
            return self._lr_dev
        return self.refresh_lr()

    def apply_flat(self, params, grads):
        m, v = self._slot(("flat", params.data_ptr()), params)
        ops.adam_step_dev(params, grads, m, v, self._lr_for_launch(), self.beta_1, self.beta_2, self.epsilon,
                          self.grad_scale, self._counter())
        ops.increment(self._counter())
        if not (torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()):
            self._host_iterations += 1

    def apply_gradients(self, grads_and_vars):
        gv = [(g, v) for g, v in grads_and_vars if g is not None]
        if not gv:
            return
        # fast path: contiguous slices of one packed buffer, in order
        first_g, first_v = gv[0]
        packed = getattr(first_v, "_pack", None)
        if packed is not None and len(gv) == len(packed.vars) and all(
                v is pv and g is v.grad for (g, v), pv in zip(gv, packed.vars)):
            self.apply_flat(packed.params, packed.grads)
            return
        lr_dev = self._lr_for_launch()
        for g, v in gv:
            m, s = self._slot(id(v), v.value)
            ops.adam_step_dev(v.value, g.contiguous(), m, s, lr_dev, self.beta_1, self.beta_2, self.epsilon,
                              self.grad_scale, self._counter())
        ops.increment(self._counter())
        if not (torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()):
            self._host_iterations += 1


class Packed:
    """All trainable variables of a model as views of ONE parameter buffer and ONE gradient buffer (so that the
    gradient all-reduce and the Adam step are single launches)."""

    def __init__(self, variables: Sequence[Variable], extra_floats: int = 0):
        self.vars = list(variables)
        al = lambda k: (k + 3) & ~3  # every variable starts on a 16-byte boundary (vector loads in the kernels)
        n = sum(al(v.value.numel()) for v in self.vars)
        self.n_params = n
        self.params = ops.zeros(n)
        self.comm = ops.zeros(n + extra_floats)  # [grads | extra (EMA statistics, loss scalars)]
        self.grads = self.comm[:n]
        self.extra = self.comm[n:]
        off = 0
        for v in self.vars:
            k = v.value.numel()
            v._rebind(self.params[off:off + k].view(v.shape), self.grads[off:off + k].view(v.shape))
            v._pack = self
            off += al(k)


class MeanSquaredError:
    """keras.losses.MeanSquaredError (vqvae.py:91).  The call is lazy; `reduce_mean` of it launches the fused
    squared-error reduction (and records its gradient)."""

    def __init__(self, reduction="none", name=None):
        self.reduction = reduction

    def __call__(self, y_true, y_pred):
        return _LazyMSE(convert_to_tensor(y_true), y_pred)


class _LazyMSE:
    def __init__(self, x, r):
        self.x, self.r = x, r


def reduce_mean(v):
    """tf.reduce_mean over a lazy loss expression -> Scalar (vqvae.py:125,127)."""
    if isinstance(v, _LazyMSE):
        x, r = v.x, v.r
        loss, _ = ops.mse(x, r)

        def bwd(g, needs):
            _, dr = ops.mse(x, r, loss_scale=float(g[0]))
            return [dr]

        record([r], [loss], bwd)
        return Scalar.leaf(loss)
    if hasattr(v, "_reduce_mean"):
        return v._reduce_mean()
    raise TypeError(f"reduce_mean: unsupported operand {type(v).__name__}")


class _Namespace:
    def __init__(self, **kw):
        self.__dict__.update(kw)


layers = _Namespace(Layer=Layer, Conv1D=Conv1D, Conv1DTranspose=Conv1DTranspose, ReLU=ReLU, Add=Add, add=add)
metrics = _Namespace(Mean=Mean)
optimizers = _Namespace(Adam=Adam)
losses = _Namespace(MeanSquaredError=MeanSquaredError)
models = _Namespace(Model=Model, Sequential=Sequential)
