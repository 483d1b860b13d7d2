"""resnet — drop-in for the reference's `resnet.py`: `ResnetConv1DBlock` (:7-29) and `DilatedResnet1D` (:40-59).

`ResnetConv1DBlock.model` is still the `Sequential([ReLU, Conv1D(dil), ReLU, Conv1D])` the reference builds (callers
walk it: encdec.py:7-14, src/conditioner/conditioners.py:108-118), but `call` runs the whole pre-activation block
as ONE fused operation (`vqb_resblock_fwd`), and its backward as `vqb_resblock_bwd_data` + `vqb_resblock_wgrad`.
"""
from __future__ import annotations

from . import _lib, ops
from .keras_compat import KerasTensor, Sequential, _is_symbolic, grad_buffer, layers, record, write_grad


class ResnetConv1DBlock(layers.Layer):
    def __init__(self, input_dim, filters, dilation=1, **kwargs):
        super(ResnetConv1DBlock, self).__init__(**kwargs)
        self.input_dim, self.filters, self.dilation = input_dim, filters, dilation
        self.model = Sequential([
            layers.ReLU(),
            layers.Conv1D(filters, 3, dilation_rate=dilation, padding="same",
                          name="dilated_cov1d_dr-{}".format(dilation)),  # resnet.py:13-15 (name kept, typo included)
            layers.ReLU(),
            layers.Conv1D(input_dim, 3, dilation_rate=1, padding="same"),
        ])
        self.precision = _lib.PREC_FP32

    def call(self, input_tensor, **kwargs):
        x = input_tensor
        if _is_symbolic(x):
            y = self.model(x)  # builds the four sub-layers
            return KerasTensor(x.shape)
        conv1, conv2 = self.model.layers[1], self.model.layers[3]
        if not conv1.built:
            conv1.build(tuple(x.shape)); conv1.built = True
            conv2.build(tuple(x.shape[:-1]) + (self.filters,)); conv2.built = True
        if x.shape[-1] != self.input_dim:
            raise ValueError(f"{self.name}: input has {x.shape[-1]} channels, block was built for {self.input_dim}")
        x = x if x.is_contiguous() else x.contiguous()
        d = self.dilation
        prec = ops.resblock_precision(self.input_dim, self.filters, d, self.precision)
        # y = x + conv2(relu(conv1(relu(x))))   (resnet.py:11-18,29)
        from .keras_compat import GradientTape
        taping = GradientTape.current() is not None
        xbits = hbits = None
        if prec != _lib.PREC_FP32 and taping:
            # tensor-core path under a tape: the forward also emits the sign masks of x and h (8 bytes per position), which is
            # all the data gradient needs of them (x and h themselves stay the operands of the weight gradients)
            y, h, xbits, hbits = ops.resblock_fwd_masks(x, conv1.kernel.value, conv1.bias.value, conv2.kernel.value,
                                                        conv2.bias.value, d, prec)
        else:  # no tape (inference): nobody needs h, the tensor-core kernel does not store it
            y, h = ops.resblock_fwd(x, conv1.kernel.value, conv1.bias.value, conv2.kernel.value, conv2.bias.value, d, prec,
                                    want_h=taping)

        def bwd(g, needs):
            dy = g[0].contiguous()
            if xbits is not None:
                dx, dh = ops.resblock_bwd_data_masks(xbits, hbits, dy, conv1.kernel.value, conv2.kernel.value, d, prec)
            else:
                dx, dh = ops.resblock_bwd_data(x, h, dy, conv1.kernel.value, conv2.kernel.value, d, prec)
            self._weight_gradients(x, h, dy, dh, prec)
            return [dx if needs[0] else None]

        record([input_tensor], [y], bwd)
        return y

    def _convs(self, in_shape=None):
        conv1, conv2 = self.model.layers[1], self.model.layers[3]
        if not conv1.built and in_shape is not None:
            conv1.build(tuple(in_shape)); conv1.built = True
            conv2.build(tuple(in_shape[:-1]) + (self.filters,)); conv2.built = True
        return conv1, conv2

    def _weight_gradients(self, x, h, dy, dh, prec):
        """both weight gradients in one call (one launch on the tensor-core paths; ops.resblock_wgrad may hold it back to launch
        the blocks of a stack together): tape.gradient wrt the four variables"""
        conv1, conv2 = self._convs()
        d = self.dilation
        if getattr(conv1.kernel, "_grad_written", False) or getattr(conv2.kernel, "_grad_written", False):
            ops.wg_flush()  # a block used twice under one tape accumulates: needs its gradient now (write_grad)
            wg = ops._resblock_wgrad_now
        else:
            wg = ops.resblock_wgrad
        write_grad(conv1.kernel, lambda buf1: write_grad(conv2.kernel, lambda buf2: wg(
            x, h, dy, dh, buf1, grad_buffer(conv1.bias), buf2, grad_buffer(conv2.bias), d, prec)))
        conv1.bias._grad_written = True
        conv2.bias._grad_written = True


class DilatedResnet1D(layers.Layer):
    def __init__(self, input_dim, depth, dilation_factor=1, reverse_dilation=False, dilation_cycle=None, **kwargs):
        super(DilatedResnet1D, self).__init__(**kwargs)

        def _get_dilation(cur_depth):
            if dilation_cycle is None:
                return dilation_factor ** cur_depth
            return dilation_factor ** (cur_depth % dilation_cycle)  # cyclic dilation (resnet.py:44-48)

        blocks = [ResnetConv1DBlock(input_dim, input_dim, dilation=_get_dilation(d)) for d in range(depth)]
        if reverse_dilation:  # decoder stacks contract the dilation down to 1 (resnet.py:54-55)
            blocks = blocks[::-1]
        self.model = Sequential(blocks)

        self.use_fused_stack = True  # one launch per <= 4 blocks where libvqvae_b200 has the fused kernel (vqb_resstack_*)

    def call(self, input, **kwargs):
        x = input
        blocks = self.model.layers
        if _is_symbolic(x) or not self.use_fused_stack or not blocks:
            return self.model(x)
        prec = blocks[0].precision
        C_ = blocks[0].input_dim
        if (x.shape[-1] != C_ or any(b.precision != prec or b.input_dim != C_ or b.filters != C_ for b in blocks)
                or not ops.resstack_supported(C_, [b.dilation for b in blocks[:ops.RESSTACK_MAX]], prec)):
            return self.model(x)
        # resnet.py:51-59 as fused launches: chains of up to RESSTACK_MAX blocks, the activation staying on chip inside a chain
        for i in range(0, len(blocks), ops.RESSTACK_MAX):
            chunk = blocks[i:i + ops.RESSTACK_MAX]
            if not ops.resstack_supported(C_, [b.dilation for b in chunk], prec):
                for b in chunk:
                    x = b(x)
                continue
            x = self._fused_chain(chunk, x, prec)
        return x

    @staticmethod
    def _fused_chain(chunk, x_in, prec):
        from .keras_compat import GradientTape
        x = x_in if x_in.is_contiguous() else x_in.contiguous()
        convs = [b._convs(x.shape) for b in chunk]
        w1 = [c1.kernel.value for c1, _ in convs]; b1 = [c1.bias.value for c1, _ in convs]
        w2 = [c2.kernel.value for _, c2 in convs]; b2 = [c2.bias.value for _, c2 in convs]
        dils = [b.dilation for b in chunk]
        taping = GradientTape.current() is not None
        # the chain's own workspace (packed operand images of its weights), kept on its first block: memory nothing else uses, so
        # the packing launch may overlap the previous kernel (vqb_resstack_fwd_private_ws)
        own = getattr(chunk[0], "_stack_ws", None)
        if own is None or own.device != x.device or getattr(chunk[0], "_stack_ws_key", None) != (tuple(dils), prec):
            own = ops.resstack_workspace(x.shape[-1], dils, prec)
            chunk[0]._stack_ws, chunk[0]._stack_ws_key = own, (tuple(dils), prec)
        ys, hs, xbits, hbits, ws = ops.resstack_fwd(x, w1, b1, w2, b2, dils, prec, train=taping, ws=own)
        y = ys[-1]
        if not taping:
            return y

        def bwd(g, needs):
            dy = g[0].contiguous()
            dxs, dhs = ops.resstack_bwd_data(dy, w1, w2, xbits, hbits, dils, prec, fwd_ws=ws)  # images packed by the forward
            n = len(chunk)
            for i in reversed(range(n)):  # operands of the weight gradients: block input, h, gradient at its output, dh
                chunk[i]._weight_gradients(x if i == 0 else ys[i - 1], hs[i], dy if i == n - 1 else dxs[i + 1], dhs[i], prec)
            return [dxs[0] if needs[0] else None]

        record([x_in], [y], bwd)
        return y
