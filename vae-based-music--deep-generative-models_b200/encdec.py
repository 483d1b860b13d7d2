"""encdec — drop-in for the reference's `encdec.py`: `EncoderConvBlock` (:17-41), `DecoderConvBlock` (:44-71),
`Encoder` (:74-108), `Decoder` (:114-151), `print_dec_layer` (:7-14).  Same constructors, same `.model` Sequentials,
same layer creation (= variable) order."""
from __future__ import annotations

from .keras_compat import Sequential, _is_symbolic, decoder_tail_fusable, decoder_tail_op, layers
from .resnet import DilatedResnet1D


def print_dec_layer(decoder):
    for dec_conv in decoder.model.layers[:-1]:
        print("-----{}-----".format(dec_conv.name))
        for l in dec_conv.model.layers[1::2]:  # take only dilated layers
            for layer in l.model.layers:
                print("---------{}---------".format(layer.name))
                for layer_ in layer.model.layers:
                    print(layer_.name)


class EncoderConvBlock(layers.Layer):
    """@embed_width: width of the down-sampling and residual stacks; @down_depth: number of down-sampling layers."""

    def __init__(self, output_dim, embed_width, embed_depth, dilation_factor=1, stride=2, down_depth=4, **kwargs):
        super(EncoderConvBlock, self).__init__(**kwargs)
        self.model = Sequential()
        self.kernel_size = stride * 2
        for i in range(down_depth):
            self.model.add(layers.Conv1D(embed_width, self.kernel_size, strides=stride, padding="same"))
            self.model.add(DilatedResnet1D(embed_width, embed_depth, dilation_factor=dilation_factor))
        self.model.add(layers.Conv1D(output_dim, 3, strides=1, padding="same"))

    def call(self, inputs, **kwargs):
        return self.model(inputs)


class DecoderConvBlock(layers.Layer):
    """@reverse_dilation: normally true for decoder blocks; @dilation_cycle: cyclic dilation (conditioner use)."""

    def __init__(self, output_dim, embed_width, embed_depth, dilation_factor=1, reverse_dilation=True,
                 dilation_cycle=None, stride=2, down_depth=4, **kwargs):
        super(DecoderConvBlock, self).__init__(**kwargs)
        self.model = Sequential()
        self.kernel_size = stride * 2
        self.model.add(layers.Conv1D(embed_width, 3, strides=1, padding="same"))
        for i in range(down_depth):
            self.model.add(DilatedResnet1D(embed_width, embed_depth, dilation_factor=dilation_factor,
                                           reverse_dilation=reverse_dilation, dilation_cycle=dilation_cycle))
            # remap to output_dim on the last up-sampling layer (encdec.py:66-68)
            self.model.add(layers.Conv1DTranspose(output_dim if i == (down_depth - 1) else embed_width,
                                                  self.kernel_size, strides=stride, padding="same"))

    def call(self, inputs, **kwargs):
        return self.model(inputs)


class Encoder(layers.Layer):
    def __init__(self, output_dim, residual_width, residual_depth, depth, down_depth, strides, dilation_factor=1,
                 **kwargs):
        super(Encoder, self).__init__(**kwargs)
        assert depth == len(down_depth), f"Depth {depth} not Legit"
        assert depth == len(strides), f"Depth {depth} not Legit"
        self.depth = depth
        self.down_depth = down_depth
        self.strides = strides
        self.model = Sequential()
        for layer, down_sampling_depth, stride in zip(list(range(self.depth)), down_depth, strides):
            self.model.add(EncoderConvBlock(output_dim, residual_width, residual_depth, stride=stride,
                                            dilation_factor=dilation_factor, down_depth=down_sampling_depth))

    def call(self, inputs, **kwargs):
        return self.model(inputs)


class Decoder(layers.Layer):
    """Mirrors the encoder while up-sampling (Conv1DTranspose); blocks are added in REVERSED order (encdec.py:142-145)."""

    def __init__(self, output_dim, embed_width, residual_width, residual_depth, depth, down_depth, strides,
                 dilation_factor=1, reverse_dilation=True, **kwargs):
        super(Decoder, self).__init__(**kwargs)
        assert depth == len(down_depth), f"Depth {depth} not Legit"
        assert depth == len(strides), f"Depth {depth} not Legit"
        self.depth = depth
        self.down_depth = down_depth
        self.strides = strides
        self.embed_width = embed_width
        self.model = Sequential()
        for layer, up_sampling_depth, stride in reversed(list(zip(list(range(self.depth)), down_depth, strides))):
            self.model.add(DecoderConvBlock(embed_width, residual_width, residual_depth, stride=stride,
                                            dilation_factor=dilation_factor, reverse_dilation=reverse_dilation,
                                            down_depth=up_sampling_depth))
        self.model.add(layers.Conv1D(output_dim, 3, strides=1, padding="same"))  # encdec.py:148

    fuse_tail = True  # run the last Conv1DTranspose + the final Conv1D as one composed operator when the shapes allow

    def call(self, inputs, **kwargs):
        seq = self.model.layers
        if self.fuse_tail and not _is_symbolic(inputs) and len(seq) >= 2 and isinstance(seq[-2], DecoderConvBlock):
            last_block = seq[-2].model.layers
            if decoder_tail_fusable(last_block[-1], seq[-1]):
                x = inputs
                for l in seq[:-2]:
                    x = l(x)
                for l in last_block[:-1]:
                    x = l(x)
                return decoder_tail_op(x, last_block[-1], seq[-1])
        return self.model(inputs)
