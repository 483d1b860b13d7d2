"""vqvae_b200 — B200-native (sm_100a) implementation of the VQ-VAE audio hot path of
sunzeyucmu/VAE-based-Music--Deep-Generative-Models behind the reference's own Keras-style layer/model API.

    from vqvae_b200 import VQVAE, VectorQuantizer, Encoder, Decoder, keras
    model = VQVAE((28160, 1), levels=2, latent_dim=64, num_embeddings=512, down_depth=[5, 3], strides=[2, 2],
                  dilation_factor=3, residual_width=32)
    model.compile(optimizer=keras.optimizers.Adam())
    model.fit(x, y, batch_size=8, epochs=4)

Module names mirror the reference files: VectorQuantizer.py, resnet.py, encdec.py, vqvae.py, data_utils.py."""
from . import _lib, dist, keras_compat, ops
from . import keras_compat as keras
from .VectorQuantizer import VectorQuantizer
from .encdec import Decoder, DecoderConvBlock, Encoder, EncoderConvBlock, print_dec_layer
from .keras_compat import GradientTape, set_seed
from .resnet import DilatedResnet1D, ResnetConv1DBlock
from .vqvae import VQVAE, get_vqvae

SMALL_VQ_VAE = dict(levels=2, latent_dim=64, num_embeddings=512, down_depth=[5, 3], strides=[2, 2],
                    dilation_factor=3, residual_width=32)  # vqvae.py:352-353

__all__ = ["VQVAE", "get_vqvae", "VectorQuantizer", "Encoder", "Decoder", "EncoderConvBlock", "DecoderConvBlock",
           "DilatedResnet1D", "ResnetConv1DBlock", "print_dec_layer", "keras", "GradientTape", "set_seed",
           "SMALL_VQ_VAE", "dist", "ops"]
