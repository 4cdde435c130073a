"""vae_play_b200 -- B200-native (sm_100a) implementation of the kungyao/vae-play VAE training step.

Python host code (this package) mirrors the reference's nn.Module API; all arithmetic runs in
``lib/libvaeplay_b200.so`` (hand-written CUDA, C ABI in ``include/vaeplay_b200.h``).
There is no CPU path: tensors must live on a CUDA device and the library must be present.
"""
from . import _lib  # noqa: F401
from .functional import (act_dtype, bce_dice_loss, get_precision, l1_loss, mse_loss, philox_normal, reparam_kl,  # noqa: F401
                         set_engine, set_epilogue_stats, set_precision, vae_loss)

__version__ = "0.1.0"
