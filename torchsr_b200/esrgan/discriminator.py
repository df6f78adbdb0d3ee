"""ESRGAN discriminator - drop-in for torchsr/esrgan/discriminator.py (reference lines 17-95)."""
from torch import nn

from ..engine import B200Module, Plan
from .. import nets
from ..srgan.discriminator import _features

CONV_IDX = (0, 2, 5, 8, 11, 14, 17, 20, 23, 26)


class Discriminator(B200Module):
    """10 strided-conv stages (128x128 input -> 512 x 4 x 4), Linear 8192->100, LeakyReLU, Linear 100->1; returns
    logits (no sigmoid), as the relativistic loss of the reference trainer expects."""

    def __init__(self, image_size: int = 128) -> None:
        super().__init__()
        self.image_size = image_size
        feature_map_size = image_size // 32
        self.features = _features([(64, 2), (128, 1), (128, 2), (256, 1), (256, 2), (512, 1), (512, 2), (512, 1),
                                   (512, 2)])
        self.classifier = nn.Sequential(
            nn.Linear(512 * feature_map_size * feature_map_size, 100),
            nn.LeakyReLU(negative_slope=0.2, inplace=True),
            nn.Linear(100, 1))

    def _records(self):
        return nets.discriminator_records(self, CONV_IDX), nets.discriminator_linears(self, self.image_size)

    def _define(self, plan: Plan, shape):
        nets.define_discriminator(self, plan, shape, CONV_IDX, sigmoid=False)
