"""ESRGAN generator - drop-in for torchsr/esrgan/generator.py (reference lines 20-81)."""
from torch import nn

from ..engine import B200Module, Plan
from .. import nets
from .residual import ResidualInResidualDenseBlock

NUM_RESIDUAL = 23


class Generator(B200Module):
    """conv 3->64 -> 23 RRDB -> conv + skip -> 2 x (nearest x2, conv, LeakyReLU) -> conv + LeakyReLU -> conv 64->3.
    No BatchNorm; forward(x: [N,3,H,W]) -> [N,3,4H,4W]."""

    def __init__(self, num_rrdb_blocks: int = NUM_RESIDUAL) -> None:
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, kernel_size=3, stride=1, padding=1)
        self.blocks = nn.Sequential(*[ResidualInResidualDenseBlock(channels=64, growth_channels=32, scale_ratio=0.2)
                                      for _ in range(num_rrdb_blocks)])
        self.conv2 = nn.Conv2d(64, 64, kernel_size=3, stride=1, padding=1)
        self.upsample1 = nn.Conv2d(64, 64, kernel_size=3, stride=1, padding=1)
        self.upsample2 = nn.Conv2d(64, 64, kernel_size=3, stride=1, padding=1)
        self.conv3 = nn.Sequential(nn.Conv2d(64, 64, kernel_size=3, stride=1, padding=1),
                                   nn.LeakyReLU(negative_slope=0.2, inplace=True))
        self.conv4 = nn.Conv2d(64, 3, kernel_size=3, stride=1, padding=1)

    def _records(self):
        return nets.esrgan_generator_records(self), []

    def _define(self, plan: Plan, shape):
        nets.define_esrgan_generator(self, plan, shape)
