"""SRGAN building blocks - drop-in for torchsr/srgan/residual.py (reference lines 16-92).

The classes keep the reference's child modules (names, shapes, default initialisation) so that state_dict() keys
match, but forward() never calls them: the arithmetic runs in the sm_100a kernels behind the C ABI."""
from torch import Tensor, nn

from .. import _lib as L
from ..engine import B200Module, ConvRec, Plan
from .. import nets


class SubpixelConvolutionLayer(B200Module):
    """conv3x3 (C -> 4C, bias) -> PixelShuffle(2) -> PReLU   (reference residual.py:16-48)."""

    def __init__(self, channels: int = 64) -> None:
        super().__init__()
        self.conv = nn.Conv2d(channels, channels * 4, kernel_size=3, stride=1, padding=1)
        self.pixel_shuffle = nn.PixelShuffle(upscale_factor=2)
        self.prelu = nn.PReLU()

    def _records(self):
        return [ConvRec("conv", self.conv, shuffle=True)], []

    def _define(self, plan: Plan, shape):
        nets.define_standalone(self, plan, shape, lambda x: nets.subpixel_stage(
            plan, plan.fwd, "sub", self, plan.store.convs[0], x))


class ResidualBlock(B200Module):
    """x + BN(conv(PReLU(BN(conv(x)))))   (reference residual.py:51-92)."""

    def __init__(self, channels: int = 64) -> None:
        super().__init__()
        self.conv1 = nn.Conv2d(channels, channels, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(channels)
        self.prelu = nn.PReLU()
        self.conv2 = nn.Conv2d(channels, channels, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(channels)

    def _records(self):
        return [ConvRec("conv1", self.conv1), ConvRec("conv2", self.conv2)], []

    def _define(self, plan: Plan, shape):
        c = plan.store.convs
        nets.define_standalone(self, plan, shape, lambda x: nets.residual_block_stage(
            plan, plan.fwd, "blk", self, c[0], c[1], x))
