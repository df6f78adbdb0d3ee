"""SRGAN generator - drop-in for torchsr/srgan/generator.py (reference lines 20-81)."""
import math

from torch import nn

from ..engine import B200Module, Plan
from .. import nets
from .residual import ResidualBlock, SubpixelConvolutionLayer

NUM_RESIDUAL = 16


class Generator(B200Module):
    """9x9 conv + PReLU -> 16 residual blocks -> 3x3 conv + BN -> skip add -> log2(scale) sub-pixel stages -> 9x9 conv.

    Same constructor, children and state_dict keys as the reference class; forward(x: [N,3,H,W] fp32) returns
    [N,3,scale*H,scale*W] fp32 and is differentiable w.r.t. every parameter."""

    def __init__(self, scale_factor: int = 4) -> None:
        super().__init__()
        num_conv_layers = int(math.log(scale_factor, 2))
        self.conv1 = nn.Sequential(nn.Conv2d(3, 64, kernel_size=9, stride=1, padding=4), nn.PReLU())
        self.blocks = nn.Sequential(*[ResidualBlock(channels=64) for _ in range(NUM_RESIDUAL)])
        self.conv2 = nn.Sequential(nn.Conv2d(64, 64, kernel_size=3, stride=1, padding=1, bias=False),
                                   nn.BatchNorm2d(64))
        self.conv_layers = nn.Sequential(*[SubpixelConvolutionLayer(64) for _ in range(num_conv_layers)])
        self.conv3 = nn.Conv2d(64, 3, kernel_size=9, stride=1, padding=4)

    def _records(self):
        return nets.srgan_generator_records(self), []

    def _define(self, plan: Plan, shape):
        nets.define_srgan_generator(self, plan, shape)
