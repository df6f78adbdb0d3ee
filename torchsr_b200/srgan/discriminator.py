"""SRGAN discriminator - drop-in for torchsr/srgan/discriminator.py (reference lines 17-88)."""
from torch import nn

from ..engine import B200Module, Plan
from .. import nets

CONV_IDX = (0, 2, 5, 8, 11, 14, 17, 20)


def _features(widths_strides):
    layers = [nn.Conv2d(3, 64, kernel_size=3, stride=1, padding=1), nn.LeakyReLU(negative_slope=0.2, inplace=True)]
    cin = 64
    for cout, stride in widths_strides:
        layers += [nn.Conv2d(cin, cout, kernel_size=3, stride=stride, padding=1, bias=False), nn.BatchNorm2d(cout),
                   nn.LeakyReLU(negative_slope=0.2, inplace=True)]
        cin = cout
    return nn.Sequential(*layers)


class Discriminator(B200Module):
    """8 strided-conv stages, flatten (C,H,W order), Linear 18432->1024, LeakyReLU, Linear 1024->1, Sigmoid."""

    def __init__(self, image_size: int = 96) -> None:
        super().__init__()
        self.image_size = image_size
        feature_map_size = int(image_size // 16)
        self.features = _features([(64, 2), (128, 1), (128, 2), (256, 1), (256, 2), (512, 1), (512, 2)])
        self.classifier = nn.Sequential(
            nn.Linear(512 * feature_map_size * feature_map_size, 1024),
            nn.LeakyReLU(negative_slope=0.2, inplace=True),
            nn.Linear(1024, 1),
            nn.Sigmoid())

    def _records(self):
        return nets.discriminator_records(self, CONV_IDX), nets.discriminator_linears(self, self.image_size)

    def _define(self, plan: Plan, shape):
        nets.define_discriminator(self, plan, shape, CONV_IDX, sigmoid=True)
