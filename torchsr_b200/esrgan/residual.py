"""ESRGAN building blocks - drop-in for torchsr/esrgan/residual.py (reference lines 17-129).

Same children, names and initialisation as the reference (kaiming_normal * 0.1, zero bias). Inside the generator the
blocks run as part of its launch lists; called on their own they build the same zero-copy dense-concatenation chain
(torchsr_b200/nets.py: rdb_chain) between two layout-conversion kernels."""
from torch import nn

from .. import nets
from ..engine import B200Module, Plan


class ResidualDenseBlock(B200Module):
    def __init__(self, channels: int = 64, growth_channels: int = 32, scale_ratio: float = 0.2) -> None:
        super().__init__()
        for k in range(1, 5):
            setattr(self, f"conv{k}", nn.Sequential(
                nn.Conv2d(channels + (k - 1) * growth_channels, growth_channels, kernel_size=3, stride=1, padding=1),
                nn.LeakyReLU(negative_slope=0.2, inplace=True)))
        self.conv5 = nn.Conv2d(channels + 4 * growth_channels, channels, kernel_size=3, stride=1, padding=1)
        self.scale_ratio = scale_ratio
        for module in self.modules():               # reference residual.py:58-63
            if isinstance(module, nn.Conv2d):
                nn.init.kaiming_normal_(module.weight)
                module.weight.data *= 0.1
                if module.bias is not None:
                    module.bias.data.zero_()

    # forward(x: [N,64,H,W]) -> conv5(cat(x, conv1..4)) * 0.2 + x   (reference residual.py:81-86)
    def _records(self):
        return nets.rdb_records("", self), []

    def _define(self, plan: Plan, shape):
        nets.define_rdb_standalone(self, plan, shape, [""], rrdb=False)


class ResidualInResidualDenseBlock(B200Module):
    def __init__(self, channels: int = 64, growth_channels: int = 32, scale_ratio: float = 0.2) -> None:
        super().__init__()
        self.RDB1 = ResidualDenseBlock(channels, growth_channels, scale_ratio)
        self.RDB2 = ResidualDenseBlock(channels, growth_channels, scale_ratio)
        self.RDB3 = ResidualDenseBlock(channels, growth_channels, scale_ratio)

    # forward(x) -> RDB3(RDB2(RDB1(x))) * 0.2 + x   (reference residual.py:124-129; 0.2 hard-coded there)
    def _records(self):
        recs = []
        for r in nets.RDB_NAMES:
            recs += nets.rdb_records(f"{r}.", getattr(self, r))
        return recs, []

    def _define(self, plan: Plan, shape):
        nets.define_rdb_standalone(self, plan, shape, list(nets.RDB_NAMES), rrdb=True)
