"""ESRGAN building blocks - drop-in for torchsr/esrgan/residual.py (reference lines 17-129).

Parameter containers with the reference's children, names and initialisation (kaiming_normal * 0.1, zero bias);
the arithmetic runs inside the generator's launch lists (torchsr_b200/nets.py: zero-copy dense concatenation)."""
from torch import nn


class ResidualDenseBlock(nn.Module):
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

    def forward(self, x):
        raise NotImplementedError("ResidualDenseBlock runs as part of torchsr_b200.esrgan.generator.Generator; a "
                                  "standalone forward is not provided on the B200 path")


class ResidualInResidualDenseBlock(nn.Module):
    def __init__(self, channels: int = 64, growth_channels: int = 32, scale_ratio: float = 0.2) -> None:
        super().__init__()
        self.RDB1 = ResidualDenseBlock(channels, growth_channels, scale_ratio)
        self.RDB2 = ResidualDenseBlock(channels, growth_channels, scale_ratio)
        self.RDB3 = ResidualDenseBlock(channels, growth_channels, scale_ratio)

    def forward(self, x):
        raise NotImplementedError("ResidualInResidualDenseBlock runs as part of torchsr_b200.esrgan.generator."
                                  "Generator; a standalone forward is not provided on the B200 path")
