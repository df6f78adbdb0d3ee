"""VGG19 perceptual loss - mirror of torchsr/srgan/loss.py:18-54.

Adjacent to the hot path: a frozen torchvision VGG19 `features[:36]`, L1 between the features
of the generated and the target image, inputs not ImageNet-normalised (as in the reference). The reference downloads
ImageNet weights (`vgg19(pretrained=True)`); offline they come from the torch-hub cache, from a file named by
TORCHSR_VGG_WEIGHTS, or - for benchmarks and parity runs, where only the arithmetic matters - from a seeded random
initialisation (TORCHSR_VGG_WEIGHTS=random).

The feature extractor runs on this repo's kernels (nets.define_vgg: implicit-GEMM convs with the ReLU fused into the
epilogue, NHWC bf16 activations, max-pool / ReLU backward kernels, data gradients only - the network is frozen) and the
feature L1 on losses.l1. There is no PyTorch/cuDNN or CPU execution path: inputs on another device raise."""
import os

import torch
from torch import Tensor, nn

from .. import losses, nets
from ..engine import B200Module, Plan


class VGGFeaturesB200(B200Module):
    """The frozen feature extractor as a launch-list module. Shares the Parameters of the torchvision layers it is
    given (no copy, same state_dict entries); input NCHW fp32, output NCHW fp32 features."""

    def __init__(self, features: nn.Sequential) -> None:
        super().__init__()
        self.features = features
        self.eval()

    def _records(self):
        return nets.vgg_records(self), []

    def _define(self, plan: Plan, shape):
        nets.define_vgg(self, plan, shape)


def _vgg19_features(feature_layer: int) -> nn.Sequential:
    import torchvision
    src = os.environ.get("TORCHSR_VGG_WEIGHTS", "")
    if src == "random":
        gen_state = torch.random.get_rng_state()
        torch.manual_seed(1234)
        model = torchvision.models.vgg19(weights=None)
        torch.random.set_rng_state(gen_state)
    elif src:
        model = torchvision.models.vgg19(weights=None)
        model.load_state_dict(torch.load(src, map_location="cpu"))
    else:
        model = torchvision.models.vgg19(weights=torchvision.models.VGG19_Weights.IMAGENET1K_V1)
    return nn.Sequential(*list(model.features.children())[:feature_layer]).eval()


class VGGLoss(nn.Module):
    def __init__(self, feature_layer: int = 36) -> None:
        super().__init__()
        self.features = _vgg19_features(feature_layer)
        for param in self.features.parameters():
            param.requires_grad = False

    def _apply(self, fn, *a, **kw):
        r = super()._apply(fn, *a, **kw)
        object.__setattr__(self, "_b200", None)
        return r

    def _features_b200(self) -> VGGFeaturesB200:
        if getattr(self, "_b200", None) is None:
            # not registered as a child module: the parameters already belong to self.features
            object.__setattr__(self, "_b200", VGGFeaturesB200(self.features))
        return self._b200

    def _extract(self, x: Tensor) -> Tensor:
        return self._features_b200()(x)      # raises TorchSRB200Error off a CUDA sm_100 device

    def target_features(self, target: Tensor) -> Tensor:
        """Features of the (constant) target image; lets a trainer compute them early, on another stream."""
        with torch.no_grad():
            return self._extract(target)

    def from_features(self, source: Tensor, target_features: Tensor) -> Tensor:
        return losses.l1(self._extract(source), target_features)

    def forward(self, source: Tensor, target: Tensor) -> Tensor:
        return self.from_features(source, self.target_features(target))
