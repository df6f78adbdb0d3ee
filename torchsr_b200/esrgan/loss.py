"""VGG19 perceptual loss - mirror of torchsr/esrgan/loss.py:18-54 (identical to the SRGAN copy, as in the reference).

Adjacent to the hot path: a frozen torchvision VGG19 `features[:36]`, L1 between the features
of the generated and the target image, inputs not ImageNet-normalised (as in the reference). The reference downloads
ImageNet weights (`vgg19(pretrained=True)`); offline they come from the torch-hub cache, from a file named by
TORCHSR_VGG_WEIGHTS, or - for benchmarks and parity runs, where only the arithmetic matters - from a seeded random
initialisation (TORCHSR_VGG_WEIGHTS=random). This module is executed by PyTorch (cuDNN, bf16 autocast on CUDA)."""
import os

import torch
from torch import Tensor, nn


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
        self._bf16 = None
        return r

    def _features_bf16(self):
        """bf16 channels_last copy of the frozen feature extractor (made once; the weights never change)."""
        if getattr(self, "_bf16", None) is None:
            import copy
            f = copy.deepcopy(self.features).to(dtype=torch.bfloat16, memory_format=torch.channels_last)
            object.__setattr__(self, "_bf16", f.eval())
        return self._bf16

    def target_features(self, target: Tensor) -> Tensor:
        """Features of the (constant) target image; lets a trainer compute them early, on another stream."""
        with torch.no_grad():
            if target.is_cuda:
                return self._features_bf16()(target.to(dtype=torch.bfloat16, memory_format=torch.channels_last))
            return self.features(target)

    def from_features(self, source: Tensor, target_features: Tensor) -> Tensor:
        if source.is_cuda:
            fs = self._features_bf16()(source.to(dtype=torch.bfloat16, memory_format=torch.channels_last))
            return torch.nn.functional.l1_loss(fs.float(), target_features.float())
        return torch.nn.functional.l1_loss(self.features(source), target_features)

    def forward(self, source: Tensor, target: Tensor) -> Tensor:
        if source.is_cuda:
            f = self._features_bf16()
            fs = f(source.to(dtype=torch.bfloat16, memory_format=torch.channels_last))
            with torch.no_grad():
                ft = f(target.to(dtype=torch.bfloat16, memory_format=torch.channels_last))
            return torch.nn.functional.l1_loss(fs.float(), ft.float())
        return torch.nn.functional.l1_loss(self.features(source), self.features(target))
