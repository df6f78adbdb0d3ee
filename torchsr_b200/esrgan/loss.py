"""VGG19 perceptual loss - mirror of torchsr/esrgan/loss.py:18-54 (identical to the SRGAN copy, as in the reference)."""
from ..srgan.loss import VGGFeaturesB200, VGGLoss  # noqa: F401
