"""ESRGAN modules on the B200 path (mirror of torchsr/esrgan/ in the reference)."""
