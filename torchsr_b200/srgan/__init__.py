"""SRGAN modules on the B200 path (mirror of torchsr/srgan/ in the reference)."""
