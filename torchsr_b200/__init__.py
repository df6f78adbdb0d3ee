"""torchsr_b200: the SRGAN / ESRGAN generator + discriminator hot path of roclark/torchsr on hand-written sm_100a
kernels. Sub-packages mirror the reference layout: torchsr_b200.srgan, torchsr_b200.esrgan."""
__version__ = "0.1.0"
