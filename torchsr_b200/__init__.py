"""torchsr_b200: the SRGAN / ESRGAN generator + discriminator hot path of roclark/torchsr on hand-written sm_100a
kernels. Sub-packages mirror the reference layout: torchsr_b200.srgan, torchsr_b200.esrgan."""
__version__ = "0.1.0"


def install_as_torchsr() -> None:
    """Makes `import torchsr...` resolve to this package (reference import paths: torchsr.srgan.generator.Generator,
    torchsr.models.select_trainer_model, torchsr.torchsr.main, ... - /root/reference/setup.py:39-41 installs the package
    under that name). Opt-in and in-process only: nothing is written to disk, and an already imported `torchsr` (the
    real reference) is left alone."""
    import importlib
    import sys
    if "torchsr" in sys.modules and sys.modules["torchsr"] is not sys.modules[__name__]:
        raise RuntimeError("a different `torchsr` package is already imported in this process")
    sys.modules["torchsr"] = sys.modules[__name__]
    for name in ("constants", "models", "dataset", "test", "torchsr", "srgan", "srgan.generator", "srgan.residual",
                 "srgan.discriminator", "srgan.loss", "srgan.trainer", "esrgan", "esrgan.generator", "esrgan.residual",
                 "esrgan.discriminator", "esrgan.loss", "esrgan.trainer"):
        sys.modules["torchsr." + name] = importlib.import_module(__name__ + "." + name)
