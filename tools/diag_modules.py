"""Module-level parity printout (GPU box): python tools/diag_modules.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import module_checks as MC  # noqa: E402
from torchsr_b200 import ops  # noqa: E402
from torchsr_b200.srgan.discriminator import Discriminator  # noqa: E402
from torchsr_b200.srgan.generator import Generator  # noqa: E402
from torchsr_b200.srgan.residual import ResidualBlock, SubpixelConvolutionLayer  # noqa: E402


def show(name, r, errs, top=6):
    print(name, "  ".join(f"{k}={v:.3e}" for k, v in r.items()), flush=True)
    for k, v in sorted(errs.items(), key=lambda kv: -kv[1])[:top]:
        print(f"      {k:40s} {v:.3e}")


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    torch.manual_seed(1234)
    if which in ("all", "block"):
        m = ResidualBlock()
        MC.randomize_bn(m)
        r, e = MC.check_module(m, lambda sd, x, tr, buf: MC.O.srgan_residual_block(
            {("." + k): v for k, v in sd.items()}, "", x, tr, buf), torch.randn(4, 64, 24, 24), input_grad=True)
        show("ResidualBlock", r, e)
        ops.check_watchdog()
    if which in ("all", "sub"):
        m = SubpixelConvolutionLayer()
        r, e = MC.check_module(m, lambda sd, x, tr, buf: MC.O.srgan_subpixel({("." + k): v for k, v in sd.items()}, "", x),
                               torch.randn(2, 64, 12, 12), input_grad=True)
        show("Subpixel", r, e)
        ops.check_watchdog()
    if which in ("all", "G"):
        G = Generator()
        MC.randomize_bn(G)
        r, e = MC.check_module(G, MC.O.srgan_generator, torch.rand(4, 3, 24, 24))
        show("Generator", r, e, top=12)
        ops.check_watchdog()
    if which in ("all", "D"):
        D = Discriminator()
        MC.randomize_bn(D)
        r, e = MC.check_module(D, MC.O.srgan_discriminator, torch.rand(4, 3, 96, 96), input_grad=True)
        show("Discriminator", r, e, top=12)
        ops.check_watchdog()


if __name__ == "__main__":
    main()
