"""Deep conv layers (Cin >= 128) with and without the activation multicast across clusters of two N tiles
(TSR_CONV_CLUSTER): per-launch time, back-to-back launches, CUDA events. Usage: python tools/microbench_cluster.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from torchsr_b200 import ops  # noqa: E402
from tools.microbench_wgrad import time_prog  # noqa: E402


def case(B, H, W, Cin, Cout, stride, block_n, cluster):
    os.environ["TSR_CONV_CLUSTER"] = "1" if cluster else "0"
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(9, Cout, Cin, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    geom = ops.fwd_geometry(H, W, 3, 3, 1, 1, stride)
    Ho, Wo = geom["Ho"], geom["Wo"]
    out = torch.empty(B, Ho, Wo, Cout, device="cuda", dtype=torch.bfloat16)
    d = ops.conv_desc(x=x, N=B, H=H, W=W, C=Cin, x_ld=Cin, geom=geom, w=w, cout_pad=Cout, w_ld=Cin, n_slots=9,
                      block_n=block_n, out=out, os_n=Ho * Wo * Cout, os_h=Wo * Cout, os_w=Cout, n_valid=Cout)
    us = time_prog([d])
    return us, 2.0 * B * Ho * Wo * Cin * Cout * 9 / us / 1e6, out


SHAPES = [  # B, H, W, Cin, Cout, stride, block_n      (VGG / discriminator layers at B=16 and B=64)
    (32, 24, 24, 256, 256, 1, 128), (32, 12, 12, 512, 512, 1, 64), (32, 12, 12, 512, 512, 1, 128),
    (32, 48, 48, 128, 128, 1, 64), (32, 24, 24, 128, 256, 1, 128), (32, 24, 24, 256, 256, 2, 128),
    (128, 24, 24, 256, 256, 1, 128), (128, 12, 12, 512, 512, 1, 128), (128, 24, 24, 128, 256, 1, 128),
]

if __name__ == "__main__":
    for sh in SHAPES:
        u0, t0, o0 = case(*sh, cluster=False)
        u1, t1, o1 = case(*sh, cluster=True)
        same = torch.equal(o0, o1)
        print(f"{sh}: one CTA per box {u0:7.1f} us {t0:6.0f} TFLOP/s | multicast {u1:7.1f} us {t1:6.0f} TFLOP/s | "
              f"x{u0 / u1:.2f} | identical output: {same}", flush=True)
