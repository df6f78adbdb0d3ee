"""Staged-epilogue attribution: two launches whose main loop is cheap relative to their epilogue - the K=32 first layer
(one K iteration per tile) and a 3x3 64->64 halo conv - timed with whatever library TSR_LIB_PATH selects (variants built
with -DTSR_EXP=n switch one step of the epilogue off: 1 proxy fence, 2 bulk store + its wait, 3 arithmetic, 4 CTA barrier;
results of the variants are wrong by construction, only the time counts). Usage: python tools/microbench_epilogue.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from torchsr_b200 import _lib as L  # noqa: E402
from torchsr_b200 import ops  # noqa: E402
from tools.microbench_wgrad import time_prog  # noqa: E402


def first_layer(B=64, H=96, W=96):
    x = torch.randn(B, H, W, 32, device="cuda").to(torch.bfloat16)
    w = (torch.randn(1, 64, 32, device="cuda") * 0.1).to(torch.bfloat16)
    bias = torch.randn(64, device="cuda")
    geom = ops.fwd_geometry(H, W, 1, 1, 0, 0, 1)
    out = torch.empty(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
    d = ops.conv_desc(x=x, N=B, H=H, W=W, C=32, x_ld=32, geom=geom, w=w, cout_pad=64, w_ld=32, n_slots=1, block_n=64,
                      out=out, os_n=H * W * 64, os_h=W * 64, os_w=64, n_valid=64, bias=bias, act=L.ACT_LEAKY)
    return time_prog([d], reps=20), B * H * W


def trunk_like(B=1, H=512, W=512, res=False):
    x = torch.randn(B, H, W, 64, device="cuda").to(torch.bfloat16)
    w = (torch.randn(9, 64, 64, device="cuda") * 0.05).to(torch.bfloat16)
    bias = torch.randn(64, device="cuda")
    r = torch.randn(B, H, W, 64, device="cuda").to(torch.bfloat16) if res else None
    geom = ops.fwd_geometry(H, W, 3, 3, 1, 1, 1)
    out = torch.empty(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
    kw = dict(res=r, aux=(H * W * 64, W * 64, 64)) if res else {}
    d = ops.conv_desc(x=x, N=B, H=H, W=W, C=64, x_ld=64, geom=geom, w=w, cout_pad=64, w_ld=64, n_slots=9, block_n=64,
                      out=out, os_n=H * W * 64, os_h=W * 64, os_w=64, n_valid=64, bias=bias, act=L.ACT_PRELU,
                      prelu=torch.tensor([0.25], device="cuda"), **kw)
    return time_prog([d], reps=20), B * H * W


if __name__ == "__main__":
    tag = os.path.basename(os.environ.get("TSR_LIB_PATH", "libtorchsr_b200.so"))
    u1, m1 = first_layer()
    u2, m2 = trunk_like()
    u3, m3 = trunk_like(res=True)
    t = lambda us, m: us * 1e-6 * 1.965e9 / (m / 128 / 148)      # noqa: E731   clocks per tile per CTA
    print(f"{tag:22s} first layer K=32: {u1:6.1f} us ({t(u1, m1):5.0f} clk/tile) | 3x3 64->64 @512^2: {u2:6.1f} us "
          f"({t(u2, m2):5.0f} clk/tile) | + residual: {u3:6.1f} us ({t(u3, m3):5.0f} clk/tile)")
