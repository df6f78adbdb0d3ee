"""Attribution experiment for the large convs (the shapes of VGG19 / the discriminator / the sub-pixel layer at B=16,
10.9 GFLOP each): time of the production (FAST) kernel and of the instrumented instantiation with parts switched off
(TSR_CONV_DEBUG bits: 16 = nothing (baseline of the instrumented build), 1 = epilogue skips read-out and stores,
2 = epilogue computes but does not store, 4 = no activation TMA loads, 8 = no UMMAs)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from torchsr_b200 import _lib as L  # noqa: E402
from torchsr_b200 import ops  # noqa: E402


def time_desc(d, reps=20):
    prog = ops.Program()
    for _ in range(reps):
        prog.add(d)
    prog.run()
    prog.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    prog.run()
    prog.run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (2 * reps)


def case(B, H, W, Cin, Cout, stride=1, act=L.ACT_RELU, stats=False):
    x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(9, Cout, Cin, device="cuda") * 0.05).to(torch.bfloat16)
    geom = ops.fwd_geometry(H, W, 3, 3, 1, 1, stride)
    Ho, Wo = geom["Ho"], geom["Wo"]
    out = torch.empty(B, Ho, Wo, Cout, device="cuda", dtype=torch.bfloat16)
    st = torch.zeros(Cout, 2, device="cuda") if stats else None
    kw = dict(x=x, N=B, H=H, W=W, C=Cin, x_ld=Cin, geom=geom, w=w, cout_pad=Cout, w_ld=Cin, n_slots=9,
              block_n=min(Cout, 128), out=out, os_n=Ho * Wo * Cout, os_h=Wo * Cout, os_w=Cout, n_valid=Cout, act=act,
              stats_partial=st, stats_ld=Cout)
    flop = 2.0 * B * Ho * Wo * Cout * Cin * 9
    return kw, flop, (x, w, out, st)


def main():
    print(torch.cuda.get_device_name(0))
    shapes = [("64->64 @96", dict(B=16, H=96, W=96, Cin=64, Cout=64)),
              ("128->128 @48", dict(B=16, H=48, W=48, Cin=128, Cout=128)),
              ("256->256 @24", dict(B=16, H=24, W=24, Cin=256, Cout=256)),
              ("512->512 @12", dict(B=16, H=12, W=12, Cin=512, Cout=512)),
              ("64->256 @48", dict(B=16, H=48, W=48, Cin=64, Cout=256)),
              ("64->64 @96 +stats", dict(B=16, H=96, W=96, Cin=64, Cout=64, stats=True, act=L.ACT_NONE))]
    for name, a in shapes:
        kw, flop, keep = case(**a)
        row = []
        for dbg in ("0", "16", "1", "2", "4", "8", "12"):
            os.environ["TSR_CONV_DEBUG"] = dbg
            d = ops.conv_desc(**kw)
            us = time_desc(d)
            row.append(f"dbg{dbg}: {us:6.1f}us {flop / us / 1e6:5.0f}TF")
        os.environ["TSR_CONV_DEBUG"] = "0"
        print(f"{name:20s} " + " | ".join(row))


if __name__ == "__main__":
    main()
