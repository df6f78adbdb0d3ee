"""Attribution microbenchmark for the implicit-GEMM conv kernel on the large 3x3 shapes of the workload.

For each shape x {im2col, halo patch} x TSR_CONV_DEBUG {0: normal, 2: no stores, 1: no accumulator read-out at all} it
reports the time per launch (20 back-to-back launches, CUDA events) and, from the in-kernel clock64 stamps of one
traced launch, where the first tile of a CTA spends its cycles. Usage: python tools/trace_conv.py [shape-name ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from torchsr_b200 import _lib as L  # noqa: E402
from torchsr_b200 import ops  # noqa: E402

SHAPES = {
    # name: (N, H, W, Cin, Cout, block_n, stats, act)
    "vgg1_2  64->64  @96": (16, 96, 96, 64, 64, 64, False, L.ACT_RELU),
    "vgg2_2 128->128 @48": (16, 48, 48, 128, 128, 128, False, L.ACT_RELU),
    "vgg3_2 256->256 @24": (16, 24, 24, 256, 256, 128, False, L.ACT_RELU),
    "up2     64->256 @48": (16, 48, 48, 64, 256, 128, False, L.ACT_PRELU),
    "up1     64->256 @24": (16, 24, 24, 64, 256, 128, False, L.ACT_PRELU),
    "trunk   64->64  @24": (16, 24, 24, 64, 64, 32, True, L.ACT_NONE),
    "trunk64 64->64  @24": (16, 24, 24, 64, 64, 64, True, L.ACT_NONE),
    "infer   64->64 @512": (1, 512, 512, 64, 64, 64, False, L.ACT_PRELU),
    "inferup 64->256@512": (1, 512, 512, 64, 256, 128, False, L.ACT_PRELU),
    "vgg2_1  64->128 @48": (16, 48, 48, 64, 128, 128, False, L.ACT_RELU),
    "n256 up2 64->256@48": (16, 48, 48, 64, 256, 256, False, L.ACT_PRELU),
    "n256 inferup    @512": (1, 512, 512, 64, 256, 256, False, L.ACT_PRELU),
    "n256 vgg3_2     @24": (16, 24, 24, 256, 256, 256, False, L.ACT_RELU),
    "n256 vgg4 512  @12": (16, 12, 12, 512, 512, 256, False, L.ACT_RELU),
    "n128 vgg4 512  @12": (16, 12, 12, 512, 512, 128, False, L.ACT_RELU),
    "n64  vgg4 512  @12": (16, 12, 12, 512, 512, 64, False, L.ACT_RELU),
    "vgg2_2n 128->128@48": (16, 48, 48, 128, 128, 64, False, L.ACT_RELU),
}


def build(shape, trace=None):
    N, H, W, C, Co, bn, stats, act = shape
    x = torch.randn(N, H, W, C, device="cuda").to(torch.bfloat16)
    w = (torch.randn(9, Co, C, device="cuda") * 0.05).to(torch.bfloat16)
    out = torch.empty(N, H, W, Co, device="cuda", dtype=torch.bfloat16)
    bias = torch.randn(Co, device="cuda")
    prelu = torch.full((1,), 0.25, device="cuda")
    st = torch.zeros(Co, 2, device="cuda") if stats else None
    d = ops.conv_desc(x=x, N=N, H=H, W=W, C=C, x_ld=C, geom=ops.fwd_geometry(H, W, 3, 3, 1, 1, 1), w=w, cout_pad=Co,
                      w_ld=C, n_slots=9, block_n=bn, out=out, os_n=H * W * Co, os_h=W * Co, os_w=Co, n_valid=Co,
                      bias=bias, prelu=prelu, act=act, stats_partial=st, stats_ld=Co if stats else 0)
    keep = (x, w, out, bias, prelu, st)
    return d, keep


def time_it(shape):
    d, keep = build(shape)
    prog = ops.Program()
    reps = 20
    for _ in range(reps):
        prog.add(d)
    prog.run()
    prog.run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    prog.run()
    prog.run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (2 * reps)


def trace_it(shape):
    N, H, W, C, Co, bn, stats, act = shape
    d, keep = build(shape)
    slots = (N * H * W // 96 + 1024) * (Co // bn)      # upper bound on CTAs in either mode
    tr = torch.zeros(slots, 40, dtype=torch.int64, device="cuda")
    d.trace = tr.data_ptr()
    ops.run_now(d)
    ops.run_now(d)
    torch.cuda.synchronize()
    t = tr.cpu()
    t = t[t[:, 7] != 0]
    rel = lambda i: float((t[:, i] - t[:, 0]).float().median())  # noqa: E731
    ch = [(rel(32 + 2 * c), rel(33 + 2 * c)) for c in range(4) if (t[:, 32 + 2 * c] != 0).any()]
    return dict(ctas=t.shape[0], setup=rel(1), acc_full=rel(5), chunks=ch, epi_done=rel(6), exit=rel(7))


def main():
    names = [n for n in SHAPES if not sys.argv[1:] or any(a in n for a in sys.argv[1:])]
    print(torch.cuda.get_device_name(0))
    for name in names:
        shape = SHAPES[name]
        N, H, W, C, Co = shape[:5]
        gflop = 2 * N * H * W * 9 * C * Co / 1e9
        for halo in ("0", "1"):
            os.environ["TSR_CONV_HALO"] = halo
            row = []
            for dbg in (("0",) if os.environ.get("TRACE_CONV_QUICK") else ("0", "2", "1", "5", "9", "13")):
                os.environ["TSR_CONV_DEBUG"] = dbg
                row.append(time_it(shape))
            os.environ["TSR_CONV_DEBUG"] = "0"
            if os.environ.get("TRACE_CONV_QUICK"):
                print(f"{name} bn={shape[5]:3d} halo={halo}: {row[0]:7.2f} us ({gflop / row[0] * 1e3:6.1f} TF/s)")
                continue
            tr = trace_it(shape)
            chunks = " ".join(f"[{int(a)}->{int(b)}]" for a, b in tr["chunks"])
            print(f"{name} bn={shape[5]:3d} halo={halo}: {row[0]:7.2f} us ({gflop / row[0] * 1e3:6.1f} TF/s) | no-store "
                  f"{row[1]:7.2f} | no-epi {row[2]:7.2f} | no-epi,no-TMA {row[3]:7.2f} | no-epi,no-MMA {row[4]:7.2f} | "
                  f"sync only {row[5]:7.2f} | ctas {tr['ctas']} setup {int(tr['setup'])} "
                  f"acc_full {int(tr['acc_full'])} chunks(ld->st) {chunks} epi_done {int(tr['epi_done'])} exit {int(tr['exit'])}")
    ops.check_watchdog()


if __name__ == "__main__":
    main()
