"""Timing experiments on the implicit-GEMM kernel: per-launch time for several shapes/modes plus the in-kernel
clock64 trace (setup / first TMA / first full barrier / last MMA issue / accumulator ready / epilogue done / exit)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from torchsr_b200 import _lib as L  # noqa: E402
from torchsr_b200 import ops  # noqa: E402


def time_prog(descs, reps=200):
    prog = ops.Program()
    for _ in range(reps):
        for d in descs:
            prog.add(d)
    prog.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    prog.run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def conv_case(B, H, W, Cin, Cout, k=3, stride=1, block_n=None, stats=True, trace=False):
    x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(k * k, Cout, Cin, device="cuda") * 0.05).to(torch.bfloat16)
    geom = ops.fwd_geometry(H, W, k, k, k // 2, k // 2, stride)
    Ho, Wo = geom["Ho"], geom["Wo"]
    out = torch.empty(B, Ho, Wo, Cout, device="cuda", dtype=torch.bfloat16)
    block_n = block_n or min(Cout, 128)
    tiles = (B * Ho * Wo + 127) // 128
    st = torch.empty(tiles, Cout, 2, device="cuda") if stats else None
    d = ops.conv_desc(x=x, N=B, H=H, W=W, C=Cin, x_ld=Cin, geom=geom, w=w, cout_pad=Cout, w_ld=Cin, n_slots=k * k,
                      block_n=block_n, out=out, os_n=Ho * Wo * Cout, os_h=Wo * Cout, os_w=Cout, n_valid=Cout,
                      stats_partial=st, stats_ld=Cout)
    keep = (x, w, out, st)
    tr = None
    if trace:
        grid = tiles * (Cout // block_n)
        tr = torch.zeros(grid, 40, dtype=torch.int64, device="cuda")
        d.trace = tr.data_ptr()
    return d, keep, tr


def gemm_case(M, N, K, block_n=64):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    d = ops.gemm_desc(a=a, M=M, K=K, a_ld=K, w=w, n_rows=N, block_n=block_n, out=out, out_ld=N, n_valid=N, out_f32=False)
    return d, (a, w, out)


def main():
    print(torch.cuda.get_device_name(0))
    # empty-ish kernel launch floor through the same path
    z = torch.zeros(64, device="cuda")
    d0 = ops.elt(L.E_ZERO, p=[z], i=[256])
    print(f"memset node             {time_prog([d0]):7.2f} us")
    c = torch.empty(64, device="cuda", dtype=torch.bfloat16)
    d1 = ops.elt(L.E_CAST, p=[z, c], i=[64, 0])
    print(f"tiny cast kernel        {time_prog([d1]):7.2f} us")
    for name, args in [("trunk 3x3 64->64 B16", dict(B=16, H=24, W=24, Cin=64, Cout=64)),
                       ("trunk no-stats", dict(B=16, H=24, W=24, Cin=64, Cout=64, stats=False)),
                       ("trunk bn=32", dict(B=16, H=24, W=24, Cin=64, Cout=64, block_n=32)),
                       ("trunk B64", dict(B=64, H=24, W=24, Cin=64, Cout=64)),
                       ("1x1 64->64 B16", dict(B=16, H=24, W=24, Cin=64, Cout=64, k=1)),
                       ("up1 64->256 @48 B16", dict(B=16, H=48, W=48, Cin=64, Cout=256)),
                       ("D 128->256 @24", dict(B=16, H=24, W=24, Cin=128, Cout=256)),
                       ("D 512->512 s2 @12", dict(B=16, H=12, W=12, Cin=512, Cout=512, stride=2))]:
        d, keep, _ = conv_case(**args)
        print(f"{name:24s}{time_prog([d]):7.2f} us")
    d, keep = gemm_case(9216, 64, 576)
    print(f"gemm 9216x64x576        {time_prog([d]):7.2f} us")
    d, keep = gemm_case(9216, 64, 64)
    print(f"gemm 9216x64x64         {time_prog([d]):7.2f} us")
    # trace of one warm launch
    d, keep, tr = conv_case(16, 24, 24, 64, 64, trace=True)
    for _ in range(3):
        ops.run_now(d)
    torch.cuda.synchronize()
    t = tr.cpu()
    rel = (t - t[:, :1])[:, 1:8]
    prod = (t[:, 8:17] - t[:, :1]).float().median(0).values.tolist()
    cons = (t[:, 24:33] - t[:, :1]).float().median(0).values.tolist()
    print("producer issue per k-iter:", [int(v) for v in prod])
    print("consumer full  per k-iter:", [int(v) for v in cons])
    epi = (t[:, 32:36] - t[:, :1]).float().median(0).values.tolist()
    print("epilogue warp2 (ld done, stored) x2 chunks:", [int(v) for v in epi])
    names = ["setup_done", "first_tma_issued", "first_full", "last_mma_issued", "accum_ready", "epilogue_done", "exit"]
    print("trace (cycles since CTA start; median / max over CTAs):")
    for i, n in enumerate(names):
        col = rel[:, i].float()
        print(f"   {n:18s} {col.median().item():8.0f} {col.max().item():8.0f}")
    ops.check_watchdog()


if __name__ == "__main__":
    main()
