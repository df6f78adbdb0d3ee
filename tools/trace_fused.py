"""Where a fused conv + training-BatchNorm launch of the generator trunk (3x3 64->64 @24x24, B=16, 144 CTAs) spends its
cycles: clock64 stamps of the instrumented kernel instantiation inside a back-to-back chain of identical launches with
programmatic dependent launch (the real situation), median over CTAs, relative to the CTA's start."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from torchsr_b200 import _lib as L  # noqa: E402
from torchsr_b200 import ops  # noqa: E402


def main():
    B, H, W, C = 16, 24, 24, 64
    x = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
    w = (torch.randn(9, C, C, device="cuda") * 0.05).to(torch.bfloat16)
    y, raw, res = (torch.empty(B, H, W, C, device="cuda", dtype=torch.bfloat16) for _ in range(3))
    gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    coef = torch.empty(4 * C, device="cuda")
    alpha = torch.full((1,), 0.25, device="cuda")
    n = 12
    zero = torch.zeros(n * 256, device="cuda")
    tr = torch.zeros(n, 512, 40, dtype=torch.int64, device="cuda")
    prog = ops.Program()
    prog.add(ops.elt(L.E_ZERO, p=[zero], i=[zero.numel() * 4]))
    geom = ops.fwd_geometry(H, W, 3, 3, 1, 1, 1)
    for k in range(n):
        d = ops.conv_desc(x=x, N=B, H=H, W=W, C=C, x_ld=C, geom=geom, w=w, cout_pad=C, w_ld=C, n_slots=9, block_n=32,
                          out=y, os_n=H * W * C, os_h=W * C, os_w=C, n_valid=C, act=L.ACT_PRELU, prelu=alpha,
                          out_preact=raw, res=res, aux=(H * W * C, W * C, C), stats_partial=ops.ptr(zero, k * 256),
                          stats_ld=C, bnf_mode=1, bnf_c=C, bnf_counter=ops.ptr(zero, k * 256 + 128), bnf_gamma=gamma,
                          bnf_beta=beta, bnf_coef=coef, bnf_count=B * H * W)
        if os.environ.get("TRACE", "1") == "1":
            d.trace = tr[k].data_ptr()
        prog.add(d)
    for _ in range(3):
        prog.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    prog.run()
    e1.record()
    torch.cuda.synchronize()
    print(f"chain of {n} fused launches: {e0.elapsed_time(e1) * 1e3 / n:.2f} us per launch (instrumented={os.environ.get('TRACE', '1')})")
    if os.environ.get("TRACE", "1") != "1":
        return
    t = tr.cpu()
    names = {1: "setup done", 2: "pdl wait returned", 5: "accumulator ready", 17: "phase A (stats) done",
             18: "column sums flushed", 19: "grid barrier passed", 20: "coefficients ready", 21: "phase B stored",
             6: "tile done", 7: "exit"}
    for k in (4, 8):
        tk = t[k]
        tk = tk[tk[:, 7] != 0]
        base = tk[:, 0]
        print(f"launch {k}: {tk.shape[0]} CTAs; first CTA start -> last CTA exit: {(tk[:, 7].max() - base.min()).item()} clk; "
              f"start spread {(base.max() - base.min()).item()} clk")
        for i in (1, 2, 5, 17, 18, 19, 20, 21, 6, 7):
            rel = (tk[:, i] - base).float()
            print(f"   {names[i]:24s} median {rel.median():8.0f}  min {rel.min():8.0f}  max {rel.max():8.0f} clk")
        # stamps relative to a common origin (globaltimer-free: clock64 differs per SM, so only per-CTA deltas are exact)
    ops.check_watchdog()


if __name__ == "__main__":
    main()
