"""Per-iteration producer/consumer timing of the implicit-GEMM pipeline under different grid sizes / tile widths."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from torchsr_b200 import ops  # noqa: E402


def run(name, M, N, K, block_n):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    d = ops.gemm_desc(a=a, M=M, K=K, a_ld=K, w=w, n_rows=N, block_n=block_n, out=out, out_ld=N, n_valid=N, out_f32=False)
    grid = ((M + 127) // 128) * (N // block_n)
    tr = torch.zeros(grid, 40, dtype=torch.int64, device="cuda")
    d.trace = tr.data_ptr()
    for _ in range(3):
        ops.run_now(d)
    torch.cuda.synchronize()
    t = tr.cpu()
    n = min(16, K // 64)
    prod = (t[:, 8:8 + n] - t[:, :1]).float().median(0).values
    cons = (t[:, 24:24 + n] - t[:, :1]).float().median(0).values
    tot = (t[:, 7] - t[:, 0]).float().median().item()
    print(f"{name:34s} grid={grid:4d} total={tot:7.0f}  producer d={[int(v) for v in (prod[1:] - prod[:-1]).tolist()]}")
    print(f"{'':34s} first_full={int(cons[0])} consumer d={[int(v) for v in (cons[1:] - cons[:-1]).tolist()]}")


def main():
    run("1 CTA  N=64", 128, 64, 576, 64)
    run("1 CTA  N=16", 128, 16, 576, 16)
    run("1 CTA  N=256", 128, 256, 576, 256)
    run("8 CTAs N=64", 1024, 64, 576, 64)
    run("72 CTAs N=64", 9216, 64, 576, 64)
    run("72 CTAs N=16", 9216, 16, 576, 16)
    run("144 CTAs N=64 (2 n-tiles of 32)", 9216, 64, 576, 32)
    run("288 CTAs N=64 (M=36864)", 36864, 64, 576, 64)
    ops.check_watchdog()


if __name__ == "__main__":
    main()
