"""x4 inference throughput of the SRGAN generator (BASELINE configs[4]): synthetic 512x512 LR -> 2048x2048, eval mode,
no_grad, output Mpx/s with the input resident in HBM; plus the B=64 training step (configs[2]) when asked."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from torchsr_b200.srgan.generator import Generator  # noqa: E402


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    torch.manual_seed(1234)
    G = Generator().cuda().eval()
    x = torch.rand(b, 3, size, size, device="cuda")
    with torch.no_grad():
        for _ in range(3):
            y = G(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            y = G(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    mpx = b * (4 * size) ** 2 / 1e6
    gflop = 277.3 * mpx          # 277.3 GFLOP per output Mpx
    print(f"SRGAN x4 inference: batch {b} of {size}x{size} LR -> {tuple(y.shape)}: {ms:.2f} ms/batch, "
          f"{mpx / (ms * 1e-3):.1f} output Mpx/s, {gflop / ms:.1f} TFLOP/s algorithmic, "
          f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")


if __name__ == "__main__":
    main()
