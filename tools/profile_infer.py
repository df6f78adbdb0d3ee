"""Per-kernel GPU timeline (torch.profiler / CUPTI, warm caches) of the x4 inference pass.
Usage: python tools/profile_infer.py [batch] [size]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from torchsr_b200.srgan.generator import Generator  # noqa: E402


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    torch.manual_seed(1234)
    G = Generator().cuda().eval()
    x = torch.rand(b, 3, size, size, device="cuda")
    with torch.no_grad():
        for _ in range(3):
            G(x)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                G(x)
            torch.cuda.synchronize()
    kev = [k for k in prof.profiler.kineto_results.events() if k.device_type() == torch.autograd.DeviceType.CUDA]
    kev.sort(key=lambda k: k.start_ns())
    last = kev[len(kev) * 2 // 3:]
    z = last[0].start_ns()
    tot = 0.0
    for k in last:
        tot += k.duration_ns() / 1e3
        print(f"{(k.start_ns() - z) / 1e3:9.1f} {k.duration_ns() / 1e3:8.1f}  {k.name()[:80]}")
    print(f"sum of durations {tot:.1f} us, span {(last[-1].start_ns() + last[-1].duration_ns() - z) / 1e3:.1f} us")


if __name__ == "__main__":
    main()
