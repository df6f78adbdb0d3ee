"""One x4 inference pass (BASELINE configs[4] shape) between cudaProfilerStart/Stop, every kernel launched individually
(TSR_GRAPHS=0 set by the caller): ncu launch list of the inference path. Usage: python tools/ncu_infer.py [batch] [size]"""
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
        for _ in range(2):
            y = G(x)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        y = G(x)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    print("ncu_infer done", tuple(y.shape), float(y.float().mean()))


if __name__ == "__main__":
    main()
