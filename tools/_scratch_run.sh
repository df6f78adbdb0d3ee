timeout 300 python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
from torchsr_b200.esrgan.generator import Generator
torch.manual_seed(0)
G = Generator().cuda().eval()
for size in (256, 512):
    x = torch.rand(1, 3, size, size, device='cuda')
    with torch.no_grad():
        for _ in range(2): y = G(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): y = G(x)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    mpx = (4 * size) ** 2 / 1e6
    print(f"ESRGAN x4 inference {size}x{size} -> {tuple(y.shape)}: {ms:.2f} ms, {mpx / ms * 1e3:.1f} output Mpx/s, peak {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
PY
