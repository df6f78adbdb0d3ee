"""Is the step CPU- or GPU-bound? Compares host enqueue time per step with the device time per step."""
import os
import sys
import time
from argparse import Namespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TORCHSR_VGG_WEIGHTS", "random")
import torch  # noqa: E402

from torchsr_b200.srgan.trainer import SRGANTrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
targs = Namespace(disable_amp=False, batch_size=B, epochs=8, pretrain_epochs=1, gan_checkpoint=None,
                  psnr_checkpoint=None, skip_image_save=True, local_rank=0, rank=-1, world_size=1)
tr = SRGANTrainer(torch.device("cuda"), targs, [], [], 0, 0, False)
lr, hr = torch.rand(B, 3, 24, 24, device="cuda"), torch.rand(B, 3, 96, 96, device="cuda")
for s in range(5):
    tr._gan_loop(lr, hr, s)
torch.cuda.synchronize()
N = 20
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(N):
    tr._gan_loop(lr, hr, s)
e1.record()
t_enq = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"B={B}: host enqueue {t_enq / N * 1e3:.3f} ms/step, device {e0.elapsed_time(e1) / N:.3f} ms/step, wall {t_all / N * 1e3:.3f} ms/step")
# split the host time by phase
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for s in range(10):
    tr._gan_loop(lr, hr, s)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
