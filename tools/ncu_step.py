"""One eager SRGAN GAN step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`): every kernel of the step
is launched individually through the library (TSR_GRAPHS=0 set by the caller keeps the per-program CUDA graphs off), so
an ncu launch list maps one line to one kernel of the step. Usage: python tools/ncu_step.py [batch] [steps_profiled]"""
import os
import sys
from argparse import Namespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TORCHSR_VGG_WEIGHTS", "random")
import torch  # noqa: E402

from torchsr_b200.srgan.trainer import SRGANTrainer  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    torch.manual_seed(0)
    targs = Namespace(disable_amp=False, batch_size=B, epochs=8, pretrain_epochs=1, gan_checkpoint=None,
                      psnr_checkpoint=None, skip_image_save=True, local_rank=0, rank=-1, world_size=1)
    tr = SRGANTrainer(torch.device("cuda"), targs, [], [], 0, 0, False)
    lr, hr = torch.rand(B, 3, 24, 24, device="cuda"), torch.rand(B, 3, 96, 96, device="cuda")
    for s in range(3):
        tr._gan_loop(lr, hr, s)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for s in range(n):
        loss = tr._gan_loop(lr, hr, s)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("ncu_step done, loss", float(loss))


if __name__ == "__main__":
    main()
