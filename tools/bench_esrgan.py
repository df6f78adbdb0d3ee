"""ESRGAN GAN training step (BASELINE configs[3] shape: 23 RRDB generator + VGG loss, 128x128 HR crops) on one GPU."""
import os
import sys
from argparse import Namespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TORCHSR_VGG_WEIGHTS", "random")
import torch  # noqa: E402

from torchsr_b200.esrgan.trainer import ESRGANTrainer  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    torch.manual_seed(0)
    args = Namespace(disable_amp=False, batch_size=B, epochs=8, pretrain_epochs=1, gan_checkpoint=None,
                     psnr_checkpoint=None, skip_image_save=True, local_rank=0, rank=-1, world_size=1)
    tr = ESRGANTrainer(torch.device("cuda"), args, [], [], 0, 0, False)
    lr, hr = torch.rand(B, 3, 32, 32, device="cuda"), torch.rand(B, 3, 128, 128, device="cuda")
    for s in range(3):
        tr.graph_step(lr, hr, s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for s in range(n):
        loss = tr.graph_step(lr, hr, s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"ESRGAN GAN step, batch {B} of 128x128 HR crops: {ms:.2f} ms/step, {B / ms * 1e3:.1f} crops/s, "
          f"{141.1 * B / ms:.1f} TFLOP/s (G+D algorithmic 141.1 GFLOP/crop), loss {float(loss):.4f}, "
          f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")


if __name__ == "__main__":
    main()
