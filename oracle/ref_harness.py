"""Drives the UNMODIFIED reference (oracle/_ref/torchsr, staged by oracle/make_ref.py) for baselines and parity runs.

TEST / BASELINE INFRASTRUCTURE ONLY. The harness supplies what the reference needs to run offline without touching its
code: a torch-hub cache holding a seeded random-init vgg19-dcbb9e9d.pth (D7: no network), WANDB
disabled (D8), a working directory containing media/waterfalls-low-res.png (D6) and --epochs >= 8 (D5)."""
import contextlib
import importlib
import os
import sys
import tempfile
from argparse import Namespace

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "torchsr"))


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def _hub_with_random_vgg():
    import torch
    import torchvision
    tmp = tempfile.mkdtemp(prefix="tsr_ref_hub_")
    os.environ["TORCH_HOME"] = tmp
    os.environ["WANDB_MODE"] = "disabled"
    os.makedirs(os.path.join(tmp, "hub", "checkpoints"))
    state = torch.random.get_rng_state()
    torch.manual_seed(1234)
    vgg = torchvision.models.vgg19(weights=None)
    torch.random.set_rng_state(state)
    torch.save(vgg.state_dict(), os.path.join(tmp, "hub", "checkpoints", "vgg19-dcbb9e9d.pth"))
    return tmp


def reference_trainer(kind: str, device, batch_size: int):
    """The reference's SRGANTrainer / ESRGANTrainer (torchsr/{srgan,esrgan}/trainer.py), constructed as its CLI would
    (torchsr.py:263-270) with empty loaders."""
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/make_ref.py` where /root/reference exists")
    import warnings
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    _hub_with_random_vgg()
    with _cwd(REF_DIR), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        T = importlib.import_module(f"torchsr.{kind}.trainer")
        T.wandb = None
        args = Namespace(disable_amp=False, batch_size=batch_size, epochs=8, pretrain_epochs=1, gan_checkpoint=None,
                         psnr_checkpoint=None, skip_image_save=True, local_rank=-1, rank=-1, world_size=1)
        cls = T.SRGANTrainer if kind == "srgan" else T.ESRGANTrainer
        return cls(device, args, [], [], 0, 0, False)


def reference_generator(kind: str = "srgan"):
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    return importlib.import_module(f"torchsr.{kind}.generator").Generator
