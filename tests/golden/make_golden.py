"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) on the CPU.

    PYTHONPATH=/root/reference python tests/golden/make_golden.py

For each of the four networks: synthetic weights (synth.py) are loaded into the reference module, one training-mode
forward + backward is run for a synthetic input and upstream gradient, and the fixture stores the input, the output,
the BatchNorm buffers after the forward, the parameter shapes and a digest of every parameter gradient (L2 norm and
the first 4 values). Step fixtures: one full SRGANTrainer._gan_loop step, one full ESRGANTrainer._gan_loop step
(relativistic GAN, 23 RRDB blocks) and two PSNR-phase (pretrain) steps of each trainer - post-step parameter digests
(and losses for the pretrain steps). `--steps-only` regenerates just the ESRGAN / pretrain step fixtures.
"""
import json
import os
import sys
import tempfile
from argparse import Namespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = os.environ.get("TORCHSR_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

from synth import synth_input, synth_state_dict  # noqa: E402


def digest(t: torch.Tensor):
    f = t.detach().double().flatten()
    return [float(f.norm())] + [float(x) for x in f[:4]]


def run_module(name, module, in_shape, seed):
    torch.set_num_threads(4)
    sd = synth_state_dict(module.state_dict(), seed)
    module.load_state_dict(sd)
    module.train()
    x = synth_input(in_shape, seed + 1)
    y = module(x)
    gout = synth_input(tuple(y.shape), seed + 2) - 0.5
    y.backward(gout)
    after = module.state_dict()
    arrays = {"input": x.numpy(), "output": y.detach().numpy(), "gout": gout.numpy()}
    meta = {"seed": seed, "shapes": {k: list(v.shape) for k, v in sd.items()},
            "dtypes": {k: str(v.dtype) for k, v in sd.items()},
            "grad_digest": {k: digest(p.grad) for k, p in module.named_parameters()},
            "buffers_after": {k: digest(v.float()) for k, v in after.items() if "running" in k or "num_batches" in k}}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=json.dumps(meta), **arrays)
    print(name, "out", tuple(y.shape), "keys", len(sd))


def run_gan_step():
    """One SRGANTrainer._gan_loop on the CPU with synthetic G/D weights and a seeded random VGG19."""
    import torchvision
    tmp = tempfile.mkdtemp()
    os.environ["TORCH_HOME"] = tmp
    os.environ["WANDB_MODE"] = "disabled"
    os.makedirs(os.path.join(tmp, "hub", "checkpoints"))
    state = torch.random.get_rng_state()
    torch.manual_seed(1234)
    vgg = torchvision.models.vgg19(weights=None)
    torch.random.set_rng_state(state)
    torch.save(vgg.state_dict(), os.path.join(tmp, "hub", "checkpoints", "vgg19-dcbb9e9d.pth"))
    cwd = os.getcwd()
    os.chdir(REF)     # the trainer opens media/waterfalls-low-res.png relative to the CWD
    try:
        import torchsr.srgan.trainer as T
        T.wandb = None
        args = Namespace(disable_amp=False, batch_size=2, epochs=8, pretrain_epochs=1, gan_checkpoint=None,
                         psnr_checkpoint=None, skip_image_save=True, local_rank=-1, rank=-1, world_size=1)
        tr = T.SRGANTrainer(torch.device("cpu"), args, [], [], 0, 0, False)
    finally:
        os.chdir(cwd)
    tr.generator.load_state_dict(synth_state_dict(tr.generator.state_dict(), 11))
    tr.discriminator.load_state_dict(synth_state_dict(tr.discriminator.state_dict(), 12))
    lr, hr = synth_input((2, 3, 24, 24), 13), synth_input((2, 3, 96, 96), 14)
    tr._gan_loop(lr, hr, 0)
    meta = {"g_after": {k: digest(v.float()) for k, v in tr.generator.state_dict().items()},
            "d_after": {k: digest(v.float()) for k, v in tr.discriminator.state_dict().items()}}
    np.savez_compressed(os.path.join(HERE, "srgan_gan_step.npz"), meta=json.dumps(meta), low_res=lr.numpy(),
                        high_res=hr.numpy())
    print("srgan_gan_step recorded")


def _reference_trainer(kind):
    """Builds the reference's SRGANTrainer / ESRGANTrainer on the CPU (seeded random VGG19 in a private hub cache)."""
    import importlib
    import torchvision
    tmp = tempfile.mkdtemp()
    os.environ["TORCH_HOME"] = tmp
    os.environ["WANDB_MODE"] = "disabled"
    os.makedirs(os.path.join(tmp, "hub", "checkpoints"))
    state = torch.random.get_rng_state()
    torch.manual_seed(1234)
    vgg = torchvision.models.vgg19(weights=None)
    torch.random.set_rng_state(state)
    torch.save(vgg.state_dict(), os.path.join(tmp, "hub", "checkpoints", "vgg19-dcbb9e9d.pth"))
    cwd = os.getcwd()
    os.chdir(REF)     # the trainer opens media/waterfalls-low-res.png relative to the CWD
    try:
        T = importlib.import_module(f"torchsr.{kind}.trainer")
        T.wandb = None
        args = Namespace(disable_amp=False, batch_size=2, epochs=8, pretrain_epochs=1, gan_checkpoint=None,
                         psnr_checkpoint=None, skip_image_save=True, local_rank=-1, rank=-1, world_size=1)
        cls = T.SRGANTrainer if kind == "srgan" else T.ESRGANTrainer
        return cls(torch.device("cpu"), args, [], [], 0, 0, False)
    finally:
        os.chdir(cwd)


def _after(tr):
    return {"g_after": {k: digest(v.float()) for k, v in tr.generator.state_dict().items()},
            "d_after": {k: digest(v.float()) for k, v in tr.discriminator.state_dict().items()}}


def run_esrgan_gan_step():
    """One ESRGANTrainer._gan_loop (relativistic GAN step, 23-RRDB generator) on the CPU."""
    torch.set_num_threads(8)
    tr = _reference_trainer("esrgan")
    tr.generator.load_state_dict(synth_state_dict(tr.generator.state_dict(), 21))
    tr.discriminator.load_state_dict(synth_state_dict(tr.discriminator.state_dict(), 22))
    lr, hr = synth_input((2, 3, 32, 32), 23), synth_input((2, 3, 128, 128), 24)
    tr._gan_loop(lr, hr, 0)
    np.savez_compressed(os.path.join(HERE, "esrgan_gan_step.npz"), meta=json.dumps(_after(tr)), low_res=lr.numpy(),
                        high_res=hr.numpy())
    print("esrgan_gan_step recorded")


def run_pretrain_steps():
    """The PSNR-phase step of both trainers. The reference has no per-step method for it (the statements sit inside
    the epoch loop of `_pretrain`, srgan/trainer.py:376-388, esrgan/trainer.py:378-390), so the fixture drives the
    reference trainer's own generator, loss module, GradScaler and psnr_optimizer with that statement sequence."""
    torch.set_num_threads(8)
    for kind, seed, lr_shape, hr_shape in [("srgan", 31, (2, 3, 24, 24), (2, 3, 96, 96)),
                                           ("esrgan", 41, (1, 3, 32, 32), (1, 3, 128, 128))]:
        tr = _reference_trainer(kind)
        tr.generator.load_state_dict(synth_state_dict(tr.generator.state_dict(), seed))
        lr, hr = synth_input(lr_shape, seed + 1), synth_input(hr_shape, seed + 2)
        tr.generator.train()
        losses = []
        for _ in range(2):
            tr.psnr_optimizer.zero_grad()
            crit = tr.mse_loss if kind == "srgan" else tr.l1_loss
            loss = crit(tr.generator(lr), hr)
            tr.scaler.scale(loss).backward()
            tr.scaler.step(tr.psnr_optimizer)
            tr.scaler.update()
            losses.append(float(loss.detach()))
        meta = {"losses": losses, "g_after": {k: digest(v.float()) for k, v in tr.generator.state_dict().items()}}
        np.savez_compressed(os.path.join(HERE, f"{kind}_pretrain_step.npz"), meta=json.dumps(meta), low_res=lr.numpy(),
                            high_res=hr.numpy())
        print(kind, "pretrain steps recorded", losses)


def main():
    if "--steps-only" in sys.argv:      # the module fixtures and srgan_gan_step.npz are already committed
        run_esrgan_gan_step()
        run_pretrain_steps()
        return
    from torchsr.esrgan.discriminator import Discriminator as ED
    from torchsr.esrgan.generator import Generator as EG
    from torchsr.srgan.discriminator import Discriminator as SD
    from torchsr.srgan.generator import Generator as SG
    run_module("srgan_generator", SG(), (2, 3, 12, 12), 1)
    run_module("srgan_discriminator", SD(), (2, 3, 96, 96), 2)
    run_module("esrgan_generator", EG(num_rrdb_blocks=2), (1, 3, 12, 12), 3)
    run_module("esrgan_discriminator", ED(), (2, 3, 128, 128), 4)
    run_gan_step()
    run_esrgan_gan_step()
    run_pretrain_steps()


if __name__ == "__main__":
    main()
