"""ESRGAN training loop - mirror of the hot parts of torchsr/esrgan/trainer.py.

Same class / method names as the reference's ESRGANTrainer for the hot path: construction (:136-196), the L1
pretrain step (:378-390), the relativistic GAN step `_gan_loop` (:418-484), evaluation and checkpoints (shared with
the SRGAN mirror). Math-identical differences: the generator forward is not recomputed for the
generator step (the reference runs it twice on unchanged weights, :462), the discriminator is frozen while it scores
images for the generator step (its discarded weight gradients are not computed), and bf16 kernels replace fp16
autocast + GradScaler."""
import torch
from torch import Tensor, nn

from .. import dist as tdist
from ..srgan.trainer import SRGANTrainer
from .discriminator import Discriminator
from .generator import Generator
from .loss import VGGLoss


class ESRGANTrainer(SRGANTrainer):
    PREFIX = 'esrgan'

    def _initialize_models(self) -> None:
        self.generator = Generator().to(self.device)
        self.discriminator = Discriminator().to(self.device)
        self._configure_modules()
        if self.distributed:
            tdist.attach(self.generator, broadcast_buffers=True)
            tdist.attach(self.discriminator, broadcast_buffers=False)

    def _initialize_loss(self) -> None:
        self.l1_loss = nn.L1Loss().to(self.device)
        self.bce_loss = nn.BCEWithLogitsLoss().to(self.device)
        self.vgg_loss = VGGLoss().to(self.device)

    def _pretrain_step(self, low_res: Tensor, high_res: Tensor) -> Tensor:
        """reference :378-390: G forward, L1, backward, Adam."""
        low_res = low_res.to(self.device, non_blocking=True)
        high_res = high_res.to(self.device, non_blocking=True)
        self.psnr_optimizer.zero_grad()
        loss = self.l1_loss(self.generator(low_res), high_res)
        loss.backward()
        self.psnr_optimizer.step()
        return loss.detach()

    def _gan_loop(self, low_res: Tensor, high_res: Tensor, step: int) -> Tensor:
        """reference :435-484 (relativistic average GAN)."""
        low_res = low_res.to(self.device, non_blocking=True)
        high_res = high_res.to(self.device, non_blocking=True)
        batch_size = low_res.size(0)
        real_label = torch.full((batch_size, 1), 1, dtype=low_res.dtype, device=self.device)
        fake_label = torch.full((batch_size, 1), 0, dtype=low_res.dtype, device=self.device)

        self.disc_optimizer.zero_grad()
        super_res = self.generator(low_res)
        real_output = self.discriminator(high_res)
        fake_output = self.discriminator(super_res.detach())
        disc_loss_real = self.bce_loss(real_output - torch.mean(fake_output), real_label)
        disc_loss_fake = self.bce_loss(fake_output - torch.mean(real_output), fake_label)
        disc_loss = (disc_loss_real + disc_loss_fake) / 2
        disc_loss.backward()
        self.disc_optimizer.step()

        self.gen_optimizer.zero_grad()
        with tdist.frozen(self.discriminator):
            with torch.no_grad():
                real_output = self.discriminator(high_res)
            fake_output = self.discriminator(super_res)
        pixel_loss = self.l1_loss(super_res, high_res)
        content_loss = self.vgg_loss(super_res, high_res)
        adversarial_loss = self.bce_loss(fake_output - torch.mean(real_output), real_label)
        gen_loss = 0.01 * pixel_loss + 1 * content_loss + 0.005 * adversarial_loss
        gen_loss.backward()
        self.gen_optimizer.step()
        self.generator.zero_grad()
        return gen_loss.detach()
