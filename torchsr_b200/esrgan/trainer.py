"""ESRGAN training loop - mirror of the hot parts of torchsr/esrgan/trainer.py.

Same class / method names as the reference's ESRGANTrainer for the hot path: construction (:136-196), the L1
pretrain step (:378-390), the relativistic GAN step `_gan_loop` (:418-484), evaluation and checkpoints (shared with
the SRGAN mirror). Math-identical differences: the generator forward is not recomputed for the
generator step (the reference runs it twice on unchanged weights, :462), the discriminator is frozen while it scores
images for the generator step (its discarded weight gradients are not computed), and bf16 kernels replace fp16
autocast + GradScaler."""
import torch
from torch import Tensor

from .. import dist as tdist
from .. import losses
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
        # same attribute names as the reference (:163-165); the criteria run on this repo's reduction kernels
        self.l1_loss = losses.L1Loss()
        self.bce_loss = losses.BCEWithLogitsLoss()
        self.vgg_loss = VGGLoss().to(self.device)
        self._one = torch.ones((), dtype=torch.float32, device=self.device)

    def _pretrain_step(self, low_res: Tensor, high_res: Tensor) -> Tensor:
        """reference :378-390: G forward, L1, backward, Adam."""
        low_res = low_res.to(self.device, non_blocking=True)
        high_res = high_res.to(self.device, non_blocking=True)
        self.psnr_optimizer.zero_grad()
        loss = self.l1_loss(self.generator(low_res), high_res)
        loss.backward(self._one)
        self.psnr_optimizer.step()
        tdist.join_buffers(self.generator)
        return loss.detach()

    def _gan_loop(self, low_res: Tensor, high_res: Tensor, step: int) -> Tensor:
        """reference :435-484 (relativistic average GAN)."""
        low_res = low_res.to(self.device, non_blocking=True)
        high_res = high_res.to(self.device, non_blocking=True)

        self.disc_optimizer.zero_grad()
        super_res = self.generator(low_res)
        # :447-449 two separate D calls (BatchNorm statistics per call, real first) -> one pass, statistics per half
        real_output, fake_output = self.discriminator.forward_pair(high_res, super_res.detach())
        # (BCEwL(real - mean(fake), 1) + BCEwL(fake - mean(real), 0)) / 2   (:451-453), one reduction launch
        disc_loss = losses.relativistic_d(real_output, fake_output, scale=0.5)
        with self.disc_optimizer.late_in_backward(self.discriminator):
            disc_loss.backward(self._one)
        self.disc_optimizer.step()

        self.gen_optimizer.zero_grad()
        with tdist.frozen(self.discriminator):
            # :463-464 D(high_res.detach()) then D(super_res), updated D weights: one pass again; the real half only
            # feeds a mean that is a constant of the generator step
            real_output, fake_output = self.discriminator.forward_pair(high_res, super_res, grad_halves=(False, True))
        pixel_loss = losses.l1(super_res, high_res, scale=0.01)
        content_loss = self.vgg_loss(super_res, high_res)
        adversarial_loss = losses.relativistic_g(fake_output, real_output, scale=0.005)
        gen_loss = losses.total(pixel_loss, content_loss, adversarial_loss)     # 0.01 L1 + 1 VGG + 0.005 adv (:466-469)
        gen_loss.backward(self._one)
        self.gen_optimizer.step()
        self.disc_optimizer.join()
        tdist.join_buffers(self.generator)
        self.generator.zero_grad()
        return gen_loss.detach()
