"""CPU oracle of the SRGAN training steps (TEST / BASELINE INFRASTRUCTURE ONLY - see torchsr_oracle.py header).

Restates SRGANTrainer._gan_loop (torchsr/srgan/trainer.py:416-469) and the pretrain step (:376-388) over the
functional module oracles, with the same optimizers (Adam lr 1e-4, betas (0.9, 0.999), :171-185) and the VGG19
perceptual loss (torchsr/srgan/loss.py:18-54; torchvision vgg19.features[:36], frozen, L1 on features).
Used (a) by tests to check the CUDA path's losses and post-step parameters, (b) by bench.py as the CPU baseline
("port": the unmodified reference cannot travel to the GPU box, /root/reference does not exist there).
"""
from typing import Dict, Optional

import torch
import torch.nn.functional as F

import torchsr_oracle as O


def vgg19_features(seed: int = 1234, feature_layer: int = 36) -> torch.nn.Module:
    """torchvision VGG19 features[:36], eval mode, frozen. The ImageNet weights the reference downloads are not
    available offline; a seeded random initialisation exercises identical arithmetic."""
    import torchvision
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    model = torchvision.models.vgg19(weights=None)
    torch.random.set_rng_state(state)
    feats = torch.nn.Sequential(*list(model.features.children())[:feature_layer]).eval()
    for p in feats.parameters():
        p.requires_grad = False
    return feats


class OracleSRGAN:
    """Holds generator / discriminator state dicts as leaf tensors and steps them like the reference trainer."""

    def __init__(self, g_sd: Dict[str, torch.Tensor], d_sd: Dict[str, torch.Tensor], vgg: Optional[torch.nn.Module]):
        self.g = O.with_grad(g_sd)
        self.d = O.with_grad(d_sd)
        self.vgg = vgg
        gp = [v for v in self.g.values() if v.requires_grad]
        dp = [v for v in self.d.values() if v.requires_grad]
        self.psnr_opt = torch.optim.Adam(gp, lr=1e-4, betas=(0.9, 0.999))     # trainer.py:171-175
        self.disc_opt = torch.optim.Adam(dp, lr=1e-4, betas=(0.9, 0.999)) if dp else None   # :176-180 (no D: pretrain only)
        self.gen_opt = torch.optim.Adam(gp, lr=1e-4, betas=(0.9, 0.999))      # :181-185
        self.gp, self.dp = gp, dp

    def _update_buffers(self, sd, buffers):
        for k, v in buffers.items():
            sd[k] = v.detach()

    def content_loss(self, sr, hr):
        if self.vgg is None:
            return F.mse_loss(sr, hr)
        return F.l1_loss(self.vgg(sr), self.vgg(hr))                          # loss.py:52-53

    def pretrain_step(self, low_res, high_res) -> float:
        """trainer.py:376-388 without autocast/GradScaler (they are disabled on CPU)."""
        self.psnr_opt.zero_grad()
        buf = {}
        sr = O.srgan_generator(self.g, low_res, True, buf)
        loss = F.mse_loss(sr, high_res)
        loss.backward()
        self.psnr_opt.step()
        self._update_buffers(self.g, buf)
        return float(loss.detach())

    def gan_step(self, low_res, high_res):
        """trainer.py:435-469. Returns (disc_loss, gen_loss) as floats."""
        for p in self.dp:                                                      # :442 discriminator.zero_grad()
            p.grad = None
        gbuf, dbuf = {}, {}
        sr = O.srgan_generator(self.g, low_res, True, gbuf)                    # :444
        d_real = O.srgan_discriminator(self.d, high_res, True, dbuf)           # :446
        d_fake = O.srgan_discriminator(self.d, sr.detach(), True, dbuf)        # :447
        disc_loss = O.bce(d_real, 1.0) + O.bce(d_fake, 0.0)                    # :446-448
        disc_loss.backward()                                                   # :450
        self.disc_opt.step()                                                   # :451
        for p in self.gp:                                                      # :453 generator.zero_grad()
            p.grad = None
        content = self.content_loss(sr, high_res.detach())                     # :455
        adv = O.bce(O.srgan_discriminator(self.d, sr, True, dbuf), 1.0)        # :456 (updated D weights)
        gen_loss = content + 0.001 * adv                                       # :457
        gen_loss.backward()                                                    # :468
        self.gen_opt.step()                                                    # :469
        self._update_buffers(self.g, gbuf)
        self._update_buffers(self.d, dbuf)
        return float(disc_loss.detach()), float(gen_loss.detach())


class OracleESRGAN:
    """ESRGANTrainer steps (torchsr/esrgan/trainer.py): L1 pretrain step (:378-390) and the relativistic-average GAN
    step `_gan_loop` (:435-484), with the reference's three Adam optimizers (:171-185). autocast / GradScaler are
    disabled on the CPU in the reference too (`amp` needs CUDA), so the arithmetic is plain fp32."""

    def __init__(self, g_sd: Dict[str, torch.Tensor], d_sd: Dict[str, torch.Tensor], vgg: Optional[torch.nn.Module]):
        self.g = O.with_grad(g_sd)
        self.d = O.with_grad(d_sd)
        self.vgg = vgg
        self.gp = [v for v in self.g.values() if v.requires_grad]
        self.dp = [v for v in self.d.values() if v.requires_grad]
        self.psnr_opt = torch.optim.Adam(self.gp, lr=1e-4, betas=(0.9, 0.999))
        self.disc_opt = torch.optim.Adam(self.dp, lr=1e-4, betas=(0.9, 0.999)) if self.dp else None
        self.gen_opt = torch.optim.Adam(self.gp, lr=1e-4, betas=(0.9, 0.999))

    def _disc(self, x):
        """Train-mode discriminator forward; the BatchNorm running statistics advance with every call, as in the
        reference (three calls in the D step + two in the G step share one module)."""
        buf = {}
        y = O.esrgan_discriminator(self.d, x, True, buf)
        for k, v in buf.items():
            self.d[k] = v.detach()
        return y

    def content_loss(self, sr, hr):
        if self.vgg is None:
            return F.mse_loss(sr, hr)
        return F.l1_loss(self.vgg(sr), self.vgg(hr))                          # esrgan/loss.py (same as srgan/loss.py)

    def pretrain_step(self, low_res, high_res) -> float:
        """:378-390."""
        self.psnr_opt.zero_grad()
        loss = F.l1_loss(O.esrgan_generator(self.g, low_res), high_res)
        loss.backward()
        self.psnr_opt.step()
        return float(loss.detach())

    def gan_step(self, low_res, high_res):
        """:435-484. Returns (disc_loss, gen_loss) as floats."""
        bcel = F.binary_cross_entropy_with_logits
        n = low_res.shape[0]
        real_label, fake_label = torch.ones(n, 1), torch.zeros(n, 1)          # :440-441
        self.disc_opt.zero_grad()                                             # :443
        sr = O.esrgan_generator(self.g, low_res)                              # :446
        real_out = self._disc(high_res)                                       # :447
        fake_out = self._disc(sr.detach())                                    # :448
        disc_loss = (bcel(real_out - fake_out.mean(), real_label) +
                     bcel(fake_out - real_out.mean(), fake_label)) / 2        # :450-452
        disc_loss.backward()                                                  # :454
        self.disc_opt.step()                                                  # :455
        self.gen_opt.zero_grad()                                              # :458
        sr = O.esrgan_generator(self.g, low_res)                              # :461 (unchanged weights: same value)
        real_out = self._disc(high_res.detach())                              # :462
        fake_out = self._disc(sr)                                             # :463
        pixel = F.l1_loss(sr, high_res)                                       # :465
        content = self.content_loss(sr, high_res)                             # :466
        adv = bcel(fake_out - real_out.mean(), real_label)                    # :467
        gen_loss = 0.01 * pixel + 1 * content + 0.005 * adv                   # :468
        gen_loss.backward()                                                   # :479
        self.gen_opt.step()                                                   # :480
        return float(disc_loss.detach()), float(gen_loss.detach())
