"""CPU oracle of the SRGAN / ESRGAN generator + discriminator hot path of roclark/torchsr.

TEST INFRASTRUCTURE ONLY. Nothing in the product package (torchsr_b200/) may import this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, and only as the checker or as the
timed CPU baseline.

What it is: a plain restatement, as pure functions over a ``state_dict``, of the reference's nn.Module graphs -
which layer feeds which, with which stride / padding / activation / residual - executed by torch's CPU fp32
operators (torch.nn.functional). The arithmetic of each operator lives in PyTorch (requirements.txt:7 pins
torch==1.11.0, setup.py:46 asks torch>=1.10; this image has 2.11.0), which is the third-party dependency the
reference itself calls; the oracle calls the same operators.

Pinning: the reference ships no tests, golden vectors or fixtures for this path, so the oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF: tests/golden/make_golden.py imports the unmodified modules from
/root/reference, loads synthetic weights into them, runs forward/backward (and one full ``_gan_loop``) on the CPU
and commits inputs + outputs + gradient digests under tests/golden/. tests/test_oracle.py checks every function here
against those vectors, and tests/test_reference_live.py re-checks against the live reference when /root/reference
is present.

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
BN_EPS = 1e-5        # nn.BatchNorm2d default
BN_MOMENTUM = 0.1    # nn.BatchNorm2d default


def _bn(sd: SD, prefix: str, x: torch.Tensor, training: bool, buffers: Optional[SD]) -> torch.Tensor:
    """nn.BatchNorm2d: batch statistics (biased variance) in training mode, running statistics in eval mode; the
    running estimates are updated with momentum 0.1 and the UNBIASED variance, num_batches_tracked += 1.
    `buffers`, when given, receives the updated running statistics (the caller's sd is never mutated)."""
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if buffers is not None and training:
        rm = buffers.get(prefix + ".running_mean", rm).clone()
        rv = buffers.get(prefix + ".running_var", rv).clone()
        out = F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], True, BN_MOMENTUM, BN_EPS)
        buffers[prefix + ".running_mean"] = rm
        buffers[prefix + ".running_var"] = rv
        nbt = buffers.get(prefix + ".num_batches_tracked", sd[prefix + ".num_batches_tracked"])
        buffers[prefix + ".num_batches_tracked"] = nbt + 1
        return out
    if training:
        return F.batch_norm(x, None, None, sd[prefix + ".weight"], sd[prefix + ".bias"], True, BN_MOMENTUM, BN_EPS)
    return F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], False, BN_MOMENTUM, BN_EPS)


# ------------------------------------------------------------------------------------------------------ SRGAN
def srgan_residual_block(sd: SD, p: str, x: torch.Tensor, training=True, buffers=None) -> torch.Tensor:
    """torchsr/srgan/residual.py:51-92 - x + BN2(conv2(PReLU(BN1(conv1(x))))), 3x3 convs without bias."""
    out = F.conv2d(x, sd[p + ".conv1.weight"], None, 1, 1)                 # residual.py:64,86
    out = _bn(sd, p + ".bn1", out, training, buffers)                      # :65,87
    out = F.prelu(out, sd[p + ".prelu.weight"])                            # :66,88
    out = F.conv2d(out, sd[p + ".conv2.weight"], None, 1, 1)               # :67,89
    out = _bn(sd, p + ".bn2", out, training, buffers)                      # :68,90
    return out + x                                                         # :91


def srgan_subpixel(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """torchsr/srgan/residual.py:16-48 - PReLU(PixelShuffle2(conv3x3 C->4C + bias))."""
    out = F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], 1, 1)  # :27,45
    out = F.pixel_shuffle(out, 2)                                          # :28,46
    return F.prelu(out, sd[p + ".prelu.weight"])                           # :29,47


def srgan_generator(sd: SD, x: torch.Tensor, training=True, buffers=None) -> torch.Tensor:
    """torchsr/srgan/generator.py:33-81."""
    c1 = F.prelu(F.conv2d(x, sd["conv1.0.weight"], sd["conv1.0.bias"], 1, 4), sd["conv1.1.weight"])   # :37-40,76
    out = c1
    n_blocks = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    for i in range(n_blocks):                                              # :42-45,77
        out = srgan_residual_block(sd, f"blocks.{i}", out, training, buffers)
    out = F.conv2d(out, sd["conv2.0.weight"], None, 1, 1)                  # :47-50,78
    out = _bn(sd, "conv2.1", out, training, buffers)
    out = c1 + out                                                         # :79
    n_up = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("conv_layers."))
    for j in range(n_up):                                                  # :52-56,80
        out = srgan_subpixel(sd, f"conv_layers.{j}", out)
    return F.conv2d(out, sd["conv3.weight"], sd["conv3.bias"], 1, 4)       # :58,81


SRGAN_D_STRIDES = {0: 1, 2: 2, 5: 1, 8: 2, 11: 1, 14: 2, 17: 1, 20: 2}   # discriminator.py:31-62


def _disc_features(sd: SD, x: torch.Tensor, conv_idx, strides, training, buffers) -> torch.Tensor:
    out = F.leaky_relu(F.conv2d(x, sd["features.0.weight"], sd["features.0.bias"], 1, 1), 0.2)
    for k in conv_idx[1:]:
        out = F.conv2d(out, sd[f"features.{k}.weight"], None, strides[k], 1)
        out = _bn(sd, f"features.{k + 1}", out, training, buffers)
        out = F.leaky_relu(out, 0.2)
    return out


def srgan_discriminator(sd: SD, x: torch.Tensor, training=True, buffers=None) -> torch.Tensor:
    """torchsr/srgan/discriminator.py:27-88 - 8 convs (BN + LeakyReLU 0.2 after all but the first), flatten in
    (C,H,W) order, Linear -> LeakyReLU -> Linear -> Sigmoid."""
    out = _disc_features(sd, x, sorted(SRGAN_D_STRIDES), SRGAN_D_STRIDES, training, buffers)   # :31-62,85
    out = torch.flatten(out, 1)                                            # :86
    out = F.leaky_relu(F.linear(out, sd["classifier.0.weight"], sd["classifier.0.bias"]), 0.2)  # :65-66
    out = F.linear(out, sd["classifier.2.weight"], sd["classifier.2.bias"])                     # :67
    return torch.sigmoid(out)                                              # :68


# ------------------------------------------------------------------------------------------------------ ESRGAN
def esrgan_rdb(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """torchsr/esrgan/residual.py:17-86 - dense block: conv_k sees cat(x, conv_1..conv_{k-1}); conv5*0.2 + x."""
    feats = [x]
    for k in range(1, 5):                                                  # :81-85
        w, b = sd[f"{p}.conv{k}.0.weight"], sd[f"{p}.conv{k}.0.bias"]
        feats.append(F.leaky_relu(F.conv2d(torch.cat(feats, 1), w, b, 1, 1), 0.2))
    c5 = F.conv2d(torch.cat(feats, 1), sd[p + ".conv5.weight"], sd[p + ".conv5.bias"], 1, 1)
    return c5 * 0.2 + x                                                    # :86 (scale_ratio = 0.2)


def esrgan_rrdb(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """torchsr/esrgan/residual.py:89-129 - RDB3(RDB2(RDB1(x))) * 0.2 + x (0.2 hard-coded at :129)."""
    out = esrgan_rdb(sd, p + ".RDB1", x)
    out = esrgan_rdb(sd, p + ".RDB2", out)
    out = esrgan_rdb(sd, p + ".RDB3", out)
    return out * 0.2 + x


def esrgan_generator(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """torchsr/esrgan/generator.py:32-81 (no BatchNorm anywhere; nearest x2 before each upsample conv)."""
    c1 = F.conv2d(x, sd["conv1.weight"], sd["conv1.bias"], 1, 1)           # :34,70
    out = c1
    n_blocks = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    for i in range(n_blocks):                                              # :36-39,71
        out = esrgan_rrdb(sd, f"blocks.{i}", out)
    out = c1 + F.conv2d(out, sd["conv2.weight"], sd["conv2.bias"], 1, 1)   # :41,72-73
    for name in ("upsample1", "upsample2"):                                # :74-79
        out = F.interpolate(out, scale_factor=2, mode="nearest")
        out = F.leaky_relu(F.conv2d(out, sd[name + ".weight"], sd[name + ".bias"], 1, 1), 0.2)
    out = F.leaky_relu(F.conv2d(out, sd["conv3.0.weight"], sd["conv3.0.bias"], 1, 1), 0.2)   # :47-50,80
    return F.conv2d(out, sd["conv4.weight"], sd["conv4.bias"], 1, 1)       # :52,81


ESRGAN_D_STRIDES = {0: 1, 2: 2, 5: 1, 8: 2, 11: 1, 14: 2, 17: 1, 20: 2, 23: 1, 26: 2}   # esrgan/discriminator.py:31-70


def esrgan_discriminator(sd: SD, x: torch.Tensor, training=True, buffers=None) -> torch.Tensor:
    """torchsr/esrgan/discriminator.py:27-95 - 10 convs, Linear 8192->100 -> LeakyReLU -> Linear 100->1, logits."""
    out = _disc_features(sd, x, sorted(ESRGAN_D_STRIDES), ESRGAN_D_STRIDES, training, buffers)
    out = torch.flatten(out, 1)
    out = F.leaky_relu(F.linear(out, sd["classifier.0.weight"], sd["classifier.0.bias"]), 0.2)
    return F.linear(out, sd["classifier.2.weight"], sd["classifier.2.bias"])


# ------------------------------------------------------------------------------------------------------ losses / steps
def bce(p: torch.Tensor, y: float) -> torch.Tensor:
    """nn.BCELoss (mean) against a constant label, with PyTorch's clamp of log at -100 (srgan/trainer.py:164)."""
    return F.binary_cross_entropy(p, torch.full_like(p, y))


def srgan_gan_step_losses(g_sd: SD, d_sd: SD, low_res: torch.Tensor, high_res: torch.Tensor,
                          content_loss=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """The forward half of SRGANTrainer._gan_loop (srgan/trainer.py:435-457) with D held fixed: returns
    (super_res, disc_loss, gen_loss). `content_loss(sr, hr)` stands in for VGGLoss (srgan/loss.py:18-54, out of
    scope: torchvision VGG19 with downloaded weights); None means MSE, the pretrain criterion (trainer.py:384).
    Separate D calls for real and fake keep separate BatchNorm batch statistics, as in the reference (:446-447)."""
    sr = srgan_generator(g_sd, low_res, True)                                       # :444
    d_loss = bce(srgan_discriminator(d_sd, high_res, True), 1.0) + \
        bce(srgan_discriminator(d_sd, sr.detach(), True), 0.0)                      # :446-448
    content = F.mse_loss(sr, high_res) if content_loss is None else content_loss(sr, high_res)
    g_loss = content + 0.001 * bce(srgan_discriminator(d_sd, sr, True), 1.0)        # :455-457
    return sr, d_loss, g_loss


def psnr(sr: torch.Tensor, hr: torch.Tensor) -> float:
    """10 * log10(1 / mse) as in SRGANTrainer._test (srgan/trainer.py:296)."""
    return float(10.0 * torch.log10(1.0 / F.mse_loss(sr, hr)))


def with_grad(sd: SD) -> SD:
    """Detached copy of a state dict whose floating-point entries require grad (for autograd parity checks)."""
    out = {}
    for k, v in sd.items():
        v = v.detach().clone()
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
        out[k] = v
    return out
