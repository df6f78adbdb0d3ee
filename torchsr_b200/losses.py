"""Loss reductions of the training steps on this repo's kernels (csrc/eltwise.cu: loss_kernel, gan_loss_kernel,
axpby_f32_kernel), bound to autograd. They replace the ATen launches behind nn.MSELoss / nn.L1Loss / nn.BCELoss /
nn.BCEWithLogitsLoss, the label-tensor fills, `torch.mean` of the relativistic criterion and the scalar adds /
multiplies around them (reference torchsr/srgan/trainer.py:163-165,384,446-457; esrgan/trainer.py:163-165,451-453,
466-469), so that a training step launches no PyTorch kernel.

Every function takes fp32 CUDA tensors and returns a 0-dim fp32 tensor with a grad_fn. The forward launch also
produces the gradient w.r.t. the inputs for an upstream gradient of 1; backward multiplies it by the upstream
gradient read from device memory (one launch, no host round trip: the whole step stays capturable in a CUDA graph).
There is no CPU path."""
from typing import Optional

import torch
from torch import Tensor, nn

from . import _lib as L
from . import ops

F32 = torch.float32
MSE, L1 = 0, 1
GAN_BCE, GAN_RELATIVISTIC, GAN_RELATIVISTIC_G = 0, 1, 2


def _check(t: Tensor, what: str) -> Tensor:
    if t.device.type != "cuda" and not ops.DRY:
        raise L.TorchSRB200Error(f"torchsr_b200.losses: {what} is on '{t.device}'; the loss kernels run only on a "
                                 "CUDA sm_100 device (no CPU fallback)")
    t = t.detach()
    if t.dtype != F32:
        t = t.float()
    return t.contiguous()


def _scaled(grad: Tensor, gout: Tensor) -> Tensor:
    """grad * gout for a 0-dim device tensor gout (the upstream gradient of a scalar loss)."""
    out = torch.empty_like(grad)
    g = gout.detach()
    if g.dtype != F32 or not g.is_contiguous():
        g = g.float().contiguous()
    ops.run_now(ops.elt(L.E_AXPBY_F32, p=[grad, None, out, g], i=[grad.numel()], f=[1.0, 0.0]))
    return out


def add_(dst: Tensor, src: Tensor, alpha: float = 1.0) -> Tensor:
    """dst += alpha * src on fp32 tensors of equal size (flat-gradient merges, gradient sums at a fan-in)."""
    assert dst.dtype == F32 and src.dtype == F32 and dst.numel() == src.numel()
    assert dst.is_contiguous() and src.is_contiguous()
    ops.run_now(ops.elt(L.E_AXPBY_F32, p=[dst, src, dst, None], i=[dst.numel()], f=[1.0, alpha]))
    return dst


def add(a: Tensor, b: Tensor, alpha: float = 1.0) -> Tensor:
    """a + alpha * b into a new tensor, no autograd (used on gradients and on detached loss values)."""
    a, b = _check(a, "a"), _check(b, "b")
    out = torch.empty_like(a)
    ops.run_now(ops.elt(L.E_AXPBY_F32, p=[a, b, out, None], i=[a.numel()], f=[1.0, alpha]))
    return out


class _Pixel(torch.autograd.Function):
    """scale * mean((a - b)^2) or scale * mean(|a - b|); gradient w.r.t. `a` only (b is a target)."""

    @staticmethod
    def forward(ctx, a: Tensor, b: Tensor, kind: int, scale: float):
        if a.shape != b.shape:
            raise RuntimeError(f"loss inputs differ in shape: {tuple(a.shape)} vs {tuple(b.shape)}")
        x, y = _check(a, "input"), _check(b, "target")
        n = x.numel()
        blocks = max(1, min(148 * 4, (n + 1023) // 1024))
        partial = torch.empty(blocks, dtype=F32, device=x.device)
        grad = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        ops.run_now(ops.elt(L.E_LOSS, p=[x, y, partial, grad], i=[n, kind, blocks], f=[scale / n]))
        out = torch.empty((), dtype=F32, device=x.device)
        ops.run_now(ops.elt(L.E_SUM_FINALIZE, p=[partial, out], i=[blocks, 0], f=[scale / n]))
        ctx.grad = grad
        ctx.shape = a.shape
        return out

    @staticmethod
    def backward(ctx, gout: Tensor):
        g = _scaled(ctx.grad, gout).view(ctx.shape) if ctx.grad is not None else None
        return g, None, None, None


class _Gan(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a: Tensor, b: Optional[Tensor], mode: int, ya: float, yb: float, scale: float,
                target: Optional[Tensor]):
        x = _check(a, "input")
        y = _check(b, "second input") if b is not None else None
        t = _check(target, "target") if target is not None else None
        if t is not None and t.numel() != x.numel():
            raise RuntimeError("target size differs from input size")
        need_b = y is not None and mode != GAN_RELATIVISTIC_G and ctx.needs_input_grad[1]
        ga = torch.empty_like(x)
        gb = torch.empty_like(y) if (y is not None and mode != GAN_RELATIVISTIC_G) else None
        out = torch.empty((), dtype=F32, device=x.device)
        ops.run_now(ops.elt(L.E_GAN_LOSS, p=[x, y, out, ga, gb, t], i=[x.numel(), y.numel() if y is not None else 0, mode],
                            f=[ya, yb, scale]))
        ctx.ga = ga if ctx.needs_input_grad[0] else None
        ctx.gb = gb if need_b else None
        ctx.shapes = (a.shape, b.shape if b is not None else None)
        return out

    @staticmethod
    def backward(ctx, gout: Tensor):
        ga = _scaled(ctx.ga, gout).view(ctx.shapes[0]) if ctx.ga is not None else None
        gb = _scaled(ctx.gb, gout).view(ctx.shapes[1]) if ctx.gb is not None else None
        return ga, gb, None, None, None, None, None


class _Sum(torch.autograd.Function):
    """Sum of 0-dim losses on the device (one launch per extra term); backward hands the upstream gradient to each."""

    @staticmethod
    def forward(ctx, *xs: Tensor):
        out = _check(xs[0], "loss term")
        for x in xs[1:]:
            out = add(out, x)
        if len(xs) == 1:
            out = out.clone()
        ctx.n = len(xs)
        return out

    @staticmethod
    def backward(ctx, gout: Tensor):
        return tuple(gout for _ in range(ctx.n))


def mse(a: Tensor, b: Tensor, scale: float = 1.0) -> Tensor:
    """scale * nn.MSELoss()(a, b)   (reference srgan/trainer.py:163,384)."""
    return _Pixel.apply(a, b, MSE, float(scale))


def l1(a: Tensor, b: Tensor, scale: float = 1.0) -> Tensor:
    """scale * nn.L1Loss()(a, b)   (reference esrgan/trainer.py:163,386,466; VGG feature L1, */loss.py:52)."""
    return _Pixel.apply(a, b, L1, float(scale))


def bce(p: Tensor, label, q: Optional[Tensor] = None, q_label: float = 0.0, scale: float = 1.0) -> Tensor:
    """scale * (nn.BCELoss()(p, label) [+ nn.BCELoss()(q, q_label)]) on probabilities, labels constant (a float) or, for
    `p` only, a tensor of per-element targets; PyTorch's clamp of log at -100 (reference srgan/trainer.py:446-448,456)."""
    if isinstance(label, Tensor):
        return _Gan.apply(p, q, GAN_BCE, 0.0, float(q_label), float(scale), label)
    return _Gan.apply(p, q, GAN_BCE, float(label), float(q_label), float(scale), None)


def bce_with_logits(x: Tensor, label, scale: float = 1.0) -> Tensor:
    """scale * nn.BCEWithLogitsLoss()(x, label)."""
    if isinstance(label, Tensor):
        return _Gan.apply(x, None, GAN_RELATIVISTIC, 0.0, 0.0, float(scale), label)
    return _Gan.apply(x, None, GAN_RELATIVISTIC, float(label), 0.0, float(scale), None)


def relativistic_d(real: Tensor, fake: Tensor, scale: float = 0.5) -> Tensor:
    """scale * (BCEwL(real - mean(fake), 1) + BCEwL(fake - mean(real), 0))   (reference esrgan/trainer.py:451-453)."""
    return _Gan.apply(real, fake, GAN_RELATIVISTIC, 1.0, 0.0, float(scale), None)


def relativistic_g(fake: Tensor, real: Tensor, scale: float = 1.0) -> Tensor:
    """scale * BCEwL(fake - mean(real), 1), `real` constant   (reference esrgan/trainer.py:463-468)."""
    return _Gan.apply(fake, real.detach(), GAN_RELATIVISTIC_G, 1.0, 0.0, float(scale), None)


def total(*terms: Tensor) -> Tensor:
    """Sum of scalar loss terms (reference srgan/trainer.py:448,457; esrgan/trainer.py:453,469)."""
    return _Sum.apply(*terms)


class MSELoss(nn.Module):
    def forward(self, input: Tensor, target: Tensor) -> Tensor:
        return mse(input, target)


class L1Loss(nn.Module):
    def forward(self, input: Tensor, target: Tensor) -> Tensor:
        return l1(input, target)


class BCELoss(nn.Module):
    """nn.BCELoss(reduction='mean'); `target` may be a tensor (the reference's label tensors) or a float."""

    def forward(self, input: Tensor, target) -> Tensor:
        return bce(input, target)


class BCEWithLogitsLoss(nn.Module):
    def forward(self, input: Tensor, target) -> Tensor:
        return bce_with_logits(input, target)
