"""Teacher-forced, stage-by-stage parity over WHOLE networks (SURVEY.md section 4 / App. E iii).

End-to-end gradients of a bf16 network cannot be held tighter than ~1e-1 against fp32 (a rounding flip of one
near-zero pre-activation changes an activation mask; PyTorch's own bf16 autocast measures the same), so a whole-network
number cannot tell a correct kernel from one with a wrong BatchNorm-gamma or PReLU-slope gradient. These checks can:
after ONE forward + backward of the real module on the GPU, every stage is recomputed in fp32 (torch ops, TF32 off) from
the tensors the CUDA path itself stored around that stage - its bf16 input activation, its bf16 raw conv output, the
bf16 gradient it received - with the parameters rounded to bf16 exactly as the kernels see them. What is left between
the two sides is accumulation order plus the one bf16 rounding of each stored result, so every stage output, every
data gradient and EVERY parameter gradient (conv weights, BatchNorm gamma / beta, PReLU slopes) of every layer is held
to a few 1e-3. Sums that cancel (gamma / beta / slope gradients) are normalised by the sum of |terms|, as App. E asks.

Reference semantics restated here: torchsr/srgan/residual.py:86-91 (residual block), srgan/generator.py:76-79,
srgan/discriminator.py:31-62 (conv / BatchNorm / LeakyReLU stages) and their autograd.
"""
import torch
import torch.nn.functional as F


def _nchw(plan, name, B, H, W, C):
    t = plan.bufs[name]
    return t[:B * H * W * C].view(B, H, W, C).permute(0, 3, 1, 2).float()


def _bf(w):
    return w.detach().bfloat16().float()


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def rel_terms(got, ref, terms_abs):
    """|got - ref| normalised by the sum of |terms| behind each entry (cancelling reductions)."""
    return float(((got.double() - ref.double()).abs() / terms_abs.double().clamp_min(1e-30)).max())


def _bn_fwd(conv, gamma, beta, lo, hi, eps=1e-5):
    """Batch statistics over rows [lo, hi) of the batch axis (one statistics group), applied to those rows."""
    x = conv[lo:hi]
    mean = x.mean(dim=(0, 2, 3))
    var = x.var(dim=(0, 2, 3), unbiased=False)
    invstd = (var + eps).rsqrt()
    sc = gamma * invstd
    sh = beta - mean * sc
    return x * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1), mean, invstd


def _bn_bwd(dz, x_bf, gamma, mean, invstd):
    """dx, dgamma, dbeta of training-mode BatchNorm for one statistics group, from dz and the (bf16) BN input."""
    v = lambda t: t.view(1, -1, 1, 1)  # noqa: E731
    xhat = (x_bf - v(mean)) * v(invstd)
    n = dz.numel() / dz.shape[1]
    s1 = dz.sum(dim=(0, 2, 3))
    s2 = (dz * xhat).sum(dim=(0, 2, 3))
    dx = v(gamma * invstd) * (dz - v(s1 / n) - xhat * v(s2 / n))
    return dx, s2, s1, (dz * xhat).abs().sum(dim=(0, 2, 3)), dz.abs().sum(dim=(0, 2, 3))


def _act(z, kind, slope):
    return z if kind == "none" else torch.where(z > 0, z, z * slope)


class StageReport:
    def __init__(self):
        self.rows = []

    def add(self, stage, what, err, tol):
        self.rows.append((stage, what, err, tol))

    def worst(self):
        return max(self.rows, key=lambda r: r[2] / r[3])

    def failures(self):
        return [r for r in self.rows if not r[2] <= r[3]]

    def summary(self):
        by = {}
        for _, what, err, tol in self.rows:
            a = by.setdefault(what, [0.0, tol, 0])
            a[0] = max(a[0], err)
            a[2] += 1
        return {k: (round(v[0], 6), v[1], v[2]) for k, v in by.items()}


OUT_TOL = 4e-3      # one bf16 rounding of a stored activation / gradient (2^-9 relative per element) + accumulation order
WGRAD_TOL = 2e-3    # fp32 accumulators, bf16 operands identical on both sides
SUM_TOL = 2e-3      # gamma / beta / slope gradients, normalised by sum |terms|
STAT_TOL = 2e-4     # batch mean / invstd published by the kernels


def _check_conv_bn_stage(rep, stage, plan, x, conv, bn, raw_name, out_name, coef_name, act, slope, res, groups,
                         stride=1):
    """Forward of conv -> BatchNorm(train) -> act (+ res): raw conv output, published statistics, stage output."""
    w = _bf(conv.weight)
    c = F.conv2d(x, w, None, stride, 1)
    B, C, H, W = c.shape
    rep.add(stage, "raw conv output", rel(_nchw(plan, raw_name, B, H, W, C), c), OUT_TOL)
    coef = plan.bufs[coef_name].view(groups, 4, C)
    outs, stats = [], []
    for g in range(groups):
        lo, hi = g * B // groups, (g + 1) * B // groups
        z, mean, invstd = _bn_fwd(c, bn.weight.detach(), bn.bias.detach(), lo, hi, bn.eps)
        rep.add(stage, "batch mean", float((coef[g, 2] - mean).abs().max() / mean.abs().max().clamp_min(1e-6)), STAT_TOL)
        rep.add(stage, "batch invstd", rel(coef[g, 3], invstd), STAT_TOL)
        outs.append(_act(z, act, slope))
        stats.append((coef[g, 0].clone(), coef[g, 1].clone(), coef[g, 2].clone(), coef[g, 3].clone()))
    y = torch.cat(outs)
    if res is not None:
        y = y + res
    rep.add(stage, "stage output", rel(_nchw(plan, out_name, B, H, W, C), y), OUT_TOL)
    return stats


def _check_bn_stage_bwd(rep, stage, plan, module_grads, g_in, x_in, conv, bn, raw_name, dx_name, stats, act, slope,
                        slope_grad_name, groups, stride=1, skip=None, dgrad_name=None, conv_name=None, bn_name=None):
    """Backward of one conv -> BN -> act stage given the gradient g_in w.r.t. the stage output (fp32, as the preceding
    data-gradient conv accumulates it): dz, BatchNorm input gradient dx (stored), dgamma / dbeta / dslope, the conv's
    weight gradient from (x_in, dx) and - when dgrad_name is given - the data gradient conv_input(dx) (+ skip)."""
    B, C, H, W = g_in.shape
    raw = _nchw(plan, raw_name, B, H, W, C)
    dxs, dgam, dbet, tg, tb = [], 0, 0, 0, 0
    dslope = g_in.new_zeros(())
    tslope = g_in.new_zeros(())
    for g in range(groups):
        lo, hi = g * B // groups, (g + 1) * B // groups
        sc, sh, mean, invstd = stats[g]
        z = raw[lo:hi] * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
        dz = g_in[lo:hi]
        if act != "none":
            neg = z <= 0
            if slope_grad_name is not None:
                dslope = dslope + (dz * z)[neg].sum()
                tslope = tslope + (dz * z)[neg].abs().sum()
            dz = torch.where(neg, dz * slope, dz)
        dx, s2, s1, a2, a1 = _bn_bwd(dz, raw[lo:hi], bn.weight.detach(), mean, invstd)
        dxs.append(dx)
        dgam, dbet, tg, tb = dgam + s2, dbet + s1, tg + a2, tb + a1
    dx = torch.cat(dxs)
    got_dx = _nchw(plan, dx_name, B, H, W, C)
    rep.add(stage, "BatchNorm input gradient", rel(got_dx, dx), OUT_TOL)
    if module_grads is not None:
        rep.add(stage, "dgamma", rel_terms(module_grads[bn_name + ".weight"], dgam, tg), SUM_TOL)
        rep.add(stage, "dbeta", rel_terms(module_grads[bn_name + ".bias"], dbet, tb), SUM_TOL)
        if slope_grad_name is not None:
            rep.add(stage, "dslope", rel_terms(module_grads[slope_grad_name].reshape(()), dslope, tslope), SUM_TOL)
        # weight gradient from the tensors the kernel read: bf16 stage input and the STORED bf16 dx
        dw = torch.nn.grad.conv2d_weight(x_in, conv.weight.shape, got_dx, stride=stride, padding=1)
        rep.add(stage, "conv weight gradient", rel(module_grads[conv_name + ".weight"], dw), WGRAD_TOL)
    if dgrad_name is not None:
        gi = torch.nn.grad.conv2d_input(x_in.shape, _bf(conv.weight), got_dx, stride=stride, padding=1)
        if skip is not None:
            gi = gi + skip
        rep.add(stage, "data gradient", rel(_nchw(plan, dgrad_name, *[x_in.shape[i] for i in (0, 2, 3, 1)]), gi), OUT_TOL)
    return got_dx


def srgan_generator_stagewise(G, x, gout):
    """Every residual block and the trunk-closing conv2 + BN of the SRGAN generator (33 conv + BatchNorm stages)."""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        G.zero_grad()
        y = G(x)
        plan = next(p for pool in G._tsr["plans"].values() for p in pool if p.busy)
        y.backward(gout)
        torch.cuda.synchronize()
        grads = {k: p.grad.detach().clone() for k, p in G.named_parameters()}
        B, _, H, W = x.shape
        C = 64
        rep = StageReport()
        n = len(G.blocks)
        with torch.no_grad():
            c1 = _nchw(plan, "c1", B, H, W, C)
            # ---- forward, block by block, each fed the CUDA path's own input activation
            xin = c1
            stats = {}
            for i, blk in enumerate(G.blocks):
                p = f"blocks.{i}"
                a = float(blk.prelu.weight)
                stats[p + ".1"] = _check_conv_bn_stage(rep, p + ".conv1", plan, xin, blk.conv1, blk.bn1, p + ".c1.raw",
                                                       p + ".a1", p + ".c1.bn.coef", "prelu", a, None, 1)
                a1 = _nchw(plan, p + ".a1", B, H, W, C)
                stats[p + ".2"] = _check_conv_bn_stage(rep, p + ".conv2", plan, a1, blk.conv2, blk.bn2, p + ".c2.raw",
                                                       p + ".y", p + ".c2.bn.coef", "none", 0.0, xin, 1)
                xin = _nchw(plan, p + ".y", B, H, W, C)
            stats["conv2"] = _check_conv_bn_stage(rep, "conv2", plan, xin, G.conv2[0], G.conv2[1], "conv2.raw", "trunk",
                                                  "conv2.bn.coef", "none", 0.0, c1, 1)
            # ---- backward. The gradient w.r.t. `trunk` arrives from the first sub-pixel stage's data gradient.
            g = _nchw(plan, "conv_layers.0.dgrad", B, H, W, C)
            xt = _nchw(plan, f"blocks.{n - 1}.y", B, H, W, C)
            _check_bn_stage_bwd(rep, "conv2", plan, grads, g, xt, G.conv2[0], G.conv2[1], "conv2.raw", "conv2.bn.dx",
                                stats["conv2"], "none", 0.0, None, 1, dgrad_name="conv2.dgrad", conv_name="conv2.0",
                                bn_name="conv2.1")
            g = _nchw(plan, "conv2.dgrad", B, H, W, C)
            for i in reversed(range(n)):
                blk, p = G.blocks[i], f"blocks.{i}"
                a = float(blk.prelu.weight)
                xin = c1 if i == 0 else _nchw(plan, f"blocks.{i - 1}.y", B, H, W, C)
                a1 = _nchw(plan, p + ".a1", B, H, W, C)
                d2 = _check_bn_stage_bwd(rep, p + ".conv2", plan, grads, g, a1, blk.conv2, blk.bn2, p + ".c2.raw",
                                         p + ".bn2.dx", stats[p + ".2"], "none", 0.0, None, 1, conv_name=p + ".conv2",
                                         bn_name=p + ".bn2")
                da1 = torch.nn.grad.conv2d_input(a1.shape, _bf(blk.conv2.weight), d2, padding=1)     # fp32 accumulator
                _check_bn_stage_bwd(rep, p + ".conv1", plan, grads, da1, xin, blk.conv1, blk.bn1, p + ".c1.raw",
                                    p + ".bn1.dx", stats[p + ".1"], "prelu", a, p + ".prelu.weight", 1, skip=g,
                                    dgrad_name=p + ".c1.dgrad", conv_name=p + ".conv1", bn_name=p + ".bn1")
                g = _nchw(plan, p + ".c1.dgrad", B, H, W, C)
        return rep
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def discriminator_stagewise(D, conv_idx, xs, pair: bool):
    """The seven (SRGAN) / nine (ESRGAN) conv + BatchNorm + LeakyReLU stages of a discriminator, forward and backward,
    for one call (xs = [x]) or one paired call over (real | fake) with per-half statistics (xs = [a, b])."""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        D.zero_grad()
        if pair:
            pa, pb = D.forward_pair(*xs)
            plan = next(p for k, pool in D._tsr["plans"].items() if len(k) == 3 for p in pool if p.busy)
            (pa.sum() + 2.0 * pb.sum()).backward()
        else:
            out = D(xs[0])
            plan = next(p for pool in D._tsr["plans"].values() for p in pool if p.busy)
            out.sum().backward()
        torch.cuda.synchronize()
        grads = {k: p.grad.detach().clone() for k, p in D.named_parameters()}
        groups = 2 if pair else 1
        B = sum(t.shape[0] for t in xs)
        size = xs[0].shape[-1]
        rep = StageReport()
        with torch.no_grad():
            acts = {conv_idx[0]: _nchw(plan, "f0", B, size, size, 64)}
            stats, shapes = {}, {}
            prev_k, h = conv_idx[0], size
            for k in conv_idx[1:]:
                conv, bn = D.features[k], D.features[k + 1]
                s = conv.stride[0]
                stats[k] = _check_conv_bn_stage(rep, f"features.{k}", plan, acts[prev_k], conv, bn, f"f{k}.raw",
                                                f"f{k}.act", f"f{k}.bn.coef", "leaky", 0.2, None, groups, stride=s)
                h = (h + 2 - 3) // s + 1
                shapes[k] = (B, conv.out_channels, h, h)
                acts[k] = _nchw(plan, f"f{k}.act", B, h, h, conv.out_channels)
                prev_k = k
            # backward: the head hands `dflat` (gradient w.r.t. the last activation); every earlier stage receives the
            # fp32 data gradient of the stage after it, formed from that stage's STORED dx
            last = conv_idx[-1]
            Bc, Cc, hc, _ = shapes[last]
            g = _nchw(plan, "dflat", Bc, hc, hc, Cc)
            order = list(conv_idx[1:])
            for j in reversed(range(len(order))):
                k = order[j]
                conv, bn = D.features[k], D.features[k + 1]
                x_in = acts[conv_idx[0]] if j == 0 else acts[order[j - 1]]
                dx = _check_bn_stage_bwd(rep, f"features.{k}", plan, grads, g, x_in, conv, bn, f"f{k}.raw", f"f{k}.dx",
                                         stats[k], "leaky", 0.2, None, groups, stride=conv.stride[0],
                                         conv_name=f"features.{k}", bn_name=f"features.{k + 1}")
                g = torch.nn.grad.conv2d_input(x_in.shape, _bf(conv.weight), dx, stride=conv.stride[0], padding=1)
        return rep
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
