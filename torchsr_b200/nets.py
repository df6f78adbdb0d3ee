"""Network definitions: how each reference module graph is laid out as native launch lists.

Each ``define_*`` function fills a Plan (engine.py): it emits the forward program in the reference's layer order and
pushes one backward closure per stage on the plan's tape. Reference graphs restated here:
  SRGAN generator      torchsr/srgan/generator.py:33-81, residual.py:16-92
  SRGAN discriminator  torchsr/srgan/discriminator.py:27-88
  ESRGAN discriminator torchsr/esrgan/discriminator.py:27-95 (same stage structure, 10 convs, logits)
"""
from typing import Dict, List

import torch

from . import _lib as L
from . import ops
from .engine import BF16, F32, Act, ConvRec, LinearRec, Plan, _round_up

CHANSUM_SPLITS = 64


def _recs(plan: Plan) -> Dict[str, ConvRec]:
    return {r.name: r for r in plan.store.convs}


def _geom1(H, W):
    return ops.fwd_geometry(H, W, 1, 1, 0, 0, 1)


# ------------------------------------------------------------------------------------------------ SRGAN generator
def srgan_generator_records(m) -> List[ConvRec]:
    recs = [ConvRec("conv1.0", m.conv1[0], "fullk", need_dgrad=False)]
    for i, blk in enumerate(m.blocks):
        recs.append(ConvRec(f"blocks.{i}.conv1", blk.conv1))
        recs.append(ConvRec(f"blocks.{i}.conv2", blk.conv2))
    recs.append(ConvRec("conv2.0", m.conv2[0]))
    for j, layer in enumerate(m.conv_layers):
        recs.append(ConvRec(f"conv_layers.{j}.conv", layer.conv, shuffle=True))
    recs.append(ConvRec("conv3", m.conv3, "rown"))
    return recs


def residual_block_stage(plan: Plan, prog, name: str, blk, ra: ConvRec, rb: ConvRec, x: Act, inplace: bool = False) -> Act:
    """conv-BN-PReLU-conv-BN + x (torchsr/srgan/residual.py:86-91). `inplace` (inference-only plans, when nothing else
    reads x afterwards): the block's output replaces x in its buffer - the second conv's epilogue then adds its tile to
    global memory with a reducing bulk store instead of loading x (ConvParams::res_reduce)."""
    B, H, W, C = x.B, x.H, x.W, x.C
    alpha = blk.prelu.weight
    a1 = plan.act(name + ".a1", B, H, W, C)
    raw1, coef1 = plan.conv_bn_act(prog, name + ".c1", ra, x, a1, bn=blk.bn1, act=L.ACT_PRELU, alpha=alpha)
    y = x if (inplace and plan.infer_only) else plan.act(name + ".y", B, H, W, C)
    raw2, coef2 = plan.conv_bn_act(prog, name + ".c2", rb, a1, y, bn=blk.bn2, act=L.ACT_NONE, res=x)

    def bwd(bp, g, want_x, want_w):
        d2 = plan.norm_act_bwd(bp, name + ".bn2", g, raw2, coef=coef2, bn=blk.bn2, act=L.ACT_NONE, want_w=want_w)
        if want_w:
            plan.conv_wgrad(bp, rb, a1, d2)
        da1 = plan.conv_dgrad(bp, name + ".c2", rb, d2, a1)
        d1 = plan.norm_act_bwd(bp, name + ".bn1", da1, raw1, coef=coef1, bn=blk.bn1, act=L.ACT_PRELU, alpha=alpha,
                               want_w=want_w)
        if want_w:
            plan.conv_wgrad(bp, ra, x, d1)
        return plan.conv_dgrad(bp, name + ".c1", ra, d1, x, res=g)

    plan.tape.append(bwd)
    return y


def subpixel_stage(plan: Plan, prog, name: str, layer, rec: ConvRec, x: Act, consumer_block_n: int = 64) -> Act:
    """PReLU(PixelShuffle2(conv(x)+b)) with the shuffle folded into the conv store (residual.py:45-47)."""
    B, H, W, C = x.B, x.H, x.W, rec.cout // 4
    store = plan.grads
    out = plan.act(name + ".out", B, 2 * H, 2 * W, C)
    pre = out if plan.infer_only else plan.act(name + ".pre", B, 2 * H, 2 * W, C)
    alpha = layer.prelu.weight
    plan.conv_fwd(prog, rec, x, out, act=L.ACT_PRELU, prelu=alpha, preact=None if plan.infer_only else pre.t,
                  shuffle_out=True)
    dconv = None if plan.infer_only else plan.act(name + ".dconv", B, H, W, rec.cout)
    dap = plan.zbuf("bwd", name + ".dalpha", 1)     # accumulated by the consumer's dgrad epilogue (one red per CTA)
    out.hook = dict(bwd_z=pre, bwd_act=L.ACT_PRELU, prelu=alpha, unshuffle_to=dconv, dalpha_partial=dap)

    def bwd(bp, g, want_x, want_w):
        assert g is dconv, "the consumer of a sub-pixel stage must honour its gradient hook"
        if want_w:
            bp.add(ops.elt(L.E_SUM_FINALIZE, p=[dap, store.grad_slice(alpha)], i=[1, 0], f=[1.0], side=True))   # parameter gradient only
            plan.colsum(bp, name + ".db", g, store.grad_slice(rec.bias), shuffle_c4=rec.cout // 4)
            plan.conv_wgrad(bp, rec, x, g)
        return plan.conv_dgrad(bp, name, rec, g, x)

    plan.tape.append(bwd)
    return out


def define_srgan_generator(m, plan: Plan, shape):
    B, cin, H, W = shape
    if cin != 3:
        raise RuntimeError(f"Generator expects 3 input channels, got {cin}")
    R = _recs(plan)
    store, fwd = plan.grads, plan.fwd
    r1 = R["conv1.0"]
    alpha1 = m.conv1[1].weight
    E1 = None if (plan.infer_only and r1.k == 9 and r1.cin == 3 and r1.cout == 64) else plan.act("E1", B, H, W, r1.epad)
    c1 = plan.act("c1", B, H, W, 64)
    c1_pre = c1 if plan.infer_only else plan.act("c1.pre", B, H, W, 64)
    g1 = _geom1(H, W)
    rowk = plan.infer_only and r1.k == 9 and r1.cin == 3 and r1.cout == 64
    if rowk:
        # Inference plans: the 9x9 3-channel input layer row-decomposed like the output layer - the im2row kernel expands
        # only the nine HORIZONTAL taps (27 -> 32 columns, 64 bytes per pixel instead of the 512 of the full 243-column
        # expansion: 17 MB instead of 134 MB per 512x512 image), the nine vertical taps are im2col taps of the conv
        E1r = plan.act("E1r", B, H, W, 32)
        w1r = plan.buf("conv1.w_rowk", 9 * 64 * 32, BF16)
        # [kh][co][kw*3 + c] <- W[co][c][kh][kw]; packed by the plan's own launch list (always the current weights)
        src = torch.arange(64 * 3 * 9 * 9, dtype=torch.int32).view(64, 3, 9, 9).permute(2, 0, 3, 1).reshape(9, 64, 27)
        idx1 = torch.full((9, 64, 32), -1, dtype=torch.int32)
        idx1[:, :, :27] = src
        idx1_d = plan.buf("conv1.w_rowk.idx", idx1.numel(), torch.int32)
        idx1_d.copy_(idx1.view(-1))
        fwd.add(ops.elt(L.E_PACK_GATHER, p=[r1.weight, idx1_d, w1r], i=[idx1.numel()]))
        g1r = ops.fwd_geometry(H, W, r1.k, 1, r1.pad, 0, 1)
        # w_static=False: the pack is written by a kernel of this very launch list - the conv must not fetch weight tiles
        # before its programmatic-launch wait
        plan.conv(fwd, E1r, w1r, 32, 9, g1r, 64, 64, c1.t, c1.strides(), 64, bias=r1.bias, act=L.ACT_PRELU, prelu=alpha1,
                  w_static=False)
    else:
        plan.conv(fwd, E1, r1.w_fwd, r1.cols, 1, g1, r1.cout_pad, r1.block_n, c1.t, c1.strides(), r1.cout_pad, bias=r1.bias,
                  act=L.ACT_PRELU, prelu=alpha1, out_preact=None if plan.infer_only else c1_pre.t)

    def bwd_conv1(bp, g, want_x, want_w):
        if want_w:
            d = plan.norm_act_bwd(bp, "c1", g, c1_pre, act=L.ACT_PRELU, alpha=alpha1, g2=plan.slots["skip"],
                                  bias_grad=store.grad_slice(r1.bias), want_w=True)
            plan.conv_wgrad(bp, r1, E1, d, geom=g1)
        return None

    plan.tape.append(bwd_conv1)

    x = c1
    for i, blk in enumerate(m.blocks):
        # block 0 reads c1, which the skip connection behind the trunk needs again: every later block may work in place
        x = residual_block_stage(plan, fwd, f"blocks.{i}", blk, R[f"blocks.{i}.conv1"], R[f"blocks.{i}.conv2"], x,
                                 inplace=i >= 1)

    rc2, bn2 = R["conv2.0"], m.conv2[1]
    xt = x
    s = c1 if plan.infer_only else plan.act("trunk", B, H, W, 64)     # inference: added into c1 in place (see the blocks)
    raw, coef = plan.conv_bn_act(fwd, "conv2", rc2, xt, s, bn=bn2, act=L.ACT_NONE, res=c1)

    def bwd_conv2(bp, g, want_x, want_w):
        plan.slots["skip"] = g          # out = conv1 + conv2 (generator.py:79): the same gradient reaches conv1
        d = plan.norm_act_bwd(bp, "conv2.bn", g, raw, coef=coef, bn=bn2, act=L.ACT_NONE, want_w=want_w)
        if want_w:
            plan.conv_wgrad(bp, rc2, xt, d)
        return plan.conv_dgrad(bp, "conv2", rc2, d, xt)

    plan.tape.append(bwd_conv2)

    u = s
    for j, layer in enumerate(m.conv_layers):
        u = subpixel_stage(plan, fwd, f"conv_layers.{j}", layer, R[f"conv_layers.{j}.conv"], u)

    r3 = R["conv3"]
    Hf, Wf = u.H, u.W
    # 9x9 64->3 output conv (srgan/generator.py:58), row-decomposed: the GEMM runs over the 9 vertical taps with
    # N' = 9 horizontal taps x 3 channels (27 -> 32 columns); its epilogue sums the horizontally shifted columns inside
    # the tile and adds them (+ bias) into the zero-initialised fp32 NCHW result (OUT_GATHER_W) - no [M][32] fp32 round
    # trip through HBM, no gather kernel.
    Y = plan.buf("Y", B * r3.cout * Hf * Wf, F32)
    fwd.add(ops.elt(L.E_ZERO, p=[Y], i=[Y.numel() * 4]))
    geom3 = ops.fwd_geometry(Hf, Wf, r3.k, 1, r3.pad, 0, 1)
    if r3.k == 9 and r3.cout == 3 and r3.cin == 64 and Hf % 2 == 0:
        # TWO output rows per GEMM row (forward only; backward keeps the one-row packs). The GEMM row (n, y2, x) reads input rows 2*y2 - 4 .. 2*y2 + 5 (ten
        # vertical taps, traversal stride 2 along H) against N = 2 x 32 columns (r, kw, c): output row 2*y2 + r uses tap
        # kh' with the weights of kh = kh' - r. Twice the columns per fetched activation tile: an N = 64 UMMA reads 6 KB of
        # operands for twice the work of the 5 KB an N = 32 UMMA reads (the launch is bound by exactly that).
        w2 = plan.buf("conv3.w2", 10 * 64 * 64, BF16)
        # [kh'][r*32 + kw*3 + c][ci] <- W[c][ci][kh' - r][kw]; packed by the plan's own launch list
        src = torch.arange(3 * 64 * 9 * 9, dtype=torch.int32).view(3, 64, 9, 9).permute(2, 3, 0, 1).reshape(9, 27, 64)
        idx2 = torch.full((10, 2, 32, 64), -1, dtype=torch.int32)
        idx2[0:9, 0, :27] = src
        idx2[1:10, 1, :27] = src
        idx2_d = plan.buf("conv3.w2.idx", idx2.numel(), torch.int32)
        idx2_d.copy_(idx2.view(-1))
        fwd.add(ops.elt(L.E_PACK_GATHER, p=[r3.weight, idx2_d, w2], i=[idx2.numel()]))
        geom2 = dict(lower_h=-r3.pad, lower_w=0, upper_h=-(r3.pad + 1), upper_w=0, Ho=Hf // 2, Wo=Wf, stride=2, stride_w=1,
                     taps=[(kh, 0, kh) for kh in range(10)])
        plan.conv(fwd, u, w2, 64, 10, geom2, 64, 64, Y, (0, 0, 0), 64, w_static=False,
                  gather=dict(k=r3.k, pad=r3.pad, c=r3.cout, bias=r3.bias, rows=2))
    else:
        plan.conv(fwd, u, r3.w_fwd, r3.cols, r3.k, geom3, r3.npad, r3.npad, Y, (0, 0, 0), r3.npad,
                  gather=dict(k=r3.k, pad=r3.pad, c=r3.cout, bias=r3.bias))
    E3 = None if plan.infer_only else plan.act("E3", B, Hf, Wf, r3.npad)
    cs = plan.buf("conv3.cs", CHANSUM_SPLITS * r3.cout * 2, F32)
    u_last = u

    def input_fn(x_nchw):
        if rowk:
            ops.run_now(ops.elt(L.E_IM2ROW, p=[x_nchw, E1r.t], i=[B, 3, H, W, 1, r1.k, 0, r1.pad, 1, 32]))
        else:
            ops.run_now(ops.elt(L.E_IM2ROW, p=[x_nchw, E1.t], i=[B, 3, H, W, r1.k, r1.k, r1.pad, r1.pad, 1, r1.epad]))

    def output_fn():
        # a fresh tensor per call (the plan's buffer is overwritten by the next forward): one device-to-device copy
        out = torch.empty(B, r3.cout, Hf, Wf, dtype=F32, device=plan.device)
        out.view(-1).copy_(Y)
        return out

    def ingest_fn(gout):
        # dOut (NCHW fp32) -> row-expanded E3[n,h,w,kw*3+co] = dOut[n,co,h,w-(kw-4)], shared by conv3's dgrad and wgrad
        ops.run_now(ops.elt(L.E_IM2ROW, p=[gout, E3.t], i=[B, r3.cout, Hf, Wf, 1, r3.k, 0, r3.pad, -1, r3.npad]))
        ops.run_now(ops.elt(L.E_CHANSUM_NCHW, p=[gout, cs], i=[B, r3.cout, Hf * Wf, CHANSUM_SPLITS]))
        return E3

    def bwd_conv3(bp, g, want_x, want_w):
        if want_w:
            bp.add(ops.elt(L.E_COLSUM_FINALIZE, p=[cs, store.grad_slice(r3.bias)], i=[CHANSUM_SPLITS, r3.cout, r3.cout, 0, 0],
                           side=True))
            plan.conv_wgrad(bp, r3, u_last, E3, geom=geom3)
        return plan.conv_dgrad(bp, "conv3", r3, E3, u_last)

    plan.tape.append(bwd_conv3)

    def grad_input_fn():
        raise NotImplementedError("torchsr_b200: the gradient w.r.t. the generator's low-resolution input is not "
                                  "implemented (the reference training loops never request it)")

    plan.input_fn, plan.output_fn, plan.ingest_fn, plan.grad_input_fn = input_fn, output_fn, ingest_fn, grad_input_fn


# ------------------------------------------------------------------------------------------------ standalone blocks
class _IdentityRec:
    """Looks like the transposed pack of a 1x1 identity conv: lets conv_dgrad run a block's gradient hook (activation
    backward + PixelShuffle inverse) when no real consumer conv exists to fuse it into."""
    kind, need_dgrad, stride, k, pad = "fullk", True, 1, 1, 0

    def __init__(self, plan: Plan, C: int):
        self.t_rows = self.t_cols = C
        self.w_t = plan.buf("identity", C * C, BF16)
        self.w_t.view(C, C).copy_(torch.eye(C, dtype=BF16, device=plan.device))


def define_standalone(m, plan: Plan, shape, stage_fn, xin: Act = None, gy: Act = None):
    """A building block called on its own (e.g. ResidualBlock()(x)): NCHW fp32 <-> NHWC bf16 conversion kernels at
    both ends, the block's stage in between. `xin` / `gy` override where the converted input / output gradient live
    (channel slices of wider buffers for the ESRGAN dense blocks)."""
    B, C, H, W = shape
    if C % 16:
        raise RuntimeError(f"{type(m).__name__} needs a channel count that is a multiple of 16, got {C}")
    if xin is None:
        xin = plan.act("in", B, H, W, C)
    y = stage_fn(xin)
    if gy is None:
        gy = plan.act("gy", y.B, y.H, y.W, y.C)
    if y.hook is not None:
        ident = _IdentityRec(plan, y.C)
        plan.tape.append(lambda bp, g, want_x, want_w: plan.conv_dgrad(bp, "hook", ident, g, y))

    def input_fn(x):
        ops.run_now(ops.elt(L.E_NCHW2NHWC, p=[x, xin.t], i=[B, C, H, W, xin.ld, xin.c0]))

    def output_fn():
        out = torch.empty(B, y.C, y.H, y.W, dtype=F32, device=plan.device)
        ops.run_now(ops.elt(L.E_NHWC2NCHW, p=[y.t, out], i=[B, y.C, y.H, y.W, y.ld, y.c0, 0]))
        return out

    def ingest_fn(gout):
        ops.run_now(ops.elt(L.E_NCHW2NHWC, p=[gout, gy.t], i=[B, y.C, y.H, y.W, gy.ld, gy.c0]))
        return gy

    def grad_input_fn():
        g = plan.cur_g
        gx = torch.empty(B, C, H, W, dtype=F32, device=plan.device)
        ops.run_now(ops.elt(L.E_NHWC2NCHW, p=[g.t, gx], i=[B, C, H, W, g.ld, g.c0, 0]))
        return gx

    plan.input_fn, plan.output_fn, plan.ingest_fn, plan.grad_input_fn = input_fn, output_fn, ingest_fn, grad_input_fn


# ------------------------------------------------------------------------------------------------ discriminators
def discriminator_records(m, conv_idx):
    recs = [ConvRec("features.0", m.features[0], "fullk", need_dgrad=True)]
    for k in conv_idx[1:]:
        recs.append(ConvRec(f"features.{k}", m.features[k]))
    return recs


def discriminator_linears(m, image_size: int):
    fm = image_size // (2 ** (sum(1 for c in m.features if isinstance(c, torch.nn.Conv2d) and c.stride[0] == 2)))
    return [LinearRec("classifier.0", m.classifier[0], 512, fm, fm)]


def linear_wgrad_gemm(plan: Plan, a: torch.Tensor, xt: torch.Tensor, k_rows: int, scale: float, out=None):
    """Descriptor of dW = scale * A^T . X over `k_rows` batch rows (a multiple of 64; the gathered factors of all ranks
    in the data-parallel case), written into the flat gradient slice of the module's first Linear weight (or `out`)."""
    l1 = plan.store.linears[0]
    return ops.gemm_desc(a=a, M=l1.nout, K=k_rows, a_ld=l1.nout_pad, a_mn_major=True, w=xt, n_rows=l1.K, block_n=128,
                         out=out if out is not None else plan.grads.grad_slice(l1.weight), out_ld=l1.K, n_valid=l1.K,
                         out_f32=True, acc_scale=scale, w_chunked=True, side=True)


def define_discriminator(m, plan: Plan, shape, conv_idx, sigmoid: bool):
    """3x3 conv + LeakyReLU, then (conv, BN, LeakyReLU) stages with strides from the module, flatten, Linear,
    LeakyReLU, Linear (+ Sigmoid for SRGAN)."""
    B, cin, H, W = shape
    if cin != 3 or H != m.image_size or W != m.image_size:
        raise RuntimeError(f"Discriminator(image_size={m.image_size}) expects [N,3,{m.image_size},{m.image_size}] "
                           f"inputs, got {tuple(shape)}")
    R = _recs(plan)
    store, fwd = plan.grads, plan.fwd
    r0 = R["features.0"]
    E0 = plan.act("E0", B, H, W, r0.epad)
    f0 = plan.act("f0", B, H, W, r0.cout)
    g1 = _geom1(H, W)
    plan.conv(fwd, E0, r0.w_fwd, r0.cols, 1, g1, r0.cout_pad, r0.block_n, f0.t, f0.strides(), r0.cout_pad, bias=r0.bias,
              act=L.ACT_LEAKY)

    def bwd_f0(bp, g, want_x, want_w):
        if not (want_x or want_w):
            return None
        d = plan.norm_act_bwd(bp, "f0", g, f0, act=L.ACT_LEAKY, bias_grad=store.grad_slice(r0.bias), want_w=want_w)
        if want_w:
            plan.conv_wgrad(bp, r0, E0, d, geom=g1)
        if want_x:
            return plan.conv_dgrad(bp, "f0", r0, d, E0, out_f32=True)   # d/dE0, fp32 [M][epad]
        return None

    plan.tape.append(bwd_f0)

    prev = f0
    plan.has_bn = True
    l1 = plan.store.linears[0]
    early_k = conv_idx[-3]
    for k in conv_idx[1:]:
        rk, bn = R[f"features.{k}"], m.features[k + 1]
        Ho = (prev.H + 2 * rk.pad - rk.k) // rk.stride + 1
        Wo = (prev.W + 2 * rk.pad - rk.k) // rk.stride + 1
        a = plan.act(f"f{k}.act", B, Ho, Wo, rk.cout)
        raw, coef = plan.conv_bn_act(fwd, f"f{k}", rk, prev, a, bn=bn, act=L.ACT_LEAKY)

        def bwd(bp, g, want_x, want_w, rk=rk, bn=bn, raw=raw, coef=coef, xin=prev, k=k):
            d = plan.norm_act_bwd(bp, f"f{k}", g, raw, coef=coef, bn=bn, act=L.ACT_LEAKY, want_w=want_w)
            if want_w and k == conv_idx[-4]:
                # data parallel: the three deepest convs hold 88 % of the conv parameters and their weight gradients
                # were launched at least one stage ago - unpack and all-reduce that slice now (engine.Plan.early_unpack;
                # emitted before this stage's own weight-gradient launch: the unpack joins the side branch)
                plan.early_unpack(bp, [R[f"features.{j}"] for j in conv_idx if j >= early_k],
                                  R[f"features.{early_k}"].weight, l1.weight)
            if want_w:
                plan.conv_wgrad(bp, rk, xin, d)
            return plan.conv_dgrad(bp, f"f{k}", rk, d, xin)

        plan.tape.append(bwd)
        prev = a

    l1 = plan.store.linears[0]
    lin2 = m.classifier[2]
    N1, K = l1.nout, l1.K
    assert prev.M // B * prev.C == K
    Bpad = _round_up(B, 16)
    pre1t = plan.buf("pre1t", l1.nout_pad * B, F32)
    fwd.add(ops.elt(L.E_ZERO, p=[pre1t], i=[l1.nout_pad * B * 4]))
    tiles_n = l1.nout_pad // 128
    k_iters = K // 64
    splits = max(1, min(k_iters, (2 * 148) // tiles_n))
    fwd.mark("late_weights")       # first launch that reads the classifier weight (optim.FusedAdam late launch)
    fwd.add(ops.gemm_desc(a=prev.t, M=B, K=K, a_ld=K, w=l1.w_fwd, n_rows=l1.nout_pad, block_n=128, out=pre1t, out_ld=B,
                          n_valid=l1.nout_pad, splits=splits, atomic_t=True, w_static=True))
    out_b = plan.buf("out", B, F32)
    h1 = plan.buf("h1", B * N1, F32)
    fwd.add(ops.elt(L.E_HEAD, p=[pre1t, l1.bias, lin2.weight, lin2.bias, out_b, h1], i=[B, N1, int(sigmoid)], f=[0.2]))
    gbuf = plan.buf("gout", B, F32)
    flat_act = prev

    def bwd_head(bp, g, want_x, want_w):
        Bp64 = _round_up(B, 64)
        chunks = Bp64 // 64
        dpre1 = plan.buf("dpre1", B * N1, F32)
        dpre1_bf = plan.buf("dpre1_bf", Bp64 * l1.nout_pad, BF16, zero=True)   # rows >= B stay zero (GEMM K padding)
        dw2 = store.grad_slice(lin2.weight) if want_w else plan.buf("dw2.scratch", N1, F32)
        db2 = store.grad_slice(lin2.bias) if want_w else plan.buf("db2.scratch", 4, F32)
        db1 = store.grad_slice(l1.bias) if want_w else None
        bp.add(ops.elt(L.E_HEAD_BWD, p=[gbuf, out_b, h1, lin2.weight, dpre1, dpre1_bf, dw2, db2, db1],
                       i=[B, N1, int(sigmoid), l1.nout_pad], f=[0.2]))
        dflat32 = plan.buf("dflat32", Bpad * K, F32)
        bp.add(ops.elt(L.E_ZERO, p=[dflat32], i=[Bpad * K * 4]))
        bn_ = next(b for b in (128, 64, 32, 16) if Bpad % b == 0)
        # dX^T[(h,w,c)][b] = sum_n Wp[n][(h,w,c)] * dpre1[b][n]: the packed weight is the MN-major A operand as stored
        bp.add(ops.gemm_desc(a=l1.w_fwd, M=K, K=l1.nout_pad, a_ld=K, a_mn_major=True, w=dpre1_bf, n_rows=Bpad,
                             block_n=bn_, out=dflat32, out_ld=K, n_valid=Bpad, atomic_t=True))
        if want_w:
            # Weight gradient of the first Linear (75 MB for SRGAN) as a tensor-core GEMM over the batch:
            #   dW[n][k] = sum_b dpre1[b][n] * X[b][k],  k in the parameter's (c,h,w) column order
            # A = dpre1 (bf16, [batch][n], MN-major as stored), B = the transposed features cut into 64-row batch
            # chunks (feat_t_kernel), fp32 result stored straight into the flat gradient. Both launches only feed
            # gradient outputs: they run on the weight-gradient side branch, emitted AFTER the data-gradient GEMM above
            # so that its CTAs (critical path) are resident first.
            # Data parallel (engine.Plan.run_backward): instead of all-reducing the 75 MB product, every rank gathers the
            # two bf16 factors of all ranks (appending K chunks) and forms the averaged gradient itself.
            xt = plan.buf("feat_t", chunks * K * 64, BF16)
            bp.add(ops.elt(L.E_FEAT_T, p=[flat_act.t, xt], i=[B, l1.C, l1.Hf * l1.Wf, flat_act.ld, flat_act.c0],
                           side=True))
            plan.factors = dict(a=dpre1_bf, xt=xt, rows=Bp64, chunks=chunks, lin=l1, K=K)
            if getattr(plan, "_building_dist", False):
                bp.mark("factors")         # both factors are final here; the GEMM runs over the gathered ones
            else:
                bp.add(linear_wgrad_gemm(plan, dpre1_bf, xt, Bp64, 1.0))
                opt = getattr(plan, "_building_inline", None)
                adam = opt.late_desc_for(plan) if opt is not None else None
                if adam is not None:
                    # optim.FusedAdam.late_in_backward: the classifier weight is updated right behind the GEMM that
                    # produced its gradient, on the same side branch (nothing later in this backward reads the weight)
                    bp.add(adam)
                    bp.mark("inline_adam")
        dflat = plan.act("dflat", B, flat_act.H, flat_act.W, flat_act.C)
        bp.add(ops.elt(L.E_CAST, p=[dflat32, dflat.t], i=[B * K, 0]))
        return dflat

    plan.tape.append(bwd_head)

    def input_fn(x_nchw):
        ops.run_now(ops.elt(L.E_IM2ROW, p=[x_nchw, E0.t], i=[B, 3, H, W, r0.k, r0.k, r0.pad, r0.pad, 1, r0.epad]))

    def output_fn():
        return out_b[:B].clone().view(B, 1)

    def ingest_fn(gout):
        gbuf[:B].copy_(gout.reshape(-1))
        return None

    # forward_pair: the (real | fake) batches land in the two halves of the im2row buffer / of the output gradient
    def input_pair_fn(xa, xb):
        h = B // 2
        for t, off in ((xa, 0), (xb, h)):
            ops.run_now(ops.elt(L.E_IM2ROW, p=[t, ops.ptr(E0.t, off * H * W * r0.epad)],
                                i=[h, 3, H, W, r0.k, r0.k, r0.pad, r0.pad, 1, r0.epad]))

    def ingest_pair_fn(ga, gb):
        h = B // 2
        for g, off in ((ga, 0), (gb, h)):
            if g is None:      # that half's output did not reach the loss
                ops.run_now(ops.elt(L.E_ZERO, p=[ops.ptr(gbuf, off)], i=[h * 4]))
            else:
                gbuf[off:off + h].copy_(g.reshape(-1))
        return None

    plan.input_pair_fn, plan.ingest_pair_fn = input_pair_fn, ingest_pair_fn

    def grad_input_fn():
        dE0 = plan.cur_g
        gx = torch.empty(B, 3, H, W, dtype=F32, device=plan.device)
        ops.run_now(ops.elt(L.E_GATHER_OUT, p=[dE0.t, gx, None], i=[B, 3, H, W, r0.k, r0.k, r0.pad, r0.pad, -1, dE0.ld, 0]))
        return gx

    plan.input_fn, plan.output_fn, plan.ingest_fn, plan.grad_input_fn = input_fn, output_fn, ingest_fn, grad_input_fn


# ------------------------------------------------------------------------------------------------ ESRGAN generator
RDB_NAMES = ("RDB1", "RDB2", "RDB3")


def esrgan_generator_records(m) -> List[ConvRec]:
    recs = [ConvRec("conv1", m.conv1, "fullk", need_dgrad=False)]
    for b, blk in enumerate(m.blocks):
        for r in RDB_NAMES:
            recs += rdb_records(f"blocks.{b}.{r}.", getattr(blk, r))
    recs += [ConvRec("conv2", m.conv2), ConvRec("upsample1", m.upsample1), ConvRec("upsample2", m.upsample2),
             ConvRec("conv3.0", m.conv3[0]), ConvRec("conv4", m.conv4)]
    return recs


RDB_GRAD_RING = 5   # ring of gradient buffers: an RRDB's incoming gradient must outlive its three RDB backward passes


def rdb_records(prefix: str, rdb) -> List[ConvRec]:
    """The five convs of one ResidualDenseBlock (conv1-4 are nn.Sequential(conv, LeakyReLU) -> '.0.', conv5 is bare)."""
    recs = [ConvRec(f"{prefix}conv{k}.0", getattr(rdb, f"conv{k}")[0]) for k in range(1, 5)]
    recs.append(ConvRec(f"{prefix}conv5", rdb.conv5))
    return recs


def rdb_chain(plan: Plan, fwd, R: Dict[str, ConvRec], names: List[str], B: int, H: int, W: int, rrdb: bool):
    """A chain of ResidualDenseBlocks (torchsr/esrgan/residual.py:81-86), every three of them closed by the RRDB
    residual when `rrdb` (:124-129). Zero-copy dense concatenation: RDB j owns the [B,H,W,192] NHWC buffer `cat{j}`;
    conv_k reads the channel prefix [0, 64+32(k-1)) and writes its 32 channels right behind it, conv5 reads all 192
    and writes (conv5+b)*0.2 + x into the first 64 channels of the NEXT buffer (for the third RDB of an RRDB the outer
    `*0.2 + x` is folded into the same epilogue: (conv5+b)*0.04 + 0.2*x_rdb3 + x_rrdb). The caller fills
    cat0[..., :64]; the chain's output is cat{len(names)}[..., :64]. Backward mirrors this with [B,H,W,192] gradient
    buffers that the data-gradient convs accumulate into in place; the incoming gradient must live in a 192-wide
    buffer too. `names[j]` is the record prefix of RDB j ('' or 'blocks.3.RDB2')."""
    C, G, CT = 64, 32, 192
    store = plan.grads
    n_rdb = len(names)
    assert not rrdb or n_rdb % 3 == 0

    def cat_buf(j):
        return plan.buf(f"cat{j}", B * H * W * CT, BF16)

    def sl(t, c, c0=0):
        return Act(t, B, H, W, c, ld=CT, c0=c0)

    def gbuf(i):
        return plan.buf(f"dcat{i % RDB_GRAD_RING}", B * H * W * CT, BF16)

    def rn(name, leaf):
        return f"{name}.{leaf}" if name else leaf

    for j in range(n_rdb):
        r = j % 3 if rrdb else 0
        name = names[j]
        cat, nxt = cat_buf(j), cat_buf(j + 1)
        rk = [R[rn(name, f"conv{k}.0")] for k in range(1, 5)]
        r5 = R[rn(name, "conv5")]
        for k in range(1, 5):
            plan.conv_fwd(fwd, rk[k - 1], sl(cat, C + G * (k - 1)), sl(cat, G, C + G * (k - 1)), act=L.ACT_LEAKY)
        x_rdb = sl(cat, C)
        if rrdb and r == 2:
            x_rrdb = sl(cat_buf(j - 2), C)
            plan.conv_fwd(fwd, r5, sl(cat, CT), sl(nxt, C), acc_scale=0.04, res=x_rdb, res_scale=0.2, res2=x_rrdb,
                          res2_scale=1.0)
        else:
            plan.conv_fwd(fwd, r5, sl(cat, CT), sl(nxt, C), acc_scale=0.2, res=x_rdb)

        def bwd(bp, g, want_x, want_w, j=j, r=r, name=name, cat=cat, rk=rk, r5=r5):
            # g: gradient w.r.t. this RDB's output (first 64 channels of a 192-wide buffer). For the third RDB of an
            # RRDB it is the RRDB-level gradient: the block output was 0.2 * rdb3_out + x_rrdb.
            outer = 0.2 if (rrdb and r == 2) else 1.0
            if rrdb and r == 2:
                plan.slots[f"rrdb{j // 3}"] = g
            dcat = gbuf(n_rdb - j)
            d5 = plan.norm_act_bwd(bp, rn(name, "c5"), g, g, act=L.ACT_NONE, gscale=0.2 * outer,
                                   bias_grad=store.grad_slice(r5.bias), want_w=want_w)
            if want_w:
                plan.conv_wgrad(bp, r5, sl(cat, CT), d5)
            plan.conv_dgrad(bp, rn(name, "c5"), r5, d5, sl(cat, CT), out=sl(dcat, CT), res=sl(g.t, C), res_scale=outer,
                            res_cols=C)
            for k in range(4, 0, -1):
                ck = C + G * (k - 1)
                dk = plan.norm_act_bwd(bp, rn(name, f"c{k}"), sl(dcat, G, ck), sl(cat, G, ck), act=L.ACT_LEAKY,
                                       bias_grad=store.grad_slice(rk[k - 1].bias), want_w=want_w)
                if want_w:
                    plan.conv_wgrad(bp, rk[k - 1], sl(cat, ck), dk)
                last = rrdb and k == 1 and r == 0
                plan.conv_dgrad(bp, rn(name, f"c{k}"), rk[k - 1], dk, sl(cat, ck), out=sl(dcat, ck), res=sl(dcat, ck),
                                res2=sl(plan.slots[f"rrdb{j // 3}"].t, ck) if last else None)
            return sl(dcat, C)

        plan.tape.append(bwd)
    return sl(cat_buf(n_rdb), C)


def define_rdb_standalone(m, plan: Plan, shape, names: List[str], rrdb: bool):
    """ResidualDenseBlock()(x) / ResidualInResidualDenseBlock()(x) on their own (torchsr/esrgan/residual.py:81-86,
    124-129): layout conversion kernels around rdb_chain; input, output and their gradients live in 192-wide buffers
    because the chain addresses them that way."""
    B, C, H, W = shape
    if C != 64:
        raise RuntimeError(f"{type(m).__name__} on the B200 path is built for 64 channels (growth 32), got {C}")
    R = _recs(plan)
    CT = 192
    xin = Act(plan.buf("cat0", B * H * W * CT, BF16), B, H, W, C, ld=CT)
    gy = Act(plan.buf("gy192", B * H * W * CT, BF16), B, H, W, C, ld=CT)
    define_standalone(m, plan, shape, lambda x: rdb_chain(plan, plan.fwd, R, names, B, H, W, rrdb), xin=xin, gy=gy)


def define_esrgan_generator(m, plan: Plan, shape):
    """torchsr/esrgan/generator.py:54-81 and residual.py:81-86,124-129.

    Dense blocks are zero-copy: every RDB owns one [B,H,W,192] NHWC buffer `cat`; conv_k reads the channel prefix
    [0, 64+32(k-1)) and writes its 32 channels right behind it, conv5 reads all 192 and writes (conv5+b)*0.2 + x into
    the first 64 channels of the NEXT block's buffer (for the third RDB of an RRDB the outer `*0.2 + x` is folded
    into the same epilogue: (conv5+b)*0.04 + 0.2*x_rdb3 + x_rrdb). Backward mirrors this with [B,H,W,192] gradient
    buffers that the data-gradient convs accumulate into in place."""
    B, cin, H, W = shape
    if cin != 3:
        raise RuntimeError(f"Generator expects 3 input channels, got {cin}")
    R = _recs(plan)
    store, fwd = plan.grads, plan.fwd
    C, G, CT = 64, 32, 192
    n_rrdb = len(m.blocks)
    n_rdb = 3 * n_rrdb

    def cat_buf(j):
        return plan.buf(f"cat{j}", B * H * W * CT, BF16)

    def sl(t, c, c0=0, h=H, w=W):
        return Act(t, B, h, w, c, ld=CT, c0=c0)

    # ---- conv1 (3 -> 64, no activation) straight into the first dense buffer
    r1 = R["conv1"]
    E1 = plan.act("E1", B, H, W, r1.epad)
    g1 = _geom1(H, W)
    c1 = sl(cat_buf(0), C)
    plan.conv(fwd, E1, r1.w_fwd, r1.cols, 1, g1, r1.cout_pad, r1.block_n, c1.t, c1.strides(), r1.cout_pad, bias=r1.bias)

    def bwd_conv1(bp, g, want_x, want_w):
        if want_w:
            d = plan.norm_act_bwd(bp, "conv1", g, g, act=L.ACT_NONE, g2=plan.slots["skip"],
                                  bias_grad=store.grad_slice(r1.bias), want_w=True)
            plan.conv_wgrad(bp, r1, E1, d, geom=g1)
        return None

    plan.tape.append(bwd_conv1)

    # ---- 23 x 3 dense blocks
    names = [f"blocks.{b}.{r}" for b in range(n_rrdb) for r in RDB_NAMES]
    rdb_chain(plan, fwd, R, names, B, H, W, rrdb=True)

    def gbuf(i):
        return plan.buf(f"dcat{i % RDB_GRAD_RING}", B * H * W * CT, BF16)

    trunk = sl(cat_buf(n_rdb), C)

    # ---- conv2 + skip, two nearest-x2 upsample stages, conv3 + LeakyReLU, conv4
    r2 = R["conv2"]
    s = plan.act("trunk", B, H, W, C)
    # F.interpolate(scale_factor=2) in front of upsample1 / upsample2 (esrgan/generator.py:73,76) is folded into the
    # PRODUCER's store: conv2 and upsample1 write every output pixel to the 2 x 2 positions of the next conv's input as
    # well (out_rep2x) - no upsample kernel, no extra read of the low-resolution tensor
    ups_in = {"upsample1": plan.act("upsample1.in", B, 2 * H, 2 * W, C),
              "upsample2": plan.act("upsample2.in", B, 4 * H, 4 * W, C)}
    plan.conv_fwd(fwd, r2, trunk, s, res=c1, rep2x=ups_in["upsample1"])

    def bwd_conv2(bp, g, want_x, want_w):
        # g: gradient w.r.t. s = conv1 + conv2(trunk), living in a 192-wide buffer; it also reaches conv1
        plan.slots["skip"] = g
        if want_w:
            plan.colsum_strided(bp, "conv2.db", g, store.grad_slice(r2.bias))
            plan.conv_wgrad(bp, r2, trunk, g)          # dY may be a channel slice (row stride 192)
        return plan.conv_dgrad(bp, "conv2", r2, g, trunk, out=sl(gbuf(0), C))

    plan.tape.append(bwd_conv2)

    prev, h, w = s, H, W
    for name in ("upsample1", "upsample2"):
        rec = R[name]
        up = ups_in[name]
        out = plan.act(name + ".out", B, 2 * h, 2 * w, C)
        plan.conv_fwd(fwd, rec, up, out, act=L.ACT_LEAKY, rep2x=ups_in["upsample2"] if name == "upsample1" else None)
        first = name == "upsample1"

        def bwd_up(bp, g, want_x, want_w, rec=rec, up=up, out=out, h=h, w=w, name=name, first=first):
            d = plan.norm_act_bwd(bp, name, g, out, act=L.ACT_LEAKY, bias_grad=store.grad_slice(rec.bias), want_w=want_w)
            if want_w:
                plan.conv_wgrad(bp, rec, up, d)
            gup = plan.conv_dgrad(bp, name, rec, d, up)
            # gradient of nearest-neighbour x2: 2x2 sums; the first stage writes into a 192-wide buffer because the
            # dense-block backward (and the skip into conv1) address their gradients that way
            tgt = Act(plan.buf("s.grad", B * h * w * CT, BF16), B, h, w, C, ld=CT) if first else \
                plan.act(name + ".gin", B, h, w, C)
            bp.add(ops.elt(L.E_UPSAMPLE2X_BWD, p=[gup.t, tgt.t], i=[B, h, w, C, C, tgt.ld]))
            return tgt

        plan.tape.append(bwd_up)
        prev, h, w = out, 2 * h, 2 * w

    r3 = R["conv3.0"]
    u3 = plan.act("conv3.out", B, h, w, C)
    u2 = prev
    plan.conv_fwd(fwd, r3, u2, u3, act=L.ACT_LEAKY)

    def bwd_conv3(bp, g, want_x, want_w):
        d = plan.norm_act_bwd(bp, "conv3", g, u3, act=L.ACT_LEAKY, bias_grad=store.grad_slice(r3.bias), want_w=want_w)
        if want_w:
            plan.conv_wgrad(bp, r3, u2, d)
        return plan.conv_dgrad(bp, "conv3", r3, d, u2)

    plan.tape.append(bwd_conv3)

    r4 = R["conv4"]
    Hf, Wf = h, w
    T = plan.act("T", B, Hf, Wf, r4.cout_pad, F32)
    plan.conv_fwd(fwd, r4, u3, T, out_f32=True, use_bias=False)     # bias (3 values) is added by the gather kernel
    E4 = plan.act("E4", B, Hf, Wf, r4.cout_pad)
    cs = plan.buf("conv4.cs", CHANSUM_SPLITS * r4.cout * 2, F32)

    def bwd_conv4(bp, g, want_x, want_w):
        if want_w:
            bp.add(ops.elt(L.E_COLSUM_FINALIZE, p=[cs, store.grad_slice(r4.bias)], i=[CHANSUM_SPLITS, r4.cout, r4.cout, 0, 0],
                           side=True))
            plan.conv_wgrad(bp, r4, u3, E4)
        return plan.conv_dgrad(bp, "conv4", r4, E4, u3)

    plan.tape.append(bwd_conv4)

    def input_fn(x_nchw):
        ops.run_now(ops.elt(L.E_IM2ROW, p=[x_nchw, E1.t], i=[B, 3, H, W, r1.k, r1.k, r1.pad, r1.pad, 1, r1.epad]))

    def output_fn():
        out = torch.empty(B, r4.cout, Hf, Wf, dtype=F32, device=plan.device)
        ops.run_now(ops.elt(L.E_GATHER_OUT, p=[T.t, out, r4.bias], i=[B, r4.cout, Hf, Wf, 1, 1, 0, 0, 1, r4.cout_pad, 0]))
        return out

    def ingest_fn(gout):
        ops.run_now(ops.elt(L.E_IM2ROW, p=[gout, E4.t], i=[B, r4.cout, Hf, Wf, 1, 1, 0, 0, 1, r4.cout_pad]))
        ops.run_now(ops.elt(L.E_CHANSUM_NCHW, p=[gout, cs], i=[B, r4.cout, Hf * Wf, CHANSUM_SPLITS]))
        return E4

    def grad_input_fn():
        raise NotImplementedError("torchsr_b200: the gradient w.r.t. the generator's low-resolution input is not "
                                  "implemented (the reference training loops never request it)")

    plan.input_fn, plan.output_fn, plan.ingest_fn, plan.grad_input_fn = input_fn, output_fn, ingest_fn, grad_input_fn


# ------------------------------------------------------------------------------------------------ VGG19 features
def vgg_records(m) -> List[ConvRec]:
    """Every nn.Conv2d of the (frozen) torchvision feature extractor; the 3-channel first conv runs as a full-K GEMM
    over im2row columns like the discriminators' first layer."""
    return [ConvRec(f"features.{i}", layer, "fullk" if layer.in_channels == 3 else "std", need_dgrad=True)
            for i, layer in enumerate(m.features) if isinstance(layer, torch.nn.Conv2d)]


def define_vgg(m, plan: Plan, shape):
    """torchvision vgg19.features[:36] (torchsr/srgan/loss.py:30-33): (3x3 conv + ReLU) x 16 with four 2x2 max-pools,
    output = post-ReLU conv5_4 features. ReLU is fused into the conv epilogue; its backward into the epilogue of the
    data-gradient conv that produces the gradient (or into the max-pool backward). Frozen: no weight gradients."""
    B, cin, H, W = shape
    if cin != 3 or H % 16 or W % 16:
        raise RuntimeError(f"VGG feature extractor expects [N,3,H,W] with H, W multiples of 16, got {tuple(shape)}")
    R = _recs(plan)
    fwd = plan.fwd
    layers = list(m.features)
    relu_hook = lambda a: dict(bwd_z=a, bwd_act=L.ACT_RELU)  # noqa: E731
    x, h, w = None, H, W
    E0 = None
    i = 0
    while i < len(layers):
        layer = layers[i]
        if isinstance(layer, torch.nn.Conv2d):
            assert i + 1 < len(layers) and isinstance(layers[i + 1], torch.nn.ReLU), "conv without ReLU in VGG features"
            assert layer.kernel_size == (3, 3) and layer.stride == (1, 1) and layer.padding == (1, 1)
            rec = R[f"features.{i}"]
            out = plan.act(f"a{i}", B, h, w, rec.cout)
            if rec.kind == "fullk":
                E0 = plan.act("E0", B, h, w, rec.epad)
                plan.conv(fwd, E0, rec.w_fwd, rec.cols, 1, _geom1(h, w), rec.cout_pad, rec.block_n, out.t, out.strides(),
                          rec.cout_pad, bias=rec.bias, act=L.ACT_RELU)

                def bwd(bp, g, want_x, want_w, rec=rec, E0=E0):
                    return plan.conv_dgrad(bp, "c0", rec, g, E0, out_f32=True) if want_x else None
            else:
                plan.conv_fwd(fwd, rec, x, out, act=L.ACT_RELU)

                def bwd(bp, g, want_x, want_w, rec=rec, xin=x, name=f"c{i}"):
                    return plan.conv_dgrad(bp, name, rec, g, xin)
            out.hook = relu_hook(out)
            plan.tape.append(bwd)
            x = out
            i += 2
        elif isinstance(layer, torch.nn.MaxPool2d):
            assert layer.kernel_size == 2 and layer.stride == 2
            y = plan.act(f"p{i}", B, h // 2, w // 2, x.C)
            fwd.add(ops.elt(L.E_MAXPOOL2, p=[x.t, y.t], i=[B, h, w, x.C]))

            def bwd_pool(bp, g, want_x, want_w, xin=x, y=y, h=h, w=w, name=f"p{i}"):
                d = plan.act(name + ".dx", B, h, w, xin.C)
                bp.add(ops.elt(L.E_MAXPOOL2_BWD, p=[xin.t, y.t, g.t, d.t], i=[B, h, w, xin.C, 1]))   # + ReLU backward
                return d

            plan.tape.append(bwd_pool)
            x, h, w = y, h // 2, w // 2
            i += 1
        else:
            raise RuntimeError(f"unexpected layer in VGG features: {layer}")
    feats = x
    gy = plan.act("gy", B, h, w, feats.C)

    def bwd_out(bp, g, want_x, want_w):
        # gradient w.r.t. the post-ReLU output features: apply the last ReLU's derivative (no consumer conv to fuse into)
        return plan.norm_act_bwd(bp, "out", g, feats, act=L.ACT_RELU, want_w=False)

    if feats.hook is not None:
        plan.tape.append(bwd_out)
    r0 = R["features.0"]

    def input_fn(x_nchw):
        ops.run_now(ops.elt(L.E_IM2ROW, p=[x_nchw, E0.t], i=[B, 3, H, W, r0.k, r0.k, r0.pad, r0.pad, 1, r0.epad]))

    def output_fn():
        out = torch.empty(B, feats.C, feats.H, feats.W, dtype=F32, device=plan.device)
        ops.run_now(ops.elt(L.E_NHWC2NCHW, p=[feats.t, out], i=[B, feats.C, feats.H, feats.W, feats.ld, feats.c0, 0]))
        return out

    def ingest_fn(gout):
        ops.run_now(ops.elt(L.E_NCHW2NHWC, p=[gout, gy.t], i=[B, feats.C, feats.H, feats.W, feats.C, 0]))
        return gy

    def grad_input_fn():
        dE0 = plan.cur_g
        gx = torch.empty(B, 3, H, W, dtype=F32, device=plan.device)
        ops.run_now(ops.elt(L.E_GATHER_OUT, p=[dE0.t, gx, None], i=[B, 3, H, W, r0.k, r0.k, r0.pad, r0.pad, -1, dE0.ld, 0]))
        return gx

    plan.input_fn, plan.output_fn, plan.ingest_fn, plan.grad_input_fn = input_fn, output_fn, ingest_fn, grad_input_fn
