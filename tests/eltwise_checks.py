"""Parity checks of the HBM-bound kernels (csrc/eltwise.cu) against torch fp32 math on the CPU, same bf16-rounded
inputs. Each returns {name: rel_l2}. Used by tests/test_kernels_gpu.py and tools/diag_kernels.py."""
import torch
import torch.nn.functional as F

from torchsr_b200 import _lib as L
from torchsr_b200 import ops
from kernel_checks import DEV, bf16_round, nhwc_bf16, rel_l2, sync_check


def check_layout(B=3, C=64, H=10, W=12, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, W, generator=g)
    xd = x.to(DEV)
    y = torch.full((B, H, W, C), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.run_now(ops.elt(L.E_NCHW2NHWC, p=[xd, y], i=[B, C, H, W, C, 0]))
    z = torch.full((B, C, H, W), float("nan"), device=DEV)
    ops.run_now(ops.elt(L.E_NHWC2NCHW, p=[y, z], i=[B, C, H, W, C, 0, 0]))
    sync_check()
    ref = bf16_round(x)
    return {"nhwc": rel_l2(y.float().permute(0, 3, 1, 2), ref), "nchw": rel_l2(z, ref)}


def check_im2row(B=2, C=3, H=9, W=11, KH=3, KW=3, sign=1, seed=1):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, W, generator=g)
    ph, pw = KH // 2, KW // 2
    cols = KH * KW * C
    Epad = (cols + 31) // 32 * 32
    E = torch.full((B, H, W, Epad), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.run_now(ops.elt(L.E_IM2ROW, p=[x.to(DEV), E], i=[B, C, H, W, KH, KW, ph, pw, sign, Epad]))
    sync_check()
    xp = F.pad(x, (pw, pw, ph, ph))
    ref = torch.zeros(B, H, W, Epad)
    for kh in range(KH):
        for kw in range(KW):
            dh, dw = sign * (kh - ph), sign * (kw - pw)
            sl = xp[:, :, ph + dh:ph + dh + H, pw + dw:pw + dw + W]
            t = kh * KW + kw
            ref[..., t * C:(t + 1) * C] = sl.permute(0, 2, 3, 1)
    return {"E": rel_l2(E.float(), bf16_round(ref))}


def check_gather_out(B=2, C=3, H=8, W=10, KW=9, seed=2):
    g = torch.Generator().manual_seed(seed)
    Tld = (KW * C + 31) // 32 * 32
    T = torch.randn(B, H, W, Tld, generator=g)
    bias = torch.randn(C, generator=g)
    out = torch.full((B, C, H, W), float("nan"), device=DEV)
    ops.run_now(ops.elt(L.E_GATHER_OUT, p=[T.to(DEV), out, bias.to(DEV)], i=[B, C, H, W, 1, KW, 0, KW // 2, 1, Tld, 0]))
    sync_check()
    ref = bias.view(1, C, 1, 1).repeat(B, 1, H, W)
    for kw in range(KW):
        dw = kw - KW // 2
        lo, hi = max(0, -dw), min(W, W - dw)
        ref[:, :, :, lo:hi] += T[:, :, lo + dw:hi + dw, kw * C:(kw + 1) * C].permute(0, 3, 1, 2)
    return {"out": rel_l2(out, ref)}


def check_gather_out_3x3(B=3, H=19, W=45, sign=-1, seed=4):
    """3x3 taps x 3 channels in fp32 rows of 32 columns (the tiled kernel): out[n,c,h,w] = bias[c] + sum over the taps of
    T[n, h + sign*(kh-1), w + sign*(kw-1), (kh*3+kw)*3 + c]."""
    g = torch.Generator().manual_seed(seed)
    C, K = 3, 3
    T = torch.randn(B, H, W, 32, generator=g)
    bias = torch.randn(C, generator=g)
    out = torch.full((B, C, H, W), float("nan"), device=DEV)
    ops.run_now(ops.elt(L.E_GATHER_OUT, p=[T.to(DEV), out, bias.to(DEV)], i=[B, C, H, W, K, K, 1, 1, sign, 32, 0]))
    sync_check()
    ref = bias.view(1, C, 1, 1).repeat(B, 1, H, W)
    Tp = torch.nn.functional.pad(T.permute(0, 3, 1, 2), (1, 1, 1, 1))      # [B, 32, H+2, W+2]
    for kh in range(K):
        for kw in range(K):
            dh, dw = sign * (kh - 1), sign * (kw - 1)
            t = kh * K + kw
            ref += Tp[:, t * C:(t + 1) * C, 1 + dh:1 + dh + H, 1 + dw:1 + dw + W]
    return {"out": rel_l2(out, ref)}


def check_bn_train(M=1000, C=64, act=L.ACT_PRELU, residual=True, seed=3):
    """column sums (as the conv epilogue leaves them) -> fused BN_ACT, then BN_BWD_REDUCE + BN_BWD_APPLY, against
    autograd through F.batch_norm / prelu / leaky_relu."""
    g = torch.Generator().manual_seed(seed)
    x = bf16_round(torch.randn(M, C, generator=g) * 1.5 + 0.3)
    res = bf16_round(torch.randn(M, C, generator=g)) if residual else None
    gamma = torch.rand(C, generator=g) + 0.5
    beta = torch.randn(C, generator=g) * 0.1
    alpha = torch.tensor([0.25])
    gout = bf16_round(torch.randn(M, C, generator=g))
    rm, rv = torch.zeros(C), torch.ones(C)
    # ---- device
    stats = torch.stack([x.sum(0), (x * x).sum(0)], -1).contiguous().to(DEV)     # [C][2]
    coef = torch.full((4, C), float("nan"), device=DEV)
    rm_d, rv_d = rm.to(DEV), rv.to(DEV)
    nbt = torch.zeros((), dtype=torch.int64, device=DEV)
    gam_d, bet_d, alp_d = gamma.to(DEV), beta.to(DEV), alpha.to(DEV)
    x_d = x.to(torch.bfloat16).to(DEV)
    res_d = res.to(torch.bfloat16).to(DEV) if residual else None
    y_d = torch.full((M, C), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.run_now(ops.elt(L.E_BN_ACT, p=[x_d, stats, y_d, res_d, alp_d, gam_d, bet_d, rm_d, rv_d, nbt, coef],
                        i=[M, C, C, C, C, act, 0, 0, 0, 1, M], f=[0.2, 1.0, 1.0, 1e-5, 0.1]))
    # eval-mode twin of the same layer (running statistics as just updated)
    y_eval = torch.full((M, C), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.run_now(ops.elt(L.E_BN_ACT, p=[x_d, None, y_eval, res_d, alp_d, gam_d, bet_d, rm_d, rv_d, None, None],
                        i=[M, C, C, C, C, act, 0, 0, 0, 2, M], f=[0.2, 1.0, 1.0, 1e-5, 0.1]))
    # backward
    g_d = gout.to(torch.bfloat16).to(DEV)
    sums = torch.zeros(C, 2, device=DEV)
    dacc = torch.zeros(1, device=DEV)
    ops.run_now(ops.elt(L.E_BN_BWD_REDUCE, p=[g_d, x_d, coef, alp_d, sums, dacc, None],
                        i=[M, C, act, 128, C, C, 1], f=[0.2]))
    dgamma, dbeta, dalpha = torch.empty(C, device=DEV), torch.empty(C, device=DEV), torch.zeros(1, device=DEV)
    dx_d = torch.full((M, C), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.run_now(ops.elt(L.E_BN_BWD_APPLY, p=[g_d, x_d, coef, sums, alp_d, dx_d, None, gam_d, dgamma, dbeta, dalpha, dacc],
                        i=[M, C, act, C, C, C, 1], f=[0.2]))
    sync_check()
    # ---- reference
    xr = x.clone().requires_grad_(True)
    gr, br, ar = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True), alpha.clone().requires_grad_(True)
    rm_r, rv_r = rm.clone(), rv.clone()

    def act_fn(z):
        if act == L.ACT_PRELU:
            return F.prelu(z, ar)
        if act == L.ACT_LEAKY:
            return F.leaky_relu(z, 0.2)
        return z

    z = F.batch_norm(xr, rm_r, rv_r, gr, br, training=True, momentum=0.1, eps=1e-5)
    a = act_fn(z)
    y = a + res if residual else a
    y.backward(gout)
    ye = act_fn(F.batch_norm(x, rm_r, rv_r, gamma, beta, training=False, eps=1e-5))
    ye = ye + res if residual else ye
    r = {
        "y": rel_l2(y_d.float(), y), "y_eval": rel_l2(y_eval.float(), ye), "running_mean": rel_l2(rm_d, rm_r),
        "running_var": rel_l2(rv_d, rv_r), "nbt": float(abs(int(nbt.item()) - 1)),
        "dx": rel_l2(dx_d.float(), xr.grad), "dgamma": rel_l2(dgamma, gr.grad), "dbeta": rel_l2(dbeta, br.grad),
    }
    if act == L.ACT_PRELU:
        r["dalpha"] = rel_l2(dalpha, ar.grad)
    return r


def check_act_bwd_bias(M=700, C=64, seed=9):
    """Activation-only backward (no BatchNorm): dx = g * leaky'(y), bias gradient = column sums of dx."""
    g = torch.Generator().manual_seed(seed)
    pre = torch.randn(M, C, generator=g)
    yv = bf16_round(F.leaky_relu(pre, 0.2))
    gout = bf16_round(torch.randn(M, C, generator=g))
    y_d, g_d = yv.to(torch.bfloat16).to(DEV), gout.to(torch.bfloat16).to(DEV)
    sums = torch.zeros(C, 2, device=DEV)
    ops.run_now(ops.elt(L.E_BN_BWD_REDUCE, p=[g_d, y_d, None, None, sums, None, None], i=[M, C, L.ACT_LEAKY, 64, C, C, 0],
                        f=[0.2]))
    dx = torch.full((M, C), float("nan"), device=DEV, dtype=torch.bfloat16)
    db = torch.full((C,), float("nan"), device=DEV)
    ops.run_now(ops.elt(L.E_BN_BWD_APPLY, p=[g_d, y_d, None, sums, None, dx, None, None, None, db, None, None],
                        i=[M, C, L.ACT_LEAKY, C, C, C, 0], f=[0.2]))
    sync_check()
    ref = gout * torch.where(yv > 0, torch.ones_like(yv), torch.full_like(yv, 0.2))
    return {"dx": rel_l2(dx.float(), ref), "dbias": rel_l2(db, ref.sum(0))}


def check_pack_unpack(seed=4):
    """PACK_W table kernel vs torch packing for every mode, and UNPACK_G round trip."""
    from kernel_checks import pack_fwd, pack_t
    g = torch.Generator().manual_seed(seed)
    w33 = torch.randn(64, 32, 3, 3, generator=g)     # FWD / T
    wsh = torch.randn(256, 64, 3, 3, generator=g)    # FWD + shuffle
    w99i = torch.randn(64, 3, 9, 9, generator=g)     # FULLK
    w99o = torch.randn(3, 64, 9, 9, generator=g)     # ROWN / ROWN_T
    wl = torch.randn(24, 8 * 2 * 3, generator=g)     # LINEAR: C=8, Hf=2, Wf=3
    dev = {k: v.to(DEV) for k, v in dict(w33=w33, wsh=wsh, w99i=w99i, w99o=w99o, wl=wl).items()}
    ent = []

    def add(src, mode, cout, cin, kh, kw, rows_pad, cols_pad, slots, shuffle=0):
        dst = torch.full((slots * rows_pad * cols_pad,), float("nan"), device=DEV, dtype=torch.bfloat16)
        ent.append(dict(src=dev[src], dst=dst, mode=mode, cout=cout, cin=cin, kh=kh, kw=kw, rows_pad=rows_pad,
                        cols_pad=cols_pad, shuffle=shuffle, count=dst.numel()))
        return dst

    d_fwd = add("w33", L.PK_FWD, 64, 32, 3, 3, 64, 32, 9)
    d_t = add("w33", L.PK_T, 64, 32, 3, 3, 32, 64, 9)
    d_sh = add("wsh", L.PK_FWD, 256, 64, 3, 3, 256, 64, 9, shuffle=1)
    d_fullk = add("w99i", L.PK_FULLK, 64, 3, 9, 9, 64, 256, 1)
    d_rown = add("w99o", L.PK_ROWN, 3, 64, 9, 9, 32, 64, 9)
    d_rownt = add("w99o", L.PK_ROWN_T, 3, 64, 9, 9, 64, 32, 9)
    d_lin = add("wl", L.PK_LINEAR, 24, 8, 2, 3, 24, 48, 1)
    tab, n, blocks = ops.pack_table(ent, DEV)
    ops.run_now(ops.elt(L.E_PACK_W, p=[tab], i=[n, blocks]))
    sync_check()
    r = {}
    r["fwd"] = rel_l2(d_fwd.float().view(9, 64, 32), pack_fwd(w33).float())
    r["t"] = rel_l2(d_t.float().view(9, 32, 64), pack_t(w33).float())
    r["shuffle"] = rel_l2(d_sh.float().view(9, 256, 64), pack_fwd(wsh, shuffle=True).float())
    ref = torch.zeros(64, 256)
    ref[:, :243] = w99i.permute(0, 2, 3, 1).reshape(64, 243)
    r["fullk"] = rel_l2(d_fullk.float().view(64, 256), bf16_round(ref))
    ref = torch.zeros(9, 32, 64)
    ref[:, :27] = w99o.permute(2, 3, 0, 1).reshape(9, 27, 64)       # [kh][kw*3+co][ci]
    r["rown"] = rel_l2(d_rown.float().view(9, 32, 64), bf16_round(ref))
    ref = torch.zeros(9, 64, 32)
    ref[:, :, :27] = w99o.permute(2, 1, 3, 0).reshape(9, 64, 27)    # [kh][ci][kw*3+co]
    r["rown_t"] = rel_l2(d_rownt.float().view(9, 64, 32), bf16_round(ref))
    ref = wl.view(24, 8, 6).permute(0, 2, 1).reshape(24, 48)        # (c,hw) -> (hw,c)
    r["linear"] = rel_l2(d_lin.float().view(24, 48), bf16_round(ref))
    # unpack round trip: accumulator [taps][cols_pad (ci)][rows_pad (co-like)] fp32 -> OIHW
    acc = torch.randn(9, 32, 64, generator=g)
    dst = torch.full((64, 32, 3, 3), float("nan"), device=DEV)
    ent2 = [dict(src=acc.to(DEV), dst=dst, mode=L.PK_FWD, cout=64, cin=32, kh=3, kw=3, rows_pad=64, cols_pad=32,
                 shuffle=0, count=acc.numel())]
    acc2 = torch.randn(9, 64, 32, generator=g)      # ROWN: [kh][ci][kw*3+co (27 valid)]
    dst2 = torch.full((3, 64, 9, 9), float("nan"), device=DEV)
    ent2.append(dict(src=acc2.to(DEV), dst=dst2, mode=L.PK_ROWN, cout=3, cin=64, kh=9, kw=9, rows_pad=32, cols_pad=64,
                     shuffle=0, count=acc2.numel()))
    tab2, n2, blocks2 = ops.pack_table(ent2, DEV)
    ops.run_now(ops.elt(L.E_UNPACK_G, p=[tab2], i=[n2, blocks2]))
    sync_check()
    r["unpack_fwd"] = rel_l2(dst, acc.view(3, 3, 32, 64).permute(3, 2, 0, 1))
    r["unpack_rown"] = rel_l2(dst2, acc2[:, :, :27].reshape(9, 64, 9, 3).permute(3, 1, 0, 2))  # [kh][ci][kw][co] -> OIHW
    return r


def check_loss(n=16 * 3 * 96 * 96 + 3, kind=0, seed=5):
    g = torch.Generator().manual_seed(seed)
    a, b = torch.rand(n, generator=g), torch.rand(n, generator=g)
    blocks = 148
    partial = torch.empty(blocks, device=DEV)
    grad = torch.full((n,), float("nan"), device=DEV)
    out = torch.zeros(1, device=DEV)
    ops.run_now(ops.elt(L.E_LOSS, p=[a.to(DEV), b.to(DEV), partial, grad], i=[n, kind, blocks], f=[1.0 / n]))
    ops.run_now(ops.elt(L.E_SUM_FINALIZE, p=[partial, out], i=[blocks, 0], f=[1.0 / n]))
    sync_check()
    ar = a.clone().requires_grad_(True)
    loss = F.mse_loss(ar, b) if kind == 0 else F.l1_loss(ar, b)
    loss.backward()
    return {"loss": rel_l2(out, loss.view(1)), "grad": rel_l2(grad, ar.grad)}


def check_head(B=16, N1=1024, sigmoid=1, seed=6):
    g = torch.Generator().manual_seed(seed)
    pre1 = torch.randn(B, N1, generator=g)
    b1, w2, b2 = torch.randn(N1, generator=g) * 0.1, torch.randn(N1, generator=g) * 0.05, torch.randn(1, generator=g)
    gout = torch.randn(B, generator=g)
    pre1_t = pre1.t().contiguous().to(DEV)
    out, h1 = torch.empty(B, device=DEV), torch.empty(B, N1, device=DEV)
    ops.run_now(ops.elt(L.E_HEAD, p=[pre1_t, b1.to(DEV), w2.to(DEV), b2.to(DEV), out, h1], i=[B, N1, sigmoid], f=[0.2]))
    dpre1 = torch.empty(B, N1, device=DEV)
    dpre1_bf = torch.empty(B, N1, device=DEV, dtype=torch.bfloat16)
    dw2, db2 = torch.empty(N1, device=DEV), torch.empty(1, device=DEV)
    ops.run_now(ops.elt(L.E_HEAD_BWD, p=[gout.to(DEV), out, h1, w2.to(DEV), dpre1, dpre1_bf, dw2, db2],
                        i=[B, N1, sigmoid], f=[0.2]))
    sync_check()
    p1 = pre1.clone().requires_grad_(True)
    w2r, b2r = w2.clone().requires_grad_(True), b2.clone().requires_grad_(True)
    h = F.leaky_relu(p1 + b1, 0.2)
    o = h @ w2r + b2r
    if sigmoid:
        o = torch.sigmoid(o)
    o.backward(gout)
    return {"out": rel_l2(out, o), "dpre1": rel_l2(dpre1, p1.grad), "dpre1_bf": rel_l2(dpre1_bf.float(), p1.grad),
            "dw2": rel_l2(dw2, w2r.grad), "db2": rel_l2(db2, b2r.grad)}


def check_linear_wgrad(B=16, Nf=40, C=16, HW=6, seed=7):
    g = torch.Generator().manual_seed(seed)
    K = C * HW
    dpre = torch.randn(B, Nf, generator=g)
    x_chw = torch.randn(B, K, generator=g)                        # already in the parameter's (c,h,w) order, fp32
    dW, db = torch.full((Nf, K), float("nan"), device=DEV), torch.empty(Nf, device=DEV)
    ops.run_now(ops.elt(L.E_LINEAR_WGRAD, p=[dpre.to(DEV), x_chw.to(DEV), dW, db], i=[B, Nf, K]))
    sync_check()
    return {"dW": rel_l2(dW, dpre.t() @ x_chw), "db": rel_l2(db, dpre.sum(0))}


def check_upsample(B=2, H=5, W=7, C=64, seed=8):
    g = torch.Generator().manual_seed(seed)
    x = bf16_round(torch.randn(B, C, H, W, generator=g))
    dy = bf16_round(torch.randn(B, C, 2 * H, 2 * W, generator=g))
    y = torch.empty(B, 2 * H, 2 * W, C, device=DEV, dtype=torch.bfloat16)
    dx = torch.empty(B, H, W, C, device=DEV, dtype=torch.bfloat16)
    ops.run_now(ops.elt(L.E_UPSAMPLE2X, p=[nhwc_bf16(x), y], i=[B, H, W, C, C, C]))
    ops.run_now(ops.elt(L.E_UPSAMPLE2X_BWD, p=[nhwc_bf16(dy), dx], i=[B, H, W, C, C, C]))
    sync_check()
    xr = x.clone().requires_grad_(True)
    yr = F.interpolate(xr, scale_factor=2, mode="nearest")
    yr.backward(dy)
    return {"y": rel_l2(y.float().permute(0, 3, 1, 2), yr), "dx": rel_l2(dx.float().permute(0, 3, 1, 2), xr.grad)}
