"""Single-kernel parity checks of libtorchsr_b200.so against torch fp32 math on the same bf16-rounded operands
(oracle O2 of SURVEY.md section 4: differences are accumulation order only, so the fp32-output tolerance is 1e-4).

Every function returns a dict of relative-L2 errors; tests/test_kernels_gpu.py asserts on them and
tools/diag_kernels.py prints them all without stopping at the first failure. The reference side runs on the CPU
(torch.nn.functional, fp32) - nothing here uses a GPU library as the checker.
"""
import torch
import torch.nn.functional as F

from torchsr_b200 import _lib as L
from torchsr_b200 import ops

DEV = "cuda"


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


def describe_mismatch(out_nhwc: torch.Tensor, ref_nhwc: torch.Tensor) -> str:
    """Short text localising an error pattern: by row within the 128-row tile, by channel octet, and NaN count."""
    o = out_nhwc.detach().float().cpu().reshape(-1, out_nhwc.shape[-1])
    r = ref_nhwc.detach().float().cpu().reshape(-1, ref_nhwc.shape[-1])
    nan = int(torch.isnan(o).sum())
    o = torch.nan_to_num(o, nan=0.0)
    e = (o - r).abs()
    rows = e.mean(1)
    M = rows.numel()
    pad = (128 - M % 128) % 128
    by_row = torch.cat([rows, torch.zeros(pad)]).view(-1, 128)
    tile_err = by_row.mean(1)
    row_err = by_row.mean(0)
    ch_err = e.mean(0)
    ratio = (o.norm() / (r.norm() + 1e-30)).item()
    return (f"nan={nan} |out|/|ref|={ratio:.3f} ref_mean_abs={r.abs().mean():.3e} tiles[{tile_err.numel()}] first="
            f"{[round(x, 3) for x in tile_err[:6].tolist()]} rows8={[round(x, 3) for x in row_err.view(16, 8).mean(1).tolist()]} "
            f"ch8={[round(x, 3) for x in ch_err.view(-1, 8).mean(1)[:16].tolist()]}")


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def nhwc_bf16(x_nchw: torch.Tensor) -> torch.Tensor:
    """fp32 NCHW (CPU) -> bf16 NHWC contiguous on the device."""
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(DEV)


def pack_fwd(w: torch.Tensor, cout_pad=None, cin_pad=None, shuffle=False) -> torch.Tensor:
    """OIHW fp32 -> [KH*KW][cout_pad][cin_pad] bf16 (PK_FWD). shuffle: PixelShuffle(2) row permutation."""
    co, ci, kh, kw = w.shape
    cout_pad = cout_pad or co
    cin_pad = cin_pad or ci
    if shuffle:
        c4 = co // 4
        rows = torch.arange(co)
        src = 4 * (rows % c4) + rows // c4
        w = w[src]
    out = torch.zeros(kh * kw, cout_pad, cin_pad)
    out[:, :co, :ci] = w.permute(2, 3, 0, 1).reshape(kh * kw, co, ci)
    return out.to(torch.bfloat16).to(DEV)


def pack_t(w: torch.Tensor) -> torch.Tensor:
    """OIHW -> [KH*KW][ci][co] bf16 (PK_T, data-gradient operand)."""
    co, ci, kh, kw = w.shape
    return w.permute(2, 3, 1, 0).reshape(kh * kw, ci, co).contiguous().to(torch.bfloat16).to(DEV)


def sync_check():
    torch.cuda.synchronize()
    ops.check_watchdog()


# ------------------------------------------------------------------------------------------------ GEMM (tiled TMA)
def check_gemm(M=200, N=128, K=256, block_n=64, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = bf16_round(torch.randn(M, K, generator=g))
    w = bf16_round(torch.randn(N, K, generator=g) * 0.1)
    a_dev, w_dev = a.to(torch.bfloat16).to(DEV), w.to(torch.bfloat16).to(DEV)
    out = torch.full((M, N), float("nan"), device=DEV)
    d = ops.gemm_desc(a=a_dev, M=M, K=K, a_ld=K, w=w_dev, n_rows=N, block_n=block_n, out=out, out_ld=N, n_valid=N,
                      out_f32=True)
    ops.run_now(d)
    sync_check()
    return {"out": rel_l2(out, a @ w.t())}


def check_gemm_splitk_t(M=16, N=256, K=1024, block_n=128, splits=4, seed=1):
    """D^T accumulated with fp32 atomics (the Linear forward path)."""
    g = torch.Generator().manual_seed(seed)
    a = bf16_round(torch.randn(M, K, generator=g))
    w = bf16_round(torch.randn(N, K, generator=g) * 0.1)
    a_dev, w_dev = a.to(torch.bfloat16).to(DEV), w.to(torch.bfloat16).to(DEV)
    out = torch.zeros(N, M, device=DEV)
    d = ops.gemm_desc(a=a_dev, M=M, K=K, a_ld=K, w=w_dev, n_rows=N, block_n=block_n, out=out, out_ld=M, n_valid=N,
                      splits=splits, atomic_t=True)
    ops.run_now(d)
    sync_check()
    return {"out": rel_l2(out, (a @ w.t()).t())}


def check_gemm_mn_major(M=192, N=32, K=128, seed=2):
    """a_mode 2: A supplied as [K][M] (M contiguous) - the Linear data-gradient path."""
    g = torch.Generator().manual_seed(seed)
    at = bf16_round(torch.randn(K, M, generator=g))          # [K][M]
    w = bf16_round(torch.randn(N, K, generator=g) * 0.1)     # [N][K]
    at_dev, w_dev = at.to(torch.bfloat16).to(DEV), w.to(torch.bfloat16).to(DEV)
    out = torch.full((M, N), float("nan"), device=DEV)
    d = ops.gemm_desc(a=at_dev, M=M, K=K, a_ld=M, a_mn_major=True, w=w_dev, n_rows=N, block_n=N, out=out, out_ld=N,
                      n_valid=N, out_f32=True)
    ops.run_now(d)
    sync_check()
    return {"out": rel_l2(out, at.t() @ w.t())}


# ------------------------------------------------------------------------------------------------ conv forward
def check_conv_fwd(B=2, H=24, W=24, Cin=64, Cout=64, k=3, stride=1, block_n=None, bias=False, act=L.ACT_NONE,
                   stats=False, shuffle=False, residual=False, seed=3, out_f32=True, splits=1, repeat=1):
    g = torch.Generator().manual_seed(seed)
    pad = k // 2
    x = bf16_round(torch.randn(B, Cin, H, W, generator=g))
    w = bf16_round(torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5)
    b = torch.randn(Cout, generator=g) if bias else None
    alpha = torch.tensor([0.25])
    geom = ops.fwd_geometry(H, W, k, k, pad, pad, stride)
    Ho, Wo = geom["Ho"], geom["Wo"]
    block_n = block_n or min(Cout, 128)
    x_dev = nhwc_bf16(x)
    w_dev = pack_fwd(w, shuffle=shuffle)
    res = bf16_round(torch.randn(B, Cout, Ho, Wo, generator=g)) if residual else None
    res_dev = nhwc_bf16(res) if residual else None
    if shuffle:
        c4 = Cout // 4
        out = torch.full((B, 2 * Ho, 2 * Wo, c4), float("nan"), device=DEV,
                         dtype=torch.float32 if out_f32 else torch.bfloat16)
        os_n, os_h, os_w = 4 * Ho * Wo * c4, 2 * Wo * c4, c4
        bias_dev = b.to(DEV) if bias else None     # the epilogue indexes the OIHW-order bias through the shuffle map
    else:
        out = torch.full((B, Ho, Wo, Cout), float("nan"), device=DEV,
                         dtype=torch.float32 if out_f32 else torch.bfloat16)
        os_n, os_h, os_w = Ho * Wo * Cout, Wo * Cout, Cout
        bias_dev = b.to(DEV) if bias else None
    alpha_dev = alpha.to(DEV)
    stats_buf = torch.zeros(Cout, 2, device=DEV) if stats else None   # accumulated with atomics by the epilogue
    d = ops.conv_desc(x=x_dev, N=B, H=H, W=W, C=Cin, x_ld=Cin, geom=geom, w=w_dev, cout_pad=Cout, w_ld=Cin,
                      n_slots=k * k, block_n=block_n, out=out, os_n=os_n, os_h=os_h, os_w=os_w, n_valid=Cout,
                      out_mode=L.OUT_SHUFFLE if shuffle else L.OUT_LINEAR, out_f32=out_f32, bias=bias_dev,
                      prelu=alpha_dev if act == L.ACT_PRELU else None, act=act, res=res_dev,
                      aux=(Ho * Wo * Cout, Wo * Cout, Cout), stats_partial=stats_buf, stats_ld=Cout,
                      shuf_c=Cout // 4 if shuffle else 64)
    if splits > 1:
        # split-K with in-kernel finalize: zeroed fp32 workspace + tile counters, which the kernel leaves zeroed
        M = B * Ho * Wo
        ws = torch.zeros((M + 127) // 128 * 128, Cout, device=DEV)
        cnt = torch.zeros(((M + 127) // 128) * (Cout // block_n), dtype=torch.int32, device=DEV)
        d.splits, d.ws, d.tile_counters, d.ws_ld = splits, ops.ptr(ws), ops.ptr(cnt), Cout
    for rep in range(repeat):     # repeated launches must give the same result (self-cleaning workspace)
        if stats and rep:
            stats_buf.zero_()
        ops.run_now(d)
    sync_check()
    if splits > 1:
        assert float(ws.abs().max()) == 0.0 and int(cnt.abs().max()) == 0, "split-K workspace not left zeroed"
    ref = F.conv2d(x, w, b, stride=stride, padding=pad)
    pre = ref.clone()
    if residual:
        ref = ref + res
    if shuffle:
        ref = F.pixel_shuffle(ref, 2)
    if act == L.ACT_PRELU:
        ref = F.prelu(ref, alpha)
    elif act == L.ACT_LEAKY:
        ref = F.leaky_relu(ref, 0.2)
    r = {"out": rel_l2(out.float().permute(0, 3, 1, 2), ref)}
    if r["out"] > 1e-2:
        print("   mismatch:", describe_mismatch(out.float(), ref.permute(0, 2, 3, 1)), flush=True)
    if stats:
        tot = stats_buf.cpu()
        r["sum"] = rel_l2(tot[:, 0], pre.sum((0, 2, 3)))
        r["sumsq"] = rel_l2(tot[:, 1], (pre * pre).sum((0, 2, 3)))
    return r


# ------------------------------------------------------------------------------------------------ data gradient
def check_dgrad_s1(B=2, H=24, W=24, Cin=64, Cout=64, k=3, seed=4):
    g = torch.Generator().manual_seed(seed)
    pad = k // 2
    dy = bf16_round(torch.randn(B, Cout, H, W, generator=g))
    w = bf16_round(torch.randn(Cout, Cin, k, k, generator=g) / (Cout * k * k) ** 0.5)
    geom = ops.dgrad_s1_geometry(H, W, k, k, pad, pad)
    dy_dev, wt_dev = nhwc_bf16(dy), pack_t(w)
    out = torch.full((B, H, W, Cin), float("nan"), device=DEV)
    d = ops.conv_desc(x=dy_dev, N=B, H=H, W=W, C=Cout, x_ld=Cout, geom=geom, w=wt_dev, cout_pad=Cin, w_ld=Cout,
                      n_slots=k * k, block_n=min(Cin, 128), out=out, os_n=H * W * Cin, os_h=W * Cin, os_w=Cin,
                      n_valid=Cin, out_f32=True)
    ops.run_now(d)
    sync_check()
    ref = F.conv_transpose2d(dy, w, stride=1, padding=pad)
    return {"out": rel_l2(out.permute(0, 3, 1, 2), ref)}


def check_dgrad_s2(B=2, H=24, W=24, Cin=64, Cout=64, seed=5, grouped=False):
    """Stride-2 3x3 pad-1 conv: dX assembled from 4 output-parity classes (H, W = input dims, even)."""
    g = torch.Generator().manual_seed(seed)
    k, pad = 3, 1
    Hy, Wy = H // 2, W // 2
    dy = bf16_round(torch.randn(B, Cout, Hy, Wy, generator=g))
    w = bf16_round(torch.randn(Cout, Cin, k, k, generator=g) / (Cout * k * k) ** 0.5)
    dy_dev, wt_dev = nhwc_bf16(dy), pack_t(w)
    out = torch.full((B, H, W, Cin), float("nan"), device=DEV)
    descs = ops.dgrad_s2_descs(dy=dy_dev, N=B, Hy=Hy, Wy=Wy, Cout=Cout, dy_ld=Cout, wt=wt_dev, Cin=Cin, cin_pad=Cin,
                               block_n=min(Cin, 128), out=out, Hx=H, Wx=W, out_ld=Cin, n_valid=Cin, out_f32=True)
    if grouped:
        # the four parity classes as ONE grouped launch inside a recorded program (twice: plain launch, then the
        # captured graph of the range)
        prog = ops.Program()
        prog.add(ops.elt(L.E_ZERO, p=[out], i=[out.numel() * 4]))
        prog.add_group(descs)
        prog.add(ops.elt(L.E_CAST, p=[out, torch.empty(out.numel(), device=DEV, dtype=torch.bfloat16)], i=[out.numel(), 0]))
        prog.add(ops.elt(L.E_CAST, p=[out, torch.empty(out.numel(), device=DEV, dtype=torch.bfloat16)], i=[out.numel(), 0]))
        prog.run()
        prog.run()
        prog.run()
    else:
        for d in descs:
            ops.run_now(d)
    sync_check()
    ref = F.conv_transpose2d(dy, w, stride=2, padding=pad, output_padding=1)
    return {"out": rel_l2(out.permute(0, 3, 1, 2), ref)}


# ------------------------------------------------------------------------------------------------ weight gradient
def check_wgrad(B=2, H=24, W=24, Cin=64, Cout=64, k=3, stride=1, seed=6, block_n=None):
    g = torch.Generator().manual_seed(seed)
    pad = k // 2
    geom = ops.fwd_geometry(H, W, k, k, pad, pad, stride)
    Ho, Wo = geom["Ho"], geom["Wo"]
    x = bf16_round(torch.randn(B, Cin, H, W, generator=g))
    dy = bf16_round(torch.randn(B, Cout, Ho, Wo, generator=g))
    x_dev, dy_dev = nhwc_bf16(x), nhwc_bf16(dy)
    acc = torch.zeros(k * k, Cin, Cout, device=DEV)   # [tap][ci][co], accumulated with vector reductions
    d = ops.wgrad_desc(x=x_dev, N=B, H=H, W=W, C=Cin, x_ld=Cin, geom=geom, dy=dy_dev, dy_ld=Cout, dy_c=Cout, out=acc,
                       cout_valid=Cout, block_n=block_n or min(Cout, 128))
    ops.run_now(d)
    sync_check()
    xr = x.clone().requires_grad_(False)
    wz = torch.zeros(Cout, Cin, k, k, requires_grad=True)
    F.conv2d(xr, wz, None, stride=stride, padding=pad).backward(dy)
    ref = wz.grad.permute(2, 3, 1, 0).reshape(k * k, Cin, Cout)
    return {"out": rel_l2(acc, ref)}
