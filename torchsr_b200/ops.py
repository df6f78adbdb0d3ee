"""Descriptor builders over the C ABI: turns convolution geometry into tsr_conv_desc_t / tsr_wgrad_desc_t /
tsr_elt_desc_t and records them into native programs. Pure host logic (no GPU needed to build descriptors
until they are handed to the library)."""
import ctypes as C
import os
from typing import Iterable, List, Optional, Sequence, Tuple

from . import _lib as L
from ._lib import ConvDesc, EltDesc, PackEntry, WgradDesc


DRY = bool(int(os.environ.get("TSR_DRY", "0")))   # build and validate descriptors without a GPU (host-logic tests)

# Address ranges of every tensor handed to a descriptor: the extent checks below refuse a descriptor whose kernel
# would touch memory outside the tensor it points into (an out-of-bounds access is a device fault, not an exception).
_RANGES: dict = {}


def ptr(t, offset_elems: int = 0) -> int:
    """Device address of a torch tensor (plus an element offset) or a plain int."""
    if t is None:
        return 0
    if isinstance(t, int):
        return t
    base = t.data_ptr()
    try:
        nbytes = t.untyped_storage().nbytes() - t.storage_offset() * t.element_size()
    except Exception:  # noqa: BLE001
        nbytes = t.numel() * t.element_size()
    if nbytes > _RANGES.get(base, 0):
        _RANGES[base] = nbytes
    return base + offset_elems * t.element_size()


class ExtentError(ValueError):
    pass


def _room(addr: int) -> int:
    """Bytes available from `addr` to the end of the registered tensor containing it (-1 if unknown)."""
    best = -1
    for base, n in _RANGES.items():
        if base <= addr < base + n:
            best = max(best, base + n - addr)
    return best


def _need(what: str, addr: int, nbytes: int):
    if addr == 0 or nbytes <= 0:
        return
    room = _room(addr)
    if room < 0:
        raise ExtentError(f"{what}: address {addr:#x} is not inside any tensor known to torchsr_b200.ops")
    if nbytes > room:
        raise ExtentError(f"{what}: kernel would touch {nbytes} bytes but only {room} remain in the tensor")


def forget_ranges():
    _RANGES.clear()


def current_stream() -> int:
    import torch
    if DRY:
        return 0
    return torch.cuda.current_stream().cuda_stream


def round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def pick_block_k(c: int) -> int:
    for bk in (64, 32, 16):
        if c % bk == 0:
            return bk
    raise ValueError(f"channel count {c} must be a multiple of 16")


# --------------------------------------------------------------------------------------------- geometry
def fwd_geometry(H: int, W: int, KH: int, KW: int, ph: int, pw: int, stride: int):
    """Bounding box and tap offsets of a forward cross-correlation (nn.Conv2d semantics)."""
    Ho = (H + 2 * ph - KH) // stride + 1
    Wo = (W + 2 * pw - KW) // stride + 1
    taps = [(kh, kw, kh * KW + kw) for kh in range(KH) for kw in range(KW)]
    return dict(lower_h=-ph, lower_w=-pw, upper_h=ph - (KH - 1), upper_w=pw - (KW - 1), Ho=Ho, Wo=Wo, stride=stride,
                taps=taps)


def dgrad_s1_geometry(H: int, W: int, KH: int, KW: int, ph: int, pw: int):
    """Data gradient of a stride-1 'same' conv, as a conv over dY with transposed (unflipped) weight slots:
    dX[h,w] = sum_{kh,kw} dY[h+ph-kh, w+pw-kw] * Wt[kh,kw]."""
    taps = [(KH - 1 - kh, KW - 1 - kw, kh * KW + kw) for kh in range(KH) for kw in range(KW)]
    return dict(lower_h=-(KH - 1 - ph), lower_w=-(KW - 1 - pw), upper_h=-ph, upper_w=-pw, Ho=H, Wo=W, stride=1,
                taps=taps)


def dgrad_s2_classes(K: int, p: int):
    """Output-parity decomposition of the data gradient of a stride-2 conv (kernel K, padding p).
    Returns {parity r: [(offset, k), ...]} per dimension: dX[2a+r] = sum dY[a+offset] * W[k]."""
    out = {}
    for r in (0, 1):
        lst = []
        for k in range(K):
            if (r + p - k) % 2 == 0 and (r + p - k) >= 0:
                lst.append(((r + p - k) // 2, k))
        out[r] = lst
    return out


# --------------------------------------------------------------------------------------------- descriptors
def _set_taps(d, taps: Sequence[Tuple[int, int, int]]):
    d.num_taps = len(taps)
    for i, (oh, ow, slot) in enumerate(taps):
        assert 0 <= oh < 256 and 0 <= ow < 256
        d.tap_off[i] = (oh << 8) | ow
        if hasattr(d, "tap_wrow"):
            d.tap_wrow[i] = slot


def conv_desc(*, x, N, H, W, C, x_ld, geom, w, cout_pad, w_ld, n_slots, block_n, out, os_n, os_h, os_w, n_valid,
              block_k=None, a_c0=0, out_mode=L.OUT_LINEAR, out_f32=False, out_ch_off=0, bias=None, prelu=None,
              act=L.ACT_NONE, out_preact=None, res=None, bwd_z=None, bwd_act=L.ACT_NONE, aux=(0, 0, 0), aux_ch_off=0,
              dalpha_partial=None, stats_partial=None, stats_ld=0, acc_scale=1.0, leaky=0.2, shuf_c=64, res2=None,
              res_scale=1.0, res2_scale=1.0, res_cols=0, w_static=True, bnr_x=None, bnr_coef=None, bnr_prelu=None,
              bnr_act=L.ACT_NONE, bnr_c=0, splits=1, ws=None, tile_counters=None, ws_ld=0, group_rows=0, bnf_mode=0,
              bnf_c=0, bnf_counter=None, bnf_gamma=None, bnf_beta=None, bnf_rm=None, bnf_rv=None, bnf_nbt=None,
              bnf_coef=None, bnf_count=0, bnf_eps=1e-5, bnf_momentum=0.1, gather=None, rep2x=None,
              stride_w=0) -> ConvDesc:
    d = ConvDesc()
    d.x, d.w = ptr(x), ptr(w)
    d.N, d.H, d.W, d.C, d.x_ld = N, H, W, C, x_ld
    d.Ho, d.Wo = geom["Ho"], geom["Wo"]
    d.a_mode = 0
    d.stride = geom["stride"]
    d.stride_w = geom.get("stride_w", stride_w)
    d.lower_h, d.lower_w, d.upper_h, d.upper_w = geom["lower_h"], geom["lower_w"], geom["upper_h"], geom["upper_w"]
    _set_taps(d, geom["taps"])
    d.block_k = block_k or pick_block_k(C - a_c0)
    d.block_n = block_n
    d.cout_pad = cout_pad
    d.w_rows, d.w_ld = n_slots * cout_pad, w_ld
    d.a_c0 = a_c0
    d.splits = splits
    d.ws, d.tile_counters, d.ws_ld = ptr(ws), ptr(tile_counters), ws_ld
    d.out, d.out_preact, d.bias, d.prelu = ptr(out), ptr(out_preact), ptr(bias), ptr(prelu)
    d.res, d.bwd_z, d.dalpha_partial, d.stats_partial = ptr(res), ptr(bwd_z), ptr(dalpha_partial), ptr(stats_partial)
    d.os_n, d.os_h, d.os_w = os_n, os_h, os_w
    d.aux_n, d.aux_h, d.aux_w = aux
    d.out_mode, d.out_f32, d.out_ch_off, d.aux_ch_off = out_mode, int(out_f32), out_ch_off, aux_ch_off
    d.n_valid, d.act, d.bwd_act, d.stats_ld, d.shuf_c = n_valid, act, bwd_act, stats_ld, shuf_c
    d.acc_scale, d.leaky_slope = acc_scale, leaky
    d.res2, d.res_scale, d.res2_scale, d.res_cols = ptr(res2), res_scale, res2_scale, res_cols
    d.w_static = int(w_static)
    d.bnr_x, d.bnr_coef, d.bnr_prelu, d.bnr_act, d.bnr_c = ptr(bnr_x), ptr(bnr_coef), ptr(bnr_prelu), bnr_act, bnr_c
    d.group_rows, d.bnf_mode, d.bnf_c, d.bnf_count = group_rows, bnf_mode, bnf_c, bnf_count
    d.bnf_counter, d.bnf_gamma, d.bnf_beta = ptr(bnf_counter), ptr(bnf_gamma), ptr(bnf_beta)
    d.bnf_rm, d.bnf_rv, d.bnf_nbt, d.bnf_coef = ptr(bnf_rm), ptr(bnf_rv), ptr(bnf_nbt), ptr(bnf_coef)
    d.bnf_eps, d.bnf_momentum = bnf_eps, bnf_momentum
    if rep2x is not None:       # dict(t, strides (n, h, w) of the x2 grid, ch_off): nearest x2 copy of the result
        d.out_rep2x = ptr(rep2x["t"])
        d.rep_n, d.rep_h, d.rep_w = rep2x["strides"]
        d.rep_ch_off = rep2x.get("ch_off", 0)
    if gather is not None:      # dict(k, pad, c, bias): OUT_GATHER_W, `out` is the zero-initialised fp32 NCHW result
        d.out_mode, d.out_f32 = L.OUT_GATHER_W, 1
        d.gather_k, d.gather_pad, d.gather_c, d.gather_bias = gather["k"], gather["pad"], gather["c"], ptr(gather.get("bias"))
        d.gather_rows = gather.get("rows", 1)
    return d


def conv_coresident_capacity(d: ConvDesc) -> int:
    """CTAs of this conv's kernel instantiation / shared-memory footprint the device holds at once."""
    lib = L.load()
    ctas, cap = C.c_int(0), C.c_int(0)
    L.check(lib.tsr_conv_bnf_capacity(C.byref(d), C.byref(ctas), C.byref(cap)))
    return cap.value


def conv_is_coresident(d: ConvDesc) -> bool:
    """True when every CTA of this conv's launch fits on the device at once (needed by the fused training-mode
    BatchNorm, whose CTAs meet at a grid barrier). Dry mode (no device): 2 CTAs on each of 148 SMs."""
    tiles = ((d.N * d.Ho * d.Wo + 127) // 128) * (d.cout_pad // d.block_n)
    if DRY:
        return tiles <= 2 * 148
    lib = L.load()
    ctas, cap = C.c_int(0), C.c_int(0)
    L.check(lib.tsr_conv_bnf_capacity(C.byref(d), C.byref(ctas), C.byref(cap)))
    return 0 < ctas.value <= cap.value


def dgrad_s2_descs(*, dy, N, Hy, Wy, Cout, dy_ld, wt, Cin, cin_pad, block_n, out, Hx, Wx, out_ld, n_valid,
                   out_f32=False, out_ch_off=0, **epi) -> List[ConvDesc]:
    """Data gradient of a 3x3 / stride-2 / pad-1 conv as four stride-1 convs over dY, one per output-parity class
    (rh, rw): dX[n, 2a+rh, 2b+rw, :] = sum_taps dY[n, a+oh, b+ow, :] . Wt[kh*3+kw]. `wt` is the PK_T pack
    [9][cin_pad][Cout]; extra epilogue keywords are forwarded to conv_desc (aux strides must be the caller's)."""
    assert Hx == 2 * Hy and Wx == 2 * Wy, "stride-2 data gradient expects even input sizes"
    cls = dgrad_s2_classes(3, 1)
    descs = []
    for rh in (0, 1):
        for rw in (0, 1):
            taps = [(oh, ow, kh * 3 + kw) for (oh, kh) in cls[rh] for (ow, kw) in cls[rw]]
            geom = dict(lower_h=0, lower_w=0, upper_h=0, upper_w=0, Ho=Hy, Wo=Wy, stride=1, taps=taps)
            elem = 4 if out_f32 else 2
            base = ptr(out) + (rh * Wx + rw) * out_ld * elem
            d = conv_desc(x=dy, N=N, H=Hy, W=Wy, C=Cout, x_ld=dy_ld, geom=geom, w=wt, cout_pad=cin_pad,
                          w_ld=Cout, n_slots=9, block_n=block_n, out=base, os_n=Hx * Wx * out_ld,
                          os_h=2 * Wx * out_ld, os_w=2 * out_ld, n_valid=n_valid, out_f32=out_f32,
                          out_ch_off=out_ch_off, **epi)
            d._parity = (rh, rw)
            descs.append(d)
    return descs


def gemm_desc(*, a, M, K, a_ld, a_mn_major=False, w, n_rows, block_n, out, out_ld, n_valid, splits=1, bias=None,
              act=L.ACT_NONE, atomic_t=False, out_f32=True, acc_scale=1.0, leaky=0.2, w_static=False, w_chunked=False,
              side=False) -> ConvDesc:
    """D[M, n] = sum_k A[m,k] * Wt[n,k].  atomic_t: fp32 atomics into out[n*out_ld + m] (split-K).
    w_chunked: Wt is stored as K/64 chunks [chunk][n_rows][64] (see tsr_conv_desc_t.w_chunk_rows); side: run on the
    weight-gradient side branch of a program."""
    d = ConvDesc()
    d.x, d.w = ptr(a), ptr(w)
    d.a_mode = 2 if a_mn_major else 1
    d.gemm_M, d.gemm_K, d.x_ld = M, K, a_ld
    d.stride = 1
    d.num_taps = 1
    d.block_k = 64
    d.block_n = block_n
    d.cout_pad = n_rows
    d.w_rows, d.w_ld = n_rows, K
    if w_chunked:
        assert K % 64 == 0
        d.w_rows, d.w_ld, d.w_chunk_rows = (K // 64) * n_rows, 64, n_rows
    d.side = int(side)
    d.splits = splits
    d.out, d.bias = ptr(out), ptr(bias)
    d.out_mode = L.OUT_GEMM_T_ATOMIC if atomic_t else L.OUT_LINEAR
    d.out_f32 = int(out_f32)
    d.os_n = out_ld if atomic_t else 0
    d.os_w = out_ld
    d.n_valid, d.act = n_valid, act
    d.acc_scale, d.leaky_slope = acc_scale, leaky
    d.res_scale = d.res2_scale = 1.0
    d.shuf_c = 64
    d.w_static = int(w_static)
    return d


def wgrad_desc(*, x, N, H, W, C, x_ld, geom, dy, dy_ld, dy_c, out, cout_valid, block_n, chan_block=None, dy_block=None,
               x_c0=0, dy_c0=0, splits=0) -> WgradDesc:
    d = WgradDesc()
    d.x, d.dy, d.out = ptr(x), ptr(dy), ptr(out)
    d.N, d.H, d.W, d.C, d.x_ld = N, H, W, C, x_ld
    d.Ho, d.Wo, d.dy_ld, d.dy_c = geom["Ho"], geom["Wo"], dy_ld, dy_c
    d.stride = geom["stride"]
    d.lower_h, d.lower_w, d.upper_h, d.upper_w = geom["lower_h"], geom["lower_w"], geom["upper_h"], geom["upper_w"]
    _set_taps(d, geom["taps"])
    d.chan_block = chan_block or pick_block_k(C - x_c0)
    d.dy_block = dy_block or pick_block_k(dy_c)
    d.block_n = block_n
    d.cout_valid = cout_valid
    d.x_c0, d.dy_c0, d.splits = x_c0, dy_c0, splits
    return d


def elt(kind: int, p: Iterable = (), i: Iterable = (), f: Iterable = (), side=False) -> EltDesc:
    """side: False = main chain; True / 1 = weight-gradient side branch; 2 = lowest-priority bulk side branch."""
    d = EltDesc()
    d.kind = kind
    d.side = int(side)
    for k, v in enumerate(p):
        d.p[k] = ptr(v)
    for k, v in enumerate(i):
        d.i[k] = int(v)
    for k, v in enumerate(f):
        d.f[k] = float(v)
    return d


# --------------------------------------------------------------------------------------------- extent checks
def validate_conv(d: ConvDesc):
    esz = 4 if d.out_f32 else 2
    if d.a_mode == 0:
        M = d.N * d.Ho * d.Wo
        _need("conv x", d.x, ((d.N * d.H * d.W - 1) * d.x_ld + d.C) * 2)
        last = (d.N - 1) * d.os_n + (d.Ho - 1) * d.os_h + (d.Wo - 1) * d.os_w
        aux_last = (d.N - 1) * d.aux_n + (d.Ho - 1) * d.aux_h + (d.Wo - 1) * d.aux_w
        if d.out_mode == L.OUT_SHUFFLE:
            last = (d.N - 1) * d.os_n + (2 * d.Ho - 1) * d.os_h + (2 * d.Wo - 1) * d.os_w
            span = d.shuf_c
        elif d.out_mode == L.OUT_UNSHUFFLE:
            last = (d.N - 1) * d.os_n + ((d.Ho - 1) // 2) * d.os_h + ((d.Wo - 1) // 2) * d.os_w + 3 * d.shuf_c
            span = d.n_valid
        else:
            span = d.n_valid
        if d.out_mode == L.OUT_GATHER_W:
            g_rows = 2 if d.gather_rows == 2 else 1
            if (d.gather_k < 1 or d.gather_c < 1 or g_rows * 32 != d.block_n or d.block_n != d.cout_pad
                    or not 0 <= d.gather_pad < d.gather_k or d.splits > 1 or not d.out_f32 or d.out_preact or d.bias
                    or d.res or d.bwd_z or d.bnr_x or d.stats_partial or d.bnf_mode):
                raise ExtentError("OUT_GATHER_W: one N tile holding gather_k*gather_c columns, fp32 output, plain epilogue")
            _need("conv gather out", d.out, d.N * d.gather_c * g_rows * d.Ho * d.Wo * 4)
            if d.gather_bias:
                _need("conv gather bias", d.gather_bias, d.gather_c * 4)
        else:
            _need("conv out", d.out, (last + d.out_ch_off + span) * esz)
        if d.out_preact:
            _need("conv out_preact", d.out_preact, (last + d.out_ch_off + span) * 2)
        if d.out_rep2x:
            if d.out_mode != L.OUT_LINEAR or d.out_f32 or d.bnf_mode or d.bnr_apply or d.splits > 1:
                raise ExtentError("out_rep2x: linear bf16 store of an unsplit conv without fused BatchNorm only")
            rep_last = (d.N - 1) * d.rep_n + (2 * d.Ho - 1) * d.rep_h + (2 * d.Wo - 1) * d.rep_w
            _need("conv out_rep2x", d.out_rep2x, (rep_last + d.rep_ch_off + d.n_valid) * 2)
        rc = min(d.n_valid, d.res_cols) if d.res_cols > 0 else d.n_valid
        for name, p, cols in (("res", d.res, rc), ("res2", d.res2, rc), ("bwd_z", d.bwd_z, d.n_valid),
                              ("bnr_x", d.bnr_x, d.n_valid)):
            if p:
                _need("conv " + name, p, (aux_last + d.aux_ch_off + cols) * 2)
    else:
        M = d.gemm_M
        rows, cols = (d.gemm_M, d.gemm_K) if d.a_mode == 1 else (d.gemm_K, d.gemm_M)
        _need("gemm a", d.x, ((rows - 1) * d.x_ld + cols) * 2)
        if d.out_mode == L.OUT_GEMM_T_ATOMIC:
            _need("gemm out^T", d.out, ((d.n_valid - 1) * d.os_n + M) * 4)
        else:
            _need("gemm out", d.out, ((M - 1) * d.os_w + d.out_ch_off + d.n_valid) * esz)
    _need("conv w", d.w, d.w_rows * d.w_ld * 2)
    if d.bias:
        _need("conv bias", d.bias, d.cout_pad * 4)
    if d.stats_partial:
        if d.stats_ld < d.cout_pad:
            raise ExtentError("conv stats_ld smaller than cout_pad")
        _need("conv stats_partial", d.stats_partial, d.stats_ld * 2 * 4)
    if d.dalpha_partial:
        _need("conv dalpha_partial", d.dalpha_partial, 4)
    if d.a_mode == 0 and d.splits > 1:
        tiles = ((d.N * d.Ho * d.Wo + 127) // 128) * (d.cout_pad // d.block_n)
        if not d.ws or not d.tile_counters or d.ws_ld < d.cout_pad or d.ws_ld % 4:
            raise ExtentError("split-K conv needs ws / tile_counters and ws_ld >= cout_pad (multiple of 4)")
        _need("conv ws", d.ws, d.N * d.Ho * d.Wo * d.ws_ld * 4)
        _need("conv tile_counters", d.tile_counters, tiles * 4)
    if d.bnr_apply:
        if not (d.bnr_x and d.bnr_dx and d.bnr_coef and d.bnr_gamma and d.bnf_counter and d.bnr_count > 0) or d.out_f32:
            raise ExtentError("fused BatchNorm-backward apply: incomplete descriptor")
        last = (d.N - 1) * d.os_n + (d.Ho - 1) * d.os_h + (d.Wo - 1) * d.os_w
        _need("conv bnr_dx", d.bnr_dx, (last + d.out_ch_off + d.n_valid) * 2)
        for nm, pp, nb in (("gamma", d.bnr_gamma, d.bnr_c * 4), ("dgamma", d.bnr_dgamma, d.bnr_c * 4),
                           ("dbeta", d.bnr_dbeta, d.bnr_c * 4), ("dalpha", d.bnr_dalpha, 4),
                           ("counter", d.bnf_counter, (d.cout_pad // d.block_n) * 4)):
            if pp:
                _need("conv bnr " + nm, pp, nb)
    if d.bnr_x:
        if not d.stats_partial or d.bwd_z or d.out_mode != L.OUT_LINEAR:
            raise ExtentError("bnr_x needs stats_partial, no bwd_z hook and a linear store")
        if d.bnr_coef:
            if d.bnr_c < d.n_valid:
                raise ExtentError("bnr_c smaller than the stored column count")
            _need("conv bnr_coef", d.bnr_coef, 4 * d.bnr_c * 4)
    groups = 2 if d.group_rows else 1
    if d.group_rows and (d.group_rows % 32 or d.a_mode != 0 or not 0 < d.group_rows < d.N * d.Ho * d.Wo):
        raise ExtentError("group_rows must be a multiple of 32 inside (0, M) of an im2col conv")
    if groups == 2 and d.stats_partial:
        _need("conv stats_partial (2 groups)", d.stats_partial, 2 * d.stats_ld * 2 * 4)
    if groups == 2 and d.bnr_coef:
        _need("conv bnr_coef (2 groups)", d.bnr_coef, 2 * 4 * d.bnr_c * 4)
    if d.bnf_mode:
        if d.bnf_mode not in (1, 2) or d.a_mode != 0 or d.bias or d.bwd_z or d.bnr_x or d.out_mode != L.OUT_LINEAR:
            raise ExtentError("fused BatchNorm forward: unsupported epilogue combination")
        if d.bnf_c < d.n_valid:
            raise ExtentError("bnf_c smaller than the stored column count")
        for nm, pp, nb in (("gamma", d.bnf_gamma, d.bnf_c * 4), ("beta", d.bnf_beta, d.bnf_c * 4),
                           ("running_mean", d.bnf_rm, d.bnf_c * 4), ("running_var", d.bnf_rv, d.bnf_c * 4),
                           ("num_batches_tracked", d.bnf_nbt, 8), ("coef", d.bnf_coef, groups * 4 * d.bnf_c * 4),
                           ("counter", d.bnf_counter, (d.cout_pad // d.block_n) * 4)):
            if pp:
                _need("conv bnf " + nm, pp, nb)
        if d.bnf_mode == 1 and not (d.stats_partial and d.bnf_counter and d.bnf_count > 0):
            raise ExtentError("training-mode fused BatchNorm needs stats_partial, bnf_counter and bnf_count")
    if d.n_valid % 16 or d.n_valid > d.cout_pad:
        raise ExtentError("conv n_valid must be a multiple of 16 and <= cout_pad")
    for off in (d.out_ch_off, d.aux_ch_off):
        if off % 8:
            raise ExtentError("channel offsets must be multiples of 8 (16-byte vector accesses)")
    if d.out_mode not in (L.OUT_GEMM_T_ATOMIC, L.OUT_GATHER_W):
        q = 4 if d.out_f32 else 8
        for st in ((d.os_n, d.os_h, d.os_w) if d.a_mode == 0 else (d.os_w,)):
            if st % q:
                raise ExtentError("output strides must keep 16-byte alignment")
        if d.res or d.bwd_z or d.bnr_x:
            for st in ((d.aux_n, d.aux_h, d.aux_w) if d.a_mode == 0 else (d.aux_w,)):
                if st % 8:
                    raise ExtentError("aux strides must keep 16-byte alignment")


def validate_wgrad(d: WgradDesc):
    M = d.N * d.Ho * d.Wo
    _need("wgrad x", d.x, ((d.N * d.H * d.W - 1) * d.x_ld + d.C) * 2)
    _need("wgrad dy", d.dy, ((M - 1) * d.dy_ld + d.dy_c0 + d.dy_c) * 2)
    _need("wgrad out", d.out, d.cout_valid * d.num_taps * (d.C - d.x_c0) * 4)
    if d.cout_valid % 16:
        raise ExtentError("wgrad cout_valid (accumulator row width) must be a multiple of 16")


def _elt_extents(d: EltDesc):
    """(pointer index, bytes) pairs the kernel of this kind touches; conservative for the kinds used by the plans."""
    i, k = d.i, d.kind
    if k == L.E_IM2ROW:
        return [(0, i[0] * i[1] * i[2] * i[3] * 4), (1, i[0] * i[2] * i[3] * i[9] * 2)]
    if k == L.E_GATHER_OUT:
        return [(0, i[0] * i[2] * i[3] * i[9] * (2 if i[10] else 4)), (1, i[0] * i[1] * i[2] * i[3] * 4), (2, i[1] * 4)]
    if k == L.E_NCHW2NHWC:
        return [(0, i[0] * i[1] * i[2] * i[3] * 4), (1, ((i[0] * i[2] * i[3] - 1) * i[4] + i[5] + i[1]) * 2)]
    if k == L.E_NHWC2NCHW:
        return [(0, ((i[0] * i[2] * i[3] - 1) * i[4] + i[5] + i[1]) * 2), (1, i[0] * i[1] * i[2] * i[3] * 4)]
    if k in (L.E_BN_FINALIZE, L.E_BN_EVAL_COEF, L.E_BN_BWD_FINALIZE):
        raise ExtentError("kernel kind retired: the finalize steps are fused into BN_ACT / BN_BWD_APPLY")
    if k == L.E_BN_ACT:
        M, C = i[0], i[1]
        if 256 % (C // 8) or C % 8:
            raise ExtentError("BN_ACT needs C/8 to divide 256")
        G = 2 if i[11] else 1
        if i[11] and (i[11] % 8 or not 0 < i[11] < M):
            raise ExtentError("BN_ACT group_rows must lie inside (0, M)")
        return [(0, ((M - 1) * i[2] + i[6] + C) * 2), (1, G * 2 * C * 4), (2, ((M - 1) * i[3] + i[7] + C) * 2),
                (3, ((M - 1) * i[4] + i[8] + C) * 2), (4, 4), (5, C * 4), (6, C * 4), (7, C * 4), (8, C * 4), (9, 8),
                (10, G * 4 * C * 4)]
    if k == L.E_BN_BWD_REDUCE:
        M, C = i[0], i[1]
        return [(0, ((M - 1) * i[4] + C) * 2), (1, ((M - 1) * i[5] + C) * 2), (2, 4 * C * 4 if i[6] else 0), (3, 4),
                (4, C * 8), (5, 4), (6, ((M - 1) * i[4] + C) * 2)]
    if k == L.E_BN_BWD_APPLY:
        M, C = i[0], i[1]
        if 256 % (C // 8) or C % 8:
            raise ExtentError("BN_BWD_APPLY needs C/8 to divide 256")
        G = 2 if (i[9] and i[6]) else 1
        return [(0, ((M - 1) * i[3] + C) * 2), (1, ((M - 1) * i[4] + C) * 2), (2, G * 4 * C * 4 if i[6] else 0),
                (3, G * C * 8), (4, 4), (5, ((M - 1) * i[5] + C) * 2), (6, ((M - 1) * i[3] + C) * 2), (7, C * 4),
                (8, C * 4), (9, C * 4), (10, 4), (11, 4)]
    if k == L.E_COLSUM_FINALIZE:
        return [(0, i[0] * i[2] * 8), (1, i[1] * 4)]
    if k == L.E_SUM_FINALIZE:
        return [(0, i[0] * 4), (1, 4)]
    if k == L.E_LINEAR_WGRAD:
        return [(0, i[0] * i[1] * 4), (1, i[0] * i[2] * 4), (2, i[1] * i[2] * 4), (3, i[1] * 4)]
    if k == L.E_LOSS:
        return [(0, i[0] * 4), (1, i[0] * 4), (2, i[2] * 4), (3, i[0] * 4)]
    if k == L.E_ZERO:
        return [(0, i[0])]
    if k == L.E_PACK_GATHER:
        return [(1, i[0] * 4), (2, i[0] * 2)]
    if k == L.E_UPSAMPLE2X:
        return [(0, ((i[0] * i[1] * i[2] - 1) * i[4] + i[3]) * 2), (1, ((4 * i[0] * i[1] * i[2] - 1) * i[5] + i[3]) * 2)]
    if k == L.E_UPSAMPLE2X_BWD:
        return [(0, ((4 * i[0] * i[1] * i[2] - 1) * i[4] + i[3]) * 2), (1, ((i[0] * i[1] * i[2] - 1) * i[5] + i[3]) * 2)]
    if k == L.E_HEAD:
        B, N1 = i[0], i[1]
        return [(0, N1 * B * 4), (1, N1 * 4), (2, N1 * 4), (3, 4), (4, B * 4), (5, B * N1 * 4)]
    if k == L.E_HEAD_BWD:
        B, N1 = i[0], i[1]
        ld = i[3] if i[3] > 0 else N1
        return [(0, B * 4), (1, B * 4), (2, B * N1 * 4), (3, N1 * 4), (4, B * N1 * 4), (5, ((B - 1) * ld + N1) * 2),
                (6, N1 * 4), (7, 4), (8, N1 * 4)]
    if k in (L.E_MAXPOOL2, L.E_MAXPOOL2_BWD):
        big, small = i[0] * i[1] * i[2] * i[3] * 2, i[0] * (i[1] // 2) * (i[2] // 2) * i[3] * 2
        if i[1] % 2 or i[2] % 2 or i[3] % 8:
            raise ExtentError("MAXPOOL2 needs even H, W and C % 8 == 0")
        return [(0, big), (1, small)] + ([(2, small), (3, big)] if k == L.E_MAXPOOL2_BWD else [])
    if k == L.E_AXPBY:
        return [(0, i[0] * 2), (1, i[0] * 2), (2, i[0] * 2)]
    if k == L.E_CAST:
        return [(0, i[0] * (4 if i[1] == 0 else 2)), (1, i[0] * (2 if i[1] == 0 else 4))]
    if k == L.E_CHANSUM_NCHW:
        return [(0, i[0] * i[1] * i[2] * 4), (1, i[3] * i[1] * 8)]
    if k == L.E_CROP_LR:
        return [(1, i[0] * 6 * 8), (2, i[2] * i[3] * 4), (3, i[2] * 2 * 4), (4, i[0] * 3 * i[1] * i[1] * 4),
                (5, i[0] * 3 * i[2] * i[2] * 4)]
    if k == L.E_FEAT_T:
        chunks = (i[0] + 63) // 64
        return [(0, ((i[0] * i[2] - 1) * i[3] + i[4] + i[1]) * 2), (1, chunks * i[1] * i[2] * 64 * 2)]
    if k == L.E_GAN_LOSS:
        return [(0, i[0] * 4), (1, i[1] * 4), (2, 4), (3, i[0] * 4), (4, i[1] * 4), (5, i[0] * 4)]
    if k == L.E_AXPBY_F32:
        return [(0, i[0] * 4), (1, i[0] * 4), (2, i[0] * 4), (3, 4)]
    return []


def validate_elt(d: EltDesc):
    for idx, nbytes in _elt_extents(d):
        if d.p[idx]:
            _need(f"elt kind {d.kind} p[{idx}]", d.p[idx], int(nbytes))


def validate(d):
    if isinstance(d, ConvDesc):
        validate_conv(d)
    elif isinstance(d, WgradDesc):
        validate_wgrad(d)
    else:
        validate_elt(d)


# --------------------------------------------------------------------------------------------- programs
GROUP_LAUNCH = os.environ.get("TSR_GROUP_LAUNCH", "1") != "0"


class ConvGroupDesc:
    """Introspection record of a grouped launch (Program.descs): the member descriptors."""

    def __init__(self, members: List):
        self.members = list(members)


class Program:
    """A recorded launch list owned by the native library (tsr_prog_t)."""

    def __init__(self):
        self.keep: List = []  # tensors referenced by raw pointer
        self.marks = {}
        self.descs: List = []  # kept for introspection (tests, launch accounting)
        if DRY:
            self._lib, self._h = None, None
            return
        self._lib = L.load()
        self._h = self._lib.tsr_prog_create()

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.tsr_prog_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # Deferred emission: a data-gradient conv may be held back until the next emission so that a BatchNorm-backward
    # stage that consumes its output can first fold its column reductions into the conv's epilogue (engine.py
    # norm_act_bwd). Any other emission, run(), mark() or len() flushes the held descriptors unchanged.
    _deferred = None

    def defer(self, descs: List, tag, group: bool = False):
        """group: the held descriptors are independent convs of one tile grid (stride-2 data-gradient parity classes)
        and are emitted as ONE grouped launch."""
        self.flush()
        self._deferred = (list(descs), tag)
        self._deferred_group = group

    def add_group(self, descs: List) -> int:
        """Up to four unsplit im2col convs in one launch (tsr_prog_add_conv_group)."""
        self.flush()
        for d in descs:
            validate(d)
        if len(descs) == 1 or not GROUP_LAUNCH:
            r = -1
            for d in descs:
                r = self.add(d)
            return r
        self.descs.append(ConvGroupDesc(descs))
        if DRY:
            return len(self.descs) - 1
        arr = (ConvDesc * len(descs))(*descs)
        r = self._lib.tsr_prog_add_conv_group(self._h, arr, len(descs))
        if r < 0:
            L.check(r)
        return r

    def take_deferred(self, tag):
        """Returns and removes the held descriptors if they were deferred under `tag` (identity), else None."""
        if self._deferred is not None and self._deferred[1] is tag:
            descs = self._deferred[0]
            self._deferred = None
            return descs
        return None

    def flush(self):
        if self._deferred is not None:
            descs, self._deferred = self._deferred[0], None
            if getattr(self, "_deferred_group", False) and len(descs) > 1:
                self.add_group(descs)
            else:
                for d in descs:
                    self.add(d)

    def add(self, d) -> int:
        self.flush()
        validate(d)
        self.descs.append(d)
        if DRY:
            return len(self.descs) - 1
        if isinstance(d, ConvDesc):
            r = self._lib.tsr_prog_add_conv(self._h, C.byref(d))
        elif isinstance(d, WgradDesc):
            r = self._lib.tsr_prog_add_wgrad(self._h, C.byref(d))
        else:
            r = self._lib.tsr_prog_add_elt(self._h, C.byref(d))
        if r < 0:
            L.check(r)
        return r

    def mark(self, name: str):
        self.marks[name] = len(self)

    def __len__(self) -> int:
        self.flush()
        return len(self.descs)

    def run(self, first: int = 0, count: int = -1, stream: Optional[int] = None):
        self.flush()
        if DRY:
            return
        L.check(self._lib.tsr_prog_run(self._h, first, count, stream if stream is not None else current_stream()))


def run_now(d, stream: Optional[int] = None):
    """Immediate-mode launch of a single descriptor."""
    validate(d)
    if DRY:
        return
    lib = L.load()
    st = stream if stream is not None else current_stream()
    if isinstance(d, ConvDesc):
        L.check(lib.tsr_conv(C.byref(d), st))
    elif isinstance(d, WgradDesc):
        L.check(lib.tsr_wgrad(C.byref(d), st))
    else:
        L.check(lib.tsr_elt(C.byref(d), st))


def check_watchdog():
    if DRY:
        return
    lib = L.load()
    L.check(lib.tsr_check_watchdog(current_stream()))


# --------------------------------------------------------------------------------------------- pack tables
def pack_table(entries: List[dict], device):
    """Builds the device table for TSR_E_PACK_W / TSR_E_UNPACK_G. Each entry: src, dst, mode, cout, cin, kh, kw,
    rows_pad, cols_pad, shuffle, count. Returns (table tensor, n_entries, n_blocks)."""
    import torch
    arr = (PackEntry * len(entries))()
    blocks = 0
    for k, e in enumerate(entries):
        pe = arr[k]
        pe.src, pe.dst = ptr(e["src"]), ptr(e["dst"])
        pe.mode = e["mode"]
        pe.cout, pe.cin, pe.kh, pe.kw = e["cout"], e["cin"], e["kh"], e["kw"]
        pe.rows_pad, pe.cols_pad, pe.shuffle = e["rows_pad"], e["cols_pad"], int(e.get("shuffle", 0))
        pe.block_start = blocks
        pe.count = e["count"]
        if e["mode"] == L.PK_LINEAR and e["kh"] * e["kw"] <= 64 and e["cols_pad"] == e["cin"] * e["kh"] * e["kw"]:
            # tiled permutation: one block per (output row, 32-channel chunk); flagged through `shuffle`
            pe.shuffle = 1
            blocks += e["rows_pad"] * ((e["cin"] + 31) // 32)
        else:
            blocks += (e["count"] + 1023) // 1024
    raw = bytes(arr)
    t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
    return t, len(entries), blocks
