"""Plan compiler of the B200 path: turns one nn.Module call (generator or discriminator, at one input shape)
into recorded native launch lists (ops.Program) over pre-allocated NHWC bf16 workspaces, and binds them to autograd.

Replaces the per-op ATen dispatch of the reference's forward() methods and autograd graph
(torchsr/srgan/generator.py:60-81, discriminator.py:71-88, the esrgan twins). Layout decisions:

* activations: NHWC bf16, one buffer per tensor that backward needs (raw conv outputs, activations, pre-activations);
* parameters stay fp32 OIHW torch Parameters (the state_dict contract); bf16 packed copies ([tap][Cout][Cin] for
  forward, [tap][Cin][Cout] for the data gradient) live in one arena per module and are re-packed by ONE kernel
  whenever a parameter's version counter moved (i.e. after optimizer.step / load_state_dict);
* weight gradients: fp32 atomically-accumulated arena in packed order -> ONE unpack kernel -> flat fp32 gradient in
  parameters() order -> handed to autograd as views of a clone.
"""
import os
import weakref
from typing import Callable, Dict, List, Optional

import torch
from torch import nn

from . import _lib as L
from . import ops

BF16 = torch.bfloat16
F32 = torch.float32
NUM_SMS = 148
# fold BatchNorm-backward column reductions into the epilogue of the data-gradient conv that produces the gradient
FUSE_BN_REDUCE = os.environ.get("TSR_FUSE_BN_REDUCE", "1") != "0"
# convs with at least this many K iterations (taps x channel chunks) and an under-filled grid split K. Off by default
# (0): measured on B200 the last-CTA finalize (dependent L2 reads of the fp32 partial sums) costs more than the shorter
# main loop saves for every layer of this workload (profiles/r01_splitk_experiment.md); the mode stays parity-tested.
SPLIT_K_MIN_ITERS = int(os.environ.get("TSR_SPLITK_MIN_ITERS", "0"))
# conv + BatchNorm (+ activation, + residual) in ONE launch: training mode crosses a grid barrier inside the conv kernel
# (only for grids the device holds at once), eval mode folds the running statistics into the epilogue. "0" keeps the
# two-launch path (conv with column sums, then bn_act_kernel) everywhere - both stay parity-tested.
FUSE_BN_FWD = os.environ.get("TSR_BN_FUSE", "1") != "0"
# 64-column N tiles for the convs the staged epilogue serves (see Plan.conv_fwd); TSR_CONV_STAGED=0 also disables it in C
def _staged_n64() -> bool:     # read per plan build, like the C side reads its switches per descriptor
    return os.environ.get("TSR_CONV_STAGED", "1") != "0" and os.environ.get("TSR_CONV_PERSISTENT", "1") != "0"
# the same for backward: dx = A*dz + B*x + C formed in the epilogue of the data-gradient conv that produced dz (after a
# grid barrier on the column sums) instead of a bn_bwd_apply_kernel launch; "0" keeps the separate launch
FUSE_BN_BWD = os.environ.get("TSR_BN_BWD_FUSE", "1") != "0"
# pick_block_n halves the N tile while the grid stays at or below this many CTAs (x2)
HALVE_BELOW_CTAS = int(os.environ.get("TSR_HALVE_BELOW_CTAS", str(NUM_SMS + 20)))
ZERO_ARENA_FLOATS = 64 * 1024


def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


class GroupSplitError(RuntimeError):
    """A two-group (real | fake) BatchNorm pass was asked for a shape whose per-group row count is not a multiple of 32
    in some layer; B200Module.forward_pair then falls back to two separate calls."""


class Act:
    """A [B, H, W, C] NHWC view: tensor + geometry (ld = pixel stride in elements, c0 = first channel)."""
    __slots__ = ("t", "B", "H", "W", "C", "ld", "c0", "hook")

    def __init__(self, t, B, H, W, C, ld=None, c0=0):
        self.t, self.B, self.H, self.W, self.C = t, B, H, W, C
        self.ld = ld if ld is not None else C
        self.c0 = c0
        # Set by the producer of this activation when its own activation backward (and PixelShuffle inverse) can be
        # fused into the epilogue of the data-gradient conv that computes d/d(this): dict(bwd_z=Act, bwd_act=int,
        # prelu=Tensor, unshuffle_to=Act, dalpha_partial=Tensor).
        self.hook = None

    @property
    def M(self) -> int:
        return self.B * self.H * self.W

    def strides(self):
        return (self.H * self.W * self.ld, self.W * self.ld, self.ld)


# ------------------------------------------------------------------------------------------------ parameter store
# Fused optimizers (torch.optim.Adam(fused=True), which the trainers use) update parameters through
# torch._fused_adam_, which does NOT bump Tensor._version - so the version counters alone cannot tell that the bf16
# operand copies went stale. Every ParamStore is therefore also marked dirty by a global optimizer-step hook whenever
# an optimizer that owns one of its parameters has stepped (also while a CUDA graph is being captured: the re-pack
# kernel of the next forward is then part of the graph).
_LIVE_STORES = weakref.WeakSet()
_HOOK_INSTALLED = False


def _optimizer_step_hook(optimizer, args, kwargs):
    ids = getattr(optimizer, "_tsr_param_ids", None)
    n = sum(len(g["params"]) for g in optimizer.param_groups)
    if ids is None or ids[0] != n:
        ids = (n, {id(p) for g in optimizer.param_groups for p in g["params"]})
        optimizer._tsr_param_ids = ids
    for st in list(_LIVE_STORES):
        if not st.dirty and any(id(p) in ids[1] for p in st._watched):
            st.dirty = True


def _install_optimizer_hook():
    global _HOOK_INSTALLED
    if not _HOOK_INSTALLED:
        from torch.optim.optimizer import register_optimizer_step_post_hook
        register_optimizer_step_post_hook(_optimizer_step_hook)
        _HOOK_INSTALLED = True


class ConvRec:
    """Packing / gradient bookkeeping of one nn.Conv2d.
    kind: 'std'   KxK conv, Cin multiple of 16: implicit GEMM over the NHWC input (TMA im2col)
          'fullk' Cin == 3: the input is expanded by the im2row kernel to K*K*3 (padded) columns, then a 1-tap GEMM
          'rown'  9x9, Cout == 3 (SRGAN conv3): 9 vertical taps with N' = kw*3+co columns, horizontal taps summed by
                  the gather kernel
    """

    def __init__(self, name: str, conv: nn.Conv2d, kind: str = "std", shuffle: bool = False, need_dgrad: bool = True):
        self.name, self.conv, self.kind, self.shuffle, self.need_dgrad = name, conv, kind, shuffle, need_dgrad
        self.weight, self.bias = conv.weight, conv.bias
        self.cout, self.cin, self.k, _ = conv.weight.shape
        self.stride = conv.stride[0]
        self.pad = conv.padding[0]
        k = self.k
        if kind == "std":
            assert self.cin % 16 == 0
            self.cout_pad = _round_up(self.cout, 16)
            self.block_n = self.cout_pad if self.cout_pad <= 128 else 128
            assert self.cout_pad % self.block_n == 0
            self.slots, self.cols = k * k, self.cin
            self.fwd_mode, self.t_mode = L.PK_FWD, L.PK_T
            self.t_rows, self.t_cols = self.cin, _round_up(self.cout, 16)
            self.acc_rows, self.acc_taps, self.acc_cols = self.cout_pad, k * k, self.cin
        elif kind == "fullk":
            self.epad = _round_up(k * k * self.cin, 32)
            self.cout_pad = self.cout
            self.block_n = min(self.cout, 128)
            self.slots, self.cols = 1, self.epad
            self.fwd_mode, self.t_mode = L.PK_FULLK, L.PK_T
            self.t_rows, self.t_cols = _round_up(k * k * self.cin, 32), self.cout   # [k*k*cin (padded)][cout]
            self.acc_rows, self.acc_taps, self.acc_cols = self.cout, 1, self.epad
        elif kind == "rown":
            self.npad = _round_up(k * self.cout, 32)     # 27 -> 32 columns (kw*cout + co)
            self.cout_pad = self.npad
            self.block_n = self.npad
            self.slots, self.cols = k, self.cin
            self.fwd_mode, self.t_mode = L.PK_ROWN, L.PK_ROWN_T
            self.t_rows, self.t_cols = self.cin, self.npad
            self.acc_rows, self.acc_taps, self.acc_cols = self.npad, k, self.cin
        else:
            raise ValueError(kind)
        self.w_fwd = self.w_t = self.acc = None


class LinearRec:
    """nn.Linear applied to a flattened NHWC feature map [B][Hf*Wf*C] (weight columns permuted from (c,h,w))."""

    def __init__(self, name: str, lin: nn.Linear, C: int, Hf: int, Wf: int):
        self.name, self.lin, self.C, self.Hf, self.Wf = name, lin, C, Hf, Wf
        self.weight, self.bias = lin.weight, lin.bias
        self.nout, self.K = lin.weight.shape
        assert self.K == C * Hf * Wf
        self.nout_pad = _round_up(self.nout, 128)
        self.w_fwd = None


class ParamStore:
    """Everything derived from one module's parameters: packed bf16 weights, wgrad accumulators, flat gradient."""

    def __init__(self, module: nn.Module, convs: List[ConvRec], linears: List[LinearRec], device):
        self.module, self.convs, self.linears, self.device = module, convs, linears, device
        self.params = [p for _, p in module.named_parameters()]
        self.offsets: Dict[int, int] = {}
        off = 0
        for p in self.params:
            self.offsets[id(p)] = off
            off += _round_up(p.numel(), 4)          # keep every slice 16-byte aligned
        self.total = off
        # ---- arenas: bf16 packed operands, fp32 weight-gradient accumulators (packed order)
        w_elems, acc_elems = 0, 0
        for r in convs:
            r._fwd_off, w_elems = w_elems, w_elems + _round_up(r.slots * r.cout_pad * r.cols, 128)
            if r.need_dgrad:
                r._t_off, w_elems = w_elems, w_elems + _round_up(self._t_slots(r) * r.t_rows * r.t_cols, 128)
            r._acc_off, acc_elems = acc_elems, acc_elems + _round_up(r.acc_rows * r.acc_taps * r.acc_cols, 64)
        for r in linears:
            r._fwd_off, w_elems = w_elems, w_elems + _round_up(r.nout_pad * r.K, 128)
        self.w_arena = torch.zeros(max(w_elems, 128), dtype=BF16, device=device)
        self.acc_elems = max(acc_elems, 64)
        self._version = None
        self.dirty = True
        self.opt_fresh = False

        # set by optim.FusedAdam(late_numel=...): the event after which the weights behind the forward program's
        # "late_weights" mark (the big Linear layer) are valid
        self.late_event = None
        self._build_tables()
        _LIVE_STORES.add(self)
        _install_optimizer_hook()

    # T-pack slot counts differ per kind
    @staticmethod
    def _t_slots(r: ConvRec) -> int:
        return {"std": r.k * r.k, "fullk": 1, "rown": r.k}[r.kind]

    def _build_tables(self):
        pack = []
        for r in self.convs:
            n_fwd = r.slots * r.cout_pad * r.cols
            r.w_fwd = self.w_arena[r._fwd_off:r._fwd_off + n_fwd]
            pack.append(dict(src=r.weight, dst=r.w_fwd, mode=r.fwd_mode, cout=r.cout, cin=r.cin, kh=r.k, kw=r.k,
                             rows_pad=r.cout_pad, cols_pad=r.cols, shuffle=r.shuffle, count=n_fwd))
            if r.need_dgrad:
                if r.kind == "fullk":
                    # [k*k*cin][cout] == PK_T with rows_pad = cin: rows t*cin+ci; the padded tail rows stay zero
                    n_t = r.k * r.k * r.cin * r.t_cols
                    r.w_t = self.w_arena[r._t_off:r._t_off + r.t_rows * r.t_cols]
                    pack.append(dict(src=r.weight, dst=r.w_t, mode=L.PK_T, cout=r.cout, cin=r.cin, kh=r.k, kw=r.k,
                                     rows_pad=r.cin, cols_pad=r.t_cols, shuffle=0, count=n_t))
                else:
                    n_t = self._t_slots(r) * r.t_rows * r.t_cols
                    r.w_t = self.w_arena[r._t_off:r._t_off + n_t]
                    pack.append(dict(src=r.weight, dst=r.w_t, mode=r.t_mode, cout=r.cout, cin=r.cin, kh=r.k, kw=r.k,
                                     rows_pad=r.t_rows, cols_pad=r.t_cols, shuffle=r.shuffle, count=n_t))
        for r in self.linears:
            n = r.nout_pad * r.K
            r.w_fwd = self.w_arena[r._fwd_off:r._fwd_off + n]
            pack.append(dict(src=r.weight, dst=r.w_fwd, mode=L.PK_LINEAR, cout=r.nout, cin=r.C, kh=r.Hf, kw=r.Wf,
                             rows_pad=r.nout_pad, cols_pad=r.K, shuffle=0, count=n))
        self._pack_tab, self._pack_n, self._pack_blocks = ops.pack_table(pack, self.device)
        self._pack_desc = ops.elt(L.E_PACK_W, p=[self._pack_tab], i=[self._pack_n, self._pack_blocks])
        # the packs torchsr_b200.optim.FusedAdam does not write itself (first / last layers with 3 channels)
        std_dsts = {r.w_fwd.data_ptr() for r in self.convs if r.kind == "std"} | \
                   {r.w_t.data_ptr() for r in self.convs if r.kind == "std" and r.need_dgrad} | \
                   {r.w_fwd.data_ptr() for r in self.linears if r.Hf * r.Wf <= 64}
        special = [e for e in pack if e["dst"].data_ptr() not in std_dsts]
        self._pack_special_desc = None
        if special:
            self._sp_tab, n, blocks = ops.pack_table(special, self.device)
            self._pack_special_desc = ops.elt(L.E_PACK_W, p=[self._sp_tab], i=[n, blocks])
        self._watched = [r.weight for r in self.convs] + [r.weight for r in self.linears]

    def ensure_packed(self):
        """Re-packs the bf16 operand copies if any watched parameter changed since the last pack (one kernel)."""
        ver = 0
        for p in self._watched:
            ver += p._version
        if ver == self._version and not self.dirty:
            return
        if ver == self._version and self.opt_fresh:
            # stepped by torchsr_b200.optim.FusedAdam only: it rewrote the 'std' conv and Linear packs itself
            if self._pack_special_desc is not None:
                ops.run_now(self._pack_special_desc)
        else:
            ops.run_now(self._pack_desc)
        self.opt_fresh = False
        self._version = ver
        self.dirty = False

    def grads_from_flat(self, flat: torch.Tensor, want: List[bool]) -> List[Optional[torch.Tensor]]:
        out = []
        for p, w in zip(self.params, want):
            if not w:
                out.append(None)
                continue
            o = self.offsets[id(p)]
            out.append(flat[o:o + p.numel()].view(p.shape))
        return out


class GradBuffers:
    """Per-plan-instance gradient storage: the fp32 weight-gradient accumulators (packed order), the flat fp32 gradient
    in parameters() order and the unpack table between them. Owned by the plan instance (not the module) so that
    the backward passes of two outstanding calls of one module - D(real) and D(fake) - may run concurrently on two
    streams. Allocated on the first backward that wants weight gradients."""

    def __init__(self, store: ParamStore):
        self.store = store
        self.flat = None

    def _ensure(self):
        if self.flat is not None:
            return
        st = self.store
        self.flat = torch.zeros(max(st.total, 4), dtype=F32, device=st.device)
        self.acc_arena = torch.zeros(st.acc_elems, dtype=F32, device=st.device)
        self.unpack_desc = None      # (set below; unpack_for() calls _ensure() again while it is being built)
        self._tabs = []
        self.unpack_desc = self.unpack_for(st.convs)

    def unpack_for(self, convs: List[ConvRec]):
        """One-launch descriptor that unpacks the weight-gradient accumulators of `convs` into the flat gradient."""
        self._ensure()
        unpack = []
        for r in convs:
            n_acc = r.acc_rows * r.acc_taps * r.acc_cols
            unpack.append(dict(src=self.acc(r), dst=self.grad_slice(r.weight), mode=r.fwd_mode, cout=r.cout, cin=r.cin,
                               kh=r.k, kw=r.k, rows_pad=r.acc_rows, cols_pad=r.acc_cols, shuffle=r.shuffle, count=n_acc))
        tab, n, blocks = ops.pack_table(unpack, self.store.device)
        self._tabs.append(tab)
        return ops.elt(L.E_UNPACK_G, p=[tab], i=[n, blocks])

    def grad_slice(self, p: torch.Tensor) -> torch.Tensor:
        self._ensure()
        o = self.store.offsets[id(p)]
        return self.flat[o:o + p.numel()]

    def acc(self, r: ConvRec) -> torch.Tensor:
        self._ensure()
        return self.acc_arena[r._acc_off:r._acc_off + r.acc_rows * r.acc_taps * r.acc_cols]


# ------------------------------------------------------------------------------------------------ plan
class Plan:
    """One instance = the workspaces + recorded programs of one module call at one input shape and mode.
    A net definition fills it through the emit_* helpers; backward is emitted by replaying the tape in reverse."""

    def __init__(self, store: ParamStore, B: int, H: int, W: int, training: bool, groups: int = 1,
                 infer_only: bool = False):
        self.store, self.B, self.H, self.W, self.training = store, B, H, W, training
        # eval-mode call that autograd will never walk back through (no_grad / nothing requires grad): the stages skip
        # the pre-activation copies they would keep for backward (x4 inference: 0.7 GB of stores per 2048x2048 image)
        self.infer_only = infer_only and not training
        # BatchNorm statistics groups: 2 = the batch is (real | fake), each half normalised with its own batch
        # statistics as two separate calls of the module would (B200Module.forward_pair); only matters in training mode
        self.groups = groups if training else 1
        self.device = store.device
        self.grads = GradBuffers(store)
        self.bufs: Dict[str, torch.Tensor] = {}
        self.fwd = ops.Program()
        self.tape: List[Callable] = []
        self.bwd: Dict[tuple, ops.Program] = {}
        self.busy = False
        self.slots: Dict[str, Act] = {}     # side-channel gradients between tape entries (skip connections)
        self.has_bn = False
        # fp32 scratch that kernels accumulate into with red.global.add: zeroed by ONE memset at the head of the
        # forward (BatchNorm column sums) / backward (BN-backward column sums, PReLU slope gradients) program
        self._zarena = {"fwd": torch.zeros(ZERO_ARENA_FLOATS, dtype=F32, device=self.device),
                        "bwd": torch.zeros(ZERO_ARENA_FLOATS, dtype=F32, device=self.device)}
        self._zused = {"fwd": 0, "bwd": 0}
        self._znamed: Dict[str, torch.Tensor] = {}
        self.fwd.add(ops.elt(L.E_ZERO, p=[self._zarena["fwd"]], i=[ZERO_ARENA_FLOATS * 4]))
        self.post_backward: List[Callable] = []
        self.input_fn = self.output_fn = self.ingest_fn = self.grad_input_fn = None
        # optional two-tensor variants for forward_pair (default: concatenate / split around the single-tensor ones)
        self.input_pair_fn = self.ingest_pair_fn = None
        self.last_g: Dict[tuple, Optional[Act]] = {}

    # ---- buffers
    def buf(self, name: str, numel: int, dtype=BF16, zero: bool = False) -> torch.Tensor:
        if name in self.bufs:
            t = self.bufs[name]
            assert t.numel() >= numel and t.dtype == dtype, name
            return t
        t = (torch.zeros if zero else torch.empty)(max(int(numel), 8), dtype=dtype, device=self.device)
        self.bufs[name] = t
        return t

    def zbuf(self, which: str, name: str, numel: int) -> torch.Tensor:
        """`numel` floats inside the forward / backward zero arena (stable across program rebuilds)."""
        key = which + ":" + name
        if key not in self._znamed:
            n = _round_up(numel, 4)
            if self._zused[which] + n > ZERO_ARENA_FLOATS:
                raise RuntimeError("torchsr_b200: zero arena exhausted")
            self._znamed[key] = self._zarena[which][self._zused[which]:self._zused[which] + n]
            self._zused[which] += n
        return self._znamed[key]

    def act(self, name: str, B, H, W, C, dtype=BF16, zero=False) -> Act:
        return Act(self.buf(name, B * H * W * C, dtype, zero), B, H, W, C)

    # ---- forward emitters
    def conv(self, prog, x: Act, w: torch.Tensor, w_cols: int, n_slots: int, geom: dict, cout_pad: int, block_n: int,
             out: torch.Tensor, out_strides, n_valid: int, defer_tag=None, **kw):
        d = ops.conv_desc(x=ops.ptr(x.t, x.c0) if x.c0 else x.t, N=x.B, H=x.H, W=x.W, C=x.C, x_ld=x.ld, geom=geom, w=w,
                          cout_pad=cout_pad, w_ld=w_cols, n_slots=n_slots, block_n=block_n, out=out,
                          os_n=out_strides[0], os_h=out_strides[1], os_w=out_strides[2], n_valid=n_valid, **kw)
        if defer_tag is not None:
            prog.defer([d], defer_tag)
        else:
            prog.add(d)
        return d

    def conv_fwd(self, prog, rec: ConvRec, x: Act, out: Act, *, stats=None, act=L.ACT_NONE, prelu=None, preact=None,
                 res: Optional[Act] = None, res2: Optional[Act] = None, res_scale=1.0, res2_scale=1.0, acc_scale=1.0,
                 out_f32=False, shuffle_out=False, use_bias=True, group_rows=0, bnf: Optional[dict] = None,
                 dry_add=True, rep2x: Optional[Act] = None):
        """Forward of a 'std' conv (any stride) into `out` (OUT_LINEAR or PixelShuffle store), with the fused epilogue
        v = (acc + bias) * acc_scale + res * res_scale + res2 * res2_scale, optional activation, optional statistics.
        `x` / `out` / `res` may be channel slices of wider NHWC buffers (ld, c0)."""
        geom = ops.fwd_geometry(x.H, x.W, rec.k, rec.k, rec.pad, rec.pad, rec.stride)
        bias = None
        if rec.bias is not None and use_bias:
            bias = rec.bias      # PixelShuffle stores index it through the channel permutation inside the epilogue
        kw = dict(bias=bias, act=act, prelu=prelu, out_preact=preact, out_f32=out_f32, acc_scale=acc_scale,
                  out_ch_off=out.c0)
        if stats is not None:
            kw.update(stats_partial=stats, stats_ld=rec.cout_pad)
        if res is not None:
            assert res2 is None or (res2.ld == res.ld and res2.c0 == res.c0)
            kw.update(res=res.t, aux=res.strides(), aux_ch_off=res.c0, res_scale=res_scale,
                      res2=res2.t if res2 is not None else None, res2_scale=res2_scale)
        if shuffle_out:
            kw.update(out_mode=L.OUT_SHUFFLE, shuf_c=rec.cout // 4)
        if group_rows:
            kw.update(group_rows=group_rows)
        if bnf is not None:
            kw.update(bnf)
        if rep2x is not None:       # nearest x2 of the result written by the same epilogue (esrgan/generator.py:73,76)
            assert rep2x.H == 2 * geom["Ho"] and rep2x.W == 2 * geom["Wo"] and not shuffle_out
            kw.update(rep2x=dict(t=rep2x.t, strides=rep2x.strides(), ch_off=rep2x.c0))
        tiles = self.pick_tiles(x.B * geom["Ho"] * geom["Wo"], rec.cout_pad, rec.block_n,
                                len(geom["taps"]) * ((x.C) // ops.pick_block_k(x.C)))
        block_n = tiles.pop("block_n")
        # 3x3 / stride-1 convs on 64 input channels with a plain bf16 store and many M tiles per SM run the persistent
        # halo kernel with the staged (TMA-store) epilogue, which is built for 64-column N tiles (csrc/conv_params.h):
        # the activations are then re-read once per N tile, but from a patch that is fetched ~1.5x instead of 9x
        if (_staged_n64() and rec.k == 3 and rec.stride == 1 and rec.pad == 1 and x.C == 64 and block_n > 64
                and rec.cout_pad % 64 == 0 and "splits" not in tiles and stats is None and preact is None and res2 is None
                and not out_f32 and (bnf is None or bnf.get("bnf_mode") == 2)
                and (not shuffle_out or (rec.cout == 256 and rec.cout_pad == 256))):
            tiles_m = (x.B * geom["Ho"] * geom["Wo"] + 127) // 128
            if tiles_m >= 2 * max(1, NUM_SMS // (rec.cout_pad // 64)):
                block_n = 64
        if not dry_add:     # descriptor only (the caller decides whether this variant is launched)
            return ops.conv_desc(x=ops.ptr(x.t, x.c0) if x.c0 else x.t, N=x.B, H=x.H, W=x.W, C=x.C, x_ld=x.ld, geom=geom,
                                 w=rec.w_fwd, cout_pad=rec.cout_pad, w_ld=rec.cols, n_slots=rec.slots, block_n=block_n,
                                 out=out.t, os_n=out.strides()[0], os_h=out.strides()[1], os_w=out.strides()[2],
                                 n_valid=rec.cout_pad, **kw, **tiles)
        return self.conv(prog, x, rec.w_fwd, rec.cols, rec.slots, geom, rec.cout_pad, block_n, out.t, out.strides(),
                         rec.cout_pad, **kw, **tiles)

    def conv_bn_act(self, prog, name: str, rec: ConvRec, x: Act, y: Act, *, bn: nn.BatchNorm2d, act=L.ACT_NONE,
                    alpha=None, res: Optional[Act] = None):
        """y = act(BN(conv(x))) + res  (reference srgan/residual.py:86-91, generator.py:78, discriminator.py:33-61).
        Returns (raw, coef): the bf16 raw conv output and the published (scale, shift, mean, invstd) per statistics group,
        both needed by backward (None in eval mode).

        One launch when possible: eval mode always (running statistics folded into the conv epilogue); training mode
        when the conv's whole grid is co-resident, so that its CTAs can meet at a grid barrier between the column sums
        and the normalisation (csrc/conv_igemm.cu). Otherwise conv (with column sums) + bn_act_kernel."""
        assert rec.kind == "std" and rec.bias is None and rec.cout == rec.cout_pad
        self.has_bn = True
        C = rec.cout
        geom = ops.fwd_geometry(x.H, x.W, rec.k, rec.k, rec.pad, rec.pad, rec.stride)
        M = x.B * geom["Ho"] * geom["Wo"]
        G = self.groups
        group_rows = M // 2 if G == 2 else 0
        if G == 2 and (M % 2 or group_rows % 32):
            raise GroupSplitError("torchsr_b200: a two-group BatchNorm pass needs (rows per group) % 32 == 0 in every layer")
        eps, mom = bn.eps, (bn.momentum if bn.momentum is not None else 0.1)
        if not self.training:
            if FUSE_BN_FWD:
                self.conv_fwd(prog, rec, x, y, act=act, prelu=alpha, res=res, bnf=dict(
                    bnf_mode=2, bnf_c=C, bnf_gamma=bn.weight, bnf_beta=bn.bias, bnf_rm=bn.running_mean,
                    bnf_rv=bn.running_var, bnf_eps=eps, bnf_momentum=mom, bnf_count=M))
            else:
                raw = self.act(name + ".raw", y.B, y.H, y.W, C)
                self.conv_fwd(prog, rec, x, raw)
                self.bn_act(prog, name + ".bn", raw, y, bn=bn, act=act, alpha=alpha, res=res)
            return None, None
        raw = self.act(name + ".raw", y.B, y.H, y.W, C)
        stats = self.zbuf("fwd", name + ".s", G * rec.cout_pad * 2)
        coef = self.buf(name + ".bn.coef", G * 4 * C, F32)
        if FUSE_BN_FWD:
            ctr = self.zbuf("fwd", name + ".ctr", 16)
            kw = dict(stats=stats, act=act, prelu=alpha, res=res, preact=raw.t, group_rows=group_rows, bnf=dict(
                bnf_mode=1, bnf_c=C, bnf_counter=ctr, bnf_gamma=bn.weight, bnf_beta=bn.bias, bnf_rm=bn.running_mean,
                bnf_rv=bn.running_var, bnf_nbt=bn.num_batches_tracked, bnf_coef=coef, bnf_eps=eps, bnf_momentum=mom,
                bnf_count=M // G))
            d = self.conv_fwd(prog, rec, x, y, dry_add=False, **kw)
            if ops.conv_is_coresident(d):
                prog.add(d)
                return raw, coef
        self.conv_fwd(prog, rec, x, raw, stats=stats, group_rows=group_rows)
        self.bn_act(prog, name + ".bn", raw, y, bn=bn, stats=stats, act=act, alpha=alpha, res=res, coef=coef,
                    group_rows=group_rows, count=M // G)
        return raw, coef

    def pick_tiles(self, M: int, n_total: int, block_n: int, total_iters: int) -> dict:
        """Tile shape / split-K choice for a conv whose grid would leave most SMs idle. Deep layers (many K iterations,
        few output tiles) split K across CTAs - every split re-reads only its slice of the activations, the partial
        tiles meet in an fp32 workspace and the last CTA runs the epilogue (conv_igemm.cu). Shallow layers halve the N
        tile instead (pick_block_n). Returns the extra conv_desc keywords."""
        tiles_m = (M + 127) // 128
        ctas = tiles_m * (n_total // block_n)
        if SPLIT_K_MIN_ITERS > 0 and total_iters >= SPLIT_K_MIN_ITERS and ctas * 2 <= NUM_SMS + 20:
            splits = max(1, min(total_iters // 4, (2 * NUM_SMS) // ctas))
            if splits > 1:
                ws = self.buf(f"splitk.ws.{M}x{n_total}", _round_up(M, 128) * n_total, F32, zero=True)
                cnt = self.buf(f"splitk.cnt.{tiles_m}x{n_total // block_n}", ctas, torch.int32, zero=True)
                return dict(block_n=block_n, splits=splits, ws=ws, tile_counters=cnt, ws_ld=n_total)
        return dict(block_n=self.pick_block_n(M, n_total, block_n))

    @staticmethod
    def pick_block_n(M: int, n_total: int, block_n: int) -> int:
        """Halve the N tile while the grid cannot cover the SMs: these layers are bound by per-SM L2->smem bandwidth
        and per-CTA latency, so more, narrower CTAs win even though the A tile is then fetched once per N tile."""
        tiles_m = (M + 127) // 128
        while block_n >= 64 and block_n % 32 == 0 and tiles_m * (n_total // block_n) * 2 <= HALVE_BELOW_CTAS:
            block_n //= 2
        return block_n

    def stats_buf(self, name: str, M: int, cout_pad: int) -> torch.Tensor:
        return self.zbuf("fwd", name, cout_pad * 2)

    def bn_act(self, prog, name: str, x: Act, y: Act, *, bn: Optional[nn.BatchNorm2d] = None,
               stats: Optional[torch.Tensor] = None, act=L.ACT_NONE, alpha=None, res: Optional[Act] = None,
               leaky=0.2, res_scale=1.0, x_scale=1.0, coef: Optional[torch.Tensor] = None, group_rows=0, count=None):
        """y = act(BN(x)) + res in one pass. Training mode: the batch statistics come from the conv epilogue's column
        sums (`stats`); the same launch publishes (scale, shift, mean, invstd) for backward and updates running_mean /
        running_var / num_batches_tracked. Eval mode: running statistics. Returns the coefficient buffer (or None)."""
        C = x.C
        if bn is None:
            mode, p_bn = 0, [None] * 6
            eps = mom = 0.0
            coef = None
        else:
            if coef is None:
                coef = self.buf(name + ".coef", 4 * C, F32)
            mode = 1 if self.training else 2
            p_bn = [bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked, coef]
            eps, mom = bn.eps, (bn.momentum if bn.momentum is not None else 0.1)
        prog.add(ops.elt(L.E_BN_ACT, p=[x.t, stats if mode == 1 else None, y.t, res.t if res is not None else None,
                                        alpha] + p_bn,
                         i=[x.M, C, x.ld, y.ld, res.ld if res is not None else 0, act, x.c0, y.c0,
                            res.c0 if res is not None else 0, mode, count if count is not None else x.M,
                            group_rows if mode == 1 else 0],
                         f=[leaky, res_scale, x_scale, eps, mom]))
        return coef

    # ---- backward emitters
    def norm_act_bwd(self, prog, name: str, g: Act, x: Act, *, coef=None, bn: Optional[nn.BatchNorm2d] = None,
                     act=L.ACT_NONE, alpha: Optional[torch.Tensor] = None, g2: Optional[Act] = None,
                     bias_grad: Optional[torch.Tensor] = None, want_w=True, leaky=0.2, gscale=1.0) -> Act:
        """Backward of y = act(BN(x)) (bn given) or y = act(x) (bn None) for upstream gradient gscale * (g + g2):
        column reduction (atomics into the zero arena) -> apply, whose block 0 also publishes dgamma / dbeta (or the
        bias gradient) / dalpha into the flat gradient. g / x may be channel slices (ld, c0). Returns d/dx (dense)."""
        M, C = x.M, x.C
        assert g.C == C and (g2 is None or (g2.ld == g.ld and g2.C == C))
        has_bn = 1 if bn is not None else 0
        # statistics groups only matter for the BatchNorm terms: mean(dz), mean(dz * xhat) are per group
        G = self.groups if has_bn else 1
        group_rows = M // 2 if G == 2 else 0
        rpb = max(32, -(-M // (2 * NUM_SMS)))
        dx = self.act(name + ".dx", x.B, x.H, x.W, C)
        prelu = act == L.ACT_PRELU
        need_reduce = has_bn or (want_w and (prelu or bias_grad is not None))
        store = self.grads
        gp = ops.ptr(g.t, g.c0)
        xp = ops.ptr(x.t, x.c0)
        g2p = ops.ptr(g2.t, g2.c0) if g2 is not None else None
        sums = dacc = None
        fused = 0
        if need_reduce:
            sums = self.zbuf("bwd", name + ".sums", G * 2 * C)
            dacc = self.zbuf("bwd", name + ".dalpha", 1) if prelu else None
            # The data-gradient conv that produced g may still be held by the program (conv_dgrad defers it): fold
            # the column reductions into its epilogue instead of a separate pass over g and x. The conv then stores
            # dz = g * act'(z) in place of g - legal because nothing else reads g when an activation is involved
            # (with act == NONE the stored value is unchanged, e.g. the skip gradient of a residual block).
            held = prog.take_deferred(g) if (FUSE_BN_REDUCE and g2 is None and gscale == 1.0 and leaky == 0.2) else None
            if held is not None and self._fuse_bn_reduce(held, x, coef if has_bn else None, act,
                                                         alpha if prelu else None, sums, dacc, G):
                fused = 1
                if has_bn and FUSE_BN_BWD and self._fuse_bn_apply(held, name, g, x, dx, bn, act, alpha, want_w, G):
                    if len(held) > 1:
                        prog.add_group(held)
                    else:
                        prog.add(held[0])
                    return dx
            if held is not None:
                if len(held) > 1:
                    prog.add_group(held)
                else:
                    prog.add(held[0])
            if not fused:
                for gi in range(G):       # one launch per statistics group (row range, coefficient and sum blocks)
                    r0, rows = gi * group_rows, (group_rows if G == 2 else M)
                    prog.add(ops.elt(L.E_BN_BWD_REDUCE,
                                     p=[gp + r0 * g.ld * 2, xp + r0 * x.ld * 2,
                                        ops.ptr(coef, gi * 4 * C) if has_bn else None, alpha if prelu else None,
                                        ops.ptr(sums, gi * 2 * C), dacc, (g2p + r0 * g.ld * 2) if g2p else None],
                                     i=[rows, C, act, rpb, g.ld, x.ld, has_bn], f=[leaky, gscale]))
        dgamma = store.grad_slice(bn.weight) if has_bn and want_w else None
        dbeta = (store.grad_slice(bn.bias) if has_bn else bias_grad) if want_w else None
        dalpha = store.grad_slice(alpha) if (prelu and want_w) else None
        prog.add(ops.elt(L.E_BN_BWD_APPLY, p=[gp, xp, coef, sums, alpha if prelu else None, dx.t, g2p,
                                              bn.weight if has_bn else None, dgamma, dbeta, dalpha, dacc],
                         i=[M, C, act, g.ld, x.ld, C, has_bn, fused, fused, group_rows], f=[leaky, gscale]))
        return dx

    def _fuse_bn_apply(self, descs, name: str, g: Act, x: Act, dx: Act, bn, act, alpha, want_w: bool, groups: int) -> bool:
        """On top of _fuse_bn_reduce: the held data-gradient conv(s) also form dx = A*dz + Bx*x + Cc after a grid barrier
        and publish dgamma / dbeta / dalpha - no bn_bwd_apply launch, no HBM round trip of dz. Needs every CTA of the
        launch co-resident (one launch = one conv, or the four parity classes of a stride-2 layer) and a dense target
        with the geometry of dx. Returns False with the descriptors left as _fuse_bn_reduce made them."""
        C = x.C
        if g.ld != C or g.c0 != 0 or g.C != C or dx.ld != C:
            return False
        store = self.grads
        ctr = self.zbuf("bwd", name + ".ctr", 16)
        M = x.M
        for d in descs:
            d.bnr_apply = 1
            d.bnr_dx = ops.ptr(dx.t)
            if getattr(d, "_parity", None) is not None:
                rh, rw = d._parity          # same scatter as `out`: (rh, rw) offset into the fine grid
                d.bnr_dx = ops.ptr(dx.t) + ((rh * x.W + rw) * dx.ld) * 2
            d.bnr_gamma = ops.ptr(bn.weight)
            d.bnr_dgamma = ops.ptr(store.grad_slice(bn.weight)) if want_w else 0
            d.bnr_dbeta = ops.ptr(store.grad_slice(bn.bias)) if want_w else 0
            d.bnr_dalpha = ops.ptr(store.grad_slice(alpha)) if (want_w and act == L.ACT_PRELU) else 0
            d.bnr_count = M // groups
            d.bnf_counter = ops.ptr(ctr)
        ok = True
        total = 0
        for d in descs:
            tiles = ((d.N * d.Ho * d.Wo + 127) // 128) * (d.cout_pad // d.block_n)
            total += tiles
        if len({(d.N * d.Ho * d.Wo, d.cout_pad // d.block_n) for d in descs}) != 1:
            ok = False
        elif ops.DRY:
            ok = total <= 2 * 148
        else:
            ok = ops.conv_coresident_capacity(descs[0]) >= total
        if not ok:
            for d in descs:
                d.bnr_apply = 0
                d.bnr_dx = d.bnr_gamma = d.bnr_dgamma = d.bnr_dbeta = d.bnr_dalpha = d.bnf_counter = 0
                d.bnr_count = 0
        return ok

    @staticmethod
    def _fuse_bn_reduce(descs, x: Act, coef, act, alpha, sums, dacc, groups: int = 1) -> bool:
        """Patches the held data-gradient conv descriptor(s) (one, or the four output-parity classes of a stride-2
        layer) so that their epilogues accumulate sum(dz), sum(dz*x) into `sums`. Returns False (descriptors left
        untouched) when the epilogue's aux addressing is already bound to a tensor of a different geometry, or when a
        two-group pass does not split the descriptor's rows on a 32-row boundary."""
        xs = x.strides()
        for d in descs:
            if groups == 2 and ((d.N * d.Ho * d.Wo) % 64 or d.N % 2):
                return False
        for d in descs:
            cls = getattr(d, "_parity", None)
            want = xs if cls is None else (xs[0], 2 * xs[1], 2 * xs[2])
            if d.n_valid != x.C or d.stats_partial or d.dalpha_partial or d.bwd_z or d.out_f32 or \
                    d.out_mode != L.OUT_LINEAR:
                return False
            if d.res and ((d.aux_n, d.aux_h, d.aux_w) != want or d.aux_ch_off != x.c0 or cls is not None):
                return False
        for d in descs:
            cls = getattr(d, "_parity", None)
            base = ops.ptr(x.t)
            if cls is None:
                d.aux_n, d.aux_h, d.aux_w = xs
                d.aux_ch_off = x.c0
            else:
                rh, rw = cls
                d.aux_n, d.aux_h, d.aux_w = xs[0], 2 * xs[1], 2 * xs[2]
                d.aux_ch_off = 0
                base += ((rh * x.W + rw) * x.ld + x.c0) * 2
            d.bnr_x = base
            d.bnr_coef = ops.ptr(coef)
            d.bnr_prelu = ops.ptr(alpha)
            d.bnr_act = act
            d.bnr_c = x.C
            d.stats_partial = ops.ptr(sums)
            d.stats_ld = x.C
            d.dalpha_partial = ops.ptr(dacc)
            d.group_rows = (d.N * d.Ho * d.Wo) // 2 if groups == 2 else 0
        return True

    def conv_dgrad(self, prog, name: str, rec: ConvRec, dy: Act, x_like: Act, *, res: Optional[Act] = None,
                   out_f32: bool = False, out: Optional[Act] = None, res2: Optional[Act] = None, res_scale=1.0,
                   res2_scale=1.0, res_cols=0, acc_scale=1.0) -> Act:
        """Gradient w.r.t. the input `x_like` of conv `rec` from dY, as an implicit-GEMM conv over dY with the
        transposed weight pack. Fused epilogue: + res (residual branch gradient) and, when the producer of x_like
        left a hook, * act'(pre-activation) with the PixelShuffle inverse folded into the store.
        Returns the Act holding the result (the hook's un-shuffled target when one was used)."""
        assert rec.need_dgrad
        kw = {}
        if acc_scale != 1.0:
            kw.update(acc_scale=acc_scale)
        if res is not None:
            assert res2 is None or (res2.ld == res.ld and res2.c0 == res.c0)
            kw.update(res=res.t, aux=res.strides(), aux_ch_off=res.c0, res_scale=res_scale, res_cols=res_cols,
                      res2=res2.t if res2 is not None else None, res2_scale=res2_scale)
        hook = x_like.hook
        if hook is not None:
            assert res is None
            z = hook["bwd_z"]
            kw.update(bwd_z=z.t, bwd_act=hook["bwd_act"], prelu=hook.get("prelu"),
                      dalpha_partial=hook.get("dalpha_partial"), aux=z.strides())
        if rec.kind == "std":
            n_out, n_slots = rec.t_rows, rec.k * rec.k
            geom = ops.dgrad_s1_geometry(dy.H, dy.W, rec.k, rec.k, rec.pad, rec.pad)
        elif rec.kind == "rown":
            n_out, n_slots = rec.t_rows, rec.k
            geom = ops.dgrad_s1_geometry(dy.H, dy.W, rec.k, 1, rec.pad, 0)
        else:  # fullk: plain GEMM back to the im2row columns
            n_out, n_slots = rec.t_rows, 1
            geom = ops.fwd_geometry(dy.H, dy.W, 1, 1, 0, 0, 1)
        block_n = next(b for b in (128, 96, 64, 160, 192, 32, 16) if n_out % b == 0 and b <= n_out)
        if rec.stride == 1 and rec.kind != "fullk":
            tiles = self.pick_tiles(dy.M, n_out, block_n, len(geom["taps"]) * (dy.C // ops.pick_block_k(dy.C)))
            block_n = tiles.pop("block_n")
            kw.update(tiles)
        elif rec.stride == 1:
            block_n = self.pick_block_n(dy.M, n_out, block_n)
        dx = out
        if dx is None and (hook is None or hook.get("unshuffle_to") is None):
            dx = self.act(name + ".dgrad", x_like.B, x_like.H, x_like.W, n_out, F32 if out_f32 else BF16)
        if hook is not None and hook.get("unshuffle_to") is not None:
            target = hook["unshuffle_to"]
            kw.update(out_mode=L.OUT_UNSHUFFLE, shuf_c=n_out)
        else:
            target = dx
        # held back until the next emission: a following norm_act_bwd may fold its reductions into this epilogue
        deferrable = hook is None and not out_f32
        if rec.stride == 1:
            self.conv(prog, dy, rec.w_t, rec.t_cols, n_slots, geom, n_out, block_n, target.t, target.strides(), n_out,
                      out_ch_off=target.c0, out_f32=out_f32, defer_tag=target if deferrable else None, **kw)
        else:
            assert rec.stride == 2 and rec.k == 3 and rec.pad == 1 and not kw and rec.kind == "std"
            descs = ops.dgrad_s2_descs(dy=dy.t, N=dy.B, Hy=dy.H, Wy=dy.W, Cout=dy.C, dy_ld=dy.ld, wt=rec.w_t,
                                       Cin=n_out, cin_pad=n_out, block_n=block_n, out=target.t, Hx=x_like.H,
                                       Wx=x_like.W, out_ld=target.ld, n_valid=n_out)
            if deferrable and target.c0 == 0:
                prog.defer(descs, target, group=True)     # one launch for the four parity classes
            else:
                prog.add_group(descs)
        return target

    def colsum(self, prog, name: str, g: Act, out_vec: torch.Tensor, shuffle_c4: int = 0):
        """out_vec[c] = sum over rows of g[:, c] (bias gradients); shuffle_c4 > 0: g's columns are in PixelShuffle-packed
        order and out_vec is the parameter-order gradient (channel 4*(r % c4) + r / c4 for column r)."""
        M, C = g.M, g.C
        rpb = max(32, -(-M // (2 * NUM_SMS)))
        sums = self.zbuf("bwd", name + ".cs", 2 * C)
        # parameter gradients only: both launches ride the weight-gradient side branch (consecutive side ops share one
        # branch, in order), off the critical path of backward
        prog.add(ops.elt(L.E_BN_BWD_REDUCE, p=[g.t, g.t, None, None, sums, None, None],
                         i=[M, C, L.ACT_NONE, rpb, g.ld, g.ld, 0], f=[0.2], side=True))
        prog.add(ops.elt(L.E_COLSUM_FINALIZE, p=[sums, out_vec], i=[1, C, C, 0, 0, shuffle_c4], side=True))

    def colsum_strided(self, prog, name: str, g: Act, out_vec: torch.Tensor):
        """colsum for a channel slice of a wider buffer."""
        M, C = g.M, g.C
        rpb = max(32, -(-M // (2 * NUM_SMS)))
        sums = self.zbuf("bwd", name + ".cs", 2 * C)
        gp = ops.ptr(g.t, g.c0)
        prog.add(ops.elt(L.E_BN_BWD_REDUCE, p=[gp, gp, None, None, sums, None, None],
                         i=[M, C, L.ACT_NONE, rpb, g.ld, g.ld, 0], f=[0.2], side=True))
        prog.add(ops.elt(L.E_COLSUM_FINALIZE, p=[sums, out_vec], i=[1, C, C, 0, 0], side=True))

    def conv_wgrad(self, prog, rec: ConvRec, x: Act, dy: Act, geom: Optional[dict] = None):
        geom = geom or ops.fwd_geometry(x.H, x.W, rec.k, rec.k, rec.pad, rec.pad, rec.stride)
        block_n = dy.C if dy.C <= 128 else 128
        assert dy.C == rec.acc_rows, (rec.name, dy.C, rec.acc_rows)
        d = ops.wgrad_desc(x=x.t, N=x.B, H=x.H, W=x.W, C=x.C + x.c0, x_ld=x.ld, geom=geom, dy=dy.t, dy_ld=dy.ld,
                           dy_c=dy.C, out=self.grads.acc(rec), cout_valid=rec.acc_rows, block_n=block_n, x_c0=x.c0, dy_c0=dy.c0)
        prog.add(d)

    def early_unpack(self, prog, convs: List[ConvRec], first_param: torch.Tensor, end_param: torch.Tensor):
        """Data-parallel programs only: the weight gradients of `convs` (the deep layers, produced first in backward)
        are unpacked NOW and the flat-gradient slice [offset(first_param), offset(end_param)) - those conv weights and
        the BatchNorm parameters between them - is marked final ("early_grads"), so that its all-reduce overlaps the
        rest of backward (run_backward). The unpack launch joins the weight-gradient side branch."""
        if not getattr(self, "_building_dist", False):
            return
        prog.add(self.grads.unpack_for(convs))
        prog.mark("early_grads")
        self.early_slice = (self.store.offsets[id(first_param)], self.store.offsets[id(end_param)])
        self._early_unpacked = {r.name for r in convs}

    # ---- execution (set by the net definition: input_fn, output_fn, ingest_fn, grad_input_fn, post_backward)
    def _run_fwd_program(self):
        ev, mark = self.store.late_event, self.fwd.marks.get("late_weights")
        if ev is not None and mark is not None:
            self.fwd.run(0, mark)
            torch.cuda.current_stream(self.device).wait_event(ev)
            self.fwd.run(mark, -1)
        else:
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)
            self.fwd.run()

    def run_forward(self, x: torch.Tensor) -> torch.Tensor:
        self.input_fn(x)
        self._run_fwd_program()
        return self.output_fn()

    def run_forward_pair(self, a: torch.Tensor, b: torch.Tensor):
        if self.input_pair_fn is not None:
            self.input_pair_fn(a, b)
        else:
            self.input_fn(torch.cat([a, b]))
        self._run_fwd_program()
        out = self.output_fn()
        h = out.shape[0] // 2
        return out[:h].clone(), out[h:].clone()

    def run_backward(self, gout, want_x: bool, want_w: bool, ddp=None, alias: bool = False, defer_comm: bool = False):
        """Runs the backward launch list and, for a data-parallel module, the gradient exchange.

        defer_comm (single GPU only): the plan's own flat buffer is returned and the caller sums the gradients of
        several outstanding calls of the module first (see _PlanFn._run_backward, 'merge_pending_grads').

        Data parallel (reference */trainer.py:143-157: DDP all-reduces every gradient): the flat fp32 gradient is averaged
        in place with bucketed NCCL all-reduces - except the weight of a discriminator's first Linear layer (75 MB for
        SRGAN, 80 % of the module): it is an outer product over the batch, so the ranks all-gather its two bf16 factors
        (dpre1 [b][n], features^T in 64-row chunks; ~2.5 MB per rank) as soon as the head of backward has produced them,
        the transfer overlaps the convolutional backward, and every rank forms  mean_r(A_r^T X_r)  itself with one GEMM
        over the gathered K chunks."""
        if not self.training and self.has_bn:
            raise NotImplementedError("torchsr_b200: backward through eval-mode BatchNorm is not implemented; call "
                                      ".train() for gradient computation (the reference trainers do)")
        seed = self.ingest_pair_fn(*gout) if isinstance(gout, tuple) else self.ingest_fn(gout)
        dist_on = want_w and ddp is not None and ddp.world > 1
        inline = getattr(self, "inline_adam", None) if (want_w and not dist_on) else None
        prog = self.backward_program(want_x, want_w, seed, dist=dist_on, inline=inline)
        flat = None
        gb = self.grads
        if dist_on:
            assert not defer_comm
            mark = prog.marks.get("factors")
            early = prog.marks.get("early_grads")
            from .dist import bucket_slices
            pos = 0
            if mark is not None:
                prog.run(0, mark)
                pos = mark
                fa = self._factor_buffers(ddp.world)
                ddp.allgather_async(fa["a_all"], self.factors["a"])
                ddp.allgather_async(fa["xt_all"], self.factors["xt"])
            sent = []         # [lo, hi) slices of the flat gradient already handed to NCCL
            cap = getattr(self, "capture_local", None)
            early_snap = None
            if early is not None:
                prog.run(pos, early - pos)
                pos = early
                lo, hi = self.early_slice
                if cap is not None:
                    early_snap = (lo, hi, gb.flat[lo:hi].clone())     # before the in-place reduction starts
                for b0, b1 in bucket_slices(hi - lo, None):
                    ddp.allreduce_async(gb.flat[lo + b0:lo + b1])
                sent.append((lo, hi))
            prog.run(pos, -1)
            for fn in self.post_backward:
                fn()
            lin_lo = lin_hi = None
            if mark is not None:
                w = self.factors["lin"].weight
                lin_lo = self.store.offsets[id(w)]
                lin_hi = lin_lo + _round_up(w.numel(), 4)
                sent.append((lin_lo, lin_hi))      # formed locally from the gathered factors, never all-reduced
            if cap is not None:
                # checker hook (tools/dp_check.py, bench.py dp_parity): this rank's own gradient of THIS backward
                # execution, before the exchange (the Linear weight slice from the local factors; the early slice was
                # copied before its in-place reduction started)
                snap = gb.flat.clone()
                if early_snap is not None:
                    snap[early_snap[0]:early_snap[1]].copy_(early_snap[2])
                if mark is not None:
                    from .nets import linear_wgrad_gemm
                    f = self.factors
                    d = linear_wgrad_gemm(self, f["a"], f["xt"], f["rows"], 1.0, out=snap[lin_lo:lin_lo + w.numel()])
                    d.side = 0
                    ops.run_now(d)
                cap.append(snap)
            cur = 0
            for lo, hi in sorted(sent) + [(self.store.total, self.store.total)]:
                if lo > cur:
                    for b0, b1 in bucket_slices(lo - cur, None):
                        ddp.allreduce_async(gb.flat[cur + b0:cur + b1])
                cur = max(cur, hi)
            if mark is not None:
                ddp.wait_gathers()      # the current stream waits for the gathered factors ...
                ops.run_now(fa["gemm"])  # ... and forms the averaged product while the last all-reduces are in flight
            ddp.wait()
            flat = gb.flat if alias else gb.flat.clone()
        else:
            prog.run()
            if inline is not None and prog.marks.get("inline_adam") is not None:
                inline.mark_late_done()
            if want_w:
                for fn in self.post_backward:
                    fn()
                # alias mode (set by the trainers, which zero the gradients before every backward): hand out views
                # of this plan's flat buffer instead of a copy; valid until this plan instance runs backward again
                flat = gb.flat if (alias or defer_comm) else gb.flat.clone()
        gx = self.grad_input_fn() if want_x else None
        return gx, flat

    def _factor_buffers(self, world: int) -> dict:
        """Gathered-factor buffers and the GEMM over them (built once per plan and world size)."""
        fa = getattr(self, "_factor_cache", None)
        if fa is None or fa["world"] != world:
            from .nets import linear_wgrad_gemm
            f = self.factors
            a_all = self.buf(f"factors.a_all.{world}", world * f["a"].numel(), BF16)
            xt_all = self.buf(f"factors.xt_all.{world}", world * f["xt"].numel(), BF16)
            fa = dict(world=world, a_all=a_all, xt_all=xt_all,
                      gemm=linear_wgrad_gemm(self, a_all, xt_all, world * f["rows"], 1.0 / world))
            fa["gemm"].side = 0
            self._factor_cache = fa
        return fa

    # ---- programs
    def backward_program(self, want_x: bool, want_w: bool, seed: Act, dist: bool = False, inline=None) -> "ops.Program":
        key = (want_x, want_w, dist, id(inline) if inline is not None else 0)
        if key not in self.bwd:
            self._building_dist = dist      # read by the net definition's tape entries (nets.bwd_head)
            self._building_inline = inline  # optimizer whose late parameters are updated inside this program
            prog = ops.Program()
            prog.add(ops.elt(L.E_ZERO, p=[self._zarena["bwd"]], i=[ZERO_ARENA_FLOATS * 4]))
            if want_w:
                self.grads._ensure()
                prog.add(ops.elt(L.E_ZERO, p=[self.grads.acc_arena], i=[self.grads.acc_arena.numel() * 4]))
            g = seed
            self.slots.clear()
            for fn in reversed(self.tape):
                g = fn(prog, g, want_x, want_w)
            self.last_g[key] = g
            if want_w:
                done = getattr(self, "_early_unpacked", None)
                if done:
                    prog.add(self.grads.unpack_for([r for r in self.store.convs if r.name not in done]))
                else:
                    prog.add(self.grads.unpack_desc)
            self._early_unpacked = None
            prog.flush()
            self.bwd[key] = prog
        self.cur_g = self.last_g[key]
        return self.bwd[key]


class _Lease:
    """Returns a plan to its pool when the autograd node that holds it dies (backward done or graph dropped).
    `pending` (optional) is the module's set of outstanding calls whose backward will produce weight gradients."""

    def __init__(self, plan: Plan, pending: Optional[set] = None, state: Optional[dict] = None):
        self.plan = plan
        self.pending = pending
        self.state = state
        if pending is not None:
            pending.add(id(plan))

    def release(self):
        if self.plan is not None:
            if self.pending is not None:
                self.pending.discard(id(self.plan))
                # no outstanding call left: gradients parked by calls whose siblings never reached backward are stale
                if not self.pending and self.state is not None:
                    self.state.pop("parked", None)
            self.plan.busy = False
            self.plan = None

    def __del__(self):
        self.release()


# ------------------------------------------------------------------------------------------------ module base
class B200Module(nn.Module):
    """Base of the drop-in modules: owns the ParamStore and the plan pool, binds plans to autograd.

    Subclasses create the same nn.Conv2d / nn.BatchNorm2d / nn.PReLU / nn.Linear children, in the same order and
    under the same attribute names, as the reference classes (so default init consumes the RNG identically and
    state_dict() keys/shapes match), and implement `_records()` and `_define(plan)`."""

    def __init__(self):
        super().__init__()
        object.__setattr__(self, "_tsr", dict(store=None, plans={}))

    # -- to be provided by subclasses
    def _records(self):
        raise NotImplementedError

    def _define(self, plan: Plan, x_shape):
        """Emit the forward program and the backward tape. Must set plan.run_input / plan.run_output callables and
        plan.bwd_seed / plan.run_grad_input / plan.run_grad_output."""
        raise NotImplementedError

    # -- invalidation: .to()/.cuda()/.float() move or recreate parameter storage
    def _apply(self, fn, *a, **kw):
        r = super()._apply(fn, *a, **kw)
        st = self.__dict__.get("_tsr")
        if st is not None:
            st["store"] = None
            st["plans"] = {}
        return r

    def _store(self) -> ParamStore:
        st = self._tsr
        p0 = next(self.parameters())
        if st["store"] is None or st["store"].params[0].data_ptr() != p0.data_ptr() or st["store"].device != p0.device:
            if p0.device.type != "cuda" and not ops.DRY:
                raise L.TorchSRB200Error(
                    f"{type(self).__name__} runs only on a CUDA sm_100 device (parameters are on '{p0.device}'); "
                    "there is no CPU fallback for this path")
            convs, linears = self._records()
            st["store"] = ParamStore(self, convs, linears, p0.device)
            st["plans"] = {}
        return st["store"]

    def _acquire(self, shape, training: bool, groups: int = 1, infer_only: bool = False) -> Plan:
        store = self._store()
        infer_only = bool(infer_only) and not training and groups == 1
        key = (tuple(shape), bool(training)) if groups == 1 else (tuple(shape), bool(training), groups)
        if infer_only:
            key = key + ("infer",)
        pool = self._tsr["plans"].setdefault(key, [])
        for pl in pool:
            if not pl.busy:
                pl.busy = True
                return pl
        pl = Plan(store, shape[0], shape[2], shape[3], training, groups, infer_only)
        try:
            self._define(pl, shape)
        except GroupSplitError:
            if not pool:
                del self._tsr["plans"][key]
            raise
        pl.busy = True
        pool.append(pl)
        return pl

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 4:
            raise RuntimeError(f"expected a 4-D NCHW tensor, got shape {tuple(x.shape)}")
        store = self._store()
        if x.device != store.device:
            raise RuntimeError(f"input is on {x.device} but the module is on {store.device}")
        x = x.contiguous()
        if x.dtype != F32:
            x = x.float()
        needs_graph = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in store.params))
        return _PlanFn.apply(self, needs_graph, x, *store.params)


    def forward_pair(self, a: torch.Tensor, b: torch.Tensor, grad_halves=(True, True)):
        """(self(a), self(b)) - two calls of the module on two batches of one shape - executed as ONE pass of its kernels
        over the concatenated batch, with the reference's two-call semantics kept exactly: in training mode every
        BatchNorm normalises each half with that half's batch statistics, the running statistics take two momentum
        updates (a's first, then b's) and num_batches_tracked += 2 (reference srgan/trainer.py:446-447,
        esrgan/trainer.py:447-449,463-464: D(high_res) then D(super_res)). grad_halves[i] False returns that half
        detached (its input then also gets no gradient). Falls back to two calls in eval mode and for shapes whose
        per-half row counts do not split on 32-row boundaries."""
        if a.shape != b.shape or a.dim() != 4:
            raise RuntimeError(f"forward_pair needs two 4-D NCHW tensors of one shape, got {tuple(a.shape)} and {tuple(b.shape)}")
        store = self._store()
        unsupported = self._tsr.setdefault("pair_unsupported", set())
        shape = (2 * a.shape[0],) + tuple(a.shape[1:])
        key = (shape, self.training)
        if self.training and key not in unsupported:
            a2, b2 = a.contiguous().float(), b.contiguous().float()
            if not grad_halves[0]:
                a2 = a2.detach()
            if not grad_halves[1]:
                b2 = b2.detach()
            needs_graph = torch.is_grad_enabled() and (a2.requires_grad or b2.requires_grad or
                                                       any(p.requires_grad for p in store.params))
            try:
                pa, pb = _PairFn.apply(self, needs_graph, shape, a2, b2, *store.params)
            except GroupSplitError:
                unsupported.add(key)
            else:
                return (pa if grad_halves[0] else pa.detach()), (pb if grad_halves[1] else pb.detach())
        outs = []
        for t, g in ((a, grad_halves[0]), (b, grad_halves[1])):
            if g:
                outs.append(self(t))
            else:
                with torch.no_grad():
                    outs.append(self(t))
        return outs[0], outs[1]


class _PlanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module: B200Module, needs_graph: bool, x: torch.Tensor, *params):
        plan = module._acquire(x.shape, module.training, infer_only=not needs_graph)
        store = plan.store
        store.ensure_packed()
        if module.training and module._tsr.get("ddp") is not None:
            from .dist import sync_buffers
            sync_buffers(module)
        out = plan.run_forward(x)
        if module.training and module._tsr.get("ddp") is not None and plan.has_bn:
            from .dist import publish_buffers
            publish_buffers(module)
        if needs_graph:
            wants_w = any(p.requires_grad for p in store.params)
            ctx.lease = _Lease(plan, module._tsr.setdefault("pending", set()) if wants_w else None, module._tsr)
            ctx.module = module
        else:
            plan.busy = False
        return out

    @staticmethod
    def backward(ctx, gout: torch.Tensor):
        gx, grads = _PlanFn._run_backward(ctx, gout.contiguous().float(), ctx.needs_input_grad[2],
                                          list(ctx.needs_input_grad[3:]))
        return (None, None, gx, *grads)

    @staticmethod
    def _run_backward(ctx, gout, want_x: bool, want: List[bool]):
        """Shared by the one- and two-batch autograd bindings; `gout` is a tensor or a pair (either may be None)."""
        lease = ctx.lease
        plan = lease.plan
        if plan is None:
            raise RuntimeError("torchsr_b200: backward through the same module call twice is not supported "
                               "(activations live in a pooled workspace)")
        want_w = any(want)
        st = ctx.module._tsr
        ddp = st.get("ddp")
        dist_on = want_w and ddp is not None and ddp.world > 1
        # Aliased gradients (views of the plan's flat buffer) are only handed out when autograd will ASSIGN them: with a
        # .grad already present (gradient accumulation, zero_grad(set_to_none=False), a second call of the module
        # feeding the same loss) AccumulateGrad would add the buffer to a tensor that may alias it: private copy then.
        safe_alias = all(p.grad is None for p in plan.store.params)
        merge = want_w and st.get("merge_pending_grads", False) and not dist_on
        plan.inline_adam = st.get("inline_adam") if (st.get("alias_grads", False) and safe_alias) else None
        if not merge:
            # data parallel: every backward exchanges its own gradient (averaging is linear, autograd sums the calls)
            others = st.get("pending", set()) - {id(plan)} if want_w else set()
            gx, flat = plan.run_backward(gout, want_x, want_w, ddp,
                                         st.get("alias_grads", False) and safe_alias and not others)
        else:
            # Several calls of the module feed one loss (D(real) and D(fake) as two calls): every backward but the last
            # parks its flat gradient; the last one adds the parked ones to its own (one kernel per parked call) and
            # returns the sum - the others return no parameter gradients, which autograd treats as zero.
            cur = torch.cuda.current_stream(plan.device)
            others = st["pending"] - {id(plan)}
            parked = st.setdefault("parked", [])
            gx, flat = plan.run_backward(gout, want_x, want_w, None, True, defer_comm=True)
            if others:
                ev = torch.cuda.Event()
                ev.record(cur)
                parked.append(dict(flat=flat, ev_done=ev))
                flat = None
            else:
                for p in st.pop("parked", []):
                    cur.wait_event(p["ev_done"])
                    _add_flat(flat, p["flat"])
                if not safe_alias:
                    flat = flat.clone()
        grads = plan.store.grads_from_flat(flat, want) if (want_w and flat is not None) else [None] * len(want)
        lease.release()
        return gx, grads


def _add_flat(dst: torch.Tensor, src: torch.Tensor):
    """dst += src on flat fp32 gradient slices (axpby_f32_kernel; offsets are multiples of 4 elements)."""
    ops.run_now(ops.elt(L.E_AXPBY_F32, p=[dst, src, dst, None], i=[dst.numel()], f=[1.0, 1.0]))


class _PairFn(torch.autograd.Function):
    """Autograd binding of B200Module.forward_pair: one plan over the concatenated (a | b) batch, two outputs."""

    @staticmethod
    def forward(ctx, module: B200Module, needs_graph: bool, shape, a: torch.Tensor, b: torch.Tensor, *params):
        plan = module._acquire(shape, module.training, groups=2)
        store = plan.store
        store.ensure_packed()
        if module.training and module._tsr.get("ddp") is not None:
            from .dist import sync_buffers
            sync_buffers(module)
        pa, pb = plan.run_forward_pair(a, b)
        ctx.set_materialize_grads(False)
        if needs_graph:
            wants_w = any(p.requires_grad for p in store.params)
            ctx.lease = _Lease(plan, module._tsr.setdefault("pending", set()) if wants_w else None, module._tsr)
            ctx.module = module
        else:
            plan.busy = False
        return pa, pb

    @staticmethod
    def backward(ctx, ga, gb):
        fix = lambda g: g.contiguous().float() if g is not None else None  # noqa: E731
        want_x = ctx.needs_input_grad[3] or ctx.needs_input_grad[4]
        gx, grads = _PlanFn._run_backward(ctx, (fix(ga), fix(gb)), want_x, list(ctx.needs_input_grad[5:]))
        gxa = gxb = None
        if gx is not None:
            h = gx.shape[0] // 2
            gxa = gx[:h] if ctx.needs_input_grad[3] else None
            gxb = gx[h:] if ctx.needs_input_grad[4] else None
        return (None, None, None, gxa, gxb, *grads)
