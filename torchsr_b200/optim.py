"""Adam on this repo's kernels: one launch per parameter group updates every parameter, its
exp_avg / exp_avg_sq state AND the bf16 operand copies the conv / GEMM kernels read (csrc/eltwise.cu
adam_pack_kernel), replacing torch's multi-tensor Adam launches plus the separate re-pack pass over the weights.

Same update rule, defaults and param_group keys as torch.optim.Adam (reference torchsr/srgan/trainer.py:171-185:
lr 1e-4, betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad), so lr schedulers (StepLR) work unchanged; a tensor
learning rate and the device-side step counter make a step capturable in a CUDA graph. There is no CPU path."""
import ctypes as C
from typing import Dict, List

import torch

from . import _lib as L
from . import engine, ops


ADAM_TILES = 4     # kAdamTiles in csrc/eltwise.cu


class FusedAdam(torch.optim.Optimizer):
    """late_numel > 0: Linear weights of at least that many elements (the discriminator's 18.9 M-element classifier
    weight: 80 % of its parameters) are updated by a SECOND launch on a side stream. The layers that read them sit at
    the end of the network, so the next forward of the module starts on the freshly updated conv weights while the big
    update is still streaming through HBM; the module's forward waits for it right before the first launch that reads
    those weights (engine.Plan.run_forward, mark "late_weights"). The caller must call join() before anything else
    touches those parameters (the trainers do, at the end of every step)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, late_numel: int = 0):
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or eps < 0.0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._tables: Dict[int, Dict[tuple, dict]] = {}     # group index -> {address-set key -> table}
        self.late_numel = late_numel
        self._late_stream = None
        self._late_pending: List = []                       # (event, stores) of late launches not yet joined
        self._inline_tables: Dict[int, dict] = {}           # id(plan) -> in-backward table of the late parameters
        self._late_done_in_backward = False

    # ---- per-group device state
    def _group_state(self, gi: int, group) -> dict:
        """Device table (parameter / gradient / state / pack pointers) of one param group for the CURRENT set of
        gradient addresses. One table per distinct address set is kept for the optimizer's lifetime: a captured CUDA
        graph holds the table's device pointer in its Adam kernel node, so a table must never be freed or rewritten
        when another batch shape (a ragged last batch run eagerly) brings other gradient buffers."""
        variants = self._tables.setdefault(gi, {})
        params = [p for p in group["params"] if p.grad is not None]
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in params) + \
            tuple(id(s) for s in self._stores_of(params))
        st = variants.get(key)
        if st is not None:
            return st
        dev = params[0].device
        if dev.type != "cuda" and not ops.DRY:
            raise L.TorchSRB200Error("torchsr_b200.optim.FusedAdam runs only on CUDA parameters (no CPU fallback)")
        old = next(iter(variants.values()), None)
        st = dict(key=key, params=params)
        # optimizer state lives in two flat arenas; self.state[p] exposes per-parameter views (state_dict compatible)
        if old is not None and [id(p) for p in old["params"]] == [id(p) for p in params]:
            st.update(m=old["m"], v=old["v"], step=old["step"], counter=old["counter"], offs=old["offs"])
        else:
            offs, tot = [], 0
            for p in params:
                offs.append(tot)
                tot += (p.numel() + 3) // 4 * 4
            st.update(m=torch.zeros(tot, dtype=torch.float32, device=dev), v=torch.zeros(tot, dtype=torch.float32, device=dev),
                      step=torch.zeros(1, dtype=torch.float32, device=dev),
                      counter=torch.zeros(1, dtype=torch.int32, device=dev), offs=offs)
            for p, o in zip(params, offs):
                s = self.state[p]
                s["step"] = st["step"]      # (late-launched parameters count the same steps in st["step_late"])
                s["exp_avg"] = st["m"][o:o + p.numel()].view(p.shape)
                s["exp_avg_sq"] = st["v"][o:o + p.numel()].view(p.shape)
        # table(s): main launch + optional late launch (big Linear weights, see the class docstring)
        recs = self._recs_of(params)
        is_late = []
        for p in params:
            rec, _ = recs.get(id(p), (None, None))
            is_late.append(self.late_numel > 0 and isinstance(rec, engine.LinearRec) and rec.Hf * rec.Wf <= 64 and
                           rec.w_fwd is not None and p.numel() >= self.late_numel)
        order = [k for k in range(len(params)) if not is_late[k]] + [k for k in range(len(params)) if is_late[k]]
        n_main = sum(1 for f in is_late if not f)
        arr = (L.AdamEntry * len(params))()
        blocks = 0
        main_blocks = 0
        stores = set()
        late_stores = set()
        for k, src in enumerate(order):
            p, o = params[src], st["offs"][src]
            if k == n_main:
                main_blocks, blocks = blocks, 0        # the late table counts its blocks from zero
            e = arr[k]
            if not p.is_contiguous() or not p.grad.is_contiguous() or p.dtype != torch.float32:
                raise L.TorchSRB200Error("FusedAdam needs contiguous fp32 parameters and gradients")
            e.p, e.g = ops.ptr(p.detach()), ops.ptr(p.grad)
            e.m, e.v = ops.ptr(st["m"], o), ops.ptr(st["v"], o)
            e.numel = p.numel()
            e.block_start = blocks
            rec, store = recs.get(id(p), (None, None))
            if isinstance(rec, engine.ConvRec) and rec.kind == "std" and rec.w_fwd is not None:
                e.mode = L.AD_CONV
                e.cout, e.cin, e.kk = rec.cout, rec.cin, rec.k * rec.k
                e.dst_fwd, e.rows_fwd, e.cols_fwd = ops.ptr(rec.w_fwd), rec.cout_pad, rec.cols
                if rec.need_dgrad:
                    e.dst_t, e.rows_t, e.cols_t = ops.ptr(rec.w_t), rec.t_rows, rec.t_cols
                e.shuffle = int(rec.shuffle)
                if rec.k == 3 and rec.cin % 32 == 0 and rec.cout % 16 == 0 and not rec.shuffle:
                    e.mode = L.AD_CONV_TILE
                    blocks += (rec.cout // 16) * (rec.cin // 32)
                else:
                    blocks += (rec.cout * rec.cin + 255) // 256
                stores.add(store)
            elif isinstance(rec, engine.LinearRec) and rec.Hf * rec.Wf <= 64 and rec.w_fwd is not None:
                e.mode = L.AD_LINEAR
                e.cout, e.cin, e.kk = rec.nout, rec.C, rec.Hf * rec.Wf
                e.dst_fwd = ops.ptr(rec.w_fwd)
                blocks += rec.nout * (((rec.C + 31) // 32 + ADAM_TILES - 1) // ADAM_TILES)
                stores.add(store)
                if k >= n_main:
                    late_stores.add(store)
            else:
                e.mode = L.AD_PLAIN
                blocks += (p.numel() + 1023) // 1024
                if store is not None:
                    stores.add(store)
        # pinned staging + async copy: legal while a CUDA graph is being captured (the host buffer is kept alive)
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        if dev.type == "cuda":
            host = host.pin_memory()
        st["table_host"] = host
        st["table"] = torch.empty(host.numel(), dtype=torch.uint8, device=dev)
        st["table"].copy_(host, non_blocking=True)
        if n_main == len(params):
            main_blocks, blocks = blocks, 0
        st["n"], st["blocks"], st["stores"] = n_main, main_blocks, stores
        st["n_late"], st["blocks_late"], st["late_stores"] = len(params) - n_main, blocks, late_stores
        st["late_offset"] = n_main * C.sizeof(L.AdamEntry)
        st["params_late"] = [params[k] for k in order[n_main:]]
        if st["n_late"] and "step_late" not in st:
            src = old if (old is not None and "step_late" in old) else None
            st["step_late"] = src["step_late"] if src else st["step"].clone()
            st["counter_late"] = src["counter_late"] if src else torch.zeros(1, dtype=torch.int32, device=dev)
        variants[key] = st
        return st

    @staticmethod
    def _stores_of(params) -> List:
        ids = {id(p) for p in params}
        return [s for s in list(engine._LIVE_STORES) if any(id(q) in ids for q in s.params)]

    def _recs_of(self, params) -> dict:
        out = {}
        for store in self._stores_of(params):
            for r in store.convs:
                out[id(r.weight)] = (r, store)
            for r in store.linears:
                out[id(r.weight)] = (r, store)
            for q in store.params:
                out.setdefault(id(q), (None, store))
        return out

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            if not any(p.grad is not None for p in group["params"]):
                continue
            st = self._group_state(gi, group)
            lr = group["lr"]
            lr_t = lr if (isinstance(lr, torch.Tensor) and lr.is_cuda) else None
            b1, b2 = group["betas"]
            f = [0.0 if lr_t is not None else float(lr), b1, b2, group["eps"], 1.0 - b1, 1.0 - b2]
            if st["n"]:
                ops.run_now(ops.elt(L.E_ADAM, p=[st["table"], lr_t, st["step"], st["counter"]], i=[st["n"], st["blocks"]],
                                    f=f))
            if st["n_late"] and self._late_done_in_backward:
                self._late_done_in_backward = False          # already applied inside this step's backward
            elif st["n_late"]:
                dev = st["table"].device
                cur = torch.cuda.current_stream(dev)
                if self._late_stream is None:
                    self._late_stream = torch.cuda.Stream(device=dev)
                side = self._late_stream
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    ops.run_now(ops.elt(L.E_ADAM, p=[ops.ptr(st["table"], st["late_offset"]), lr_t, st["step_late"],
                                                     st["counter_late"]], i=[st["n_late"], st["blocks_late"]], f=f))
                    ev = side.record_event()
                for store in st["late_stores"]:
                    store.late_event = ev
                self._late_pending.append((ev, st["late_stores"]))
            for store in st["stores"]:
                store.opt_fresh = True     # its 'std' conv / Linear packs were just rewritten by the kernel
        return loss

    # ---- late parameters updated from inside the module's backward launch list
    def late_in_backward(self, module):
        """Context for ONE backward of `module` that is followed by step(): the late parameters (see the class
        docstring) are updated by a launch placed inside the backward launch list, right behind the GEMM that produces
        their gradient - for the discriminator that is the head of backward, so the 75 MB classifier update streams
        through HBM under the convolutional backward instead of after it. step() then only updates the rest. Same
        arithmetic, same step count; active once a regular step() has created the optimizer state."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            ok = bool(self._tables) and module._tsr.get("ddp") is None
            if ok:
                module._tsr["inline_adam"] = self
            try:
                yield
            finally:
                module._tsr.pop("inline_adam", None)
        return ctx()

    def late_desc_for(self, plan):
        """The launch descriptor of the in-backward update for `plan` (its gradients are slices of the plan's flat
        buffer, engine alias mode), or None when this optimizer has no late parameters."""
        ent = self._inline_tables.get(id(plan))
        if ent is None:
            st = next((v for g in self._tables.values() for v in g.values() if v.get("n_late")), None)
            if st is None:
                return None
            n = st["n_late"]
            raw = bytes(st["table_host"].numpy().tobytes())[st["late_offset"]:st["late_offset"] + n * C.sizeof(L.AdamEntry)]
            arr = (L.AdamEntry * n).from_buffer_copy(raw)
            late_params = st["params_late"]
            for e, p in zip(arr, late_params):
                e.g = ops.ptr(plan.grads.grad_slice(p))
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).pin_memory()
            tab = torch.empty(host.numel(), dtype=torch.uint8, device=st["table"].device)
            tab.copy_(host, non_blocking=True)
            group = self.param_groups[0]
            lr = group["lr"]
            lr_t = lr if (isinstance(lr, torch.Tensor) and lr.is_cuda) else None
            b1, b2 = group["betas"]
            desc = ops.elt(L.E_ADAM, p=[tab, lr_t, st["step_late"], st["counter_late"]], i=[n, st["blocks_late"]],
                           f=[0.0 if lr_t is not None else float(lr), b1, b2, group["eps"], 1.0 - b1, 1.0 - b2], side=True)
            ent = self._inline_tables[id(plan)] = dict(desc=desc, tab=tab, host=host, stores=st["late_stores"])
        return ent["desc"]

    def mark_late_done(self):
        """Called by the module's backward after it ran the in-backward update (engine.Plan.run_backward)."""
        self._late_done_in_backward = True

    def join(self):
        """Makes the current stream wait for the late launches of step() and forgets them."""
        for ev, stores in self._late_pending:
            torch.cuda.current_stream().wait_event(ev)
            for store in stores:
                if store.late_event is ev:
                    store.late_event = None
        self._late_pending.clear()
