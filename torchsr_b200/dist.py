"""Data-parallel plumbing: one process per GPU, gradients averaged with NCCL over NVLink.

Replaces the DistributedDataParallel wrappers of the reference (torchsr/srgan/trainer.py:143-157, torchsr.py:257-258):
  * attach(): parameters (and buffers) are broadcast from rank 0 once - mandatory, ranks are seeded differently
    (torchsr.py:152-153);
  * every backward of an attached module averages its flat fp32 gradient over the ranks with bucketed NCCL all-reduces,
    except the discriminator's 75 MB classifier weight gradient: that one is an outer product over the batch, so the
    ranks all-gather its two small bf16 factors at the head of backward (the transfer overlaps the convolutional
    backward) and each forms the averaged product itself with one tensor-core GEMM (engine.Plan.run_backward);
  * BatchNorm statistics stay local to each rank (the reference uses no SyncBatchNorm); with broadcast_buffers=True
    rank 0's running statistics are re-broadcast before each training forward, as DDP does for the generator.
"""
import contextlib
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def init_process_group(local_rank: int):
    """dist.init_process_group('nccl') as the reference does (torchsr.py:257-258), with NCCL's kernels on HIGH-priority
    streams. The compute graph of a training step keeps every SM slot occupied (programmatic dependent launch stages the
    next kernel's CTAs while the current one drains) on high-priority streams of its own; default-priority NCCL kernels
    then only start in the gaps - measured: the gradient all-reduces issued in the middle of the discriminator backward
    ran after its end. High priority lets their few CTAs in as soon as any slot frees."""
    import os
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    opts = None
    try:
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
    except Exception:  # noqa: BLE001
        opts = None
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=opts)


class DataParallelState:
    def __init__(self, group=None, broadcast_buffers: bool = True):
        self.group = group
        self.world = dist.get_world_size(group)
        self.broadcast_buffers = broadcast_buffers
        self.pending: List = []
        self.gathers: List = []

    def allreduce_async(self, t: torch.Tensor):
        """Average `t` over the ranks, asynchronously w.r.t. the current stream."""
        if self.world == 1:
            return
        backend = dist.get_backend(self.group)
        if backend == "nccl":
            self.pending.append((dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=True), None))
        else:   # gloo has no AVG
            self.pending.append((dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True), t))

    def allgather_async(self, out: torch.Tensor, t: torch.Tensor):
        """out[r * t.numel():(r + 1) * t.numel()] = rank r's t, asynchronously w.r.t. the current stream."""
        if self.world == 1:
            out.copy_(t)
            return
        self.gathers.append(dist.all_gather_into_tensor(out, t, group=self.group, async_op=True))

    def wait_gathers(self):
        for work in self.gathers:
            work.wait()
        self.gathers.clear()

    def wait(self):
        self.wait_gathers()
        for work, t in self.pending:
            work.wait()
            if t is not None:
                t.div_(self.world)
        self.pending.clear()


def bucket_slices(total: int, early_from: Optional[int], max_elems: int = 32 * 1024 * 1024) -> List[Tuple[int, int]]:
    """Splits the flat gradient [0, total) into buckets: the tail [early_from, total) first (it is produced first in
    backward), then the head; each piece capped at max_elems elements so that one bucket is at most 128 MB."""
    out: List[Tuple[int, int]] = []

    def split(lo, hi):
        while lo < hi:
            step = min(max_elems, hi - lo)
            out.append((lo, lo + step))
            lo += step

    if early_from is not None and 0 < early_from < total:
        split(early_from, total)
        split(0, early_from)
    else:
        split(0, total)
    return out


def allreduce_async_flat(flat: torch.Tensor, state: "DataParallelState"):
    """Launches the in-place averaging of a flat gradient (slice) over the ranks, bucketed; state.wait() completes it."""
    for lo, hi in bucket_slices(flat.numel(), None):
        state.allreduce_async(flat[lo:hi])


def allreduce_flat(flat: torch.Tensor, state: "DataParallelState"):
    allreduce_async_flat(flat, state)
    state.wait()


def attach(module, group=None, broadcast_buffers: bool = True) -> DataParallelState:
    """Makes `module` (a torchsr_b200 drop-in module) data parallel over the default process group."""
    if not dist.is_initialized():
        raise RuntimeError("torchsr_b200.dist.attach needs an initialised torch.distributed process group")
    state = DataParallelState(group, broadcast_buffers)
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    module._tsr["ddp"] = state
    return state


def _broadcast_buffers(module, state):
    with torch.no_grad():
        groups = {}
        for b in module.buffers():
            groups.setdefault(b.dtype, []).append(b)
        for bufs in groups.values():
            flat = torch.cat([b.reshape(-1) for b in bufs])
            dist.broadcast(flat, src=0, group=state.group)
            torch._foreach_copy_(bufs, [c.view(b.shape) for b, c in zip(bufs, flat.split([b.numel() for b in bufs]))])


def sync_buffers(module):
    """DDP's broadcast_buffers=True behaviour: rank 0's buffers win before a training forward. The buffers are
    coalesced per dtype into one flat tensor (one broadcast per dtype instead of one per buffer), as DDP does.
    When the previous training forward already published rank 0's buffers (publish_buffers, off the critical path)
    nothing has changed them since and the broadcast is skipped."""
    state = module._tsr.get("ddp")
    if state is None or not state.broadcast_buffers or state.world == 1:
        return
    ev = getattr(state, "buffers_event", None)
    if ev is not None:
        torch.cuda.current_stream().wait_event(ev)      # (already joined by join_buffers in the trainers: a no-op)
        state.buffers_event = None
    if getattr(state, "buffers_fresh", False):
        state.buffers_fresh = False
        return
    _broadcast_buffers(module, state)


def publish_buffers(module):
    """Right after a training forward (the only thing that changes BatchNorm buffers) rank 0's buffers are broadcast on a
    side stream, so the next forward finds them in place instead of paying two broadcasts at its head. Same values
    as DDP's pre-forward broadcast: nothing modifies the buffers in between. join_buffers() (or the next forward) makes
    the consumer stream wait."""
    state = module._tsr.get("ddp")
    if state is None or not state.broadcast_buffers or state.world == 1 or not torch.cuda.is_available():
        return
    cur = torch.cuda.current_stream()
    side = getattr(state, "buffers_stream", None)
    if side is None:
        side = state.buffers_stream = torch.cuda.Stream()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        _broadcast_buffers(module, state)
        state.buffers_event = side.record_event()
    state.buffers_fresh = True


def join_buffers(module):
    """The current stream waits for publish_buffers (end of a training step; required inside CUDA-graph capture)."""
    state = module._tsr.get("ddp")
    ev = getattr(state, "buffers_event", None) if state is not None else None
    if ev is not None:
        torch.cuda.current_stream().wait_event(ev)
        state.buffers_event = None


@contextlib.contextmanager
def frozen(module):
    """Runs a block with the module's parameters not requiring grad: its backward then produces only the gradient
    w.r.t. the input (the discriminator inside the generator step, reference trainer.py:456)."""
    params = [p for p in module.parameters() if p.requires_grad]
    for p in params:
        p.requires_grad_(False)
    try:
        yield module
    finally:
        for p in params:
            p.requires_grad_(True)
