"""Data-parallel host logic on the CPU: two gloo ranks average bucketed flat gradients exactly like one process
averaging the per-rank gradients (BatchNorm statistics are not exchanged, as in the reference)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from torchsr_b200.dist import DataParallelState, bucket_slices
    torch.manual_seed(100 + rank)
    total, early = 10_000, 7_000
    src = torch.randn(total)                       # this rank's flat gradient
    flat = torch.empty(total)
    st = DataParallelState(None, broadcast_buffers=False)
    order = []
    for lo, hi in bucket_slices(total, early, max_elems=2048):
        flat[lo:hi].copy_(src[lo:hi])
        st.allreduce_async(flat[lo:hi])
        order.append(lo)
    st.wait()
    # parameter broadcast semantics of attach(): rank 0 wins
    w = torch.full((5,), float(rank + 1))
    dist.broadcast(w, src=0)
    gathered = [torch.empty(total) for _ in range(world)]
    dist.all_gather(gathered, src)
    if rank == 0:
        torch.save({"flat": flat, "mean": torch.stack(gathered).mean(0), "order": order, "w": w}, out)
    dist.destroy_process_group()


def test_bucketed_allreduce_two_ranks(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert torch.allclose(r["flat"], r["mean"], atol=1e-6)
    assert r["order"][0] == 7_000                      # tail buckets (produced first in backward) go first
    assert torch.equal(r["w"], torch.ones(5))
