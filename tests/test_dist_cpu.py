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
    # Factor exchange of the discriminator's first Linear weight gradient (engine.Plan.run_backward): every rank holds
    # A_r = dpre1 [64 rows (batch, zero padded)][n] and X_r^T cut into 64-row batch chunks [chunk][k][64]; gathering both
    # appends K chunks, and ONE product over the gathered K, scaled by 1/world, must equal the mean over the ranks of the
    # per-rank products - the gradient DDP's all-reduce would have produced (reference */trainer.py:148-157).
    n, k, b = 24, 40, 11                           # features out, features in, this rank's batch (< 64)
    a = torch.zeros(64, n)
    a[:b] = torch.randn(b, n)
    x = torch.randn(b, k)
    xt = torch.zeros(1, k, 64)
    xt[0, :, :b] = x.t()
    a_all, xt_all = torch.empty(world * 64, n), torch.empty(world, k, 64)
    st.allgather_async(a_all.view(-1), a.view(-1))
    st.allgather_async(xt_all.view(-1), xt.view(-1))
    st.wait()
    # GEMM over the gathered chunks, as csrc/conv_igemm.cu runs it: dW[n][k] = sum_chunks sum_j A[chunk*64+j][n] * XT[chunk][k][j]
    dw = sum(a_all[c * 64:(c + 1) * 64].t() @ xt_all[c].t() for c in range(world)) / world
    local = a[:b].t() @ x
    locals_ = [torch.empty(n, k) for _ in range(world)]
    dist.all_gather(locals_, local)
    if rank == 0:
        torch.save({"flat": flat, "mean": torch.stack(gathered).mean(0), "order": order, "w": w,
                    "dw": dw, "dw_mean": torch.stack(locals_).mean(0)}, out)
    dist.destroy_process_group()


def test_bucketed_allreduce_two_ranks(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert torch.allclose(r["flat"], r["mean"], atol=1e-6)
    assert r["order"][0] == 7_000                      # tail buckets (produced first in backward) go first
    assert torch.equal(r["w"], torch.ones(5))
    assert torch.allclose(r["dw"], r["dw_mean"], atol=1e-5), (r["dw"] - r["dw_mean"]).abs().max()
