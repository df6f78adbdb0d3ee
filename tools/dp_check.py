"""Data-parallel gradient parity on real GPUs (torchrun --nproc-per-node N tools/dp_check.py): for the SRGAN generator
and discriminator, the gradient every rank holds after the in-backward exchange (bucketed NCCL all-reduce + factor
all-gather for the classifier weight) against the mean over ranks of the single-rank shard gradients, BatchNorm local
(reference torchsr/srgan/trainer.py:143-157). Also checks the initial parameter broadcast and two training steps.
Prints one JSON line on rank 0."""
import json
import os
import sys
from argparse import Namespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TORCHSR_VGG_WEIGHTS", "random")
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    from torchsr_b200.dist import init_process_group as tdist_init
    tdist_init(local)
    sys.path.insert(0, ROOT)
    import bench
    from torchsr_b200 import ops
    from torchsr_b200.srgan.trainer import SRGANTrainer
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    torch.manual_seed(1000 + rank)           # different initial weights per rank: attach() must broadcast rank 0's
    args = Namespace(disable_amp=False, batch_size=B, epochs=8, pretrain_epochs=1, gan_checkpoint=None,
                     psnr_checkpoint=None, skip_image_save=True, local_rank=local, rank=rank, world_size=world)
    tr = SRGANTrainer(torch.device("cuda"), args, [], [], 0, 0, True)
    w0 = torch.cat([p.detach().reshape(-1) for p in tr.discriminator.parameters()])
    ref = w0.clone()
    dist.broadcast(ref, src=0)
    bcast = float((w0 - ref).abs().max())
    lr, hr = bench.synthetic_batch(B, 77 + rank)
    lr, hr = lr.cuda(), hr.cuda()
    parity = bench.dp_parity(tr, lr, hr)
    losses = [float(tr._gan_loop(lr, hr, s)) for s in range(2)]        # eager steps
    losses += [float(tr.graph_step(lr, hr, s)) for s in range(3)]       # whole-step graph (NCCL captured)
    torch.cuda.synchronize()
    ops.check_watchdog()
    # after identical averaged updates every rank must hold identical weights
    w1 = torch.cat([p.detach().reshape(-1) for p in tr.discriminator.parameters()] +
                   [p.detach().reshape(-1) for p in tr.generator.parameters()])
    ref = w1.clone()
    dist.broadcast(ref, src=0)
    drift = (w1 - ref).abs().max().reshape(1)
    dist.all_reduce(drift, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"world": world, "batch_per_gpu": B, "dp_parity": parity, "broadcast_max_abs": bcast,
                          "weight_drift_max_abs": float(drift), "losses": losses}), flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()
