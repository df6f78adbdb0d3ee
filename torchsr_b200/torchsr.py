"""CLI - mirror of torchsr/torchsr.py:101-270: `torchsr train` / `torchsr test` with the reference's flags, defaults
and torchrun / Slurm rank discovery. Fix of reference defect App. D1: `--seed` is read with a default for `test`."""
import os
import sys
from argparse import ArgumentParser, Namespace

import torch
import torch.distributed as dist

from .constants import BATCH_SIZE, EPOCHS, MODEL, PRE_EPOCHS, TRAIN_DIR
from .models import select_test_model, select_trainer_model


def get_device(args: Namespace) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError('torchsr_b200 needs an NVIDIA B200 (sm_100) GPU: there is no CPU path for the kernels')
    return torch.device('cuda')


def distributed_params(args: Namespace) -> Namespace:
    """WORLD_SIZE/RANK/LOCAL_RANK (torchrun) -> SLURM_* -> single process (reference :101-154)."""
    env = os.environ
    if 'WORLD_SIZE' in env and 'RANK' in env:
        args.world_size, args.rank = int(env['WORLD_SIZE']), int(env['RANK'])
        args.local_rank = int(env.get('LOCAL_RANK', 0))
    elif 'SLURM_NTASKS' in env and 'SLURM_PROCID' in env:
        args.world_size, args.rank = int(env['SLURM_NTASKS']), int(env['SLURM_PROCID'])
        args.local_rank = int(env.get('SLURM_LOCALID', 0))
        env.setdefault('MASTER_ADDR', getattr(args, 'master_addr', None) or '127.0.0.1')
        env.setdefault('MASTER_PORT', str(getattr(args, 'master_port', None) or 29500))
        env['WORLD_SIZE'], env['RANK'] = str(args.world_size), str(args.rank)
    else:
        args.world_size, args.rank, args.local_rank = 1, -1, 0
    args.distributed = args.world_size > 1
    seed = getattr(args, 'seed', 0)
    if seed:
        torch.manual_seed(seed + max(args.rank, 0))
    return args


def parse_args(argv=None) -> Namespace:
    parser = ArgumentParser(prog='torchsr', description='Super resolution of images with SRGAN / ESRGAN on B200')
    sub = parser.add_subparsers(dest='function', required=True)
    train = sub.add_parser('train', help='Train an SRGAN or ESRGAN model')
    train.add_argument('--batch-size', type=int, default=BATCH_SIZE)
    train.add_argument('--data-workers', type=int, default=16)
    train.add_argument('--dataset-multiplier', type=int, default=1)
    train.add_argument('--disable-amp', action='store_true')
    train.add_argument('--epochs', type=int, default=EPOCHS)
    train.add_argument('--gan-checkpoint', type=str, default=None)
    train.add_argument('--master-addr', type=str, default=None)
    train.add_argument('--master-port', type=int, default=None)
    train.add_argument('--model', type=str, default=MODEL, choices=['esrgan', 'srgan', 'ESRGAN', 'SRGAN'])
    train.add_argument('--pretrain-epochs', type=int, default=PRE_EPOCHS)
    train.add_argument('--psnr-checkpoint', type=str, default=None)
    train.add_argument('--seed', type=int, default=0)
    train.add_argument('--skip-image-save', action='store_true')
    train.add_argument('--train-dir', type=str, default=TRAIN_DIR)
    test = sub.add_parser('test', help='Upscale one image with a trained generator')
    test.add_argument('image', type=str)
    test.add_argument('--model', type=str, default=MODEL, choices=['esrgan', 'srgan', 'ESRGAN', 'SRGAN'])
    return parser.parse_args(argv)


def main(argv=None) -> None:
    args = distributed_params(parse_args(argv))
    device = get_device(args)
    if args.function == 'test':
        from .test import test
        print(test(args, select_test_model(args), device))
        return
    trainer_class, crop_size = select_trainer_model(args)
    torch.cuda.set_device(args.local_rank)
    if args.distributed:
        from .dist import init_process_group
        init_process_group(args.local_rank)       # reference torchsr.py:257-258 (NCCL); high-priority NCCL streams
    from .dataset import initialize_datasets
    train_loader, test_loader, train_len, test_len = initialize_datasets(
        args.train_dir, args.batch_size, crop_size, args.dataset_multiplier, args.data_workers, args.distributed,
        args.seed, device=device, rank=max(args.rank, 0), world_size=max(args.world_size, 1))
    trainer = trainer_class(device, args, train_loader, test_loader, train_len, test_len, args.distributed)
    trainer.train()
    if args.distributed:
        dist.destroy_process_group()


if __name__ == '__main__':
    main(sys.argv[1:])
