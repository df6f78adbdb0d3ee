"""Image-folder input pipeline behind `torchsr train` (reference torchsr/dataset.py:55-428): 90/10 train/test split,
random HR crops + flips, bicubic /4 low-resolution inputs, one shard per rank.

Training batches come from the GPU pipeline (gpu_data.py: images decoded once into HBM, crop + flips + Pillow-exact
bicubic in one launch per batch) whenever the trainer runs on a CUDA device; TORCHSR_GPU_DATA=0 selects the reference's
scheme instead (PIL in DataLoader workers, one image decode per crop). The small evaluation split always uses the
DataLoader path."""
import os
import random
from typing import Tuple

import torch
from torch.utils.data import DataLoader, Dataset
from torch.utils.data.distributed import DistributedSampler

EXT = ('.png', '.jpg', '.jpeg', '.bmp')


def _files(root):
    out = []
    for d, _, fs in os.walk(root):
        out += [os.path.join(d, f) for f in fs if f.lower().endswith(EXT)]
    return sorted(out)


class TrainData(Dataset):
    def __init__(self, files, crop_size: int, multiplier: int = 1):
        self.files, self.crop, self.mult = files, crop_size, multiplier

    def __len__(self):
        return len(self.files) * self.mult

    def __getitem__(self, i):
        from PIL import Image
        from torchvision.transforms import functional as TF
        img = Image.open(self.files[i % len(self.files)]).convert('RGB')
        w, h = img.size
        c = self.crop
        if w < c or h < c:
            img = img.resize((max(w, c), max(h, c)), Image.BICUBIC)
            w, h = img.size
        x, y = random.randint(0, w - c), random.randint(0, h - c)
        hr = img.crop((x, y, x + c, y + c))
        if random.random() < 0.5:
            hr = TF.hflip(hr)
        if random.random() < 0.5:
            hr = TF.vflip(hr)
        lr = hr.resize((c // 4, c // 4), Image.BICUBIC)
        return TF.to_tensor(lr), TF.to_tensor(hr)


class TestData(Dataset):
    def __init__(self, files, crop_size: int):
        self.files, self.crop = files, crop_size

    def __len__(self):
        return len(self.files)

    def __getitem__(self, i):
        from PIL import Image
        from torchvision.transforms import functional as TF
        img = TF.center_crop(Image.open(self.files[i]).convert('RGB'), self.crop)
        lr = img.resize((self.crop // 4, self.crop // 4), Image.BICUBIC)
        bic = lr.resize((self.crop, self.crop), Image.BICUBIC)
        return TF.to_tensor(lr), TF.to_tensor(bic), TF.to_tensor(img)


def initialize_datasets(train_dir: str, batch_size: int, crop_size: int, dataset_multiplier: int = 1, workers: int = 16,
                        distributed: bool = False, seed: int = 0, device=None, rank: int = 0,
                        world_size: int = 1) -> Tuple[DataLoader, DataLoader, int, int]:
    files = _files(train_dir)
    if not files:
        raise RuntimeError(f'no images found under {train_dir}')
    rng = random.Random(seed or 0)
    rng.shuffle(files)
    n_test = max(1, len(files) // 10)
    test_files, train_files = files[:n_test], files[n_test:] or files
    train, test = TrainData(train_files, crop_size, dataset_multiplier), TestData(test_files, crop_size)
    es = DistributedSampler(test, seed=seed, shuffle=False) if distributed else None
    use_gpu = device is not None and torch.device(device).type == 'cuda' and os.environ.get('TORCHSR_GPU_DATA', '1') != '0'
    if use_gpu:
        from .gpu_data import GpuTrainLoader, ImagePool
        pool = ImagePool.from_files(train_files, torch.device(device))
        small = [f for f, (h, w) in zip(train_files, pool.shapes) if h < crop_size or w < crop_size]
        if small:      # the CPU path upsizes such images; keep them out of the pool instead of failing the run
            keep = [f for f in train_files if f not in set(small)]
            if not keep:
                raise RuntimeError(f'every image under {train_dir} is smaller than the {crop_size}x{crop_size} crop')
            pool = ImagePool.from_files(keep, torch.device(device))
        tl = GpuTrainLoader(pool, crop_size, batch_size, dataset_multiplier, seed=seed, rank=rank if distributed else 0,
                            world_size=world_size if distributed else 1)
    else:
        ts = DistributedSampler(train, seed=seed) if distributed else None
        tl = DataLoader(train, batch_size=batch_size, shuffle=ts is None, sampler=ts, num_workers=workers,
                        pin_memory=True, drop_last=True)
    el = DataLoader(test, batch_size=batch_size, shuffle=False, sampler=es, num_workers=workers, pin_memory=True)
    return tl, el, len(train), len(test)
