"""GPU-side training input pipeline - replaces the per-sample CPU work of the reference's
TrainData.__getitem__ (torchsr/dataset.py:55-136): PIL decode -> RandomCrop -> random flips -> ToTensor for the HR
crop, ToPILImage -> Resize(crop/4, BICUBIC) -> ToTensor for the LR input.

Every image is decoded ONCE (PIL, host) into a uint8 pool in HBM; after that a whole batch is one kernel launch
(csrc/eltwise.cu crop_lr_kernel): crop + flips + Pillow's antialiased 8-bit bicubic resize restated bit for bit, so
the LR inputs are byte-identical to what the reference's DataLoader workers produce (tests/test_gpu_data.py). At the
step rates of the B200 path (> 5 000 crops/s per GPU) sixteen PIL workers decoding a full image per crop are the
bottleneck by orders of magnitude; the pool costs H*W*3 bytes per image (DIV2K: ~7 GB of the 180 GB).

Randomness: crop origin and flips come from a torch.Generator (seeded per rank), not from the reference's
per-worker Python RNG streams - the augmentation distribution is the reference's (uniform origin, p = 0.5 flips), the
sample sequence is not (nor is it reproducible in the reference across worker counts).
"""
import math
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from . import ops

PRECISION_BITS = 32 - 8 - 2          # Pillow Resample.c: 8-bit coefficients carry 22 fractional bits


def _bicubic(x: float, a: float = -0.5) -> float:
    """Pillow's bicubic_filter (Resample.c): Keys cubic, a = -0.5."""
    x = -x if x < 0.0 else x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_bicubic_tables(in_size: int, out_size: int) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """Integer coefficient table kk [out][ksize] and bounds [out][2] (first tap, tap count) of Pillow's
    precompute_coeffs + normalize_coeffs_8bpc for a bicubic resize in_size -> out_size (same double arithmetic, same
    operation order: the tables are bit-identical to Pillow's)."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = torch.zeros(out_size, ksize, dtype=torch.int32)
    bounds = torch.zeros(out_size, 2, dtype=torch.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(v * (1 << PRECISION_BITS) + (-0.5 if v < 0 else 0.5))
        bounds[xx, 0], bounds[xx, 1] = xmin, xmax
    return kk, bounds, ksize


class ImagePool:
    """uint8 RGB images (HWC) packed back to back in one device buffer."""

    def __init__(self, images: Sequence[torch.Tensor], device):
        self.shapes: List[Tuple[int, int]] = []
        self.offsets: List[int] = []
        off = 0
        for im in images:
            if im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3:
                raise RuntimeError("ImagePool takes uint8 HxWx3 tensors")
            self.shapes.append((int(im.shape[0]), int(im.shape[1])))
            self.offsets.append(off)
            off += im.numel()
        self.buf = torch.empty(max(off, 16), dtype=torch.uint8, device=device)
        for im, o in zip(images, self.offsets):
            self.buf[o:o + im.numel()].copy_(im.reshape(-1), non_blocking=True)

    def __len__(self):
        return len(self.shapes)

    @classmethod
    def from_files(cls, files: Sequence[str], device):
        import numpy as np
        from PIL import Image
        images = [torch.from_numpy(np.asarray(Image.open(f).convert("RGB")).copy()) for f in files]
        return cls(images, device)


def crop_batch(pool: ImagePool, index: torch.Tensor, x0: torch.Tensor, y0: torch.Tensor, flip_h: torch.Tensor,
               flip_v: torch.Tensor, crop: int, tables=None):
    """(lr [B,3,crop/4,crop/4], hr [B,3,crop,crop]) fp32 on the pool's device for the given per-sample image indices,
    crop origins and flips (host int tensors)."""
    if crop % 4 or crop > 128:
        raise RuntimeError("crop_batch: crop must be a multiple of 4, at most 128")
    dev = pool.buf.device
    if dev.type != "cuda" and not ops.DRY:
        raise L.TorchSRB200Error("torchsr_b200.gpu_data runs only on a CUDA device (no CPU fallback)")
    B, out = int(index.numel()), crop // 4
    kk, bounds, ksize = tables if tables is not None else device_tables(crop, dev)
    rows = []
    for i in range(B):
        k = int(index[i])
        H, W = pool.shapes[k]
        xi, yi = int(x0[i]), int(y0[i])
        if not (0 <= xi <= W - crop and 0 <= yi <= H - crop):
            raise RuntimeError(f"crop ({xi},{yi})+{crop} leaves image {k} of size {W}x{H}")
        rows.append([pool.offsets[k], W, xi, yi, int(flip_h[i]), int(flip_v[i])])
    params = torch.tensor(rows, dtype=torch.int64).pin_memory() if dev.type == "cuda" else torch.tensor(rows, dtype=torch.int64)
    params = params.to(dev, non_blocking=True)
    hr = torch.empty(B, 3, crop, crop, dtype=torch.float32, device=dev)
    lr = torch.empty(B, 3, out, out, dtype=torch.float32, device=dev)
    ops.run_now(ops.elt(L.E_CROP_LR, p=[pool.buf, params, kk, bounds, hr, lr], i=[B, crop, out, ksize]))
    return lr, hr


_TABLES = {}


def device_tables(crop: int, device):
    key = (crop, str(device))
    if key not in _TABLES:
        kk, bounds, ksize = pil_bicubic_tables(crop, crop // 4)
        _TABLES[key] = (kk.to(device), bounds.to(device), ksize)
    return _TABLES[key]


class GpuTrainLoader:
    """Iterable of (low_res, high_res) device batches with the reference DataLoader's shape of an epoch: every image
    `multiplier` times, shuffled, `batch_size` per step, last ragged batch dropped (reference dataset.py:280-293 keeps
    it; the CUDA-graph step wants full batches - as the CPU loader here). With `world_size` > 1 every rank draws its
    own 1/world share of the permutation (DistributedSampler semantics, dataset.py:279)."""

    def __init__(self, pool: ImagePool, crop: int, batch_size: int, multiplier: int = 1, seed: int = 0, rank: int = 0,
                 world_size: int = 1):
        self.pool, self.crop, self.batch_size, self.multiplier = pool, crop, batch_size, multiplier
        self.rank, self.world = rank, max(world_size, 1)
        self.gen = torch.Generator().manual_seed(seed * 1000003 + rank)
        self.perm_gen = torch.Generator().manual_seed(seed)            # same permutation on every rank
        self.tables = device_tables(crop, pool.buf.device)
        for H, W in pool.shapes:
            if H < crop or W < crop:
                raise RuntimeError(f"image of size {W}x{H} is smaller than the {crop}x{crop} crop "
                                   "(torchvision RandomCrop raises for it in the reference too)")

    def __len__(self):
        return (len(self.pool) * self.multiplier // self.world) // self.batch_size

    def __iter__(self):
        n = len(self.pool) * self.multiplier
        perm = torch.randperm(n, generator=self.perm_gen) % len(self.pool)
        mine = perm[self.rank::self.world]
        for b in range(len(self)):
            idx = mine[b * self.batch_size:(b + 1) * self.batch_size]
            hs = torch.tensor([self.pool.shapes[int(k)][0] for k in idx])
            ws = torch.tensor([self.pool.shapes[int(k)][1] for k in idx])
            u = torch.rand(4, idx.numel(), generator=self.gen)
            x0 = (u[0] * (ws - self.crop + 1)).long().clamp_(max=ws - self.crop)
            y0 = (u[1] * (hs - self.crop + 1)).long().clamp_(max=hs - self.crop)
            yield crop_batch(self.pool, idx, x0, y0, u[2] < 0.5, u[3] < 0.5, self.crop, self.tables)
