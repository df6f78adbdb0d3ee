"""Input pipeline parity: the oracle's restatement of Pillow's 8-bit bicubic resize is pinned against
Pillow itself and against the reference's own transform chain (CPU, byte-exact); the CUDA kernel is held to the oracle
byte for byte (gpu)."""
import numpy as np
import pytest
import torch


def _images(rng, sizes):
    out = []
    for k, (h, w) in enumerate(sizes):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if k % 3 == 1:
            img = (img // 85 * 85).astype(np.uint8)          # flat regions / hard edges: clamping paths
        if k % 3 == 2:
            img[::2] = 255
            img[1::2] = 0                                   # maximum ringing
        out.append(img)
    return out


def test_oracle_resize_is_pillow_bit_for_bit():
    import pil_resample as R
    from PIL import Image
    rng = np.random.default_rng(0)
    for size in (96, 128):
        for img in _images(rng, [(size, size)] * 9):
            ref = np.asarray(Image.fromarray(img).resize((size // 4, size // 4), Image.BICUBIC))
            assert np.array_equal(R.resize_u8(img, size // 4), ref)


def test_oracle_sample_equals_reference_transform_chain():
    """torchsr/dataset.py:86-99: hr = ToTensor(crop/flip), lr = ToTensor(Resize(BICUBIC)(ToPILImage(hr)))."""
    import pil_resample as R
    from PIL import Image
    from torchvision.transforms import Compose, InterpolationMode, Resize, ToPILImage, ToTensor
    from torchvision.transforms import functional as TF
    rng = np.random.default_rng(1)
    lr_transform = Compose([ToPILImage(), Resize((24, 24), interpolation=InterpolationMode.BICUBIC), ToTensor()])
    for img in _images(rng, [(131, 150), (96, 96), (200, 97)]):
        for (x0, y0, fh, fv) in ((0, 0, False, False), (img.shape[1] - 96, img.shape[0] - 96, True, False), (min(1, img.shape[1] - 96), 0, True, True)):
            pil = Image.fromarray(img).crop((x0, y0, x0 + 96, y0 + 96))
            if fh:
                pil = TF.hflip(pil)
            if fv:
                pil = TF.vflip(pil)
            hr_ref = ToTensor()(pil)
            lr_ref = lr_transform(hr_ref)
            lr, hr = R.train_sample(img, x0, y0, 96, fh, fv)
            assert np.array_equal(hr, hr_ref.numpy()) and np.array_equal(lr, lr_ref.numpy())


def test_product_tables_equal_oracle_tables():
    import pil_resample as R
    from torchsr_b200.gpu_data import pil_bicubic_tables
    for size in (96, 128, 64):
        kk, bounds, ksize = pil_bicubic_tables(size, size // 4)
        rk, rb = R.coefficients(size, size // 4)
        assert ksize == rk.shape[1] == 17
        assert np.array_equal(kk.numpy(), rk) and np.array_equal(bounds.numpy(), rb)


@pytest.mark.gpu
@pytest.mark.parametrize("crop", [96, 128])
def test_crop_kernel_is_byte_exact(crop):
    import pil_resample as R
    from torchsr_b200 import gpu_data as GD
    rng = np.random.default_rng(2)
    imgs = _images(rng, [(crop, crop), (crop + 37, crop + 5), (crop + 1, 2 * crop), (crop, crop + 64), (3 * crop, crop + 9)])
    pool = GD.ImagePool([torch.from_numpy(i) for i in imgs], "cuda")
    g = torch.Generator().manual_seed(3)
    B = 24
    idx = torch.randint(0, len(imgs), (B,), generator=g)
    x0 = torch.tensor([int(torch.randint(0, imgs[int(k)].shape[1] - crop + 1, (1,), generator=g)) for k in idx])
    y0 = torch.tensor([int(torch.randint(0, imgs[int(k)].shape[0] - crop + 1, (1,), generator=g)) for k in idx])
    fh, fv = torch.rand(B, generator=g) < 0.5, torch.rand(B, generator=g) < 0.5
    lr, hr = GD.crop_batch(pool, idx, x0, y0, fh, fv, crop)
    lr, hr = lr.cpu().numpy(), hr.cpu().numpy()
    for i in range(B):
        rl, rh = R.train_sample(imgs[int(idx[i])], int(x0[i]), int(y0[i]), crop, bool(fh[i]), bool(fv[i]))
        assert np.array_equal(hr[i], rh), i
        assert np.array_equal(lr[i], rl), i


@pytest.mark.gpu
def test_gpu_loader_epoch_shapes_and_shards():
    from torchsr_b200 import gpu_data as GD
    rng = np.random.default_rng(4)
    pool = GD.ImagePool([torch.from_numpy(i) for i in _images(rng, [(100 + k, 140) for k in range(10)])], "cuda")
    seen = []
    for rank in range(2):
        loader = GD.GpuTrainLoader(pool, 96, 4, multiplier=2, seed=5, rank=rank, world_size=2)
        assert len(loader) == 2                       # 10 images x 2 / 2 ranks / batch 4
        for lr, hr in loader:
            assert lr.shape == (4, 3, 24, 24) and hr.shape == (4, 3, 96, 96) and lr.is_cuda
            assert 0.0 <= float(lr.min()) and float(hr.max()) <= 1.0
            seen.append(hr.sum().item())
    assert len(seen) == 4
