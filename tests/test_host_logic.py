"""Host-side logic that needs no GPU: the C-ABI library loads and exports what include/torchsr_b200.h declares,
descriptor geometry, extent validation, plan construction in dry mode, gradient bucketing, loud failure on CPU."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from torchsr_b200 import _lib, build
    build.build()
    hdr = open(os.path.join(ROOT, "include", "torchsr_b200.h")).read()
    declared = set(re.findall(r"\b(tsr_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert set(_lib.EXPORTS) <= declared
    assert lib.tsr_version() >= 1


def test_ctypes_structs_match_header_sizes():
    """sizeof of the ctypes mirrors equals what a C compiler computes for the header's structs."""
    from torchsr_b200 import _lib
    src = '#include <stdio.h>\n#include "torchsr_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(tsr_conv_desc_t),' \
          ' sizeof(tsr_wgrad_desc_t), sizeof(tsr_elt_desc_t), sizeof(tsr_pack_entry_t), sizeof(tsr_adam_entry_t));return 0;}\n'
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(_lib.ConvDesc), ctypes.sizeof(_lib.WgradDesc), ctypes.sizeof(_lib.EltDesc),
                     ctypes.sizeof(_lib.PackEntry), ctypes.sizeof(_lib.AdamEntry)]


def test_geometry_helpers():
    from torchsr_b200 import ops
    g = ops.fwd_geometry(96, 96, 3, 3, 1, 1, 2)
    assert (g["Ho"], g["Wo"], g["lower_h"], g["upper_h"]) == (48, 48, -1, -1)
    g = ops.fwd_geometry(24, 24, 9, 9, 4, 4, 1)
    assert len(g["taps"]) == 81 and g["lower_w"] == -4 and g["upper_w"] == -4
    g = ops.dgrad_s1_geometry(24, 24, 3, 3, 1, 1)
    assert g["taps"][0] == (2, 2, 0) and g["taps"][-1] == (0, 0, 8)
    cls = ops.dgrad_s2_classes(3, 1)
    assert cls[0] == [(0, 1)] and sorted(cls[1]) == [(0, 2), (1, 0)]
    # every (input row, kernel row) pair of a stride-2 conv appears in exactly one parity class
    for hx in range(8):
        r, a = hx % 2, hx // 2
        pairs = {(a + off, k) for off, k in cls[r]}
        want = {(hy, k) for hy in range(-1, 6) for k in range(3) if 2 * hy - 1 + k == hx}
        assert pairs == want


def test_extent_validation_rejects_out_of_bounds():
    from torchsr_b200 import _lib as L
    from torchsr_b200 import ops
    x = torch.zeros(2 * 8 * 8 * 64, dtype=torch.bfloat16)
    w = torch.zeros(9 * 64 * 64, dtype=torch.bfloat16)
    out = torch.zeros(2 * 8 * 8 * 64 - 8, dtype=torch.bfloat16)      # 8 elements too small
    geom = ops.fwd_geometry(8, 8, 3, 3, 1, 1, 1)
    d = ops.conv_desc(x=x, N=2, H=8, W=8, C=64, x_ld=64, geom=geom, w=w, cout_pad=64, w_ld=64, n_slots=9, block_n=64,
                      out=out, os_n=8 * 8 * 64, os_h=8 * 64, os_w=64, n_valid=64)
    with pytest.raises(ops.ExtentError):
        ops.validate(d)
    ok = torch.zeros(2 * 8 * 8 * 64, dtype=torch.bfloat16)
    d.out = ops.ptr(ok)
    ops.validate(d)
    e = ops.elt(L.E_CAST, p=[torch.zeros(10), torch.zeros(4, dtype=torch.bfloat16)], i=[10, 0])
    with pytest.raises(ops.ExtentError):
        ops.validate(e)


def test_extent_validation_of_the_gather_and_replicating_stores():
    """OUT_GATHER_W (one and two output rows per GEMM row) and out_rep2x descriptors: extents of the fp32 NCHW image /
    the x2 tensor are checked, incomplete combinations are rejected."""
    from torchsr_b200 import _lib as L
    from torchsr_b200 import ops
    B, H, W = 1, 8, 8
    x = torch.zeros(B * H * W * 64, dtype=torch.bfloat16)
    # two output rows per GEMM row: ten vertical taps, stride 2 along H, 64 columns
    w2 = torch.zeros(10 * 64 * 64, dtype=torch.bfloat16)
    geom2 = dict(lower_h=-4, lower_w=0, upper_h=-5, upper_w=0, Ho=H // 2, Wo=W, stride=2, stride_w=1,
                 taps=[(k, 0, k) for k in range(10)])
    img = torch.zeros(B * 3 * H * W)
    d = ops.conv_desc(x=x, N=B, H=H, W=W, C=64, x_ld=64, geom=geom2, w=w2, cout_pad=64, w_ld=64, n_slots=10, block_n=64,
                      out=img, os_n=0, os_h=0, os_w=0, n_valid=64, gather=dict(k=9, pad=4, c=3, bias=torch.zeros(3), rows=2))
    assert d.out_mode == L.OUT_GATHER_W and d.gather_rows == 2 and d.stride == 2 and d.stride_w == 1
    ops.validate(d)
    d.out = ops.ptr(torch.zeros(B * 3 * H * W - 4))          # image too small
    with pytest.raises(ops.ExtentError):
        ops.validate(d)
    # one row per GEMM row needs 32 columns
    w1 = torch.zeros(9 * 32 * 64, dtype=torch.bfloat16)
    geom1 = ops.fwd_geometry(H, W, 9, 1, 4, 0, 1)
    d1 = ops.conv_desc(x=x, N=B, H=H, W=W, C=64, x_ld=64, geom=geom1, w=w1, cout_pad=32, w_ld=64, n_slots=9, block_n=32,
                       out=img, os_n=0, os_h=0, os_w=0, n_valid=32, gather=dict(k=9, pad=4, c=3))
    ops.validate(d1)
    d1.gather_rows = 2
    with pytest.raises(ops.ExtentError):
        ops.validate(d1)
    # nearest x2 replica of the result
    w3 = torch.zeros(9 * 64 * 64, dtype=torch.bfloat16)
    out = torch.zeros(B * H * W * 64, dtype=torch.bfloat16)
    up = torch.zeros(B * 2 * H * 2 * W * 64, dtype=torch.bfloat16)
    g3 = ops.fwd_geometry(H, W, 3, 3, 1, 1, 1)
    d3 = ops.conv_desc(x=x, N=B, H=H, W=W, C=64, x_ld=64, geom=g3, w=w3, cout_pad=64, w_ld=64, n_slots=9, block_n=64, out=out,
                       os_n=H * W * 64, os_h=W * 64, os_w=64, n_valid=64,
                       rep2x=dict(t=up, strides=(4 * H * W * 64, 2 * W * 64, 64)))
    ops.validate(d3)
    d3.out_rep2x = ops.ptr(torch.zeros(B * 2 * H * 2 * W * 64 - 64, dtype=torch.bfloat16))
    with pytest.raises(ops.ExtentError):
        ops.validate(d3)


def test_modules_fail_loudly_without_cuda():
    from torchsr_b200 import _lib as L
    from torchsr_b200.srgan.generator import Generator
    with pytest.raises(L.TorchSRB200Error):
        Generator()(torch.rand(1, 3, 8, 8))


DRY_SCRIPT = r"""
import os, sys
os.environ["TSR_DRY"] = "1"
sys.path.insert(0, %r)
import torch
from torchsr_b200.srgan.generator import Generator
from torchsr_b200.srgan.discriminator import Discriminator
from torchsr_b200.esrgan.discriminator import Discriminator as ED
from torchsr_b200.esrgan.generator import Generator as EG
from torchsr_b200.srgan.residual import ResidualBlock, SubpixelConvolutionLayer
from torchsr_b200.esrgan.residual import ResidualDenseBlock, ResidualInResidualDenseBlock
for mod, shape, need_x in [(ResidualDenseBlock(), (2, 64, 8, 8), True), (ResidualInResidualDenseBlock(), (1, 64, 8, 8), True),
                           (Generator(), (2, 3, 24, 24), False), (Discriminator(), (2, 3, 96, 96), True),
                           (ED(), (1, 3, 128, 128), True), (EG(num_rrdb_blocks=2), (1, 3, 16, 16), False),
                           (ResidualBlock(), (2, 64, 8, 8), True), (SubpixelConvolutionLayer(), (2, 64, 8, 8), True)]:
    x = torch.rand(*shape, requires_grad=need_x)
    y = mod(x)
    y.sum().backward()
    assert all(p.grad is not None and p.grad.shape == p.shape for p in mod.parameters()), type(mod)
    if need_x:
        assert x.grad.shape == x.shape
    mod.eval()
    with torch.no_grad():
        mod(x.detach())
    plans = mod._tsr["plans"]
    assert len(plans) == 2, plans.keys()
    for key, pool in plans.items():
        assert len(pool) == 1 and not pool[0].busy
# one discriminator pass over (real | fake) with BatchNorm statistics per half, fused and two-launch BatchNorm paths
from torchsr_b200 import engine, losses
for fuse in (True, False):
    engine.FUSE_BN_FWD = engine.FUSE_BN_BWD = fuse
    for D, n, size in ((Discriminator(), 8, 96), (ED(), 2, 128), (Discriminator(), 3, 96)):
        a, b = torch.rand(n, 3, size, size), torch.rand(n, 3, size, size, requires_grad=True)
        pa, pb = D.forward_pair(a, b)
        assert pa.shape == (n, 1) and pb.shape == (n, 1)
        losses.bce(pa, 1.0, pb, 0.0).backward()
        assert all(p.grad is not None for p in D.parameters()) and b.grad.shape == b.shape
        paired = [k for k in D._tsr["plans"] if len(k) == 3]
        assert (len(paired) == 1) == (n != 3), (n, D._tsr["plans"].keys())      # 3 x 36 rows per half: falls back
engine.FUSE_BN_FWD = engine.FUSE_BN_BWD = True
x = torch.rand(4, 3, 8, 8, requires_grad=True)
for fn in (lambda: losses.mse(x, torch.rand(4, 3, 8, 8)), lambda: losses.l1(x, torch.rand(4, 3, 8, 8), scale=0.01),
           lambda: losses.total(losses.relativistic_d(x[:, :1, 0, 0], x[:, 1:2, 0, 0]),
                                losses.relativistic_g(x[:, :1, 0, 0], x[:, 1:2, 0, 0].detach(), scale=0.005))):
    out = fn()
    assert out.dim() == 0
    out.backward(torch.ones(()))
    assert x.grad.shape == x.shape
print("DRY_OK")
"""


def test_plans_build_and_validate_in_dry_mode():
    """Builds the forward and backward launch lists of every module on the CPU (descriptors are validated for extents
    and alignment, nothing is launched)."""
    out = subprocess.run([sys.executable, "-c", DRY_SCRIPT % ROOT], capture_output=True, text=True, timeout=600)
    assert "DRY_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_bucket_slices_cover_flat_gradient_once():
    from torchsr_b200.dist import bucket_slices
    for total, early, cap in [(1000, 600, 300), (1000, None, 400), (23_563_652, 4_689_000, 8 * 1024 * 1024), (10, 0, 4)]:
        sl = bucket_slices(total, early, cap)
        covered = sorted(sl)
        assert covered[0][0] == 0 and covered[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
        assert all(hi - lo <= cap for lo, hi in sl)
        if early:
            assert sl[0][0] == early      # the tail (produced first in backward) is sent first


def test_image_folder_pipeline_shapes_and_split(tmp_path):
    """The minimal loader behind `torchsr train` (reference torchsr/dataset.py:55-136,279): 90/10 split, HR crops of the
    model's crop size with /4 bicubic LR inputs in [0, 1], (LR, bicubic, HR) triples for evaluation."""
    import numpy as np
    import torch
    from PIL import Image
    from torchsr_b200.dataset import initialize_datasets
    rng = np.random.default_rng(0)
    for i in range(20):
        Image.fromarray(rng.integers(0, 255, (100 + i, 130, 3), dtype=np.uint8)).save(tmp_path / f"im{i:02d}.png")
    tl, el, n_train, n_test = initialize_datasets(str(tmp_path), 4, 96, dataset_multiplier=2, workers=0, seed=3)
    assert (n_train, n_test) == (36, 2)
    lr, hr = next(iter(tl))
    assert lr.shape == (4, 3, 24, 24) and hr.shape == (4, 3, 96, 96) and lr.dtype == torch.float32
    assert 0.0 <= float(lr.min()) and float(hr.max()) <= 1.0
    assert len(tl) == 9                     # drop_last: 36 crops / 4
    low, bic, high = next(iter(el))
    assert low.shape == (2, 3, 24, 24) and bic.shape == (2, 3, 96, 96) and high.shape == (2, 3, 96, 96)
    import pytest
    with pytest.raises(RuntimeError):
        initialize_datasets(str(tmp_path / "missing"), 4, 96, workers=0)


def test_trainer_resumes_at_checkpoint_epoch(tmp_path, monkeypatch):
    """reference srgan/trainer.py:357-364, :483-499: a restarted job continues at checkpoint['epoch'] (it does not
    repeat finished epochs), a GAN phase without its own checkpoint starts at 1 from the PSNR-phase weights."""
    from argparse import Namespace
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("TORCHSR_VGG_WEIGHTS", "random")
    from torchsr_b200.srgan.trainer import SRGANTrainer
    args = Namespace(disable_amp=False, batch_size=2, epochs=8, pretrain_epochs=4, gan_checkpoint=None,
                     psnr_checkpoint=None, skip_image_save=True, local_rank=-1, rank=-1, world_size=1)
    tr = SRGANTrainer(torch.device("cpu"), args, [], [], 0, 0, False)
    seen = []
    monkeypatch.setattr(tr, "_test", lambda epoch, phase, step: seen.append((phase, epoch)))
    marker = {k: torch.full_like(v, 0.5) if v.is_floating_point() else v for k, v in tr.generator.state_dict().items()}
    torch.save({"epoch": 3, "phase": "srgan-psnr", "state": {"module." + k: v for k, v in marker.items()}},
               "srgan-psnr-latest.pth")
    tr._pretrain()
    assert seen == [("srgan-psnr", 3), ("srgan-psnr", 4)], seen
    assert float(tr.generator.conv3.weight.mean()) == 0.5          # weights restored (module. prefix stripped)
    seen.clear()
    tr._gan_train()                                                 # no GAN checkpoint: epoch 1, PSNR weights
    assert [e for _, e in seen] == list(range(1, 9)), seen
    torch.save({"epoch": 7, "phase": "srgan-gan", "state": marker}, "srgan-gan-latest.pth")
    seen.clear()
    tr._gan_train()
    assert seen == [("srgan-gan", 7), ("srgan-gan", 8)], seen


def test_install_as_torchsr_gives_reference_import_paths():
    """`import torchsr...` (the name the reference installs under, setup.py:39-41) served by the drop-in package."""
    code = ("import sys; sys.path.insert(0, %r); import torchsr_b200; torchsr_b200.install_as_torchsr(); "
            "from torchsr.srgan.generator import Generator; from torchsr.esrgan.residual import ResidualDenseBlock; "
            "from torchsr.models import select_trainer_model, CROP_SIZE; from torchsr.torchsr import main; "
            "import torchsr_b200.srgan.generator as g; assert Generator is g.Generator and CROP_SIZE['srgan'] == 96; "
            "print('ALIAS_OK')") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "ALIAS_OK" in out.stdout, out.stdout + out.stderr[-2000:]
