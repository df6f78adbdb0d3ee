"""Inference command - mirror of torchsr/test.py:22-63: build the generator, load `<model>-gan-best.pth`, upscale
one image x4, write `upres-<image name>`.

Deliberate fixes of reference defects that make `torchsr test` unusable as shipped: all three
checkpoint layouts are accepted (the {"epoch","phase","state"} dict that `train` writes, a raw state dict, and a
`module.`-prefixed DDP state dict); the output name uses the image's basename; inference runs in eval mode under
no_grad (set TORCHSR_TEST_TRAIN_MODE=1 to reproduce the reference's train-mode BatchNorm behaviour)."""
import os
from argparse import Namespace

import torch


def load_generator_state(path: str, device) -> dict:
    ck = torch.load(path, map_location=device)
    state = ck["state"] if isinstance(ck, dict) and "state" in ck else ck
    return {(k[len("module."):] if k.startswith("module.") else k): v for k, v in state.items()}


def upscale(generator, low_res: torch.Tensor, train_mode: bool = False) -> torch.Tensor:
    generator.train(train_mode)
    with torch.no_grad():
        return generator(low_res)


def to_uint8_image(img: torch.Tensor) -> torch.Tensor:
    """[0,1] float image -> 8-bit, with the arithmetic of torchvision.utils.save_image (the reference's last step,
    test.py:62): mul(255).add_(0.5).clamp_(0, 255).to(uint8). In place on `img`, which must be a temporary."""
    return img.mul_(255.0).add_(0.5).clamp_(0.0, 255.0).to(torch.uint8)


def upscale_pipelined(generator, inputs, outputs, train_mode: bool = False) -> None:
    """Throughput path for a folder / stream of images (the loop of `_test`, */trainer.py:282-286, over many images):
    `inputs` are host tensors [B,3,H,W] (pinned for real overlap), `outputs` pre-allocated host tensors [B,3,4H,4W].
    The host->device copy of image i+1 and the device->host copy of image i-1 run on their own streams while image i
    is being computed; every image still takes the same three steps in order. Returns once all outputs are on the host.
    uint8 `outputs` receive the 8-bit image `torchvision.utils.save_image` would write (quantised on the device, a
    quarter of the bytes over PCIe); float outputs receive the raw generator output."""
    dev = next(generator.parameters()).device
    cur = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    generator.train(train_mode)
    with torch.no_grad():
        for x_h, y_h in zip(inputs, outputs):
            with torch.cuda.stream(s_in):
                x = x_h.to(dev, non_blocking=True)
            cur.wait_stream(s_in)
            x.record_stream(cur)
            y = generator(x)
            if y_h.dtype == torch.uint8:
                y = to_uint8_image(y)
            s_out.wait_stream(cur)
            with torch.cuda.stream(s_out):
                y_h.copy_(y, non_blocking=True)
            y.record_stream(s_out)
    cur.wait_stream(s_out)
    s_out.synchronize()


def test(args: Namespace, model_class, device) -> str:
    from PIL import Image
    from torchvision import utils
    from torchvision.transforms import ToTensor
    generator = model_class().to(device)
    ckpt = f'{args.model.lower()}-gan-best.pth'
    generator.load_state_dict(load_generator_state(ckpt, device))
    image = ToTensor()(Image.open(args.image).convert('RGB')).unsqueeze(0).to(device)
    super_res = upscale(generator, image, train_mode=os.environ.get("TORCHSR_TEST_TRAIN_MODE") == "1")
    out = f'upres-{os.path.basename(args.image)}'
    utils.save_image(super_res, out)
    return out
