"""VGG content-loss gradient: this repo's kernels and PyTorch bf16 (cuDNN) against PyTorch fp32 on the CPU."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["TORCHSR_VGG_WEIGHTS"] = "random"
import torch  # noqa: E402
import module_checks as MC  # noqa: E402
from torchsr_b200.srgan.loss import VGGLoss  # noqa: E402

torch.manual_seed(12)
x, t = torch.rand(2, 3, 96, 96), torch.rand(2, 3, 96, 96)
ref_mod = VGGLoss()
xr = x.clone().requires_grad_(True)
ref_mod(xr, t).backward()
grads = {}
for impl in ("torch", "b200"):
    os.environ["TORCHSR_VGG_IMPL"] = impl
    mod = VGGLoss().cuda()
    xg = x.cuda().requires_grad_(True)
    mod(xg, t.cuda()).backward()
    torch.cuda.synchronize()
    grads[impl] = xg.grad.cpu()
    cos = torch.nn.functional.cosine_similarity(grads[impl].flatten(), xr.grad.flatten(), dim=0)
    print(impl, "dx rel-L2 vs fp32:", MC.rel_l2(grads[impl], xr.grad), "cos", float(cos))
print("b200 vs torch-bf16:", MC.rel_l2(grads["b200"], grads["torch"]))
# fp32 on the GPU with TF32 off, to separate bf16 effects from implementation differences
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
f32 = VGGLoss().features.cuda()
xg = x.cuda().requires_grad_(True)
torch.nn.functional.l1_loss(f32(xg), f32(t.cuda()).detach()).backward()
print("gpu fp32 vs cpu fp32:", MC.rel_l2(xg.grad.cpu(), xr.grad))
