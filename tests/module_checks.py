"""Whole-module parity checks against the CPU oracle (oracle/torchsr_oracle.py) on identical fp32 weights.
Returns dicts of rel-L2 errors. Used by tests/test_modules_gpu.py and tools/diag_modules.py."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import torchsr_oracle as O  # noqa: E402

DEV = "cuda"


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


def randomize_bn(module, seed=0):
    """Non-trivial BatchNorm affine parameters and PReLU slopes so that their gradients are exercised."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in module.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 0.5 + 0.75)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
            if isinstance(m, torch.nn.PReLU):
                m.weight.copy_(torch.rand(1, generator=g) * 0.3 + 0.1)


def grad_report(module, ref_grads, prefix=""):
    """rel-L2 per parameter between module.grad and the oracle's gradient; returns (worst, median, dict)."""
    errs = {}
    for name, p in module.named_parameters():
        ref = ref_grads[name]
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        errs[prefix + name] = rel_l2(got, ref)
    vals = sorted(errs.values())
    return vals[-1], vals[len(vals) // 2], errs


def check_module(module_cpu, oracle_fn, x, gout_seed=1, train=True, input_grad=False):
    """Runs `module` (moved to the GPU) and the oracle on the same state dict and input; compares the output, the
    parameter gradients for a random upstream gradient, the input gradient and the BatchNorm buffers."""
    sd = {k: v.clone() for k, v in module_cpu.state_dict().items()}
    m = module_cpu.to(DEV)
    m.train(train)
    xg = x.detach().clone().to(DEV).requires_grad_(input_grad)
    y = m(xg)
    g = torch.Generator().manual_seed(gout_seed)
    gout = torch.randn(y.shape, generator=g)
    y.backward(gout.to(DEV))
    if DEV == "cuda":
        torch.cuda.synchronize()
    # oracle
    osd = O.with_grad(sd)
    xr = x.detach().clone().requires_grad_(input_grad)
    buffers = {}
    yr = oracle_fn(osd, xr, train, buffers) if buffers is not None else oracle_fn(osd, xr)
    yr.backward(gout)
    ref_grads = {k.lstrip("."): v.grad for k, v in osd.items() if v.requires_grad}
    worst, median, errs = grad_report(m, ref_grads)
    r = {"out": rel_l2(y, yr), "grad_worst": worst, "grad_median": median}
    if input_grad:
        r["dx"] = rel_l2(xg.grad, xr.grad)
    new_sd = m.state_dict()
    bn_err = 0.0
    for k, v in buffers.items():
        k = k.lstrip(".")
        if "num_batches" in k:
            bn_err = max(bn_err, float(abs(int(new_sd[k]) - int(v))))
        else:
            bn_err = max(bn_err, rel_l2(new_sd[k], v))
    r["bn_buffers"] = bn_err
    return r, errs
