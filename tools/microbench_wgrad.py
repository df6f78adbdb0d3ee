"""Weight-gradient kernel: per-launch time of the grid shape the host heuristic picks (csrc/api.cu:build_wgrad) against
pinned alternatives (TSR_WGRAD_GPC / TSR_WGRAD_SPLITS), for the layer shapes of the training step.
Usage: python tools/microbench_wgrad.py [sweep]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from torchsr_b200 import ops  # noqa: E402


def time_prog(descs, reps=40):
    prog = ops.Program()
    for _ in range(reps):
        for d in descs:
            prog.add(d)
    prog.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    prog.run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def case(B, H, W, Cin, Cout, stride=1, **env):
    for k in ("TSR_WGRAD_GPC", "TSR_WGRAD_SPLITS", "TSR_WGRAD_PIX"):
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ[k] = str(v)
    x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
    geom = ops.fwd_geometry(H, W, 3, 3, 1, 1, stride)
    Ho, Wo = geom["Ho"], geom["Wo"]
    dy = torch.randn(B, Ho, Wo, Cout, device="cuda").to(torch.bfloat16)
    acc = torch.zeros(9 * Cin * Cout, device="cuda")
    d = ops.wgrad_desc(x=x, N=B, H=H, W=W, C=Cin, x_ld=Cin, geom=geom, dy=dy, dy_ld=Cout, dy_c=Cout, out=acc,
                       cout_valid=Cout, block_n=min(Cout, 128))
    us = time_prog([d])
    gf = 2.0 * B * Ho * Wo * Cin * Cout * 9 / 1e9
    return us, gf / us * 1e3


SHAPES = [  # B, H, W, Cin, Cout, stride
    (16, 24, 24, 64, 64, 1), (64, 24, 24, 64, 64, 1), (16, 24, 24, 64, 256, 1), (16, 48, 48, 64, 256, 1),
    (32, 96, 96, 64, 64, 2), (32, 48, 48, 64, 128, 1), (32, 48, 48, 128, 128, 2), (32, 24, 24, 128, 256, 1),
    (32, 24, 24, 256, 256, 2), (32, 12, 12, 256, 512, 1), (32, 12, 12, 512, 512, 2),
    (128, 96, 96, 64, 64, 2), (128, 24, 24, 256, 256, 2),
]


OLD = {(16, 24, 24, 64, 64, 1): (1, 36), (64, 24, 24, 64, 64, 1): (3, 144)}   # grids of the previous heuristic (ncu lists)


def main():
    sweep = len(sys.argv) > 1
    for sh in SHAPES:
        us, tf = case(*sh)
        line = f"{sh}: default {us:7.1f} us {tf:6.1f} TFLOP/s"
        old = OLD.get(sh)
        if old:
            u2, _ = case(*sh, TSR_WGRAD_GPC=old[0], TSR_WGRAD_SPLITS=old[1])
            line += f" | round-1 grid (gpc={old[0]}, splits={old[1]}): {u2:6.1f}"
        if sweep:
            for gpc in (1, 2, 3, 5):
                try:
                    u2, _ = case(*sh, TSR_WGRAD_GPC=gpc)
                    line += f" | gpc={gpc}: {u2:6.1f}"
                except Exception:  # noqa: BLE001
                    line += f" | gpc={gpc}: n/a"
        print(line, flush=True)


if __name__ == "__main__":
    os.environ["TSR_CONV_VERBOSE"] = os.environ.get("TSR_CONV_VERBOSE", "0")
    main()
