#!/usr/bin/env python
"""Benchmark of the SRGAN generator+discriminator GAN training step (BASELINE.json configs[1]):
batch 16 per GPU of synthetic 96x96 HR / 24x24 LR crops, bf16 kernels with fp32 accumulation, random-init weights.

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun launches it for N > 1)
    python bench.py --impl reference --gpus N ...            # the reference algorithm's CPU port on the host cores

One "step" = one full SRGANTrainer._gan_loop: G forward, D(real) + D(fake) forward/backward + Adam on D, VGG
perceptual loss + D(sr) forward/backward + G backward + Adam on G (reference torchsr/srgan/trainer.py:416-469).
Rank 0 prints ONE JSON line. `value` is whole-job crops/s with inputs resident in HBM; `e2e` the same through the
public trainer API from pinned host batches with the loss read back every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from argparse import Namespace

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_PER_CROP_GD = 21.73          # G+D only, algorithmic minimum, SRGAN GAN step per 96^2 crop
GFLOP_PER_CROP_VGG = 21.50         # frozen VGG19 loss: 2 forwards + 1 data gradient
TRUNK_CONV_GFLOP_PER_CROP = 0.04247  # one 64->64 3x3 conv at 24x24


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="crops per GPU per step (configs[1]: 16; configs[2]: 64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-vgg", action="store_true", help="MSE content loss instead of VGG (NOT the headline config)")
    ap.add_argument("--eager", action="store_true", help="call _gan_loop eagerly instead of trainer.graph_step")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(burst=p["bf16_tflops"], sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    hbm=p["hbm_gbs"], source="measured")
    except Exception:  # noqa: BLE001
        return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """Samples SM clock and throttle reasons with nvidia-smi while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [x.strip() for x in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def synthetic_batch(batch, seed, pinned=False):
    import torch
    g = torch.Generator().manual_seed(seed)
    lr = torch.rand(batch, 3, 24, 24, generator=g)
    hr = torch.rand(batch, 3, 96, 96, generator=g)
    if pinned:
        lr, hr = lr.pin_memory(), hr.pin_memory()
    return lr, hr


# ---------------------------------------------------------------------------------------------- reference arm / cpu
def cpu_port_crops_per_sec(batch, steps, warmup, use_vgg=True, threads=None):
    """Times the oracle port of the reference `_gan_loop` on the host cores (fp32, torch CPU operators)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import step_oracle as S
    from torchsr_b200.srgan.discriminator import Discriminator
    from torchsr_b200.srgan.generator import Generator
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(1234)
    G, D = Generator(), Discriminator()       # parameter containers only: same default init as the reference classes
    o = S.OracleSRGAN(G.state_dict(), D.state_dict(), S.vgg19_features() if use_vgg else None)
    lr, hr = synthetic_batch(batch, 1234)
    for _ in range(warmup):
        o.gan_step(lr, hr)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.gan_step(lr, hr)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = min(args.steps, 4), min(args.warmup, 1)
    cps, spstep, threads = cpu_port_crops_per_sec(args.batch, steps, warmup, not args.no_vgg)
    sample = f"{steps} timed + {warmup} warm-up steps of batch {args.batch} (--steps/--warmup capped to keep the run short)"
    line = {
        "impl": "reference", "metric": "SRGAN GAN training crops/sec (96x96 HR)", "value": cps, "unit": "crops/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": spstep * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "SRGAN G+D _gan_loop, batch %d of 96x96 HR / 24x24 LR crops, VGG19 loss %s" %
                   (args.batch, "off (MSE)" if args.no_vgg else "on (random-init weights)")},
        "cpu_baseline": {"value": cps, "unit": "crops/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": cps, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "oracle port of the reference algorithm (oracle/step_oracle.py) on torch CPU operators; the unmodified "
                "reference lives in /root/reference, which does not exist on the GPU box",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- our arm
def trunk_conv_roofline(batch, pk):
    """Isolated timing of the dominant kernel (conv_igemm on the 64->64 3x3 trunk shape of this workload)."""
    import torch
    from torchsr_b200 import ops
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    B, H, W, C = batch, 24, 24, 64
    x = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
    w = (torch.randn(9, C, C, device="cuda") * 0.05).to(torch.bfloat16)
    out = torch.empty(B, H, W, C, device="cuda", dtype=torch.bfloat16)
    stats = torch.empty(((B * H * W + 127) // 128), C, 2, device="cuda")
    d = ops.conv_desc(x=x, N=B, H=H, W=W, C=C, x_ld=C, geom=ops.fwd_geometry(H, W, 3, 3, 1, 1, 1), w=w, cout_pad=C,
                      w_ld=C, n_slots=9, block_n=64, out=out, os_n=H * W * C, os_h=W * C, os_w=C, n_valid=C,
                      stats_partial=stats, stats_ld=C)
    prog = ops.Program()
    n = 200
    for _ in range(n):
        prog.add(d)
    prog.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    prog.run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    flops = TRUNK_CONV_GFLOP_PER_CROP * 1e9 * batch
    ach = flops / (us * 1e-6) / 1e12
    return {"kernel": "conv_igemm_kernel (3x3 64->64 @24x24, B=%d, back-to-back launches)" % batch, "us_per_launch": us,
            "achieved": ach, "peak": pk["burst"], "unit": "TFLOP/s", "frac": ach / pk["burst"], "bound": "tensor"}


# algorithmic GFLOP per crop executed by conv_igemm_kernel in one step (forward + data-gradient convs and the Linear
# GEMMs; weight gradients run in conv_wgrad_kernel): G 2.5553 + 2.5374, D 3 x 1.7683 forward + 1.7365 (real) + 1.7365
# (fake) + 1.7683 (super-res pass, full data gradient), VGG19 3 x 7.166
CONV_IGEMM_GFLOP_PER_CROP_GD = 15.639
CONV_IGEMM_GFLOP_PER_CROP_VGG = 21.50


def conv_kernel_roofline(trainer, lr_d, hr_d, batch, pk, vgg_ours):
    """The dominant kernel (conv_igemm_kernel with its persistent / grouped variants: the largest share of the step's launch time, profiles/r01d_ncu_launches_step.csv)
    over ALL of its launches in one training step: every distinct conv / GEMM descriptor of the programs a step
    executes is replayed 20x back to back (CUDA events on the launching stream) and weighted by how often the step
    runs it; achieved = algorithmic FLOPs of those launches / summed launch time."""
    import torch
    from torchsr_b200 import _lib as L
    from torchsr_b200 import ops
    counts = {}
    orig = ops.Program.run

    def counting(self, first=0, count=-1, stream=None):
        if first == 0:
            counts[id(self)] = (self, counts.get(id(self), (self, 0))[1] + 1)
        return orig(self, first, count, stream)

    ops.Program.run = counting
    try:
        trainer._gan_loop(lr_d, hr_d, 0)
        torch.cuda.synchronize()
    finally:
        ops.Program.run = orig
    cache, total_us, launches, reps = {}, 0.0, 0, 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for prog, c in counts.values():
        for d in prog.descs:
            if isinstance(d, ops.ConvGroupDesc):        # four parity-class convs in one launch
                key = b"".join(bytes(m) for m in d.members)
            elif isinstance(d, L.ConvDesc):
                key = bytes(d)
            else:
                continue
            if key not in cache:
                p2 = ops.Program()
                for _ in range(reps):
                    if isinstance(d, ops.ConvGroupDesc):
                        p2.add_group(d.members)
                    else:
                        p2.add(d)
                p2.run()
                p2.run()
                e0.record()
                p2.run()
                p2.run()
                e1.record()
                torch.cuda.synchronize()
                cache[key] = e0.elapsed_time(e1) * 1e3 / (2 * reps)
            total_us += cache[key] * c
            launches += c
    gflop = (CONV_IGEMM_GFLOP_PER_CROP_GD + (CONV_IGEMM_GFLOP_PER_CROP_VGG if vgg_ours else 0.0)) * batch
    ach = gflop * 1e9 / (total_us * 1e-6) / 1e12
    return {"kernel": "conv_igemm_kernel, all %d launches of one step (%d distinct descriptors)" % (launches, len(cache)),
            "us_per_step": total_us, "us_per_launch": total_us / max(launches, 1), "gflop_per_step": gflop,
            "achieved": ach, "peak": pk["burst"], "unit": "TFLOP/s", "frac": ach / pk["burst"], "bound": "tensor"}


def run_b200(args):
    import torch
    import torch.distributed as dist
    os.environ.setdefault("TORCHSR_VGG_WEIGHTS", "random")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl b200) needs a CUDA device: there is no CPU fallback for the kernels")
    torch.cuda.set_device(local)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from torchsr_b200 import _lib as L
    from torchsr_b200.srgan.trainer import SRGANTrainer

    torch.manual_seed(1234)           # identical initial weights on every rank (attach() broadcasts anyway)
    targs = Namespace(disable_amp=False, batch_size=args.batch, epochs=8, pretrain_epochs=1, gan_checkpoint=None,
                      psnr_checkpoint=None, skip_image_save=True, local_rank=local, rank=rank if distributed else -1,
                      world_size=world)
    trainer = SRGANTrainer(torch.device("cuda"), targs, [], [], 0, 0, distributed)
    if args.no_vgg:
        trainer.vgg_loss = lambda a, b: torch.nn.functional.mse_loss(a, b)
    lr_h, hr_h = synthetic_batch(args.batch, 1234 + rank, pinned=True)
    lr_d, hr_d = lr_h.cuda(), hr_h.cuda()

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = L.launch_count()
        e0.record()
        for s in range(steps):
            fn(s)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if distributed:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), L.launch_count() - l0

    # public step API: trainer.graph_step replays the whole _gan_loop as one CUDA graph (same arithmetic as the eager
    # call, see srgan/trainer.py); --eager times the plain Python call instead
    step_fn = trainer._gan_loop if args.eager else trainer.graph_step
    for s in range(max(args.warmup, 3)):
        step_fn(lr_d, hr_d, s)
    # kernels of ours per step (graph replays do not pass through the library's launch counter)
    l0 = L.launch_count()
    trainer._gan_loop(lr_d, hr_d, 0)
    launches_per_step = L.launch_count() - l0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, _ = timed(lambda s: step_fn(lr_d, hr_d, s), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = launches_per_step * args.steps
    # end to end: pinned host batch -> H2D inside the step, loss read back to the host every step
    ms_e2e, _ = timed(lambda s: step_fn(lr_h, hr_h, s).item(), args.steps)
    crops = args.batch * world * args.steps
    value = crops / (ms * 1e-3)
    e2e = crops / (ms_e2e * 1e-3)
    def leave():
        # captured NCCL work is still referenced by the step graph: synchronise, meet the other ranks, and leave
        # without tearing the communicator down underneath it
        sys.stdout.flush()
        if distributed:
            torch.cuda.synchronize()
            dist.barrier()
            os._exit(0)

    if rank != 0:
        leave()
        return
    pk = peaks()
    ach = value * GFLOP_PER_CROP_GD / 1e3 / world     # TFLOP/s per GPU on the G+D algorithmic FLOPs
    kern = trunk_conv_roofline(args.batch, pk)
    vgg_ours = not args.no_vgg and os.environ.get("TORCHSR_VGG_IMPL", "b200") != "torch"
    if distributed:
        dom = {"note": "measured at N=1 only (replaying a step on rank 0 alone would leave its collectives unmatched)"}
    else:
        try:
            dom = conv_kernel_roofline(trainer, lr_d, hr_d, args.batch, pk, vgg_ours)
        except Exception as exc:  # noqa: BLE001
            dom = {"error": repr(exc)}
    line = {
        "metric": "SRGAN GAN training crops/sec (96x96 HR)", "value": value, "unit": "crops/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "SRGAN G+D _gan_loop (BASELINE configs[1]), batch %d per GPU of 96x96 HR / 24x24 LR "
                               "crops, random-init weights, VGG19 loss %s" %
                               (args.batch, "replaced by MSE (--no-vgg)" if args.no_vgg else
                                ("on this repo's kernels (nets.define_vgg)" if os.environ.get("TORCHSR_VGG_IMPL", "b200") != "torch"
                                 else "executed by PyTorch/cuDNN under bf16 (TORCHSR_VGG_IMPL=torch)")),
                   "parallelism": f"dp{world}", "global_batch": args.batch * world,
                   "step_api": "SRGANTrainer._gan_loop (eager)" if args.eager else
                               "SRGANTrainer.graph_step (whole step replayed as one CUDA graph)",
                   "l2": "per-step working set (fp32 weights + Adam state + activations, > 0.5 GB) exceeds the 126 MB "
                         "L2; no explicit flush between steps"},
        "e2e": {"value": e2e, "unit": "crops/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int((lr_h.numel() + hr_h.numel()) * 4 * world), "d2h_bytes_per_step": 4 * world},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": ach, "peak": pk["sustained"], "unit": "TFLOP/s",
                     "frac": ach / pk["sustained"],
                     # dram__bytes_read + dram__bytes_write of one captured launch of the dominant kernel (persistent
                     # 3x3 64->256 conv at 48x48, B=16: profiles/r01d_ncu_full_conv_summary.csv, ID 0); its operands are
                     # 4.7 MB of activations + 0.3 MB of weights read, the 18.9 MB output stays in the 126 MB L2
                     "traffic": 5.12e6,
                     "what": "whole step: crops/s x 21.73 GFLOP/crop (G+D algorithmic minimum, SURVEY 8d) per GPU vs the "
                             f"{pk['source']} sustained bf16 peak; with the VGG19 FLOPs (+21.50/crop) "
                             f"the step sustains {value * (GFLOP_PER_CROP_GD + GFLOP_PER_CROP_VGG) / 1e3 / world:.1f} TFLOP/s",
                     "dominant_kernel": dom, "trunk_conv_only": kern,
                     "ncu": "profiles/r01d_ncu_full_conv_summary.csv (dram bytes, tensor-pipe activity, L2->SM bytes per "
                            "launch; r01c_* hold the captures before the lean main loops), r01d_ncu_launches_step.csv "
                            "(every launch of one step), r01d_conv_attribution.md (what bounds the kernel)"},
    }
    if not args.no_cpu_baseline and world == 1:
        cps, spstep, threads = cpu_port_crops_per_sec(args.batch, 3, 1, not args.no_vgg)
        line["cpu_baseline"] = {"value": cps, "unit": "crops/s", "cores": threads, "kind": "port",
                                "sample": f"3 timed + 1 warm-up steps of batch {args.batch} of the oracle port "
                                          "(oracle/step_oracle.py), fp32, all host threads"}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    leave()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
