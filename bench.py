#!/usr/bin/env python
"""Benchmark of the SRGAN generator+discriminator GAN training step (BASELINE.json configs[1]):
batch 16 per GPU of synthetic 96x96 HR / 24x24 LR crops, bf16 kernels with fp32 accumulation, random-init weights.

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun launches it for N > 1)
    python bench.py --impl reference --gpus N ...            # the UNMODIFIED reference on the host cores (oracle/_ref)

One "step" = one full SRGANTrainer._gan_loop: G forward, D(real | fake) forward/backward + Adam on D, VGG
perceptual loss + D(sr) forward/backward + G backward + Adam on G (reference torchsr/srgan/trainer.py:416-469).
Rank 0 prints ONE JSON line. `value` is whole-job crops/s with inputs resident in HBM; `e2e` the same through the
public trainer API from pinned host batches with the loss read back every step. Further blocks of the same line
(everything BASELINE.json's metric and configs name): `b64` (configs[2], every N), `dp_parity` (N > 1), and at N = 1
`inference` (configs[4], output Mpx/s), `esrgan` (configs[3] shape), `hbm_kernels` (achieved GB/s of the HBM-bound
kernels), `gpu_eager_baseline` (the reference's own modules through PyTorch/cuDNN on the same B200) and `cpu_baseline`.
"""
import argparse
import glob
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
GFLOP_PER_MPX_INFER = 277.3        # SRGAN G inference per output Mpx
GFLOP_PER_CROP_ESRGAN = 141.1      # ESRGAN GAN step per 128^2 crop, G+D algorithmic minimum


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="crops per GPU per step (configs[1]: 16)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-vgg", action="store_true", help="MSE content loss instead of VGG (NOT the headline config)")
    ap.add_argument("--eager", action="store_true", help="call _gan_loop eagerly instead of trainer.graph_step")
    ap.add_argument("--only", default="", help="comma list restricting the extra blocks: b64,inference,esrgan,hbm,"
                                               "eager_baseline,data_pipeline,cpu,roofline,dp_parity (default: all)")
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
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
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


def synthetic_batch(batch, seed, pinned=False, lr_size=24, scale=4):
    import torch
    g = torch.Generator().manual_seed(seed)
    lr = torch.rand(batch, 3, lr_size, lr_size, generator=g)
    hr = torch.rand(batch, 3, lr_size * scale, lr_size * scale, generator=g)
    if pinned:
        lr, hr = lr.pin_memory(), hr.pin_memory()
    return lr, hr


def want(args, block: str) -> bool:
    return not args.only or block in args.only.split(",")


# ---------------------------------------------------------------------------------------------- reference arm / cpu
def _ref_harness():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_harness
    return ref_harness


def cpu_reference_crops_per_sec(batch, steps, warmup, threads=None):
    """Times the reference's own SRGANTrainer._gan_loop (oracle/_ref, unmodified) on the host cores, fp32, all
    threads - or, when oracle/_ref was not staged, the oracle port of it (oracle/step_oracle.py). Returns
    (crops/s, s/step, threads, kind)."""
    import torch
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    lr, hr = synthetic_batch(batch, 1234)
    H = _ref_harness()
    if H.available():
        torch.manual_seed(1234)
        tr = H.reference_trainer("srgan", torch.device("cpu"), batch)
        step, kind = (lambda: tr._gan_loop(lr, hr, 0)), "reference"
    else:
        import step_oracle as S
        from torchsr_b200.srgan.discriminator import Discriminator
        from torchsr_b200.srgan.generator import Generator
        torch.manual_seed(1234)
        G, D = Generator(), Discriminator()       # parameter containers only: same default init as the reference
        o = S.OracleSRGAN(G.state_dict(), D.state_dict(), S.vgg19_features())
        step, kind = (lambda: o.gan_step(lr, hr)), "port"
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, threads, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = min(args.steps, 6), min(args.warmup, 1)
    cps, spstep, threads, kind = cpu_reference_crops_per_sec(args.batch, steps, warmup)
    what = ("the unmodified reference (oracle/_ref/torchsr, staged by oracle/make_ref.py): SRGANTrainer._gan_loop on "
            "torch.device('cpu')") if kind == "reference" else \
        "oracle port of the reference step (oracle/step_oracle.py): oracle/_ref was not staged"
    sample = f"{steps} timed + {warmup} warm-up steps of batch {args.batch} (--steps/--warmup capped to keep the run short)"
    line = {
        "impl": "reference", "metric": "SRGAN GAN training crops/sec (96x96 HR)", "value": cps, "unit": "crops/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": spstep * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "SRGAN G+D _gan_loop (BASELINE configs[1]), batch %d per GPU of 96x96 HR / 24x24 LR crops, "
                               "random-init weights, VGG19 loss on (random-init weights)" % args.batch},
        "cpu_baseline": {"value": cps, "unit": "crops/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": cps, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": what,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- kernel timing helpers
def _time_descs(descs, reps=20):
    """Average microseconds per launch of a descriptor (or conv group) replayed back to back (CUDA events)."""
    import torch
    from torchsr_b200 import ops
    p2 = ops.Program()
    for _ in range(reps):
        if isinstance(descs, ops.ConvGroupDesc):
            p2.add_group(descs.members)
        else:
            p2.add(descs)
    p2.run()
    p2.run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    p2.run()
    p2.run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (2 * reps)


def _programs_of_one_step(trainer, lr_d, hr_d):
    """{id: (Program, runs per step)} of every launch list one eager _gan_loop executes."""
    import torch
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
    return counts


def trunk_conv_roofline(batch, pk):
    """Isolated timing of the most frequent conv shape (3x3 64->64 at 24x24) with plain column statistics."""
    import torch
    from torchsr_b200 import ops
    B, H, W, C = batch, 24, 24, 64
    x = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
    w = (torch.randn(9, C, C, device="cuda") * 0.05).to(torch.bfloat16)
    out = torch.empty(B, H, W, C, device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros(C, 2, device="cuda")
    d = ops.conv_desc(x=x, N=B, H=H, W=W, C=C, x_ld=C, geom=ops.fwd_geometry(H, W, 3, 3, 1, 1, 1), w=w, cout_pad=C,
                      w_ld=C, n_slots=9, block_n=32, out=out, os_n=H * W * C, os_h=W * C, os_w=C, n_valid=C,
                      stats_partial=stats, stats_ld=C)
    us = _time_descs(d, reps=100)
    flops = TRUNK_CONV_GFLOP_PER_CROP * 1e9 * batch
    ach = flops / (us * 1e-6) / 1e12
    return {"kernel": "conv_igemm_kernel (3x3 64->64 @24x24, B=%d, back-to-back launches)" % batch, "us_per_launch": us,
            "achieved": ach, "peak": pk["burst"], "unit": "TFLOP/s", "frac": ach / pk["burst"], "bound": "tensor"}


# algorithmic GFLOP per crop executed by conv_igemm_kernel in one step (forward + data-gradient convs and the Linear
# GEMMs; weight gradients run in conv_wgrad_kernel): G 2.5553 + 2.5374, D 3 x 1.7683 forward + 1.7365 (real) + 1.7365
# (fake) + 1.7683 (super-res pass, full data gradient), VGG19 3 x 7.166
CONV_IGEMM_GFLOP_PER_CROP_GD = 15.639
CONV_IGEMM_GFLOP_PER_CROP_VGG = 21.50


def conv_kernel_roofline(counts, batch, pk, vgg_ours):
    """The dominant kernel (conv_igemm_kernel with its persistent / grouped variants: the largest share of the step's
    launch time, see the committed ncu launch list) over ALL of its launches in one training step: every distinct
    conv / GEMM descriptor of the programs a step executes is replayed back to back (CUDA events on the launching
    stream) and weighted by how often the step runs it; achieved = algorithmic FLOPs of those launches / summed time."""
    from torchsr_b200 import _lib as L
    from torchsr_b200 import ops
    cache, total_us, launches = {}, 0.0, 0
    for prog, c in counts.values():
        for d in prog.descs:
            if isinstance(d, ops.ConvGroupDesc):        # four parity-class convs in one launch
                key = b"".join(bytes(m) for m in d.members)
            elif isinstance(d, L.ConvDesc):
                key = bytes(d)
            else:
                continue
            if key not in cache:
                if isinstance(d, L.ConvDesc) and d.bnf_mode == 1:
                    # fused training BatchNorm: the launch needs its arrival counters zero - time it with the zero
                    # kernel it follows in the real program, and subtract that kernel's own replay time
                    cache[key] = _time_fused_bn_conv(d)
                else:
                    cache[key] = _time_descs(d)
            total_us += cache[key] * c
            launches += c
    gflop = (CONV_IGEMM_GFLOP_PER_CROP_GD + (CONV_IGEMM_GFLOP_PER_CROP_VGG if vgg_ours else 0.0)) * batch
    ach = gflop * 1e9 / (total_us * 1e-6) / 1e12
    return {"kernel": "conv_igemm_kernel, all %d launches of one step (%d distinct descriptors)" % (launches, len(cache)),
            "us_per_step": total_us, "us_per_launch": total_us / max(launches, 1), "gflop_per_step": gflop,
            "achieved": ach, "peak": pk["burst"], "unit": "TFLOP/s", "frac": ach / pk["burst"], "bound": "tensor"}


def _time_fused_bn_conv(d, reps=20):
    """A conv with the fused training-mode BatchNorm crosses a grid barrier on counters that must be zero at launch:
    replay (zero counters, conv) pairs and subtract the replay time of the zero kernel alone."""
    import torch
    from torchsr_b200 import _lib as L
    from torchsr_b200 import ops
    scratch = torch.zeros(64 + 2 * 2 * d.stats_ld, dtype=torch.float32, device="cuda")
    d2 = type(d).from_buffer_copy(bytes(d))
    d2.bnf_counter = ops.ptr(scratch)
    d2.stats_partial = ops.ptr(scratch, 64)
    d2.bnf_rm = d2.bnf_rv = d2.bnf_nbt = 0          # keep the module's running statistics untouched
    z = ops.elt(L.E_ZERO, p=[scratch], i=[scratch.numel() * 4])
    both, only_z = ops.Program(), ops.Program()
    for _ in range(reps):
        both.add(z)
        both.add(d2)
        only_z.add(z)
    out = []
    for p in (both, only_z):
        p.run()
        p.run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        p.run()
        p.run()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) * 1e3 / (2 * reps))
    return max(out[0] - out[1], 0.1)


def hbm_kernels(counts, trainer, pk):
    """Achieved GB/s of the HBM/L2-bound kernels of one step against the measured copy bandwidth: BatchNorm apply
    (forward, where it is not fused into the conv), BatchNorm backward apply, the Linear weight gradient and the fused
    Adam + re-pack. Bytes are algorithmic (each operand once); every distinct descriptor is replayed back to back."""
    import torch
    from torchsr_b200 import _lib as L
    from torchsr_b200 import ops
    acc = {}
    cache = {}
    for prog, c in counts.values():
        for d in prog.descs:
            if not isinstance(d, L.EltDesc):
                continue
            i, k = d.i, d.kind
            if k == L.E_BN_ACT:
                name, nbytes = "bn_act_kernel", i[0] * i[1] * 2 * (3 if d.p[3] else 2)
            elif k == L.E_BN_BWD_APPLY:
                need_x = bool(i[6]) or (i[2] != L.ACT_NONE and not i[8])
                name, nbytes = "bn_bwd_apply_kernel", i[0] * i[1] * 2 * (2 + (1 if need_x else 0) + (1 if d.p[6] else 0))
            elif k == L.E_LINEAR_WGRAD:
                name, nbytes = "linear_wgrad_kernel", (i[0] * (i[1] + i[2]) + i[1] * i[2] + i[1]) * 4
            else:
                continue
            key = bytes(d)
            if key not in cache:
                cache[key] = _time_descs(d)
            a = acc.setdefault(name, [0, 0.0, 0.0])
            a[0] += c
            a[1] += nbytes * c
            a[2] += cache[key] * c
    out = []
    for name, (n, nbytes, us) in sorted(acc.items()):
        gbs = nbytes / (us * 1e-6) / 1e9
        out.append({"kernel": name, "launches_per_step": n, "bytes_per_step": int(nbytes), "us_per_step": us,
                    "achieved_gbs": gbs, "peak_gbs": pk["hbm"], "frac": gbs / pk["hbm"]})
    # fused Adam + pack: 7 fp32 words per parameter (p, g, m, v read; p, m, v written) + the bf16 operand copies
    for label, opt in (("adam_pack_kernel (discriminator)", trainer.disc_optimizer),
                       ("adam_pack_kernel (generator)", trainer.gen_optimizer)):
        params = [p for g in opt.param_groups for p in g["params"] if p.grad is not None]
        if not params:
            continue
        packs = 0
        st = next(iter(opt._tables[0].values()))
        for store in st["stores"]:
            for r in store.convs:
                if r.kind == "std":
                    packs += r.w_fwd.numel() * 2 + (r.w_t.numel() * 2 if r.need_dgrad else 0)
            for r in store.linears:
                packs += r.w_fwd.numel() * 2
        nbytes = sum(p.numel() for p in params) * 28 + packs
        for _ in range(2):
            opt.step()
            opt.join()
        torch.cuda.synchronize()
        # replayed as a CUDA graph, like inside graph_step: the eager call's host work (table lookup) would dominate
        # the smaller kernel; the late launch (big Linear weight, side stream) is part of the step
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            opt.step()
            opt.join()
        graph.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 5
        gbs = nbytes / (us * 1e-6) / 1e9
        out.append({"kernel": label, "launches_per_step": 1, "bytes_per_step": int(nbytes), "us_per_step": us,
                    "achieved_gbs": gbs, "peak_gbs": pk["hbm"], "frac": gbs / pk["hbm"]})
    return out


def traffic_from_profiles():
    """dram__bytes_read + dram__bytes_write per launch of the dominant kernel, read from the newest committed ncu
    `--set full` summary of the conv kernel (profiles/r*_ncu_full_conv*summary.csv, `ncu --page raw --csv` export: a
    header row, a units row, one row per captured launch; the first launch is used); (None, None) if there is none."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full_conv*summary.csv")))
    if not files:
        return None, None
    import csv
    path = files[-1]
    try:
        rows = list(csv.reader(open(path)))
        hdr, units, row = rows[0], rows[1], rows[2]
        ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        total = float(row[ri]) * scale.get(units[ri], 1.0) + float(row[wi]) * scale.get(units[wi], 1.0)
        name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        grid = row[hdr.index("Grid Size")] if "Grid Size" in hdr else "?"
        return total, os.path.relpath(path, ROOT) + " (first captured launch: " + name[:60] + ", grid " + grid + ")"
    except Exception:  # noqa: BLE001
        return None, None


# ---------------------------------------------------------------------------------------------- extra blocks
def make_trainer(cls, batch, distributed, rank, local, world):
    import torch
    torch.manual_seed(1234)           # identical initial weights on every rank (attach() broadcasts anyway)
    targs = Namespace(disable_amp=False, batch_size=batch, epochs=8, pretrain_epochs=1, gan_checkpoint=None,
                      psnr_checkpoint=None, skip_image_save=True, local_rank=local, rank=rank if distributed else -1,
                      world_size=world)
    return cls(torch.device("cuda"), targs, [], [], 0, 0, distributed)


def inference_block(pk, steps=10):
    """BASELINE configs[4]: x4 inference of a synthetic 512x512 LR image to 2048x2048, eval mode, no_grad (the
    reference's `_test` semantics, */trainer.py:282-286); output Mpx/s resident and end to end (pinned host image in,
    pinned host image out)."""
    import torch
    from torchsr_b200.srgan.generator import Generator
    torch.manual_seed(1234)
    G = Generator().cuda().eval()
    b, size = 1, 512
    x_h = torch.rand(b, 3, size, size).pin_memory()
    y_h = torch.empty(b, 3, 4 * size, 4 * size).pin_memory()
    x = x_h.cuda()
    with torch.no_grad():
        for _ in range(3):
            y = G(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            y = G(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        # end to end through the public throughput API: every image is copied in from pinned host memory, upscaled and
        # copied back to pinned host memory; the copies of neighbouring images overlap the compute (own streams)
        from torchsr_b200.test import upscale_pipelined
        y_hs = [y_h, torch.empty_like(y_h).pin_memory()]
        upscale_pipelined(G, [x_h] * 2, y_hs)
        torch.cuda.synchronize()
        e0.record()
        upscale_pipelined(G, [x_h] * steps, [y_hs[i & 1] for i in range(steps)])
        e1.record()
        torch.cuda.synchronize()
        ms_e2e = e0.elapsed_time(e1) / steps
        e0.record()
        for _ in range(steps):
            y = G(x_h.cuda(non_blocking=True))
            y_h.copy_(y, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms_e2e_serial = e0.elapsed_time(e1) / steps
        # the same loop delivering the 8-bit image the reference's test() writes (save_image quantisation on the device)
        y8 = [torch.empty(y_h.shape, dtype=torch.uint8).pin_memory() for _ in range(2)]
        upscale_pipelined(G, [x_h] * 2, y8)
        torch.cuda.synchronize()
        e0.record()
        upscale_pipelined(G, [x_h] * steps, [y8[i & 1] for i in range(steps)])
        e1.record()
        torch.cuda.synchronize()
        ms_e2e_u8 = e0.elapsed_time(e1) / steps
    mpx = b * (4 * size) ** 2 / 1e6
    tf = GFLOP_PER_MPX_INFER * mpx / ms
    out = {"workload": "SRGAN generator x4 inference (BASELINE configs[4]): batch %d of synthetic %dx%d LR -> %dx%d, eval "
                       "mode, no_grad, random-init weights" % (b, size, size, 4 * size, 4 * size),
           "value": mpx / (ms * 1e-3), "unit": "output Mpx/s", "ms_per_image": ms / b,
           "e2e": {"value": mpx / (ms_e2e * 1e-3), "unit": "output Mpx/s", "ms_per_image": ms_e2e / b,
                   "h2d_bytes_per_step": x_h.numel() * 4, "d2h_bytes_per_step": y_h.numel() * 4,
                   "api": "torchsr_b200.test.upscale_pipelined (copies of neighbouring images overlap the compute)",
                   "serial_ms_per_image": ms_e2e_serial / b,
                   "uint8_output": {"value": mpx / (ms_e2e_u8 * 1e-3), "ms_per_image": ms_e2e_u8 / b,
                                    "d2h_bytes_per_step": y_h.numel(),
                                    "what": "same API with uint8 host outputs: the 8-bit image of torchvision.utils."
                                            "save_image (reference test.py:62), quantised on the device"}},
           "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["sustained"], "unit": "TFLOP/s",
                        "frac": tf / pk["sustained"], "gflop_per_output_mpx": GFLOP_PER_MPX_INFER},
           "peak_memory_gib": torch.cuda.max_memory_allocated() / 2 ** 30}
    del G, x, y
    torch.cuda.empty_cache()
    return out


def data_pipeline_block(batch=16, steps=50):
    """training batches from the GPU input pipeline (gpu_data.py: RandomCrop + flips + Pillow-exact
    bicubic /4 on uint8 images resident in HBM, one launch per batch) - crops/s including the host-side index draw."""
    import torch
    from torchsr_b200 import gpu_data as GD
    g = torch.Generator().manual_seed(0)
    pool = GD.ImagePool([torch.randint(0, 256, (512, 512, 3), dtype=torch.uint8, generator=g) for _ in range(64)], "cuda")
    loader = GD.GpuTrainLoader(pool, 96, batch, multiplier=-(-steps * batch // 64) + 1, seed=1)
    it = iter(loader)
    for _ in range(3):
        next(it)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    for _ in range(steps):
        lr, hr = next(it)
        n += lr.shape[0]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"workload": "GpuTrainLoader: batch %d of 96x96 crops from 64 synthetic 512x512 uint8 images in HBM" % batch,
            "value": n / dt, "unit": "crops/s", "ms_per_batch": dt / steps * 1e3}


def esrgan_block(pk, batch=16, steps=8):
    """BASELINE configs[3] shape on one GPU: ESRGAN (23 RRDB) generator + 10-conv discriminator + VGG loss, 128x128 HR
    crops; one step = ESRGANTrainer._gan_loop through graph_step."""
    import torch
    from torchsr_b200.esrgan.trainer import ESRGANTrainer
    tr = make_trainer(ESRGANTrainer, batch, False, 0, torch.cuda.current_device(), 1)
    lr, hr = synthetic_batch(batch, 4321, lr_size=32)
    lr, hr = lr.cuda(), hr.cuda()
    for s in range(3):
        tr.graph_step(lr, hr, s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        loss = tr.graph_step(lr, hr, s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    tf = GFLOP_PER_CROP_ESRGAN * batch / ms
    out = {"workload": "ESRGAN (23 RRDB) G+D _gan_loop + VGG19 loss (BASELINE configs[3] shape), batch %d of 128x128 HR / "
                       "32x32 LR crops on 1 GPU, graph_step" % batch,
           "value": batch / (ms * 1e-3), "unit": "crops/s", "ms_per_step": ms, "loss": float(loss),
           "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["sustained"], "unit": "TFLOP/s",
                        "frac": tf / pk["sustained"], "gflop_per_crop": GFLOP_PER_CROP_ESRGAN}}
    del tr
    torch.cuda.empty_cache()
    return out


def gpu_eager_baseline(batch, steps=10):
    """The reference's own modules executed by PyTorch / cuDNN on this B200 (the number a reader compares
    against): the UNMODIFIED SRGANTrainer._gan_loop from oracle/_ref as shipped (fp32 with TF32 convs, no autocast in
    that loop) and with its generator / discriminator / VGG forwards under bf16 autocast; plus the same statement
    sequence with device-side labels captured into ONE CUDA graph (the unmodified loop builds its labels on the host,
    trainer.py:439-440, which stream capture rejects)."""
    import torch
    H = _ref_harness()
    if not H.available():
        return {"unavailable": "oracle/_ref not staged (python oracle/make_ref.py where /root/reference exists)"}
    out = {"what": "reference modules (oracle/_ref, unmodified) through PyTorch %s / cuDNN %s on the same GPU, batch %d"
                   % (torch.__version__, torch.backends.cudnn.version(), batch), "unit": "crops/s", "variants": {}}
    lr, hr = synthetic_batch(batch, 1234)
    lr, hr = lr.cuda(), hr.cuda()

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"value": batch / (ms * 1e-3), "ms_per_step": ms}

    def build(autocast):
        torch.manual_seed(1234)
        tr = H.reference_trainer("srgan", torch.device("cuda"), batch)
        for opt in (tr.disc_optimizer, tr.gen_optimizer):
            for g in opt.param_groups:
                g["capturable"] = True            # device-side step counters: needed by the graphed variant
        if autocast:
            for m in (tr.generator, tr.discriminator, tr.vgg_loss):
                fwd = m.forward

                def wrapped(*a, _f=fwd, **k):
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        y = _f(*a, **k)
                    return y.float()
                m.forward = wrapped
        return tr

    def capturable_step(tr):
        """trainer.py:435-469 statement for statement, labels created on the device."""
        real = torch.ones(batch, 1, device="cuda")
        fake = torch.zeros(batch, 1, device="cuda")
        tr.discriminator.zero_grad()
        sr = tr.generator(lr)
        d_loss = tr.bce_loss(tr.discriminator(hr), real) + tr.bce_loss(tr.discriminator(sr.detach()), fake)
        d_loss.backward()
        tr.disc_optimizer.step()
        tr.generator.zero_grad()
        g_loss = tr.vgg_loss(sr, hr.detach()) + 0.001 * tr.bce_loss(tr.discriminator(sr), real)
        g_loss.backward()
        tr.gen_optimizer.step()

    for name, autocast in (("fp32_tf32", False), ("bf16_autocast", True)):
        try:
            tr = build(autocast)
            out["variants"][name + "_eager"] = timed(lambda: tr._gan_loop(lr, hr, 0))
        except Exception as exc:  # noqa: BLE001
            out["variants"][name + "_eager"] = {"error": repr(exc)[:300]}
            continue
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    capturable_step(tr)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                capturable_step(tr)
            out["variants"][name + "_cuda_graph"] = timed(graph.replay)
        except Exception as exc:  # noqa: BLE001
            out["variants"][name + "_cuda_graph"] = {"error": repr(exc)[:300]}
        del tr
        torch.cuda.empty_cache()
    return out


def dp_parity(trainer, lr_d, hr_d):
    """Data-parallel correctness (reference */trainer.py:143-157 DDP semantics, BatchNorm local): the gradient every rank
    holds after the in-backward exchange (bucketed all-reduce; factor all-gather + local GEMM for the classifier weight)
    must equal the mean over ranks of the gradients the ranks computed on their own shards. Both sides come from the
    SAME backward execution (engine.Plan.capture_local snapshots the pre-exchange gradient): two separate executions of
    this bf16 network differ by ~1e-1 in the discriminator's gradient through atomics-order noise alone (rounding flips
    of near-zero pre-activations), which would mask the exchange error being measured. Returns the max over both
    modules and all ranks of the rel-L2 difference."""
    import torch
    import torch.distributed as dist
    from torchsr_b200 import losses
    world = dist.get_world_size()
    one = torch.ones((), device="cuda")
    worst = 0.0

    def flat_grads(module):
        return torch.cat([p.grad.detach().reshape(-1).float() for p in module.parameters()])

    def unpad(module, flat):
        """engine's flat gradient pads every parameter to a multiple of 4 elements: back to the dense order."""
        st = module._tsr["store"]
        return torch.cat([flat[st.offsets[id(p)]:st.offsets[id(p)] + p.numel()] for p in st.params])

    def d_loss():
        with torch.no_grad():
            sr = trainer.generator(lr_d)
        pr, pf = trainer.discriminator.forward_pair(hr_d, sr)
        return losses.bce(pr, 1.0, pf, 0.0)

    def g_loss():
        return losses.mse(trainer.generator(lr_d), hr_d)

    for module, loss_fn in ((trainer.discriminator, d_loss), (trainer.generator, g_loss)):
        cap = []
        module.zero_grad()
        loss = loss_fn()
        for pool in module._tsr["plans"].values():
            for plan in pool:
                plan.capture_local = cap
        loss.backward(one)
        for pool in module._tsr["plans"].values():
            for plan in pool:
                plan.capture_local = None
        assert len(cap) == 1, len(cap)
        local = unpad(module, cap[0])
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        expect = torch.stack(gathered).mean(0)
        got = flat_grads(module)
        err = ((got - expect).double().norm() / expect.double().norm().clamp_min(1e-30)).reshape(1).float()
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        worst = max(worst, float(err))
        module.zero_grad()
    return worst


# ---------------------------------------------------------------------------------------------- our arm
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
        from torchsr_b200.dist import init_process_group as tdist_init
        tdist_init(local)
    from torchsr_b200 import _lib as L
    from torchsr_b200 import ops
    from torchsr_b200.srgan.trainer import SRGANTrainer

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            fn(s)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if distributed:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def train_bench(batch, steps, warmup, sampler=None):
        """value / e2e of the SRGAN GAN step at `batch` crops per GPU. Returns (dict, trainer, device batch)."""
        trainer = make_trainer(SRGANTrainer, batch, distributed, rank, local, world)
        if args.no_vgg:
            trainer.vgg_loss = lambda a, b: torch.nn.functional.mse_loss(a, b)
        lr_h, hr_h = synthetic_batch(batch, 1234 + rank, pinned=True)
        lr_d, hr_d = lr_h.cuda(), hr_h.cuda()
        # public step API: trainer.graph_step replays the whole _gan_loop as one CUDA graph (same arithmetic as the
        # eager call, see srgan/trainer.py); --eager times the plain Python call instead
        step_fn = trainer._gan_loop if args.eager else trainer.graph_step
        for s in range(max(warmup, 3)):
            step_fn(lr_d, hr_d, s)
        # kernels of ours per step (graph replays do not pass through the library's launch counter)
        l0 = L.launch_count()
        trainer._gan_loop(lr_d, hr_d, 0)
        launches_per_step = L.launch_count() - l0
        if sampler is not None:
            sampler.start()
        ms = timed(lambda s: step_fn(lr_d, hr_d, s), steps)
        clocks = sampler.stop() if sampler is not None else None
        # end to end: pinned host batch -> H2D inside the step, loss read back to the host every step
        ms_e2e = timed(lambda s: step_fn(lr_h, hr_h, s).item(), steps)
        crops = batch * world * steps
        r = {"value": crops / (ms * 1e-3), "ms_per_step": ms / steps,
             "e2e": {"value": crops / (ms_e2e * 1e-3), "unit": "crops/s", "ms_per_step": ms_e2e / steps,
                     "h2d_bytes_per_step": int((lr_h.numel() + hr_h.numel()) * 4 * world), "d2h_bytes_per_step": 4 * world},
             "gpu_launches": int(launches_per_step * steps), "launches_per_step": int(launches_per_step), "clocks": clocks}
        return r, trainer, (lr_d, hr_d)

    sampler = ClockSampler(local) if rank == 0 else None
    head, trainer, (lr_d, hr_d) = train_bench(args.batch, args.steps, args.warmup, sampler)
    parity = None
    if distributed and want(args, "dp_parity"):
        try:
            parity = dp_parity(trainer, lr_d, hr_d)
        except Exception as exc:  # noqa: BLE001
            parity = "error: " + repr(exc)[:200]
    b64 = None
    if want(args, "b64"):
        try:
            r64, tr64, _ = train_bench(64, max(5, min(args.steps, 20)), 5)
            b64 = {"workload": "SRGAN G+D _gan_loop (BASELINE configs[2]), batch 64 per GPU", "value": r64["value"],
                   "unit": "crops/s", "ms_per_step": r64["ms_per_step"], "e2e": r64["e2e"], "n_gpus": world,
                   "launches_per_step": r64["launches_per_step"]}
            del tr64
            torch.cuda.empty_cache()
        except Exception as exc:  # noqa: BLE001
            b64 = {"error": repr(exc)[:300]}

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
    value = head["value"]
    ach = value * GFLOP_PER_CROP_GD / 1e3 / world     # TFLOP/s per GPU on the G+D algorithmic FLOPs
    vgg_ours = not args.no_vgg
    extra = {}
    dom = kern = None
    if not distributed:
        counts = _programs_of_one_step(trainer, lr_d, hr_d)
        if want(args, "roofline"):
            kern = trunk_conv_roofline(args.batch, pk)
            try:
                dom = conv_kernel_roofline(counts, args.batch, pk, vgg_ours)
            except Exception as exc:  # noqa: BLE001
                dom = {"error": repr(exc)[:300]}
        if want(args, "hbm"):
            try:
                extra["hbm_kernels"] = hbm_kernels(counts, trainer, pk)
            except Exception as exc:  # noqa: BLE001
                extra["hbm_kernels"] = {"error": repr(exc)[:300]}
        del counts
    else:
        dom = {"note": "measured at N=1 only (replaying a step on rank 0 alone would leave its collectives unmatched)"}
    ops.check_watchdog()
    traffic, traffic_src = traffic_from_profiles()
    line = {
        "metric": "SRGAN GAN training crops/sec (96x96 HR)", "value": value, "unit": "crops/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "SRGAN G+D _gan_loop (BASELINE configs[1]), batch %d per GPU of 96x96 HR / 24x24 LR "
                               "crops, random-init weights, VGG19 loss %s" %
                               (args.batch, "replaced by MSE (--no-vgg)" if args.no_vgg else
                                "on (random-init weights)"),
                   "vgg": "executed by this repo's kernels (nets.define_vgg)", "parallelism": f"dp{world}",
                   "global_batch": args.batch * world,
                   "step_api": "SRGANTrainer._gan_loop (eager)" if args.eager else
                               "SRGANTrainer.graph_step (whole step replayed as one CUDA graph)",
                   "l2": "per-step working set (fp32 weights + Adam state + activations, > 0.5 GB) exceeds the 126 MB "
                         "L2; no explicit flush between steps"},
        "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "launches_per_step": head["launches_per_step"],
        "clocks": head["clocks"],
        "roofline": {"bound": "tensor", "achieved": ach, "peak": pk["sustained"], "unit": "TFLOP/s",
                     "frac": ach / pk["sustained"], "traffic": traffic, "traffic_source": traffic_src,
                     "what": "whole step: crops/s x 21.73 GFLOP/crop (G+D algorithmic minimum, SURVEY 8d) per GPU vs the "
                             f"{pk['source']} sustained bf16 peak; with the VGG19 FLOPs (+21.50/crop) "
                             f"the step sustains {value * (GFLOP_PER_CROP_GD + GFLOP_PER_CROP_VGG) / 1e3 / world:.1f} TFLOP/s",
                     "dominant_kernel": dom, "trunk_conv_only": kern,
                     "ncu": "profiles/r02d_* / r02g_ncu_launches_step.csv (launch list of one step of the final build, --set full "
                            "summaries of the conv, weight-gradient, BatchNorm and Adam kernels), r02f_* (inference kernels); "
                            "r02, r02b = earlier builds of round 2, r01* = round 1"},
    }
    if parity is not None:
        line["dp_parity"] = parity
    if b64 is not None:
        if isinstance(b64, dict) and "value" in b64:
            tf = b64["value"] * GFLOP_PER_CROP_GD / 1e3 / world
            b64["roofline"] = {"bound": "tensor", "achieved": tf, "peak": pk["sustained"], "unit": "TFLOP/s",
                               "frac": tf / pk["sustained"]}
        line["b64"] = b64
    line.update(extra)
    if not distributed:
        del trainer
        torch.cuda.empty_cache()
        for name, fn in (("inference", lambda: inference_block(pk)), ("esrgan", lambda: esrgan_block(pk)),
                         ("data_pipeline", lambda: data_pipeline_block(args.batch)),
                         ("gpu_eager_baseline", lambda: gpu_eager_baseline(args.batch))):
            key = {"gpu_eager_baseline": "eager_baseline"}.get(name, name)
            if not want(args, key):
                continue
            try:
                line[name] = fn()
            except Exception as exc:  # noqa: BLE001
                line[name] = {"error": repr(exc)[:300]}
        if isinstance(line.get("gpu_eager_baseline"), dict) and "variants" in line["gpu_eager_baseline"]:
            best = max((v["value"] for v in line["gpu_eager_baseline"]["variants"].values() if "value" in v), default=None)
            if best:
                line["gpu_eager_baseline"]["speedup_over_best_variant"] = value / best
    if not args.no_cpu_baseline and world == 1 and want(args, "cpu"):
        cps, spstep, threads, kind = cpu_reference_crops_per_sec(args.batch, 4, 1)
        line["cpu_baseline"] = {"value": cps, "unit": "crops/s", "cores": threads, "kind": kind,
                                "sample": f"4 timed + 1 warm-up steps of batch {args.batch}: " +
                                          ("the unmodified reference's SRGANTrainer._gan_loop (oracle/_ref) on "
                                           "torch.device('cpu'), fp32, all host threads" if kind == "reference" else
                                           "oracle port (oracle/step_oracle.py), fp32, all host threads")}
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
