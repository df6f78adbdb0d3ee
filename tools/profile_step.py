"""GPU timeline of the training step with torch.profiler (CUPTI): per-kernel totals and busy fraction.
Usage: python tools/profile_step.py [batch]"""
import os
import sys
from argparse import Namespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TORCHSR_VGG_WEIGHTS", "random")
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from torchsr_b200.srgan.trainer import SRGANTrainer  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    novgg = len(sys.argv) > 2 and sys.argv[2] == "novgg"
    torch.manual_seed(0)
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:       # torchrun: data-parallel timeline of rank 0
        import torch.distributed as dist
        from torchsr_b200.dist import init_process_group as tdist_init
        tdist_init(local)
    targs = Namespace(disable_amp=False, batch_size=B, epochs=8, pretrain_epochs=1, gan_checkpoint=None,
                      psnr_checkpoint=None, skip_image_save=True, local_rank=local, rank=rank if world > 1 else -1,
                      world_size=world)
    tr = SRGANTrainer(torch.device("cuda"), targs, [], [], 0, 0, world > 1)
    if novgg:
        tr.vgg_loss = lambda a, b: torch.nn.functional.mse_loss(a, b)
    lr, hr = torch.rand(B, 3, 24, 24, device="cuda"), torch.rand(B, 3, 96, 96, device="cuda")
    step_fn = tr._gan_loop if os.environ.get("EAGER") else tr.graph_step
    for s in range(5):
        step_fn(lr, hr, s)
    torch.cuda.synchronize()
    steps = 3
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for s in range(steps):
            step_fn(lr, hr, s)
        torch.cuda.synchronize()
    if rank != 0:
        return
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    agg = {}
    t0, t1, busy = None, None, 0.0
    for e in evs:
        name = e.name[:70]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += e.device_time
        s_, e_ = e.time_range.start, e.time_range.end
        t0 = s_ if t0 is None else min(t0, s_)
        t1 = e_ if t1 is None else max(t1, e_)
        busy += e.device_time
    span = (t1 - t0)
    # union of busy intervals (kernels overlapping on different streams / graph branches count once)
    iv = sorted((e.time_range.start, e.time_range.end) for e in evs)
    union, cur_s, cur_e = 0.0, None, None
    for s_, e_ in iv:
        if cur_e is None or s_ > cur_e:
            if cur_e is not None:
                union += cur_e - cur_s
            cur_s, cur_e = s_, e_
        else:
            cur_e = max(cur_e, e_)
    union += (cur_e - cur_s) if cur_e is not None else 0.0
    print(f"union-busy {union / steps / 1e3:.3f} ms/step (sum of durations {busy / steps / 1e3:.3f})")
    cpu_ops = {}
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CPU and e.name.startswith("aten::"):
            a = cpu_ops.setdefault(e.name, [0, 0.0])
            a[0] += 1
            a[1] += e.cpu_time
    print("top aten ops on the CPU (count/step, us/step):",
          [(k, round(v[0] / steps, 1), round(v[1] / steps)) for k, v in sorted(cpu_ops.items(), key=lambda kv: -kv[1][1])[:14]])
    print(f"batch {B}: {steps} steps, GPU span {span / steps / 1e3:.3f} ms/step, kernel-busy {busy / steps / 1e3:.3f} ms/step, "
          f"{len(evs) / steps:.0f} GPU activities/step")
    dump = os.environ.get("TIMELINE")
    if dump:
        # last step only: start (us since the step's first activity), duration, stream id, name
        kev = [k for k in prof.profiler.kineto_results.events() if k.device_type() == torch.autograd.DeviceType.CUDA]
        kev.sort(key=lambda k: k.start_ns())
        last = kev[len(kev) * (steps - 1) // steps:]
        z = last[0].start_ns()
        with open(dump, "w") as f:
            for k in last:
                f.write(f"{(k.start_ns() - z) / 1e3:.1f},{k.duration_ns() / 1e3:.1f},{k.device_resource_id()},{k.name()[:90]}\n")
    durs = {}
    for e in evs:
        durs.setdefault(e.name[:70], []).append(e.device_time)
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get("TOP", "40"))]:
        d = sorted(durs[name])
        q = lambda f: d[min(len(d) - 1, int(f * len(d)))]  # noqa: E731
        print(f"{t / steps:9.1f} us/step {n / steps:6.1f} x {t / n:8.2f} us  [min {d[0]:.1f} p25 {q(.25):.1f} p50 {q(.5):.1f} "
              f"p75 {q(.75):.1f} max {d[-1]:.1f}]  {name[:60]}")


if __name__ == "__main__":
    main()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.cuda.synchronize()
        torch.distributed.barrier()
        os._exit(0)
