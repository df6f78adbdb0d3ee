"""Per-program timing on the GPU: replays the recorded forward / backward launch lists of the SRGAN generator and
discriminator (batch from argv) and prints microseconds per replay and per op. Env TSR_PDL / TSR_WGRAD_BRANCH / TSR_GRAPHS
select the library's launch modes, so two runs give an A/B."""
import os
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from torchsr_b200 import _lib as L  # noqa: E402
from torchsr_b200 import dist as tdist  # noqa: E402
from torchsr_b200 import ops as ops_mod  # noqa: E402
from torchsr_b200.srgan.discriminator import Discriminator  # noqa: E402
from torchsr_b200.srgan.generator import Generator  # noqa: E402

KINDS = {v: k for k, v in vars(L).items() if k.startswith("E_")}


def time_prog(prog, reps=20):
    prog.run()
    prog.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        prog.run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def describe(prog):
    c = Counter()
    for d in prog.descs:
        if isinstance(d, ops_mod.ConvGroupDesc):
            c["conv_group"] += 1
        elif isinstance(d, L.ConvDesc):
            c["conv"] += 1
        elif isinstance(d, L.WgradDesc):
            c["wgrad"] += 1
        else:
            c[KINDS.get(d.kind, str(d.kind))] += 1
    return dict(c)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    torch.manual_seed(0)
    dev = torch.device("cuda")
    G, D = Generator().to(dev), Discriminator().to(dev)
    lr, hr = torch.rand(B, 3, 24, 24, device=dev), torch.rand(B, 3, 96, 96, device=dev)
    sr = G(lr)
    sr.mean().backward()
    D(hr).mean().backward()
    with tdist.frozen(D):
        x = hr.clone().requires_grad_(True)
        D(x).mean().backward()
    torch.cuda.synchronize()
    print(f"batch {B}  TSR_PDL={os.environ.get('TSR_PDL', '1')} TSR_WGRAD_BRANCH={os.environ.get('TSR_WGRAD_BRANCH', '1')}")
    for name, m in (("G", G), ("D", D)):
        for key, plans in m._tsr["plans"].items():
            pl = plans[0]
            progs = [("fwd", pl.fwd)] + [(f"bwd(x={k[0]},w={k[1]})", p) for k, p in pl.bwd.items()]
            for pn, prog in progs:
                us = time_prog(prog)
                n = len(prog)
                print(f"{name} {pn:22s} {us:9.1f} us  {n:4d} ops  {us / n:6.2f} us/op  {describe(prog)}", flush=True)




def op_name(d):
    if isinstance(d, ops_mod.ConvGroupDesc):
        return "group[" + " | ".join(op_name(m) for m in d.members) + "]"
    if isinstance(d, L.ConvDesc):
        if d.a_mode == 0:
            return (f"conv M={d.N * d.Ho * d.Wo} N={d.cout_pad} K={d.num_taps}x{d.C - d.a_c0} bn={d.block_n} s={d.stride} "
                    f"mode={d.out_mode} stats={int(bool(d.stats_partial))} res={int(bool(d.res))} z={int(bool(d.bwd_z))}")
        return f"gemm M={d.gemm_M} N={d.cout_pad} K={d.gemm_K} bn={d.block_n} splits={d.splits} a_mode={d.a_mode}"
    if isinstance(d, L.WgradDesc):
        return f"wgrad M={d.N * d.Ho * d.Wo} C={d.C - d.x_c0} Cout={d.cout_valid} taps={d.num_taps} bn={d.block_n}"
    return f"{KINDS.get(d.kind, d.kind)} i={[int(v) for v in d.i[:4]]}"


def per_op():
    """Each op of each program in isolation: the same descriptor 40x back to back in one graph."""
    from torchsr_b200 import ops
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    torch.manual_seed(0)
    dev = torch.device("cuda")
    G, D = Generator().to(dev), Discriminator().to(dev)
    lr, hr = torch.rand(B, 3, 24, 24, device=dev), torch.rand(B, 3, 96, 96, device=dev)
    G(lr).mean().backward()
    D(hr).mean().backward()
    torch.cuda.synchronize()
    for name, m in (("G", G), ("D", D)):
        for key, plans in m._tsr["plans"].items():
            pl = plans[0]
            progs = [("fwd", pl.fwd)] + [(f"bwd(x={k[0]},w={k[1]})", p) for k, p in pl.bwd.items()]
            for pn, prog in progs:
                print(f"---- {name} {pn}")
                seen = {}
                tot = 0.0
                for d in prog.descs:
                    nm = op_name(d)
                    if nm not in seen:
                        p2 = ops.Program()
                        for _ in range(40):
                            if isinstance(d, ops_mod.ConvGroupDesc):
                                p2.add_group(d.members)
                            else:
                                p2.add(d)
                        seen[nm] = [time_prog(p2, 5) / 40, 0]
                    seen[nm][1] += 1
                    tot += seen[nm][0]
                for nm, (us, cnt) in seen.items():
                    print(f"  {us:8.2f} us x{cnt:3d}  {nm}")
                print(f"  sum of isolated op times: {tot:.1f} us")


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[2] == "ops":
    per_op()
elif __name__ == "__main__":
    main()
