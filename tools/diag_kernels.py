"""Runs every single-kernel parity check and prints one line per check (never stops at the first failure).
Usage (GPU box):  python tools/diag_kernels.py [filter]  -> gpurun_out/diag_kernels.json"""
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import eltwise_checks as E  # noqa: E402
import kernel_checks as K  # noqa: E402
from torchsr_b200 import _lib as L  # noqa: E402

def _env(name, value, fn):
    """Runs fn() with an environment switch set (the C side reads its switches when a descriptor is built)."""
    old = os.environ.get(name)
    os.environ[name] = value
    try:
        return fn()
    finally:
        if old is None:
            os.environ.pop(name, None)
        else:
            os.environ[name] = old


CHECKS = [
    ("layout", lambda: E.check_layout()),
    ("im2row_3x3", lambda: E.check_im2row()),
    ("im2row_9x9", lambda: E.check_im2row(KH=9, KW=9)),
    ("im2row_row_bwd", lambda: E.check_im2row(KH=1, KW=9, sign=-1)),
    ("gather_out", lambda: E.check_gather_out()),
    ("gather_out_3x3_tiled", lambda: E.check_gather_out_3x3()),
    ("gather_out_3x3_tiled_fwd_sign", lambda: E.check_gather_out_3x3(B=1, H=8, W=32, sign=1)),
    ("bn_prelu_res", lambda: E.check_bn_train()),
    ("bn_leaky_c512", lambda: E.check_bn_train(M=600, C=512, act=L.ACT_LEAKY, residual=False)),
    ("bn_none", lambda: E.check_bn_train(M=333, C=64, act=L.ACT_NONE, residual=False)),
    ("act_bwd_bias", lambda: E.check_act_bwd_bias()),
    ("pack_unpack", lambda: E.check_pack_unpack()),
    ("loss_mse", lambda: E.check_loss(kind=0)),
    ("loss_l1", lambda: E.check_loss(kind=1)),
    ("head_sigmoid", lambda: E.check_head()),
    ("head_logits", lambda: E.check_head(N1=100, sigmoid=0)),
    ("linear_wgrad", lambda: E.check_linear_wgrad()),
    ("upsample", lambda: E.check_upsample()),
    ("gemm_k64", lambda: K.check_gemm()),
    ("gemm_n256", lambda: K.check_gemm(M=300, N=256, K=512, block_n=256)),
    ("gemm_n16", lambda: K.check_gemm(M=128, N=16, K=64, block_n=16)),
    ("gemm_splitk_t", lambda: K.check_gemm_splitk_t()),
    ("gemm_mn_major", lambda: K.check_gemm_mn_major()),
    ("conv3x3_64", lambda: K.check_conv_fwd()),
    ("conv3x3_64_b16", lambda: K.check_conv_fwd(B=16)),
    ("conv3x3_odd", lambda: K.check_conv_fwd(B=3, H=13, W=10)),
    ("conv3x3_stats", lambda: K.check_conv_fwd(B=4, stats=True)),
    ("conv3x3_bias_prelu", lambda: K.check_conv_fwd(bias=True, act=L.ACT_PRELU)),
    ("conv3x3_res", lambda: K.check_conv_fwd(residual=True)),
    ("conv3x3_bf16out", lambda: K.check_conv_fwd(out_f32=False)),
    ("conv3x3_shuffle", lambda: K.check_conv_fwd(Cout=256, block_n=64, bias=True, act=L.ACT_PRELU, shuffle=True)),
    ("conv3x3_shuffle_n128", lambda: K.check_conv_fwd(Cout=256, block_n=128, bias=True, shuffle=True)),
    ("conv3x3_c32", lambda: K.check_conv_fwd(Cin=32, Cout=32)),
    ("conv3x3_c96_n32", lambda: K.check_conv_fwd(Cin=96, Cout=32)),
    ("conv3x3_c16", lambda: K.check_conv_fwd(Cin=16, Cout=64)),
    ("conv3x3_128_256", lambda: K.check_conv_fwd(Cin=128, Cout=256, H=12, W=12)),
    ("conv3x3_512", lambda: K.check_conv_fwd(Cin=512, Cout=512, H=6, W=6)),
    # activation multicast across clusters of two N tiles (tiles_n even, M % 128 == 0, >= 18 K iterations)
    ("conv3x3_cluster_128_256", lambda: _env("TSR_CONV_CLUSTER", "1", lambda: K.check_conv_fwd(Cin=128, Cout=256, H=16, W=16))),
    ("conv3x3_cluster_256_512_epi", lambda: _env("TSR_CONV_CLUSTER", "1", lambda: K.check_conv_fwd(
        B=4, Cin=256, Cout=512, H=8, W=8, bias=True, act=L.ACT_LEAKY, stats=True))),
    ("conv3x3_cluster_s2", lambda: _env("TSR_CONV_CLUSTER", "1", lambda: K.check_conv_fwd(
        Cin=128, Cout=256, H=32, W=32, stride=2, stats=True))),
    ("conv3x3_cluster_n64_res", lambda: _env("TSR_CONV_CLUSTER", "1", lambda: K.check_conv_fwd(
        B=8, Cin=128, Cout=128, H=16, W=16, block_n=64, residual=True, repeat=3))),
    ("conv3x3_splitk", lambda: K.check_conv_fwd(Cin=256, Cout=256, H=12, W=12, splits=6, bias=True, act=L.ACT_LEAKY,
                                                stats=True, repeat=2)),
    ("conv3x3_splitk_odd", lambda: K.check_conv_fwd(B=3, H=13, W=10, Cin=128, Cout=64, splits=4, residual=True,
                                                    out_f32=False, repeat=2)),
    # persistent weight-stationary mode (>= 2 M tiles per CTA): resident weights, two accumulator stages
    ("conv3x3_persistent", lambda: K.check_conv_fwd(B=18, H=48, W=48, bias=True, act=L.ACT_PRELU, stats=True)),
    ("conv3x3_persistent_n256", lambda: K.check_conv_fwd(B=4, H=72, W=72, Cout=256, block_n=128, bias=True, shuffle=True,
                                                         out_f32=False)),
    ("conv3x3_persistent_odd", lambda: K.check_conv_fwd(B=21, H=45, W=43, Cin=32, Cout=32, residual=True)),
    ("conv3x3_persistent_s2", lambda: K.check_conv_fwd(B=20, H=96, W=96, stride=2, stats=True)),
    ("conv3x3_s2", lambda: K.check_conv_fwd(stride=2)),
    ("conv3x3_s2_128", lambda: K.check_conv_fwd(Cin=128, Cout=128, H=12, W=12, stride=2, stats=True)),
    ("conv9x9_64", lambda: K.check_conv_fwd(k=9, Cout=32, H=12, W=12)),
    ("dgrad_s1", lambda: K.check_dgrad_s1()),
    ("dgrad_s1_256_128", lambda: K.check_dgrad_s1(Cin=128, Cout=256, H=12, W=12)),
    ("dgrad_s2", lambda: K.check_dgrad_s2()),
    ("dgrad_s2_512", lambda: K.check_dgrad_s2(Cin=512, Cout=512, H=12, W=12)),
    ("dgrad_s2_grouped", lambda: K.check_dgrad_s2(B=3, H=12, W=20, Cin=128, Cout=64, grouped=True)),
    ("wgrad_64", lambda: K.check_wgrad()),
    ("wgrad_64_b16", lambda: K.check_wgrad(B=16)),
    ("wgrad_odd", lambda: K.check_wgrad(B=3, H=13, W=10)),
    ("wgrad_s2", lambda: K.check_wgrad(stride=2)),
    ("wgrad_128_256", lambda: K.check_wgrad(Cin=128, Cout=256, H=12, W=12)),
    ("wgrad_32_32", lambda: K.check_wgrad(Cin=32, Cout=32)),
    ("wgrad_192_64", lambda: K.check_wgrad(Cin=192, Cout=64, H=16, W=16)),
    ("wgrad_512", lambda: K.check_wgrad(Cin=512, Cout=512, H=6, W=6, B=4)),
    ("wgrad_1tap_c256", lambda: K.check_wgrad(Cin=256, Cout=64, k=1)),
]


def run_checks(names, out_path):
    print(torch.cuda.get_device_name(0), flush=True)
    table = dict(CHECKS)
    for name in names:
        fn = table[name]
        t0 = time.time()
        try:
            r = fn()
            worst = max(r.values()) if r else 0.0
            status = "ok" if worst < 1e-2 else "BAD"
            rec = dict(name=name, status=status, errs=r)
            print(f"{status:4s} {name:24s} {time.time() - t0:6.2f}s  " + "  ".join(f"{k}={v:.2e}" for k, v in r.items()),
                  flush=True)
        except Exception as ex:  # noqa: BLE001
            rec = dict(name=name, status="EXC", error=repr(ex))
            print(f"EXC  {name:24s} {ex!r}", flush=True)
            traceback.print_exc(limit=3)
        with open(out_path, "a") as f:
            f.write(json.dumps(rec) + "\n")
        if rec["status"] == "EXC":
            try:
                torch.cuda.synchronize()
            except Exception as ex2:  # noqa: BLE001
                print("CUDA context is dead:", ex2, flush=True)
                sys.exit(3)


def main():
    """Driver: runs the checks in child processes so that a faulting kernel cannot take the rest down."""
    import subprocess
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    if len(sys.argv) > 2 and sys.argv[1] == "--only":
        run_checks(sys.argv[2].split(","), sys.argv[3])
        return
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    names = [n for n, _ in CHECKS if flt in n]
    jl = os.path.join(out_dir, "diag_kernels.jsonl")
    if os.path.exists(jl):
        os.remove(jl)
    pending = list(names)
    chunk = len(pending)
    while pending:
        batch, pending = pending[:chunk], pending[chunk:]
        rc = subprocess.call([sys.executable, os.path.abspath(__file__), "--only", ",".join(batch), jl],
                             timeout=900)
        done = set()
        if os.path.exists(jl):
            done = {json.loads(l)["name"] for l in open(jl)}
        rest = [n for n in batch if n not in done]
        if rc != 0 and rest:
            # A child died mid-check: most likely a device fault. Faults are rationed on the shared boxes, so stop here
            # instead of risking a second one; the remaining checks are reported as not run.
            with open(jl, "a") as f:
                f.write(json.dumps(dict(name=rest[0], status="CRASH", error=f"child exit {rc}")) + "\n")
                for n in rest[1:] + pending:
                    f.write(json.dumps(dict(name=n, status="NOT_RUN")) + "\n")
            print(f"CRASH {rest[0]} (child exit {rc}); not run: {rest[1:] + pending}", flush=True)
            pending = []
    results = [json.loads(l) for l in open(jl)]
    bad = [r["name"] for r in results if r["status"] != "ok"]
    print(f"{len(results) - len(bad)}/{len(results)} ok; failing: {bad}")


if __name__ == "__main__":
    main()
