"""Whole-module and training-step parity on the B200, through the public nn.Module API, against the CPU oracle and the
golden vectors recorded from the reference.

Tolerances (north star): per-stage outputs rel-L2 <= 1e-2 in bf16; whole-network outputs accumulate one bf16 rounding
per layer (SURVEY.md App. E measures 1.5e-2 for PyTorch's own bf16 autocast on the 16-block generator), so the
end-to-end bound is 3e-2; gradients through BatchNorm + PReLU/LeakyReLU in bf16 flip activation signs near zero, so
whole-network gradients are held to the level PyTorch's bf16 autocast reaches (median rel-L2 <= 0.2) while the
per-kernel gradient arithmetic is pinned to 1e-4 in test_kernels_gpu.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mods():
    import module_checks as MC
    from torchsr_b200.esrgan.discriminator import Discriminator as ED
    from torchsr_b200.esrgan.generator import Generator as EG
    from torchsr_b200.srgan.discriminator import Discriminator as SD
    from torchsr_b200.srgan.generator import Generator as SG
    return MC, SG, SD, EG, ED


@pytest.fixture
def bn_fuse(request, monkeypatch):
    """Runs a test with the fused conv+BatchNorm launch (default) or with the two-launch path (conv, then bn_act)."""
    from torchsr_b200 import engine
    monkeypatch.setattr(engine, "FUSE_BN_FWD", bool(request.param))
    monkeypatch.setattr(engine, "FUSE_BN_BWD", bool(request.param))
    return bool(request.param)


BN_PATHS = pytest.mark.parametrize("bn_fuse", [True, False], indirect=True, ids=["bn-fused", "bn-two-launch"])


@BN_PATHS
def test_residual_block_stage(bn_fuse):
    import module_checks as MC
    from torchsr_b200.srgan.residual import ResidualBlock
    torch.manual_seed(1)
    m = ResidualBlock()
    MC.randomize_bn(m)
    r, errs = MC.check_module(m, lambda sd, x, tr, buf: MC.O.srgan_residual_block(
        {("." + k): v for k, v in sd.items()}, "", x, tr, buf), torch.randn(4, 64, 24, 24), input_grad=True)
    assert r["out"] <= 1e-2 and r["bn_buffers"] <= 1e-2, r
    assert r["grad_median"] <= 3e-2 and r["dx"] <= 5e-2, (r, errs)


@pytest.mark.parametrize("occupancy", ["api", "own"])
def test_residual_block_stage_batch64_more_tiles_than_sms(occupancy, monkeypatch):
    """Batch 64 of 24x24: 288 output tiles, i.e. more CTAs than SMs. With the runtime's occupancy answer (1 CTA per SM for
    every tcgen05 kernel) the stage takes the two-launch BatchNorm path; with TSR_OCCUPANCY=own (opt-in, host-side
    resource count, profiles/r02c_two_ctas_per_sm.md) the fused launches rely on two CTAs per SM being resident together
    at the grid barrier - fine for a module run on its own, which is what this test does."""
    monkeypatch.setenv("TSR_OCCUPANCY", occupancy)
    import module_checks as MC
    from torchsr_b200 import ops
    from torchsr_b200.srgan.residual import ResidualBlock
    torch.manual_seed(3)
    m = ResidualBlock()
    MC.randomize_bn(m)
    launches0 = ops.launch_count() if hasattr(ops, "launch_count") else None
    r, errs = MC.check_module(m, lambda sd, x, tr, buf: MC.O.srgan_residual_block(
        {("." + k): v for k, v in sd.items()}, "", x, tr, buf), torch.randn(64, 64, 24, 24), input_grad=True)
    assert r["out"] <= 1e-2 and r["bn_buffers"] <= 1e-2, r
    assert r["grad_median"] <= 3e-2 and r["dx"] <= 5e-2, (r, errs)
    ops.check_watchdog()
    del launches0


def test_subpixel_stage():
    import module_checks as MC
    from torchsr_b200.srgan.residual import SubpixelConvolutionLayer
    torch.manual_seed(2)
    r, errs = MC.check_module(SubpixelConvolutionLayer(), lambda sd, x, tr, buf: MC.O.srgan_subpixel(
        {("." + k): v for k, v in sd.items()}, "", x), torch.randn(2, 64, 12, 12), input_grad=True)
    assert r["out"] <= 1e-2 and r["grad_worst"] <= 5e-2 and r["dx"] <= 5e-2, (r, errs)


def _branch_gain(module, gain):
    """The reference init (kaiming_normal * 0.1) makes a dense block's residual branch ~3 % of its output, below one
    bf16 ulp of the identity path; the parity tests scale the weights up so that the branch is what gets compared."""
    with torch.no_grad():
        for m in module.modules():
            if isinstance(m, torch.nn.Conv2d):
                m.weight.mul_(gain)
                m.bias.copy_(torch.randn(m.bias.shape) * 0.05)


@pytest.mark.parametrize("kind", ["rdb", "rrdb"])
def test_esrgan_dense_blocks_standalone_residual_branch(kind):
    """ResidualDenseBlock / ResidualInResidualDenseBlock called on their own (reference esrgan/residual.py:81-86,
    124-129), teacher-forced on a bf16-representable input. The block output is dominated by the identity path
    (SURVEY.md section 4), so the RESIDUAL BRANCH (out - x) is what is compared, plus every parameter gradient."""
    import module_checks as MC
    from torchsr_b200.esrgan.residual import ResidualDenseBlock, ResidualInResidualDenseBlock
    torch.manual_seed(11)
    m = ResidualDenseBlock() if kind == "rdb" else ResidualInResidualDenseBlock()
    _branch_gain(m, 10.0 if kind == "rdb" else 15.0)
    x = torch.randn(2, 64, 16, 16).bfloat16().float()
    fn = MC.O.esrgan_rdb if kind == "rdb" else MC.O.esrgan_rrdb
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    r, errs = MC.check_module(m, lambda s_, x_, tr, buf: fn({("p." + k): v for k, v in s_.items()}, "p", x_), x,
                              input_grad=True)
    with torch.no_grad():
        y_gpu = m(x.cuda()).cpu()
        y_ref = fn({("p." + k): v for k, v in sd.items()}, "p", x)
    ident = 1.0 if kind == "rdb" else 1.2      # RRDB: out = 1.2 x + 0.04 (c5_1 + c5_2 + c5_3)
    branch = MC.rel_l2(y_gpu - ident * x, y_ref - ident * x)
    share = float((y_ref - ident * x).norm() / y_ref.norm())
    print(kind, "branch", branch, "branch share of output", share, r)
    assert share > 0.2, share
    assert branch <= 2e-2, (branch, r)
    # measured on B200: rdb median 3.3e-2 / worst 4.0e-2, rrdb 5.0e-2 / 6.9e-2 (three chained blocks, 15 LeakyReLU masks
    # from bf16 pre-activations; PyTorch's own bf16 autocast measures 2.9e-2 median on ONE block)
    assert r["grad_median"] <= 7e-2 and r["grad_worst"] <= 0.12 and r["dx"] <= 5e-2, (r, sorted(errs.items(), key=lambda kv: -kv[1])[:5])


def test_srgan_generator_vs_oracle():
    MC, SG, SD, EG, ED = _mods()
    torch.manual_seed(3)
    G = SG()
    MC.randomize_bn(G)
    r, errs = MC.check_module(G, MC.O.srgan_generator, torch.rand(4, 3, 24, 24))
    assert r["out"] <= 3e-2 and r["bn_buffers"] <= 1e-2, r
    assert r["grad_median"] <= 0.2, (r, sorted(errs.items(), key=lambda kv: -kv[1])[:5])


@BN_PATHS
def test_srgan_discriminator_vs_oracle(bn_fuse):
    MC, SG, SD, EG, ED = _mods()
    torch.manual_seed(4)
    D = SD()
    MC.randomize_bn(D)
    r, errs = MC.check_module(D, MC.O.srgan_discriminator, torch.rand(4, 3, 96, 96), input_grad=True)
    assert r["out"] <= 1e-2 and r["bn_buffers"] <= 1e-2, r
    assert r["grad_median"] <= 0.25 and r["dx"] <= 0.3, (r, sorted(errs.items(), key=lambda kv: -kv[1])[:5])


def test_esrgan_generator_vs_oracle():
    MC, SG, SD, EG, ED = _mods()
    torch.manual_seed(5)
    G = EG(num_rrdb_blocks=3)
    r, errs = MC.check_module(G, lambda sd, x, tr, buf: MC.O.esrgan_generator(sd, x), torch.rand(2, 3, 16, 16))
    assert r["out"] <= 2e-2, r
    assert r["grad_median"] <= 0.1, (r, sorted(errs.items(), key=lambda kv: -kv[1])[:5])


def test_esrgan_generator_23_rrdb_vs_oracle():
    """The full-depth generator (23 RRDB = 345 dense-block convs) against the oracle, as a module: output and every
    parameter gradient (module parity had only been run at 2-3 blocks)."""
    MC, SG, SD, EG, ED = _mods()
    torch.manual_seed(15)
    G = EG()
    assert len(G.blocks) == 23
    r, errs = MC.check_module(G, lambda sd, x, tr, buf: MC.O.esrgan_generator(sd, x), torch.rand(1, 3, 16, 16))
    print("esrgan generator, 23 RRDB:", r, sorted(errs.items(), key=lambda kv: -kv[1])[:3])
    assert r["out"] <= 2e-2, r
    # measured: median 0.114, worst 0.208 (bias of a 32-channel dense conv deep in the trunk) - 345 LeakyReLU masks from
    # bf16 pre-activations between the loss and the first block; the 3-block module sits at <= 0.1
    assert r["grad_median"] <= 0.15 and r["grad_worst"] <= 0.4, (r, sorted(errs.items(), key=lambda kv: -kv[1])[:5])


def test_esrgan_discriminator_vs_oracle():
    MC, SG, SD, EG, ED = _mods()
    torch.manual_seed(6)
    D = ED()
    MC.randomize_bn(D)
    r, errs = MC.check_module(D, MC.O.esrgan_discriminator, torch.rand(4, 3, 128, 128), input_grad=True)
    # logits of a random-init net are small sums of cancelling terms after 9 BatchNorm stages (the last one over only
    # B*4*4 samples): a looser relative bound than for the sigmoid output of the SRGAN discriminator
    assert r["out"] <= 5e-2 and r["bn_buffers"] <= 1e-2, r
    assert r["grad_median"] <= 0.3, (r, sorted(errs.items(), key=lambda kv: -kv[1])[:5])


@pytest.mark.parametrize("name", ["srgan_generator", "srgan_discriminator", "esrgan_generator", "esrgan_discriminator"])
def test_golden_vectors_through_cuda_path(name):
    """The reference's recorded outputs (tests/golden) reproduced by the CUDA path with the same synthetic weights."""
    import module_checks as MC
    from golden_util import load_fixture, synth_state_dict, template_from_meta
    MC_, SG, SD, EG, ED = _mods()
    meta, arr = load_fixture(name)
    mod = {"srgan_generator": SG, "srgan_discriminator": SD, "esrgan_generator": lambda: EG(num_rrdb_blocks=2),
           "esrgan_discriminator": ED}[name]()
    mod.load_state_dict(synth_state_dict(template_from_meta(meta), meta["seed"]))
    mod = mod.cuda().train()
    y = mod(arr["input"].cuda())
    err = MC.rel_l2(y, arr["output"])
    # logits (ESRGAN discriminator, batch 2) are small sums of cancelling terms: looser relative bound, see above
    assert err <= (5e-2 if name == "esrgan_discriminator" else 3e-2), err
    # state dict round trip is bit exact (parameters stay fp32 torch tensors)
    sd = mod.state_dict()
    ref_sd = synth_state_dict(template_from_meta(meta), meta["seed"])
    for k, v in ref_sd.items():
        if "running" not in k and "num_batches" not in k:
            assert torch.equal(sd[k].cpu(), v), k


def test_discriminator_two_outstanding_forwards_and_frozen_pass():
    """The GAN step calls D twice before one backward (separate BatchNorm statistics per call), then once more with
    frozen parameters for the generator step (input gradient only)."""
    MC, SG, SD, EG, ED = _mods()
    from torchsr_b200 import dist as tdist
    torch.manual_seed(7)
    D = SD()
    sd = {k: v.clone() for k, v in D.state_dict().items()}
    D = D.cuda().train()
    a, b = torch.rand(3, 3, 96, 96), torch.rand(3, 3, 96, 96)
    pa, pb = D(a.cuda()), D(b.cuda())
    (pa.sum() + 2 * pb.sum()).backward()
    osd = MC.O.with_grad(sd)
    buf = {}
    ra, rb = MC.O.srgan_discriminator(osd, a, True, buf), MC.O.srgan_discriminator(osd, b, True, buf)
    (ra.sum() + 2 * rb.sum()).backward()
    assert MC.rel_l2(pa, ra) <= 1e-2 and MC.rel_l2(pb, rb) <= 1e-2
    assert int(D.state_dict()["features.3.num_batches_tracked"]) == 2
    ref = {k: v.grad for k, v in osd.items() if v.requires_grad}
    worst, median, _ = MC.grad_report(D, ref)
    assert median <= 0.25, (worst, median)
    x = a.cuda().requires_grad_(True)
    with tdist.frozen(D):
        D.zero_grad()
        D(x).sum().backward()
    assert x.grad is not None and all(p.grad is None for p in D.parameters())


def _stage_assert(rep, label):
    print(label, "stagewise summary (max error, tolerance, checks):", rep.summary())
    bad = rep.failures()
    assert not bad, (label, len(bad), "of", len(rep.rows), "worst", rep.worst(), bad[:6])


@BN_PATHS
def test_srgan_generator_every_stage_teacher_forced(bn_fuse):
    """All 33 conv + BatchNorm stages of the 16-block generator, each checked against fp32 math on the tensors the CUDA
    path stored around it: raw conv output, batch statistics, stage output, BatchNorm input gradient, data gradient and
    EVERY parameter gradient (conv weights, gamma, beta, PReLU slopes) to a few 1e-3 (tests/stagewise_checks.py)."""
    import module_checks as MC
    import stagewise_checks as SC
    MC_, SG, SD, EG, ED = _mods()
    torch.manual_seed(41)
    G = SG()
    MC.randomize_bn(G)
    G = G.cuda().train()
    x = torch.rand(4, 3, 24, 24, device="cuda")
    gout = torch.randn(4, 3, 96, 96, device="cuda")
    rep = SC.srgan_generator_stagewise(G, x, gout)
    assert len(rep.rows) > 33 * 8
    _stage_assert(rep, "srgan generator " + ("fused" if bn_fuse else "two-launch"))


@BN_PATHS
@pytest.mark.parametrize("pair", [False, True], ids=["one-call", "paired-call"])
def test_srgan_discriminator_every_stage_teacher_forced(bn_fuse, pair):
    """The seven conv + BatchNorm + LeakyReLU stages of the discriminator, forward and backward, for one call and for
    one paired (real | fake) call with per-half batch statistics: statistics, outputs, gradients, parameter gradients."""
    import module_checks as MC
    import stagewise_checks as SC
    from torchsr_b200.srgan.discriminator import CONV_IDX
    MC_, SG, SD, EG, ED = _mods()
    torch.manual_seed(42)
    D = SD()
    MC.randomize_bn(D)
    D = D.cuda().train()
    xs = [torch.rand(8, 3, 96, 96, device="cuda") for _ in range(2 if pair else 1)]
    rep = SC.discriminator_stagewise(D, CONV_IDX, xs, pair)
    _stage_assert(rep, "srgan discriminator " + ("pair " if pair else "") + ("fused" if bn_fuse else "two-launch"))


def test_esrgan_discriminator_every_stage_teacher_forced_paired():
    import module_checks as MC
    import stagewise_checks as SC
    from torchsr_b200.esrgan.discriminator import CONV_IDX
    MC_, SG, SD, EG, ED = _mods()
    torch.manual_seed(43)
    D = ED()
    MC.randomize_bn(D)
    D = D.cuda().train()
    xs = [torch.rand(4, 3, 128, 128, device="cuda") for _ in range(2)]
    rep = SC.discriminator_stagewise(D, CONV_IDX, xs, True)
    _stage_assert(rep, "esrgan discriminator pair")


def _pair_case(D_cls, n, size, fuse, monkeypatch, oracle_fn):
    """forward_pair(a, b) against two separate calls of an identical module AND against the oracle's two calls."""
    import module_checks as MC
    from torchsr_b200 import engine
    monkeypatch.setattr(engine, "FUSE_BN_FWD", fuse)
    monkeypatch.setattr(engine, "FUSE_BN_BWD", fuse)
    torch.manual_seed(21)
    D1 = D_cls()
    MC.randomize_bn(D1)
    sd = {k: v.clone() for k, v in D1.state_dict().items()}
    D2 = D_cls()
    D2.load_state_dict(sd)
    D1, D2 = D1.cuda().train(), D2.cuda().train()
    a, b = torch.rand(n, 3, size, size), torch.rand(n, 3, size, size)
    xa1, xb1 = a.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    xa2, xb2 = a.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    pa1, pb1 = D1.forward_pair(xa1, xb1)
    assert any(len(k) == 3 for k in D1._tsr["plans"]), "the paired plan was not used"
    (pa1.sum() + 2 * pb1.sum()).backward()
    pa2, pb2 = D2(xa2), D2(xb2)
    (pa2.sum() + 2 * pb2.sum()).backward()
    torch.cuda.synchronize()
    r = {"pa": MC.rel_l2(pa1, pa2), "pb": MC.rel_l2(pb1, pb2), "dxa": MC.rel_l2(xa1.grad, xa2.grad),
         "dxb": MC.rel_l2(xb1.grad, xb2.grad)}
    s1, s2 = D1.state_dict(), D2.state_dict()
    for k in s1:
        if "num_batches" in k:
            assert int(s1[k]) == int(s2[k]) == 2, k
        elif "running" in k:
            r["buf"] = max(r.get("buf", 0.0), MC.rel_l2(s1[k], s2[k]))
    g = sorted(MC.rel_l2(p1.grad, p2.grad) for p1, p2 in zip(D1.parameters(), D2.parameters()))
    r["grad_median"], r["grad_worst"] = g[len(g) // 2], g[-1]
    # and against the oracle (two calls, shared buffers dict -> two running-statistics updates, a first)
    osd = MC.O.with_grad(sd)
    buf = {}
    ra, rb = oracle_fn(osd, a, True, buf), oracle_fn(osd, b, True, buf)
    (ra.sum() + 2 * rb.sum()).backward()
    r["oracle_pa"], r["oracle_pb"] = MC.rel_l2(pa1, ra), MC.rel_l2(pb1, rb)
    for k, v in buf.items():
        if "running" in k:
            r["oracle_buf"] = max(r.get("oracle_buf", 0.0), MC.rel_l2(s1[k.lstrip(".")], v))
    ref = {k: v.grad for k, v in osd.items() if v.requires_grad}
    r["oracle_grad_worst"], r["oracle_grad_median"], _ = MC.grad_report(D1, ref)
    print("forward_pair", D_cls.__module__, n, "fused" if fuse else "two-launch", r)
    return r


@pytest.mark.parametrize("fuse", [True, False], ids=["bn-fused", "bn-two-launch"])
def test_forward_pair_equals_two_calls_srgan(fuse, monkeypatch):
    """One discriminator pass over (real | fake) with BatchNorm statistics per half (srgan/trainer.py:446-447 semantics)
    equals two separate calls: outputs, input gradients, parameter gradients, running statistics (two updates, real
    first), num_batches_tracked += 2."""
    MC, SG, SD, EG, ED = _mods()
    r = _pair_case(SD, 8, 96, fuse, monkeypatch, MC.O.srgan_discriminator)
    # Both sides run the same bf16 kernels on different tile grids (M doubles): outputs and statistics agree to 1e-3.
    # Gradients through 7 BatchNorm + LeakyReLU stages at batch 8 are chaotic in bf16 (a rounding flip of one near-zero
    # pre-activation changes a mask): the paired pass sits as close to the two-call pass (median ~0.1) as either sits
    # to the fp32 oracle (~0.15, the level of PyTorch's own bf16 autocast), which is the bound that counts.
    assert r["pa"] <= 5e-3 and r["pb"] <= 5e-3 and r["buf"] <= 2e-3, r
    assert r["grad_median"] <= 0.2 and r["dxa"] <= 0.25 and r["dxb"] <= 0.25, r
    assert r["oracle_pa"] <= 1e-2 and r["oracle_pb"] <= 1e-2 and r["oracle_buf"] <= 1e-2 and r["oracle_grad_median"] <= 0.25, r


@pytest.mark.parametrize("fuse", [True, False], ids=["bn-fused", "bn-two-launch"])
def test_forward_pair_equals_two_calls_esrgan(fuse, monkeypatch):
    MC, SG, SD, EG, ED = _mods()
    r = _pair_case(ED, 4, 128, fuse, monkeypatch, MC.O.esrgan_discriminator)
    assert r["pa"] <= 5e-2 and r["pb"] <= 5e-2 and r["buf"] <= 5e-3, r     # logits: small sums of cancelling terms
    assert r["grad_median"] <= 0.25, r
    # the logits are sums of cancelling terms and the BatchNorm column sums arrive through atomics in a run-dependent
    # order: two runs of the SAME pass differ by up to ~3e-2 here (r["pa"]); measured against the oracle: 3.5e-2 .. 5.1e-2
    assert r["oracle_pa"] <= 7e-2 and r["oracle_pb"] <= 7e-2 and r["oracle_buf"] <= 1e-2 and r["oracle_grad_median"] <= 0.3, r


def test_forward_pair_falls_back_and_detaches_halves():
    """Shapes whose halves do not split on 32-row boundaries (3 x 6 x 6 rows in the last layer) run as two calls;
    grad_halves=(False, True) returns the first half detached (the relativistic generator step, esrgan/trainer.py:463)."""
    MC, SG, SD, EG, ED = _mods()
    torch.manual_seed(22)
    D = SD().cuda().train()
    a, b = torch.rand(3, 3, 96, 96, device="cuda"), torch.rand(3, 3, 96, 96, device="cuda", requires_grad=True)
    pa, pb = D.forward_pair(a, b, grad_halves=(False, True))
    assert not any(len(k) == 3 for k in D._tsr["plans"])
    assert not pa.requires_grad and pb.requires_grad
    D2 = ED().cuda().train()
    a, b = torch.rand(4, 3, 128, 128, device="cuda"), torch.rand(4, 3, 128, 128, device="cuda", requires_grad=True)
    pa, pb = D2.forward_pair(a, b, grad_halves=(False, True))
    assert any(len(k) == 3 for k in D2._tsr["plans"])
    assert not pa.requires_grad and pb.requires_grad
    pb.sum().backward()
    assert b.grad is not None and float(b.grad.abs().sum()) > 0


def test_loss_kernels_match_torch():
    """losses.* (loss_kernel / gan_loss_kernel / axpby_f32_kernel) against torch.nn.functional: values and gradients."""
    import torch.nn.functional as F
    from torchsr_b200 import losses
    torch.manual_seed(23)
    dev = "cuda"

    def both(fn_ours, fn_ref, *shapes, prob=False):
        xs = [torch.rand(s, device=dev) * 0.98 + 0.01 if prob else torch.randn(s, device=dev) for s in shapes]
        a1 = [x.clone().requires_grad_(True) for x in xs]
        a2 = [x.clone().requires_grad_(True) for x in xs]
        up = torch.tensor(0.7, device=dev)
        l1, l2 = fn_ours(*a1), fn_ref(*a2)
        l1.backward(up)
        l2.backward(up)
        assert abs(float(l1) - float(l2)) <= 1e-5 * max(1.0, abs(float(l2))), (float(l1), float(l2))
        for p, q in zip(a1, a2):
            if q.grad is None:
                assert p.grad is None
            else:
                assert torch.allclose(p.grad, q.grad, rtol=1e-4, atol=1e-7), (p.grad - q.grad).abs().max()

    t = torch.rand(5, 3, 17, 13, device=dev)
    both(lambda x: losses.mse(x, t), lambda x: F.mse_loss(x, t), (5, 3, 17, 13))
    both(lambda x: losses.l1(x, t, scale=0.01), lambda x: 0.01 * F.l1_loss(x, t), (5, 3, 17, 13))
    one = lambda x: torch.ones_like(x)  # noqa: E731
    zero = lambda x: torch.zeros_like(x)  # noqa: E731
    both(lambda p, q: losses.bce(p, 1.0, q, 0.0),
         lambda p, q: F.binary_cross_entropy(p, one(p)) + F.binary_cross_entropy(q, zero(q)), (16, 1), (16, 1), prob=True)
    both(lambda p: losses.bce(p, 1.0, scale=0.001), lambda p: 0.001 * F.binary_cross_entropy(p, one(p)), (7, 1), prob=True)
    lab = (torch.rand(9, 1, device=dev) > 0.5).float()
    both(lambda p: losses.BCELoss()(p, lab), lambda p: F.binary_cross_entropy(p, lab), (9, 1), prob=True)
    both(lambda x: losses.BCEWithLogitsLoss()(x, lab), lambda x: F.binary_cross_entropy_with_logits(x, lab), (9, 1))
    both(lambda r, f: losses.relativistic_d(r, f, scale=0.5),
         lambda r, f: (F.binary_cross_entropy_with_logits(r - f.mean(), one(r)) +
                       F.binary_cross_entropy_with_logits(f - r.mean(), zero(f))) / 2, (16, 1), (16, 1))
    rconst = torch.randn(16, 1, device=dev)
    both(lambda f: losses.relativistic_g(f, rconst, scale=0.005),
         lambda f: 0.005 * F.binary_cross_entropy_with_logits(f - rconst.mean(), one(f)), (16, 1))
    both(lambda x, y: losses.total(losses.mse(x, t), losses.l1(y, t), losses.mse(y, t, scale=2.0)),
         lambda x, y: F.mse_loss(x, t) + F.l1_loss(y, t) + 2.0 * F.mse_loss(y, t), (5, 3, 17, 13), (5, 3, 17, 13))
    # saturated probabilities: PyTorch clamps log at -100 and the gradient denominator at 1e-12
    p = torch.tensor([[0.0], [1.0], [1e-30], [0.5]], device=dev)
    both(lambda q: losses.bce(q * 0 + p, 1.0), lambda q: F.binary_cross_entropy(q * 0 + p, one(p)), (4, 1), prob=True)


def test_graph_step_survives_an_eager_ragged_batch():
    """FusedAdam's device table is referenced by the captured step graph; an eager step on another batch
    shape (ragged last batch -> other gradient buffers) must not free or rewrite it. graph, ragged eager, graph again
    must match an eager-only trainer."""
    import os
    from argparse import Namespace
    os.environ.setdefault("TORCHSR_VGG_WEIGHTS", "random")
    from torchsr_b200.srgan.trainer import SRGANTrainer

    def make():
        torch.manual_seed(31)
        a = Namespace(disable_amp=False, batch_size=8, epochs=8, pretrain_epochs=1, gan_checkpoint=None,
                      psnr_checkpoint=None, skip_image_save=True, local_rank=0, rank=-1, world_size=1)
        return SRGANTrainer(torch.device("cuda"), a, [], [], 0, 0, False)

    g = torch.Generator().manual_seed(5)
    batches = [(torch.rand(n, 3, 24, 24, generator=g).cuda(), torch.rand(n, 3, 96, 96, generator=g).cuda())
               for n in (8, 8, 5, 8, 8)]
    tr = make()
    from torchsr_b200 import ops
    for i, (lr, hr) in enumerate(batches):
        before = [p.detach().clone() for p in tr.generator.parameters()] + \
                 [p.detach().clone() for p in tr.discriminator.parameters()]
        loss = tr._train_step('gan', lr, hr, i)       # graph for full batches, eager for the ragged one
        torch.cuda.synchronize()
        ops.check_watchdog()
        assert torch.isfinite(loss).all(), (i, loss)
        after = list(tr.generator.parameters()) + list(tr.discriminator.parameters())
        moved = 0.0
        for b, a in zip(before, after):
            assert torch.isfinite(a).all(), i
            moved = max(moved, float((a - b).abs().max()))
        # Adam with lr 1e-4: |update| <= lr * (1 - beta1) / sqrt(1 - beta2) ~ 3.2e-4 per step; garbage pointers would
        # move (or not move) the parameters by anything
        assert 1e-5 <= moved <= 4e-4, (i, moved)
    for opt in (tr.disc_optimizer, tr.gen_optimizer):
        variants = opt._tables[0]
        assert len(variants) == 2, len(variants)       # full-batch plan and ragged-batch plan gradient buffers
        for st in variants.values():
            assert torch.equal(st["table"].cpu(), st["table_host"]), "a device table was rewritten or freed"
            assert float(st["step"]) == 5.0, float(st["step"])


@BN_PATHS
def test_eval_mode_no_grad_and_large_image_psnr(bn_fuse):
    """Eval-mode inference on a non-square image: output PSNR against a target agrees with the oracle within 0.05 dB."""
    MC, SG, SD, EG, ED = _mods()
    torch.manual_seed(8)
    G = SG()
    MC.randomize_bn(G)
    sd = {k: v.clone() for k, v in G.state_dict().items()}
    G = G.cuda().eval()
    x = torch.rand(1, 3, 40, 56)
    with torch.no_grad():
        y = G(x.cuda()).cpu()
        ref = MC.O.srgan_generator(sd, x, False)
    assert y.shape == (1, 3, 160, 224)
    assert MC.rel_l2(y, ref) <= 3e-2
    target = torch.nn.functional.interpolate(x, scale_factor=4, mode="bicubic").clamp(0, 1)
    scale = 0.05 / max(ref.std().item(), 1e-6)      # bring the random-init output into image range around the target
    p_ours = MC.O.psnr(target + scale * (y - y.mean()), target)
    p_ref = MC.O.psnr(target + scale * (ref - ref.mean()), target)
    assert abs(p_ours - p_ref) <= 0.05, (p_ours, p_ref)
    assert not any(p.busy for pool in G._tsr["plans"].values() for p in pool)


def test_large_image_inference_staged_halo_paths_match_oracle_and_generic_kernels(monkeypatch):
    """A 160x192 image is large enough for the persistent kernels: trunk convs run the halo-patch main loop with the
    staged (bulk tensor store) epilogue, the sub-pixel convs its PixelShuffle variant, the 9x9 output conv sums its
    horizontal taps in the epilogue. Checked against the oracle and against the generic kernels (both switches off)."""
    MC, SG, SD, EG, ED = _mods()
    torch.manual_seed(21)
    G = SG()
    MC.randomize_bn(G)
    sd = {k: v.clone() for k, v in G.state_dict().items()}
    x = torch.rand(1, 3, 160, 192)
    with torch.no_grad():
        ref = MC.O.srgan_generator(sd, x, False)
        y_new = G.cuda().eval()(x.cuda()).cpu()
        monkeypatch.setenv("TSR_CONV_STAGED", "0")
        monkeypatch.setenv("TSR_CONV_HALO", "0")
        G2 = SG()
        G2.load_state_dict(sd)
        y_old = G2.cuda().eval()(x.cuda()).cpu()
    assert y_new.shape == (1, 3, 640, 768)
    e_new, e_old, d = MC.rel_l2(y_new, ref), MC.rel_l2(y_old, ref), MC.rel_l2(y_new, y_old)
    print("large-image inference vs oracle: staged/halo/in-place", e_new, "generic kernels", e_old, "between them", d)
    assert e_new <= 3e-2 and e_old <= 3e-2, (e_new, e_old)
    # same bf16 store points and fp32 accumulation per tile, except that the in-place residual blocks of the inference
    # plan add their tile to the block input in global memory (bulk reduce store): the tile is rounded to bf16 before
    # the add, the generic path adds in fp32 and rounds once - one extra 2^-9 rounding of the branch per block
    assert d <= 1e-2 and e_new <= 1.25 * e_old + 2e-3, (e_new, e_old, d)


@pytest.mark.parametrize("shape", [(1, 3, 37, 53), (2, 3, 64, 100), (1, 3, 131, 259)], ids=lambda s: "x".join(map(str, s)))
def test_inference_odd_image_sizes(shape):
    """Eval-mode inference on sizes that divide nothing: partial halo strips / row tiles, accumulator tiles that straddle
    image rows in the two-rows-per-GEMM-row output conv, M not a multiple of 128, batch of two images."""
    MC, SG, SD, EG, ED = _mods()
    torch.manual_seed(24)
    G = SG()
    MC.randomize_bn(G)
    sd = {k: v.clone() for k, v in G.state_dict().items()}
    x = torch.rand(*shape)
    with torch.no_grad():
        ref = MC.O.srgan_generator(sd, x, False)
        y = G.cuda().eval()(x.cuda()).cpu()
    assert y.shape == ref.shape and torch.isfinite(y).all()
    assert MC.rel_l2(y, ref) <= 3e-2, MC.rel_l2(y, ref)


def test_inference_plan_packs_follow_the_weights():
    """The inference plans keep their own packs of the two 9x9 layers (row-decomposed input conv, two output rows per GEMM
    row in the output conv), derived from the fp32 parameters by a kernel of the plan's own launch list: a changed
    weight must show in the very next eval forward, also when it changed on the device only (no version bump)."""
    MC, SG, SD, EG, ED = _mods()
    torch.manual_seed(23)
    G = SG().cuda().eval()
    x = torch.rand(1, 3, 40, 48, device="cuda")

    def fresh():
        H = SG().cuda().eval()
        H.load_state_dict(G.state_dict())
        with torch.no_grad():
            return H(x)

    with torch.no_grad():
        y0 = G(x)
        assert MC.rel_l2(y0, fresh()) <= 1e-6
        G.conv3.weight.mul_(0.5)                       # version bump
        y1 = G(x)
        assert MC.rel_l2(y1, fresh()) <= 1e-6 and MC.rel_l2(y1, y0) > 1e-2
        G.conv3.weight.data.view(-1)[:].mul_(2.0)      # in place on the device through .data: no version bump
        G.conv1[0].weight.data.mul_(0.5)
        y2 = G(x)
        assert MC.rel_l2(y2, fresh()) <= 1e-6 and MC.rel_l2(y2, y1) > 1e-2
        ref = MC.O.srgan_generator({k: v.detach().cpu() for k, v in G.state_dict().items()}, x.cpu(), False)
    assert MC.rel_l2(y2.cpu(), ref) <= 3e-2


def test_upscale_pipelined_equals_sequential_upscale():
    from torchsr_b200.test import upscale, upscale_pipelined
    MC, SG, SD, EG, ED = _mods()
    torch.manual_seed(22)
    G = SG().cuda()
    xs = [torch.rand(1, 3, 24, 32).pin_memory() for _ in range(5)]
    ys = [torch.empty(1, 3, 96, 128).pin_memory() for _ in range(5)]
    upscale_pipelined(G, xs, ys)
    for x, y in zip(xs, ys):
        assert torch.equal(y, upscale(G, x.cuda()).cpu())
    y8 = [torch.empty(1, 3, 96, 128, dtype=torch.uint8).pin_memory() for _ in range(5)]
    upscale_pipelined(G, xs, y8)
    for y, q in zip(ys, y8):        # torchvision.utils.save_image's quantisation
        assert torch.equal(q, y.clone().mul_(255).add_(0.5).clamp_(0, 255).to(torch.uint8))


def test_gan_step_matches_oracle_step():
    """One SRGANTrainer._gan_loop on the B200 against the oracle port of the reference step: same losses (MSE content
    loss stands in for VGG, which is executed by PyTorch on both sides) and the same direction of the Adam update."""
    import step_oracle as S
    from argparse import Namespace
    from torchsr_b200.srgan.trainer import SRGANTrainer
    import os
    os.environ["TORCHSR_VGG_WEIGHTS"] = "random"
    torch.manual_seed(9)
    args = Namespace(disable_amp=False, batch_size=4, epochs=8, pretrain_epochs=1, gan_checkpoint=None,
                     psnr_checkpoint=None, skip_image_save=True, local_rank=0, rank=-1, world_size=1)
    tr = SRGANTrainer(torch.device("cuda"), args, [], [], 0, 0, False)
    tr.vgg_loss = lambda a, b: torch.nn.functional.mse_loss(a, b)
    g_sd = {k: v.detach().cpu().clone() for k, v in tr.generator.state_dict().items()}
    d_sd = {k: v.detach().cpu().clone() for k, v in tr.discriminator.state_dict().items()}
    lr, hr = torch.rand(4, 3, 24, 24), torch.rand(4, 3, 96, 96)
    gen_loss = float(tr._gan_loop(lr.cuda(), hr.cuda(), 0))
    o = S.OracleSRGAN(g_sd, d_sd, None)
    _, ref_gen_loss = o.gan_step(lr, hr)
    assert abs(gen_loss - ref_gen_loss) <= 2e-2 * abs(ref_gen_loss) + 1e-4, (gen_loss, ref_gen_loss)
    # Adam's first step moves every weight by ~lr * sign(grad): compare the update directions
    agree, total = 0, 0
    new_g = tr.generator.state_dict()
    for k, v in o.g.items():
        if not v.requires_grad or v.numel() < 64:
            continue
        du = (new_g[k].cpu() - g_sd[k]).flatten()
        dr = (v.detach() - g_sd[k]).flatten()
        agree += int((torch.sign(du) == torch.sign(dr)).sum())
        total += du.numel()
    assert agree / total >= 0.85, agree / total


def test_packed_weights_follow_fused_optimizer_and_graph_replay():
    """torch.optim.Adam(fused=True) does not bump Tensor._version: the bf16 operand copies must still be re-packed
    after every optimizer step - eagerly and inside a replayed whole-step CUDA graph (regression: stale packs)."""
    MC, SG, SD, EG, ED = _mods()
    torch.manual_seed(11)
    G = SG().cuda()
    x = torch.rand(2, 3, 24, 24, device="cuda")
    opt = torch.optim.Adam(G.parameters(), lr=torch.tensor(1e-2, device="cuda"), fused=True, capturable=True)

    def step():
        opt.zero_grad()
        y = G(x)
        y.square().mean().backward()
        opt.step()
        return y

    def fresh_forward():
        # an independent module instance holding the current weights packs them from scratch
        H = SG().cuda()
        H.load_state_dict(G.state_dict())
        with torch.no_grad():
            return H(x)

    y0 = step().detach().clone()
    y1 = step().detach().clone()          # must see the weights of the first update
    assert MC.rel_l2(y1, y0) > 1e-2, "the optimizer step did not reach the kernels (stale packed weights)"
    with torch.no_grad():
        now = G(x)
    assert MC.rel_l2(now, fresh_forward()) <= 2e-3
    # whole step captured into a CUDA graph and replayed
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = step()
    outs = []
    for _ in range(3):
        graph.replay()
        outs.append(out.detach().clone())
    torch.cuda.synchronize()
    assert MC.rel_l2(outs[2], outs[0]) > 1e-3, "replays do not see updated weights"
    with torch.no_grad():
        now = G(x)
    assert MC.rel_l2(now, fresh_forward()) <= 2e-3


def test_vgg_content_loss_vs_torch_fp32():
    """VGG19 perceptual loss on this repo's kernels (bf16, fused ReLU, max-pool kernels, data gradients only) against
    the same torchvision network in fp32 on the CPU: loss value, features and the gradient w.r.t. the image."""
    import os
    import module_checks as MC
    os.environ["TORCHSR_VGG_WEIGHTS"] = "random"
    from torchsr_b200.srgan.loss import VGGLoss
    torch.manual_seed(12)
    x, t = torch.rand(2, 3, 96, 96), torch.rand(2, 3, 96, 96)
    ref_mod = VGGLoss()
    xr = x.clone().requires_grad_(True)
    with torch.no_grad():
        t_feats = ref_mod.features(t)
    ref = torch.nn.functional.l1_loss(ref_mod.features(xr), t_feats)     # reference loss.py:52-53 in plain torch fp32
    ref.backward()
    mod = VGGLoss().cuda()           # seeded initialisation: identical weights
    xg = x.cuda().requires_grad_(True)
    got = mod(xg, t.cuda())
    got.backward()
    torch.cuda.synchronize()
    feats = mod.target_features(x.cuda())
    with torch.no_grad():
        ref_feats = ref_mod.features(x)
    errs = {"loss": abs(float(got) - float(ref)) / abs(float(ref)), "features": MC.rel_l2(feats, ref_feats),
            "dx": MC.rel_l2(xg.grad, xr.grad)}
    print("vgg", errs)
    assert errs["features"] <= 3e-2 and errs["loss"] <= 3e-2, errs
    # The image gradient passes 16 ReLU masks and the sign() of the L1 loss in bf16: PyTorch's own bf16 path (cuDNN,
    # the former TORCHSR_VGG_IMPL=torch path) measured rel-L2 0.416 / cosine 0.913 against fp32 on these inputs, this path 0.410 / 0.916
    # (measured in round 1); the per-kernel arithmetic is pinned to 1e-4 in test_kernels_gpu.py.
    cos = float(torch.nn.functional.cosine_similarity(xg.grad.flatten().cpu(), xr.grad.flatten(), dim=0))
    assert errs["dx"] <= 0.5 and cos >= 0.88, (errs, cos)
    assert any(p._version >= 0 for p in mod.parameters()) and mod._b200 is not None


def test_fused_adam_matches_torch_adam_and_keeps_packs_fresh():
    """torchsr_b200.optim.FusedAdam against torch.optim.Adam on identical gradients over several steps: parameters
    and state agree to fp32 rounding, and the bf16 operand copies written by the optimizer kernel equal a fresh pack."""
    import copy
    MC, SG, SD, EG, ED = _mods()
    from torchsr_b200.optim import FusedAdam
    for make, shape in ((SG, (2, 3, 24, 24)), (SD, (2, 3, 96, 96))):
        torch.manual_seed(13)
        A = make().cuda()
        B = copy.deepcopy(A)
        x = torch.rand(*shape, device="cuda")
        oa = FusedAdam(A.parameters(), lr=torch.tensor(1e-3, device="cuda"))
        ob = torch.optim.Adam(B.parameters(), lr=1e-3)
        for step in range(3):
            oa.zero_grad()
            A(x).square().mean().backward()
            # same gradients on both sides: the comparison is about the update rule, not the kernels' bf16 noise
            for pa, pb in zip(A.parameters(), B.parameters()):
                pb.grad = pa.grad.detach().clone()
            oa.step()
            ob.step()
            B.load_state_dict({**B.state_dict(), **{k: v for k, v in A.state_dict().items() if "running" in k or "num_batches" in k}})
        worst = max(MC.rel_l2(pa, pb) for pa, pb in zip(A.parameters(), B.parameters()))
        assert worst <= 1e-5, worst
        for pa, pb in zip(A.parameters(), B.parameters()):
            assert MC.rel_l2(oa.state[pa]["exp_avg_sq"], ob.state[pb]["exp_avg_sq"]) <= 1e-5
        # packs written by the optimizer == packs made from scratch by a fresh module holding the same weights
        H = make().cuda()
        H.load_state_dict(A.state_dict())
        A.eval(), H.eval()
        with torch.no_grad():
            assert MC.rel_l2(A(x), H(x)) <= 1e-6


def test_esrgan_trainer_steps_eager_and_graph():
    """ESRGANTrainer (relativistic GAN step, torchsr/esrgan/trainer.py:418-484) end to end on the B200 path: pretrain
    step, eager GAN steps and the same step replayed as a CUDA graph - finite losses, both networks updated, and the
    replayed step continues the eager trajectory (same loss scale)."""
    import os
    from argparse import Namespace
    os.environ["TORCHSR_VGG_WEIGHTS"] = "random"
    from torchsr_b200.esrgan.trainer import ESRGANTrainer
    torch.manual_seed(21)
    args = Namespace(disable_amp=False, batch_size=2, epochs=8, pretrain_epochs=1, gan_checkpoint=None,
                     psnr_checkpoint=None, skip_image_save=True, local_rank=0, rank=-1, world_size=1)
    tr = ESRGANTrainer(torch.device("cuda"), args, [], [], 0, 0, False)
    lr, hr = torch.rand(2, 3, 32, 32, device="cuda"), torch.rand(2, 3, 128, 128, device="cuda")
    g0 = {k: v.detach().clone() for k, v in tr.generator.state_dict().items()}
    d0 = {k: v.detach().clone() for k, v in tr.discriminator.state_dict().items()}
    l_pre = float(tr._pretrain_step(lr, hr))
    losses = [float(tr._gan_loop(lr, hr, s)) for s in range(2)]
    losses += [float(tr.graph_step(lr, hr, s)) for s in range(3)]
    torch.cuda.synchronize()
    assert all(l == l and abs(l) < 1e3 for l in [l_pre] + losses), (l_pre, losses)
    assert max(losses) <= 3 * min(losses) + 1e-3, losses
    moved_g = sum(int(not torch.equal(v, g0[k])) for k, v in tr.generator.state_dict().items())
    moved_d = sum(int(not torch.equal(v, d0[k])) for k, v in tr.discriminator.state_dict().items() if "num_batches" not in k)
    assert moved_g >= len(g0) - 2 and moved_d >= len(d0) // 2, (moved_g, len(g0), moved_d, len(d0))


def _trainer(kind, batch):
    import os
    from argparse import Namespace
    os.environ["TORCHSR_VGG_WEIGHTS"] = "random"     # seeded (1234) random VGG19: the oracle builds the same one
    args = Namespace(disable_amp=False, batch_size=batch, epochs=8, pretrain_epochs=1, gan_checkpoint=None,
                     psnr_checkpoint=None, skip_image_save=True, local_rank=0, rank=-1, world_size=1)
    if kind == "srgan":
        from torchsr_b200.srgan.trainer import SRGANTrainer as T
    else:
        from torchsr_b200.esrgan.trainer import ESRGANTrainer as T
    return T(torch.device("cuda"), args, [], [], 0, 0, False)


def _cpu_state(module):
    return {k: v.detach().cpu().clone() for k, v in module.state_dict().items()}


def _update_agreement(new_sd, old_sd, oracle_sd, min_numel=64):
    """Adam's first step moves every weight by ~lr * sign(grad): fraction of weights moved in the oracle's direction."""
    agree, total = 0, 0
    for k, v in oracle_sd.items():
        if not v.requires_grad or v.numel() < min_numel:
            continue
        du = (new_sd[k].cpu() - old_sd[k]).flatten()
        dr = (v.detach() - old_sd[k]).flatten()
        agree += int((torch.sign(du) == torch.sign(dr)).sum())
        total += du.numel()
    return agree / max(total, 1)


@pytest.mark.parametrize("kind", ["srgan", "esrgan"])
def test_pretrain_step_matches_oracle(kind):
    """PSNR-phase step (srgan/trainer.py:376-388 MSE; esrgan/trainer.py:378-390 L1) through `_pretrain_step` on the
    B200 against the oracle step pinned by tests/golden/*_pretrain_step.npz: same loss, same update direction."""
    import step_oracle as S
    torch.manual_seed(31)
    tr = _trainer(kind, 2)
    g_sd = _cpu_state(tr.generator)
    s = 24 if kind == "srgan" else 32
    lr, hr = torch.rand(2, 3, s, s), torch.rand(2, 3, 4 * s, 4 * s)
    loss = float(tr._pretrain_step(lr.cuda(), hr.cuda()))
    o = (S.OracleSRGAN if kind == "srgan" else S.OracleESRGAN)(g_sd, {}, None)
    ref = o.pretrain_step(lr, hr)
    assert abs(loss - ref) <= 2e-2 * abs(ref) + 1e-4, (loss, ref)
    frac = _update_agreement(tr.generator.state_dict(), g_sd, o.g)
    assert frac >= 0.85, frac


def test_esrgan_gan_step_matches_oracle_step():
    """One ESRGANTrainer._gan_loop (relativistic GAN step, esrgan/trainer.py:435-484) on the B200, WITH the VGG19
    perceptual loss on this repo's kernels, against the oracle step (pinned by tests/golden/esrgan_gan_step.npz):
    generator loss within 5 % (23 RRDB blocks + 16 VGG convs of bf16 roundings) and the same Adam update direction
    for both networks."""
    import step_oracle as S
    torch.manual_seed(41)
    tr = _trainer("esrgan", 2)
    g_sd, d_sd = _cpu_state(tr.generator), _cpu_state(tr.discriminator)
    lr, hr = torch.rand(2, 3, 32, 32), torch.rand(2, 3, 128, 128)
    gen_loss = float(tr._gan_loop(lr.cuda(), hr.cuda(), 0))
    o = S.OracleESRGAN(g_sd, d_sd, S.vgg19_features(1234))
    _, ref = o.gan_step(lr, hr)
    assert abs(gen_loss - ref) <= 5e-2 * abs(ref) + 1e-4, (gen_loss, ref)
    fg = _update_agreement(tr.generator.state_dict(), g_sd, o.g)
    fd = _update_agreement(tr.discriminator.state_dict(), d_sd, o.d)
    assert fg >= 0.8 and fd >= 0.8, (fg, fd)


def test_srgan_gan_step_with_vgg_matches_oracle_step():
    """The headline step exactly as bench.py runs it (VGG19 loss on this repo's kernels, seeded random VGG weights on
    both sides) against the oracle step pinned by tests/golden/srgan_gan_step.npz."""
    import step_oracle as S
    torch.manual_seed(51)
    tr = _trainer("srgan", 4)
    g_sd, d_sd = _cpu_state(tr.generator), _cpu_state(tr.discriminator)
    lr, hr = torch.rand(4, 3, 24, 24), torch.rand(4, 3, 96, 96)
    gen_loss = float(tr._gan_loop(lr.cuda(), hr.cuda(), 0))
    o = S.OracleSRGAN(g_sd, d_sd, S.vgg19_features(1234))
    _, ref = o.gan_step(lr, hr)
    assert abs(gen_loss - ref) <= 5e-2 * abs(ref) + 1e-4, (gen_loss, ref)
    fg = _update_agreement(tr.generator.state_dict(), g_sd, o.g)
    fd = _update_agreement(tr.discriminator.state_dict(), d_sd, o.d)
    assert fg >= 0.8 and fd >= 0.8, (fg, fd)


def _write_png(path, h, w, seed):
    """Smooth synthetic RGB image (low-frequency noise), written as PNG."""
    import numpy as np
    from PIL import Image
    rng = np.random.Generator(np.random.PCG64(seed))
    small = torch.from_numpy(rng.uniform(0, 1, (1, 3, h // 8 + 2, w // 8 + 2)).astype("float32"))
    img = torch.nn.functional.interpolate(small, size=(h, w), mode="bicubic", align_corners=False).clamp(0, 1)
    arr = (img[0].permute(1, 2, 0).numpy() * 255 + 0.5).astype("uint8")
    Image.fromarray(arr).save(path)


def _read_png(path):
    from PIL import Image
    from torchvision.transforms import ToTensor
    return ToTensor()(Image.open(path).convert("RGB")).unsqueeze(0)


def test_cli_train_then_test_end_to_end(tmp_path, monkeypatch):
    """`torchsr train` (1 pretrain + 2 GAN epochs on a tiny synthetic image folder, whole-step CUDA graphs in the epoch
    loops) writes the reference's checkpoint files ({"epoch","phase","state"}, torchsr/srgan/trainer.py:305-343);
    `torchsr test` then loads `srgan-gan-best.pth` and writes `upres-<image>` at x4 (torchsr/test.py:22-63). The image
    is compared with the oracle generator run on the same checkpoint."""
    import os
    import torchsr_oracle as O
    from torchsr_b200 import torchsr as cli
    monkeypatch.setenv("TORCHSR_VGG_WEIGHTS", "random")
    monkeypatch.chdir(tmp_path)
    data = tmp_path / "data"
    data.mkdir()
    for i in range(12):
        _write_png(str(data / f"img{i:02d}.png"), 128, 160, 100 + i)
    torch.manual_seed(61)
    cli.main(["train", "--train-dir", str(data), "--batch-size", "2", "--epochs", "2", "--pretrain-epochs", "1",
              "--data-workers", "0", "--skip-image-save", "--model", "srgan"])
    for name in ["srgan-psnr-best.pth", "srgan-psnr-latest.pth", "srgan-gan-best.pth", "srgan-gan-latest.pth"]:
        ck = torch.load(str(tmp_path / name), map_location="cpu")
        assert set(ck) == {"epoch", "phase", "state"} and ck["phase"] == name.rsplit("-", 1)[0], (name, ck.keys())
        assert all(torch.isfinite(v.float()).all() for v in ck["state"].values()), name
    _write_png(str(tmp_path / "lowres.png"), 48, 64, 7)
    cli.main(["test", "lowres.png", "--model", "srgan"])
    ours = _read_png(str(tmp_path / "upres-lowres.png"))
    assert ours.shape == (1, 3, 192, 256)
    state = {k: v.float() if v.is_floating_point() else v
             for k, v in torch.load(str(tmp_path / "srgan-gan-best.pth"), map_location="cpu")["state"].items()}
    with torch.no_grad():
        ref = O.srgan_generator(state, _read_png(str(tmp_path / "lowres.png")), False).clamp(0, 1)
    ref8 = (ref * 255 + 0.5).floor() / 255          # torchvision.utils.save_image quantisation
    assert O.psnr(ours, ref8) >= 30.0, O.psnr(ours, ref8)


def test_c1_fixture_shape_inference_psnr_both_bn_modes():
    """BASELINE configs[0] shape: one 320x480 image -> 1280x1920 through `test.upscale` with trained-like synthetic
    weights, in eval mode (this repo's default) and in the reference's train-mode-BatchNorm behaviour (torchsr/test.py
    never calls eval()). Output PSNR against a x4 bicubic target agrees with the oracle within
    0.05 dB (north star)."""
    import torchsr_oracle as O
    from golden_util import synth_input, synth_state_dict
    from torchsr_b200.srgan.generator import Generator
    from torchsr_b200.test import upscale
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    G = Generator()
    sd = synth_state_dict(G.state_dict(), 71)
    G.load_state_dict(sd)
    G = G.cuda()
    x = synth_input((1, 3, 320, 480), 72)
    target = torch.nn.functional.interpolate(x, scale_factor=4, mode="bicubic").clamp(0, 1)
    for train_mode in (False, True):
        y = upscale(G, x.cuda(), train_mode=train_mode).float().cpu()
        with torch.no_grad():
            ref = O.srgan_generator(sd, x, train_mode)
        assert y.shape == (1, 3, 1280, 1920)
        assert float((y - ref).norm() / ref.norm()) <= 3e-2, train_mode
        p_ours, p_ref = O.psnr(y.clamp(0, 1), target), O.psnr(ref.clamp(0, 1), target)
        assert abs(p_ours - p_ref) <= 0.05, (train_mode, p_ours, p_ref)
