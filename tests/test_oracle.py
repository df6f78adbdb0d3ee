"""The oracle (oracle/torchsr_oracle.py, step_oracle.py) against the golden vectors recorded from the unmodified
reference (tests/golden/make_golden.py). CPU only."""
import pytest
import torch

import step_oracle as S
import torchsr_oracle as O
from golden_util import digest, digest_close, load_fixture, synth_input, synth_state_dict, template_from_meta

CASES = [
    ("srgan_generator", lambda sd, x, b: O.srgan_generator(sd, x, True, b)),
    ("srgan_discriminator", lambda sd, x, b: O.srgan_discriminator(sd, x, True, b)),
    ("esrgan_generator", lambda sd, x, b: O.esrgan_generator(sd, x)),
    ("esrgan_discriminator", lambda sd, x, b: O.esrgan_discriminator(sd, x, True, b)),
]


@pytest.mark.parametrize("name,fn", CASES, ids=[c[0] for c in CASES])
def test_oracle_reproduces_reference_vectors(name, fn):
    torch.set_num_threads(4)
    meta, arr = load_fixture(name)
    sd = O.with_grad(synth_state_dict(template_from_meta(meta), meta["seed"]))
    buffers = {}
    y = fn(sd, arr["input"], buffers)
    assert y.shape == arr["output"].shape
    assert torch.allclose(y, arr["output"], rtol=1e-4, atol=1e-5), float((y - arr["output"]).abs().max())
    y.backward(arr["gout"])
    for k, ref in meta["grad_digest"].items():
        assert digest_close(digest(sd[k].grad), ref, rtol=2e-3), (k, digest(sd[k].grad), ref)
    for k, ref in meta["buffers_after"].items():
        assert digest_close(digest(buffers[k].float()), ref, rtol=1e-4), (k, digest(buffers[k].float()), ref)


def test_oracle_gan_step_reproduces_reference_trainer():
    """One full _gan_loop: post-step generator and discriminator parameters / buffers match the reference trainer."""
    torch.set_num_threads(4)
    meta, arr = load_fixture("srgan_gan_step")
    from torchsr_b200.srgan.discriminator import Discriminator
    from torchsr_b200.srgan.generator import Generator
    g_sd = synth_state_dict(Generator().state_dict(), 11)
    d_sd = synth_state_dict(Discriminator().state_dict(), 12)
    o = S.OracleSRGAN(g_sd, d_sd, S.vgg19_features(1234))
    o.gan_step(arr["low_res"], arr["high_res"])
    for k, ref in meta["g_after"].items():
        assert digest_close(digest(o.g[k].float()), ref, rtol=2e-4, atol=1e-6), (k, digest(o.g[k].float()), ref)
    for k, ref in meta["d_after"].items():
        assert digest_close(digest(o.d[k].float()), ref, rtol=2e-4, atol=1e-6), (k, digest(o.d[k].float()), ref)


def test_oracle_esrgan_gan_step_reproduces_reference_trainer():
    """One full ESRGANTrainer._gan_loop (relativistic GAN, 23 RRDB blocks): post-step parameters and BatchNorm buffers
    of the oracle step equal the reference trainer's."""
    torch.set_num_threads(8)
    meta, arr = load_fixture("esrgan_gan_step")
    from torchsr_b200.esrgan.discriminator import Discriminator
    from torchsr_b200.esrgan.generator import Generator
    g_sd = synth_state_dict(Generator().state_dict(), 21)
    d_sd = synth_state_dict(Discriminator().state_dict(), 22)
    o = S.OracleESRGAN(g_sd, d_sd, S.vgg19_features(1234))
    o.gan_step(arr["low_res"], arr["high_res"])
    for k, ref in meta["g_after"].items():
        assert digest_close(digest(o.g[k].float()), ref, rtol=2e-4, atol=1e-6), (k, digest(o.g[k].float()), ref)
    for k, ref in meta["d_after"].items():
        assert digest_close(digest(o.d[k].float()), ref, rtol=2e-4, atol=1e-6), (k, digest(o.d[k].float()), ref)


@pytest.mark.parametrize("kind", ["srgan", "esrgan"])
def test_oracle_pretrain_steps_reproduce_reference(kind):
    """Two PSNR-phase steps (srgan/trainer.py:376-388 MSE, esrgan/trainer.py:378-390 L1): losses and post-step
    generator state equal the reference trainer's own modules / optimizer driven through the same statements."""
    torch.set_num_threads(8)
    meta, arr = load_fixture(f"{kind}_pretrain_step")
    if kind == "srgan":
        from torchsr_b200.srgan.generator import Generator
        o = S.OracleSRGAN(synth_state_dict(Generator().state_dict(), 31), {}, None)
    else:
        from torchsr_b200.esrgan.generator import Generator
        o = S.OracleESRGAN(synth_state_dict(Generator().state_dict(), 41), {}, None)
    losses = [o.pretrain_step(arr["low_res"], arr["high_res"]) for _ in range(2)]
    for a, b in zip(losses, meta["losses"]):
        assert abs(a - b) <= 1e-4 * abs(b), (losses, meta["losses"])
    for k, ref in meta["g_after"].items():
        assert digest_close(digest(o.g[k].float()), ref, rtol=2e-4, atol=1e-6), (k, digest(o.g[k].float()), ref)


def test_state_dict_contract_matches_reference_shapes():
    """Keys, shapes and dtypes of the drop-in modules equal the reference's (recorded in the fixtures)."""
    from torchsr_b200.esrgan.discriminator import Discriminator as ED
    from torchsr_b200.esrgan.generator import Generator as EG
    from torchsr_b200.srgan.discriminator import Discriminator as SD
    from torchsr_b200.srgan.generator import Generator as SG
    for name, mod in [("srgan_generator", SG()), ("srgan_discriminator", SD()),
                      ("esrgan_generator", EG(num_rrdb_blocks=2)), ("esrgan_discriminator", ED())]:
        meta, _ = load_fixture(name)
        sd = mod.state_dict()
        assert list(sd.keys()) == list(meta["shapes"].keys()), name
        for k, v in sd.items():
            assert list(v.shape) == meta["shapes"][k], (name, k)
            assert str(v.dtype) == meta["dtypes"][k], (name, k)


def test_psnr_definition():
    a = torch.full((1, 3, 4, 4), 0.5)
    b = torch.full((1, 3, 4, 4), 0.6)
    assert abs(O.psnr(a, b) - 20.0) < 1e-4
