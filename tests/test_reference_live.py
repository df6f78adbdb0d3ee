"""Cross-check of the oracle against the live reference modules; runs only where /root/reference exists (the build
container), never on the GPU box."""
import os
import sys

import pytest
import torch

REF = os.environ.get("TORCHSR_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "torchsr")), reason="reference tree not present")


def _ref():
    # the reference package is named `torchsr`; it does not collide with `torchsr_b200`
    if REF not in sys.path:
        sys.path.insert(0, REF)


@pytest.mark.parametrize("seed", [0, 1])
def test_srgan_modules_live(seed):
    _ref()
    import torchsr_oracle as O
    from torchsr.srgan.discriminator import Discriminator
    from torchsr.srgan.generator import Generator
    torch.manual_seed(seed)
    torch.set_num_threads(4)
    G, D = Generator(), Discriminator()
    x, hr = torch.rand(2, 3, 10, 14), torch.rand(2, 3, 96, 96)
    assert torch.equal(G(x), O.srgan_generator(G.state_dict(), x, True))
    G.eval()
    assert torch.equal(G(x), O.srgan_generator(G.state_dict(), x, False))
    assert torch.equal(D(hr), O.srgan_discriminator(D.state_dict(), hr, True))


def test_esrgan_modules_live():
    _ref()
    import torchsr_oracle as O
    from torchsr.esrgan.discriminator import Discriminator
    from torchsr.esrgan.generator import Generator
    torch.manual_seed(3)
    torch.set_num_threads(4)
    G, D = Generator(num_rrdb_blocks=2), Discriminator()
    x, hr = torch.rand(1, 3, 9, 11), torch.rand(2, 3, 128, 128)
    assert torch.allclose(G(x), O.esrgan_generator(G.state_dict(), x), atol=1e-6)
    assert torch.allclose(D(hr), O.esrgan_discriminator(D.state_dict(), hr, True), atol=1e-6)


def test_default_init_matches_reference_rng_consumption():
    """Same seed -> same initial weights as the reference classes (identical children in identical order)."""
    _ref()
    from torchsr.srgan.discriminator import Discriminator as RD
    from torchsr.srgan.generator import Generator as RG
    from torchsr_b200.srgan.discriminator import Discriminator
    from torchsr_b200.srgan.generator import Generator
    for ours, ref in ((Generator, RG), (Discriminator, RD)):
        torch.manual_seed(7)
        a = ours().state_dict()
        torch.manual_seed(7)
        b = ref().state_dict()
        assert list(a) == list(b)
        for k in a:
            assert torch.equal(a[k], b[k]), k
