"""Full-size (BASELINE.json configs) checks through size-independent properties of the path: the oracle cannot run these
batches in seconds, but every image's trajectory is independent given its noise (per-sample GroupNorm / attention, no
BatchNorm - SURVEY 8e), `sample()` equals the looped `sample_step()` (two reference code paths for the same math, SURVEY
section 4), and the path is deterministic."""
import pytest
import torch

from common import EDM_IN64_CFG, build_ddpm, build_edm, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ddpm4():
    return build_ddpm(4)


def test_cifar_b256_batch_independence_and_determinism(ddpm4):
    """configs[1]: T=4, B=256. Rows 0..2 of the big batch == the same noise rolled out alone; two runs are bit-identical."""
    net, sampler, value, sd, vsd = ddpm4
    B = 256
    g = torch.Generator().manual_seed(3)
    noise = torch.randn(5, B, 3, 32, 32, generator=g).cuda()
    d = sampler.sample(B, device="cuda", noise=noise)
    e = value(d["sample"], 4)
    d2 = sampler.sample(B, device="cuda", noise=noise)
    assert torch.equal(d["sample"], d2["sample"]) and torch.equal(torch.stack(d["logp"]), torch.stack(d2["logp"]))
    ds = sampler.sample(3, device="cuda", noise=noise[:, :3].contiguous())
    es = value(ds["sample"], 4)
    assert torch.isfinite(d["sample"]).all() and d["sample"].shape == (B, 3, 32, 32)
    for i in range(5):
        assert rel_l2(d["l_sample"][i][:3], ds["l_sample"][i]) < 1e-6, i
    assert rel_l2(e[:3], es) < 1e-6
    assert e.shape == (B, 1) and len(d["logp"]) == 4 and d["logp"][0].shape == (B,)


def test_cifar_b256_sample_equals_looped_sample_step(ddpm4):
    net, sampler, value, sd, vsd = ddpm4
    B = 256
    g = torch.Generator().manual_seed(4)
    noise = torch.randn(5, B, 3, 32, 32, generator=g).cuda()
    d = sampler.sample(B, device="cuda", noise=noise)
    x = noise[0]
    for i in range(4):
        s = sampler.sample_step(x, i, noise=noise[i + 1])
        assert rel_l2(s["sample"], d["l_sample"][i + 1]) < 1e-6, i
        assert rel_l2(s["mean"], d["mean"][i]) < 1e-6 and rel_l2(s["logp"], d["logp"][i]) < 1e-5
        x = s["sample"]


def test_cifar_logp_identity_full_batch(ddpm4):
    """logp only depends on the injected z: mean_CHW(-z^2/2) - ln sigma - 0.5 ln 2 pi (Appendix E.1)."""
    import math

    net, sampler, value, sd, vsd = ddpm4
    B = 256
    noise = torch.randn(5, B, 3, 32, 32, device="cuda")
    d = sampler.sample(B, device="cuda", noise=noise)
    for i in range(3):  # the last step has sigma = 1e-3: (x' - mean) cancels catastrophically in fp32 there, like the reference
        sig = d["sigma"][i].flatten()
        want = (-(noise[i + 1] ** 2) / 2).mean((1, 2, 3)) - torch.log(sig) - 0.5 * math.log(2 * math.pi)
        assert torch.allclose(d["logp"][i], want, atol=2e-3), i


def test_imagenet64_b64_batch_independence():
    """configs[2] per-GPU shard: ImageNet-64 EDM, B=64 (T=2 to keep the test short)."""
    unet, sampler, sd = build_edm(EDM_IN64_CFG, 2)
    B = 64
    g = torch.Generator().manual_seed(5)
    noise = torch.randn(3, B, 3, 64, 64, generator=g).cuda()
    y = torch.randint(0, 1000, (B,), generator=g).cuda()
    d = sampler.sample(B, "cuda", i_class=y, x0=noise[0] * 80.0, noise=noise[1:])
    ds = sampler.sample(2, "cuda", i_class=y[:2], x0=noise[0, :2] * 80.0, noise=noise[1:, :2].contiguous())
    assert torch.isfinite(d["sample"]).all()
    for i in range(3):
        assert rel_l2(d["l_sample"][i][:2], ds["l_sample"][i]) < 1e-6, i
    x = noise[0] * 80.0
    for i in range(2):
        s = sampler.sample_step(x, torch.full((B,), i, dtype=torch.long), noise=noise[i + 1], y=y)
        assert rel_l2(s["sample"], d["l_sample"][i + 1]) < 1e-6, i
        x = s["sample"]
