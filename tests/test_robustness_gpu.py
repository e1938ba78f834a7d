"""Host-side contracts of the drop-in modules on the GPU (round-1 advisor findings + the byte-valued op):

* borrowed parameters may MOVE (p.data = ..., load_state_dict(assign=True), EMA swaps): plans are rebuilt, nothing reads the
  freed storage;
* in-place edits through `.data` (invisible to the version counters) are picked up after `mark_dirty()`;
* `dxmi_quantize_u8` is bit-exact with the reference's `((x + 1) * 127.5).clamp(0, 255).to(uint8)` (generate_large.py:43),
  including the rounding boundaries and out-of-range / non-finite inputs;
* a batch-split rollout (S sub-batches on S streams) is bit-identical to the unsplit one;
* a network living on cuda:1 works while cuda:0 is the current device (skipped with one GPU).
"""
import ctypes as C

import pytest
import torch

from common import DDPM_CFG, build_ddpm, load_synth_into

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ddpm4():
    return build_ddpm(4)


def _eps(net, B=3, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, 32, 32, generator=g).cuda()
    t = torch.tensor([300.0, 50.0, 1.0][:B]).cuda()
    with torch.no_grad():
        return net(x, t).clone()


def test_parameter_storage_may_move(ddpm4):
    net = ddpm4[0]
    ref = _eps(net)
    old = []
    for p in net.parameters():
        old.append(p.data)
        p.data = p.data.clone()  # new storage, same values
    for o in old:
        o.fill_(float("nan"))  # whoever still reads the old storage gets NaNs
    torch.cuda.synchronize()
    out = _eps(net)
    assert torch.isfinite(out).all()
    assert torch.equal(out, ref)
    # value net: same contract
    value = ddpm4[2]
    x = torch.randn(4, 3, 32, 32, device="cuda")
    with torch.no_grad():
        e0 = value(x, 0).clone()
        old = []
        for p in value.parameters():
            old.append(p.data)
            p.data = p.data.clone()
        for o in old:
            o.fill_(float("nan"))
        e1 = value(x, 0)
    assert torch.equal(e0, e1)


def test_load_state_dict_assign_rebinds(ddpm4):
    net = ddpm4[0]
    ref = _eps(net)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net.load_state_dict(sd, assign=True)
    assert torch.equal(_eps(net), ref)


def test_mark_dirty_picks_up_data_edits():
    from diffusion_by_maxentirl_b200.models.DxMI.unet_small import Model

    net = Model(**DDPM_CFG)
    load_synth_into(net, skip=())
    net.cuda().eval()
    before = _eps(net)
    w = net._param("down.0.block.0.conv1.weight")
    v0 = w._version
    w.data.mul_(1.5)  # invisible to the version counter
    assert w._version == v0
    net.mark_dirty()
    after = _eps(net)
    assert not torch.equal(before, after)
    fresh = Model(**DDPM_CFG)
    fresh.load_state_dict(net.state_dict())
    fresh.cuda().eval()
    assert torch.equal(_eps(fresh), after)


def test_quantize_u8_bit_exact():
    """fp32 -> u8 must be the reference's truncating conversion, bit for bit."""
    from diffusion_by_maxentirl_b200.dist import quantize_u8

    ulp = torch.finfo(torch.float32).eps
    special = [-1.0, 1.0, 0.0, -0.0, -1.0 - ulp, -1.0 + ulp, 1.0 - ulp, 1.0 + ulp, 1.5, -1.5, 3.0e38, -3.0e38, 1e-30]
    # every representable level boundary k / 127.5 - 1 and its float neighbours
    k = torch.arange(0, 257, dtype=torch.float64)
    edges = (k / 127.5 - 1.0).float()
    neigh = torch.cat([edges, torch.nextafter(edges, torch.full_like(edges, 2.0)), torch.nextafter(edges, torch.full_like(edges, -2.0))])
    g = torch.Generator().manual_seed(3)
    x = torch.cat([torch.tensor(special), neigh, torch.rand(1 << 20, generator=g) * 2.4 - 1.2, torch.randn(1 << 18, generator=g)])
    pad = (-x.numel()) % (3 * 32 * 32)
    x = torch.cat([x, torch.zeros(pad)]).reshape(-1, 3, 32, 32).cuda()
    want = ((x + 1) * 127.5).clamp(0, 255).to(torch.uint8)
    got = quantize_u8(x)
    assert got.dtype == torch.uint8 and got.shape == x.shape
    assert torch.equal(got, want), (got != want).sum().item()
    # odd sizes (tail handling)
    for n in (1, 3, 5, 1023, 4097):
        y = (torch.rand(n, generator=g) * 2.2 - 1.1).cuda()
        assert torch.equal(quantize_u8(y), ((y + 1) * 127.5).clamp(0, 255).to(torch.uint8))


def test_split_rollout_is_bit_identical(ddpm4):
    """S sub-batches on S streams == one batch (the path is bitwise batch-invariant)."""
    from diffusion_by_maxentirl_b200 import _lib as L

    net, sampler, value, sd, vsd = ddpm4
    lib = L.lib()
    B, T = 64, 4
    noise = torch.randn(T + 1, B, 3, 32, 32, generator=torch.Generator().manual_seed(5)).cuda()
    outs = []
    try:
        for split in (1, 2, 4):
            lib.dxmi_set_option(b"rollout_split", split)
            lib.dxmi_set_option(b"rollout_split_min", 8)
            with torch.no_grad():
                d = sampler.sample(B, device="cuda", noise=noise)
            torch.cuda.synchronize()
            outs.append((torch.stack(d["l_sample"]).clone(), torch.stack(d["mean"]).clone(), torch.stack(d["logp"]).clone()))
    finally:
        lib.dxmi_set_option(b"rollout_split", 1)
        lib.dxmi_set_option(b"rollout_split_min", 32)
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert torch.equal(a, b)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_net_on_non_current_device():
    from diffusion_by_maxentirl_b200.models.DxMI.unet_small import Model

    net = Model(**DDPM_CFG)
    load_synth_into(net, skip=())
    net0 = Model(**DDPM_CFG)
    net0.load_state_dict(net.state_dict())
    net.to("cuda:1").eval()
    net0.to("cuda:0").eval()
    torch.cuda.set_device(0)
    x = torch.randn(2, 3, 32, 32)
    t = torch.tensor([100.0, 10.0])
    with torch.no_grad():
        a = net(x.to("cuda:1"), t.to("cuda:1"))
        b = net0(x.cuda(0), t.cuda(0))
    assert torch.cuda.current_device() == 0
    assert a.device.index == 1 and torch.equal(a.cpu(), b.cpu())


def test_fused_u8_of_last_step_equals_standalone_quantise(ddpm4):
    """f3: the last transition kernel writes the u8 samples itself; must equal quantize_u8(sample) == the reference's
    ((x + 1) * 127.5).clamp(0, 255).to(uint8), bit for bit (VARSampler and OpenAIDiffusion)."""
    from common import EDM_SMALL_CFG, build_edm
    from diffusion_by_maxentirl_b200.dist import PackedRollout, quantize_u8

    net, sampler, value, sd, vsd = ddpm4
    B = 6
    pk = PackedRollout(B, (3, 32, 32), "cuda", world=1)
    with torch.no_grad():
        d = sampler.sample(B, device="cuda", u8_out=pk.samples_u8)
        e = value(d["sample"], 4, out=pk.energies)
    torch.cuda.synchronize()
    want = ((d["sample"] + 1) * 127.5).clamp(0, 255).to(torch.uint8)
    assert torch.equal(d["sample_u8"], want) and torch.equal(quantize_u8(d["sample"]), want)
    assert torch.equal(pk.energies, e.reshape(-1)) and e.data_ptr() == pk.energies.data_ptr()
    unet, esampler, _ = build_edm(EDM_SMALL_CFG, 4)
    u8 = torch.empty(B, 3, 32, 32, dtype=torch.uint8, device="cuda")
    y = torch.arange(B, device="cuda")
    with torch.no_grad():
        d = esampler.sample(B, "cuda", i_class=y, u8_out=u8)
    torch.cuda.synchronize()
    assert torch.equal(u8, ((d["sample"] + 1) * 127.5).clamp(0, 255).to(torch.uint8))
