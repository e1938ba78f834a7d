"""models/cm/fp16_util.py drop-in: the MixedPrecisionTrainer contract on a tiny torch model on the CPU -
parameter grouping, master copies, loss scaling, overflow handling, copy-back - and, where a reference checkout exists, step by
step against the reference's own class."""
import copy
import os
import sys

import pytest
import torch
import torch.nn as nn

from diffusion_by_maxentirl_b200.models.cm import fp16_util as F16


class Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(3, 4, 3, padding=1)
        self.lin = nn.Linear(4, 2)
        self.log_betas = nn.Parameter(torch.zeros(5))

    def forward(self, x):
        h = self.conv(x.to(self.conv.weight.dtype)).float().mean((2, 3))
        return self.lin(h) + self.log_betas.sum()


def _model():
    torch.manual_seed(0)
    m = Tiny()
    F16.convert_module_to_f16(m.conv)
    return m


def test_groups_masters_and_step():
    m = _model()
    assert m.conv.weight.dtype == torch.float16 and m.lin.weight.dtype == torch.float32
    mp = F16.MixedPrecisionTrainer(model=m, use_fp16=True, initial_lg_loss_scale=4.0, special_key="log_betas")
    groups = mp.param_groups_and_shapes
    assert [n for n, _ in groups[0][0]] == ["log_betas"]
    assert sorted(n for n, _ in groups[1][0]) == ["conv.bias", "lin.bias"]
    assert sorted(n for n, _ in groups[2][0]) == ["conv.weight", "lin.weight"]
    assert mp.master_params[2].shape == (1, 4 * 3 * 9 + 2 * 4) and all(p.dtype == torch.float32 for p in mp.master_params)
    opt = torch.optim.SGD(mp.master_params, lr=0.5)
    x = torch.randn(2, 3, 8, 8)
    mp.zero_grad()
    loss = m(x).pow(2).mean()
    mp.backward(loss)
    g_conv = m.conv.weight.grad.float() / 16.0   # unscaled
    w_before = mp.master_params[2].detach().clone()
    assert mp.optimize(opt) is True
    assert abs(mp.lg_loss_scale - 4.001) < 1e-12
    # master = master - lr * unscaled grad; the fp16 model parameter is the rounded master
    n = m.conv.weight.numel()
    assert torch.allclose(mp.master_params[2].detach()[0, :n], w_before[0, :n] - 0.5 * g_conv.reshape(-1), atol=1e-6)
    assert torch.equal(m.conv.weight.detach(), mp.master_params[2].detach()[0, :n].view_as(m.conv.weight).half())
    assert all(p.grad is None for p in mp.master_params)
    sd = mp.master_params_to_state_dict(mp.master_params)
    assert sd["conv.weight"].dtype == torch.float32 and sd["conv.weight"].shape == m.conv.weight.shape
    again = mp.state_dict_to_master_params(sd)
    assert len(again) == 2 and again[1].numel() == mp.master_params[2].numel()  # (no special key on this path, like the reference)


def test_overflow_skips_the_step():
    m = _model()
    mp = F16.MixedPrecisionTrainer(model=m, use_fp16=True, initial_lg_loss_scale=30.0)
    opt = torch.optim.SGD(mp.master_params, lr=0.5)
    x = torch.randn(2, 3, 8, 8) * 100
    mp.zero_grad()
    mp.backward(m(x).pow(2).mean())
    before = [p.detach().clone() for p in m.parameters()]
    assert mp.optimize(opt) is False and mp.lg_loss_scale == 29.0
    assert all(torch.equal(a, b.detach()) for a, b in zip(before, m.parameters()))
    # fp32 mode: plain step
    m2 = Tiny()
    mp2 = F16.MixedPrecisionTrainer(model=m2, use_fp16=False)
    opt2 = torch.optim.SGD(mp2.master_params, lr=0.1)
    mp2.zero_grad()
    mp2.backward(m2(torch.randn(2, 3, 8, 8)).pow(2).mean())
    assert mp2.optimize(opt2) is True and mp2.last_grad_norm > 0


def test_matches_the_reference_class():
    ref = os.environ.get("DXMI_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "models")):
        pytest.skip("no reference checkout here")
    sys.path.insert(0, ref)
    try:
        from models.cm import fp16_util as R
    except Exception as e:  # noqa: BLE001
        pytest.skip(f"reference fp16_util not importable here: {e}")
    finally:
        sys.path.remove(ref)
    ma, mb = _model(), None
    mb = copy.deepcopy(ma)
    a = F16.MixedPrecisionTrainer(model=ma, use_fp16=True, initial_lg_loss_scale=6.0, special_key="log_betas")
    b = R.MixedPrecisionTrainer(model=mb, use_fp16=True, initial_lg_loss_scale=6.0, special_key="log_betas")
    oa = torch.optim.Adam([{"params": a.master_params[1:], "lr": 1e-2}, {"params": a.master_params[0:1], "lr": 1e-1}])
    ob = torch.optim.Adam([{"params": b.master_params[1:], "lr": 1e-2}, {"params": b.master_params[0:1], "lr": 1e-1}])
    g = torch.Generator().manual_seed(4)
    for step in range(4):
        x = torch.randn(2, 3, 8, 8, generator=g) * (1e4 if step == 2 else 1.0)  # step 2 overflows in both
        for mp, m, o in ((a, ma, oa), (b, mb, ob)):
            mp.zero_grad()
            mp.backward(m(x).pow(2).mean())
        ra, rb = a.optimize(oa), b.optimize(ob)
        assert ra == rb and abs(a.lg_loss_scale - b.lg_loss_scale) < 1e-12, step
        for pa, pb in zip(ma.parameters(), mb.parameters()):
            assert torch.equal(pa.detach(), pb.detach()), step
        for qa, qb in zip(a.master_params, b.master_params):
            assert torch.equal(qa.detach(), qb.detach()), step
