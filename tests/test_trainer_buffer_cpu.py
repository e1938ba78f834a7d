"""Replay-buffer drop-ins (reference models/DxMI/trainer.py:23-70): same dict as the reference's O(T^2)
torch.cat loop, built with one concatenation per key."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rollout(T=4, B=3, views=True):
    base = torch.randn(T + 1, B, 3, 8, 8)
    l_sample = [base[i] for i in range(T + 1)] if views else [base[i].clone() for i in range(T + 1)]
    return {"l_sample": l_sample, "logp": [torch.randn(B) for _ in range(T)], "mean": list(torch.randn(T, B, 3, 8, 8)),
            "control": [torch.randn(B, 3, 8, 8) for _ in range(T)], "sigma": [torch.rand(B, 1, 1, 1) for _ in range(T)],
            "y": torch.randint(0, 10, (B,))}


@pytest.mark.parametrize("views", [True, False])
def test_append_buffer_layout(views):
    from diffusion_by_maxentirl_b200.models.DxMI.trainer import append_buffer, reset_buffer

    T, B = 4, 3
    d = _rollout(T, B, views)
    buf = append_buffer(reset_buffer("cpu"), d)
    assert buf["state"].shape == (T * B, 3, 8, 8) and buf["timestep"].dtype == torch.long and buf["y"].dtype == torch.long
    for t in range(T):  # step-major rows
        rows = slice(t * B, (t + 1) * B)
        assert torch.equal(buf["state"][rows], d["l_sample"][t]) and torch.equal(buf["next_state"][rows], d["l_sample"][t + 1])
        assert torch.equal(buf["final"][rows], d["l_sample"][-1]) and (buf["timestep"][rows] == t).all()
        assert torch.equal(buf["logp"][rows], d["logp"][t]) and torch.equal(buf["sigma"][rows], d["sigma"][t])
        assert torch.equal(buf["control"][rows], d["control"][t]) and torch.equal(buf["y"][rows], d["y"])
    assert buf["entropy"].numel() == 0  # key absent from sample(): stays empty like in the reference
    buf = append_buffer(buf, d)  # a second rollout appends
    assert buf["state"].shape[0] == 2 * T * B and torch.equal(buf["state"][T * B:], buf["state"][:T * B])
    assert not buf["state"].requires_grad


def test_append_buffer_equals_reference():
    ref = os.environ.get("DXMI_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "models")):
        pytest.skip("no reference checkout here")
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import torch\n"
        "import models.DxMI.trainer as rt\n"
        "from diffusion_by_maxentirl_b200.models.DxMI import trainer as ours\n"
        "from test_trainer_buffer_cpu import _rollout\n"
        "for views in (True, False):\n"
        "    torch.manual_seed(3)\n"
        "    d = _rollout(5, 4, views)\n"
        "    r = rt.append_buffer(rt.reset_buffer('cpu'), d); o = ours.append_buffer(ours.reset_buffer('cpu'), d)\n"
        "    r = rt.append_buffer(r, d); o = ours.append_buffer(o, d)\n"
        "    assert set(r) == set(o)\n"
        "    for k in r:\n"
        "        assert r[k].dtype == o[k].dtype and r[k].shape == o[k].shape and torch.equal(r[k], o[k]), k\n"
        "print('same')\n" % (ref, ROOT, os.path.join(ROOT, "tests"))
    )
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0 and "same" in r.stdout, r.stderr[-2000:]


def test_first_append_owns_its_memory():
    """The buffer must not alias the rollout's tensors (a GraphedRollout overwrites them on the next replay; the reference
    always copies through torch.cat)."""
    from diffusion_by_maxentirl_b200.models.DxMI.trainer import append_buffer, reset_buffer

    d = _rollout(4, 3, views=True)
    want = torch.stack(d["l_sample"][:4]).reshape(12, 3, 8, 8).clone()
    buf = append_buffer(reset_buffer("cpu"), d)
    d["l_sample"][0]._base.fill_(7.0)  # the next replay overwrites the static rollout tensor
    assert torch.equal(buf["state"], want)
    assert buf["state"]._base is None or buf["state"]._base is not d["l_sample"][0]._base


def test_value_net_load_pretrained_cpu():
    """Reference modules.py:165-180: `load_pretrained` copies conv1.* / blocks.* from ckpt['state_dict'] (keys carry a `net.`
    prefix) and leaves the head alone; the reference's TimeIndependentValue.load_pretrained forwards to it."""
    from common import VALUE_CFG
    from diffusion_by_maxentirl_b200.models.modules import IGEBMEncoderV2
    from diffusion_by_maxentirl_b200.models.value import TimeIndependentValue

    torch.manual_seed(0)
    src = TimeIndependentValue(IGEBMEncoderV2(**VALUE_CFG))
    dst = TimeIndependentValue(IGEBMEncoderV2(**VALUE_CFG))
    head_before = dst.net.linear.weight.detach().clone()
    ckpt = {"state_dict": {k: v.clone() for k, v in src.state_dict().items()}}
    dst.load_pretrained(ckpt)
    sd_s, sd_d = src.state_dict(), dst.state_dict()
    for k in sd_s:
        if k.startswith("net.conv1.") or k.startswith("net.blocks."):
            assert torch.equal(sd_s[k], sd_d[k]), k
    assert torch.equal(dst.net.linear.weight, head_before)
    bad = {"state_dict": {"net.conv1.weight": torch.zeros(1)}}
    with pytest.raises(RuntimeError):
        dst.net.load_pretrained(bad)
