"""Trainer-side fused ops (second half) against the reference trainer's own torch expressions:
running cost (trainer.py:163-169) forward + backward, clip_grad_norm_ (trainer.py:324-325, :388) and Adam with two LR groups
(train_cifar10.py:283-296) - values after several steps, the total norm, and the interplay with the drop-in networks' re-pack."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_running_cost_matches_reference_expression():
    from diffusion_by_maxentirl_b200.train_ops import running_cost

    g = torch.Generator().manual_seed(0)
    B = 37
    s = torch.randn(B, 3, 32, 32, generator=g).cuda()
    ns = (s.cpu() + 0.3 * torch.randn(B, 3, 32, 32, generator=g)).cuda().requires_grad_(True)
    betas = torch.rand(10, generator=g).cuda() * 0.5 + 0.01
    t = torch.randint(0, 10, (B,), generator=g).cuda()
    beta_next = betas[10 - t - 1]
    w = torch.randn(B, generator=g).cuda()
    rc = running_cost(s, ns, beta_next)
    (rc * w).sum().backward()
    ns2 = ns.detach().clone().requires_grad_(True)
    ref = (((ns2 - s) ** 2) / (2 * beta_next.view(-1, 1, 1, 1))).view(B, -1).mean(dim=1)  # trainer.py:167-168
    (ref * w).sum().backward()
    assert torch.allclose(rc, ref, rtol=2e-6, atol=1e-7)
    assert torch.allclose(ns.grad, ns2.grad, rtol=2e-6, atol=1e-9)


def _make_params(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(10,), (128, 3, 3, 3), (256, 256, 3, 3), (70001,), (1,), (512, 128)]
    return [torch.nn.Parameter(torch.randn(*s, generator=g).cuda()) for s in shapes]


def test_clip_and_adam_match_torch():
    from diffusion_by_maxentirl_b200.train_ops import FusedAdam, clip_grad_norm_

    a, b = _make_params(1), _make_params(1)
    opt_a = torch.optim.Adam([{"params": a[:1], "lr": 3e-3}, {"params": a[1:], "lr": 1e-4}])  # train_cifar10.py:287-290
    opt_b = FusedAdam([{"params": b[:1], "lr": 3e-3}, {"params": b[1:], "lr": 1e-4}])
    g = torch.Generator().manual_seed(2)
    for it in range(5):
        grads = [torch.randn(p.shape, generator=g).cuda() * (10.0 if it % 2 == 0 else 1e-3) for p in a]
        for pa, pb, gr in zip(a, b, grads):
            pa.grad, pb.grad = gr.clone(), gr.clone()
        if it == 3:  # a parameter without a gradient this step is skipped by both
            a[4].grad = b[4].grad = None
            continue
        na = torch.nn.utils.clip_grad_norm_(a, 0.1)
        if it < 2:
            nb = clip_grad_norm_(b, 0.1)   # separate fused clip (scales the gradients in place), then the fused step
            for pa, pb in zip(a, b):
                assert torch.allclose(pa.grad, pb.grad, rtol=1e-5, atol=1e-12)
            opt_b.step()
        else:
            opt_b.step(max_norm=0.1)       # clip folded into the step
            nb = opt_b.last_grad_norm
        opt_a.step()
        assert torch.allclose(na, nb, rtol=1e-5), (na, nb)
        for pa, pb in zip(a, b):
            assert torch.allclose(pa, pb, rtol=1e-5, atol=1e-7)
    sa, sb = opt_a.state_dict(), opt_b.state_dict()
    for k in sa["state"]:
        assert torch.allclose(sa["state"][k]["exp_avg"], sb["state"][k]["exp_avg"], rtol=1e-5, atol=1e-9)
        assert torch.allclose(sa["state"][k]["exp_avg_sq"], sb["state"][k]["exp_avg_sq"], rtol=1e-5, atol=1e-12)


def test_fused_adam_triggers_repack_of_the_dropin_network():
    """The kernel writes parameters through raw pointers: the drop-in net must still see the update (version counters)."""
    from common import DDPM_CFG, load_synth_into
    from diffusion_by_maxentirl_b200.models.DxMI.unet_small import Model
    from diffusion_by_maxentirl_b200.train_ops import FusedAdam

    net = Model(**DDPM_CFG)
    load_synth_into(net, skip=())
    net.cuda().eval()
    x = torch.randn(2, 3, 32, 32, device="cuda")
    t = torch.tensor([100.0, 10.0], device="cuda")
    with torch.no_grad():
        before = net(x, t).clone()
    opt = FusedAdam(net.parameters(), lr=1e-2)
    for p in net.parameters():
        p.grad = torch.ones_like(p)
    opt.step()
    with torch.no_grad():
        after = net(x, t).clone()
    assert not torch.equal(before, after)
    fresh = Model(**DDPM_CFG)
    fresh.load_state_dict(net.state_dict())
    fresh.cuda().eval()
    with torch.no_grad():
        assert torch.equal(fresh(x, t), after)
