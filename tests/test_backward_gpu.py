"""Backward operators of the training step against torch autograd on the same bf16-rounded operands:
tensor-core weight gradient (MN-major UMMA operands, split-K), data gradient (forward implicit GEMM on transposed, tap-flipped
weights) with the fused leaky-relu gate.  fp32 outputs must agree to fp32 accumulation round-off (rel-L2 <= 2e-5); bf16
outputs to the bf16 rounding floor (4e-3)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


@pytest.fixture(scope="module")
def ops():
    from diffusion_by_maxentirl_b200 import ops as o

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return o


@pytest.mark.parametrize("N,H,Cin,Cout,k", [(4, 32, 128, 128, 3), (6, 16, 128, 256, 3), (5, 8, 256, 256, 3), (7, 4, 256, 256, 3),
                                            (3, 32, 128, 128, 1), (4, 16, 128, 256, 1), (2, 64, 192, 128, 3), (128, 4, 64, 128, 3)])
def test_conv_wgrad(ops, N, H, Cin, Cout, k):
    torch.manual_seed(20)
    dev = "cuda"
    x = nhwc(torch.randn(N, Cin, H, H, device=dev))
    dy = nhwc(torch.randn(N, Cout, H, H, device=dev))
    g = ops.conv_wgrad(dy, x, k, scale=0.5)
    torch.cuda.synchronize()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(False)
    w = torch.zeros(Cout, Cin, k, k, device=dev, requires_grad=True)
    y = F.conv2d(xr, w, None, padding=k // 2)
    y.backward(dy.float().permute(0, 3, 1, 2))
    err = rel_l2(g, 0.5 * w.grad)
    print(f"wgrad N={N} H={H} {Cin}->{Cout} k={k}: rel-L2 {err:.2e}")
    assert err < 2e-5
    # deterministic (fixed split-K reduction order)
    g2 = ops.conv_wgrad(dy, x, k, scale=0.5)
    assert torch.equal(g, g2)


def test_conv_wgrad_channel_slice(ops):
    """Two sources (torch.cat along channels): each fills its slice of the OIHW gradient."""
    torch.manual_seed(21)
    dev = "cuda"
    N, H, Ca, Cb, Cout = 3, 16, 128, 64, 128
    xa, xb = nhwc(torch.randn(N, Ca, H, H, device=dev)), nhwc(torch.randn(N, Cb, H, H, device=dev))
    dy = nhwc(torch.randn(N, Cout, H, H, device=dev))
    g = torch.full((Cout, Ca + Cb, 1, 1), float("nan"), device=dev)
    ops.conv_wgrad(dy, xa, 1, grad=g, ci_off=0)
    ops.conv_wgrad(dy, xb, 1, grad=g, ci_off=Ca)
    ref = torch.einsum("nhwo,nhwi->oi", dy.float(), torch.cat([xa, xb], -1).float())
    assert rel_l2(g.view(Cout, -1), ref) < 2e-5


@pytest.mark.parametrize("N,H,Cin,Cout", [(4, 32, 128, 128), (5, 8, 128, 256), (6, 4, 256, 256)])
def test_conv_dgrad_with_skip_and_gate(ops, N, H, Cin, Cout):
    """dX = conv3x3^T(dZ1, W1) + conv1x1^T(dO, Ws), then the leaky-relu gate of the saved block input - one GEMM."""
    torch.manual_seed(22)
    dev = "cuda"
    w1 = torch.randn(Cout, Cin, 3, 3, device=dev) / (3 * Cin**0.5)
    ws = torch.randn(Cout, Cin, 1, 1, device=dev) / Cin**0.5
    dz1 = nhwc(torch.randn(N, Cout, H, H, device=dev))
    do = nhwc(torch.randn(N, Cout, H, H, device=dev))
    h_in = nhwc(torch.randn(N, Cin, H, H, device=dev))
    wp = ops.pack_conv_weight_dgrad([w1, ws])
    srcs, segs = [(dz1, Cout, Cout), (do, Cout, Cout)], [(0, 9), (1, 1)]
    y32 = ops.conv_gemm(srcs, segs, wp, N, H, H, out_fp32=True)
    w1q, wsq = w1.to(torch.bfloat16).float(), ws.to(torch.bfloat16).float()
    ref = F.conv_transpose2d(dz1.float().permute(0, 3, 1, 2), w1q, padding=1) + F.conv_transpose2d(do.float().permute(0, 3, 1, 2), wsq)
    ref = ref.permute(0, 2, 3, 1)
    assert rel_l2(y32.view(N, H, H, Cin), ref) < 2e-5
    yg = ops.conv_gemm(srcs, segs, wp, N, H, H, gate=h_in)
    gate = torch.where(h_in.float() > 0, 1.0, 0.2)
    assert rel_l2(yg.view(N, H, H, Cin), ref * gate) < 4e-3
    # identity skip: residual = dO, same gate, with column sums (bias gradient of the next wgrad) from the fused statistics
    if Cin == Cout:
        wp1 = ops.pack_conv_weight_dgrad([w1])
        seg = 16 if H == 4 else 64
        stats = torch.zeros(N * H * H // seg, Cin, 2, device=dev)
        yr = ops.conv_gemm([(dz1, Cout, Cout)], [(0, 9)], wp1, N, H, H, residual=do, gate=h_in, gn_stats=stats, gn_seg=seg)
        ref2 = (F.conv_transpose2d(dz1.float().permute(0, 3, 1, 2), w1q, padding=1).permute(0, 2, 3, 1) + do.float()) * gate
        assert rel_l2(yr.view(N, H, H, Cin), ref2) < 4e-3
        assert torch.allclose(stats[..., 0].sum(0), yr.float().sum(0), rtol=1e-4, atol=2e-2)


@pytest.mark.parametrize("N,H,C1,C2,silu", [(4, 32, 128, 0, 1), (3, 16, 256, 128, 1), (5, 8, 256, 256, 0), (2, 4, 256, 0, 1)])
def test_group_norm_backward(ops, N, H, C1, C2, silu):
    """GroupNorm(32, eps 1e-6)(+SiLU) backward over a channel concat, vs torch autograd on the same bf16 inputs.  dx is a
    cancellation (mean-subtracted) quantity rounded to bf16: rel-L2 <= 1e-2; the fp32 parameter gradients <= 2e-3."""
    torch.manual_seed(30)
    dev = "cuda"
    C = C1 + C2
    x = torch.randn(N, C, H, H, device=dev) * 1.7 + 0.3
    gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
    xb = nhwc(x)  # NHWC bf16
    dy = nhwc(torch.randn(N, C, H, H, device=dev))
    xr = xb.float().permute(0, 3, 1, 2).requires_grad_(True)
    g, b = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = F.group_norm(xr, 32, g, b, 1e-6)
    if silu:
        y = F.silu(y)
    y.backward(dy.float().permute(0, 3, 1, 2))
    # the forward's saved quantities
    xg = xb.float().view(N, H * H, 32, C // 32)
    mean = xg.mean(dim=(1, 3))
    var = xg.var(dim=(1, 3), unbiased=False)
    rstd = torch.rsqrt(var + 1e-6)
    mr = torch.stack([mean, rstd], -1).contiguous()
    a = rstd.repeat_interleave(C // 32, dim=1) * gamma[None]
    bb = beta[None] - mean.repeat_interleave(C // 32, dim=1) * a
    ab = torch.stack([a, bb], -1).contiguous()
    x1 = xb[..., :C1].contiguous()
    x2 = xb[..., C1:].contiguous() if C2 else None
    dx, dg, db = ops.group_norm_bwd(x1, dy.view(N, H * H, C), ab, mr, silu, x2=x2)
    torch.cuda.synchronize()
    e = rel_l2(dx.view(N, H, H, C), xr.grad.permute(0, 2, 3, 1))
    eg, eb = rel_l2(dg, g.grad), rel_l2(db, b.grad)
    print(f"GN bwd N={N} H={H} C={C1}+{C2} silu={silu}: dx {e:.2e} dgamma {eg:.2e} dbeta {eb:.2e}")
    assert e < 1e-2 and eg < 2e-3 and eb < 2e-3


def _gn_saved(xb, gamma, beta, eps):
    """(ab, mr) the forward keeps for the backward, from the bf16 input the kernels read."""
    N, H, W, C = xb.shape
    xg = xb.float().view(N, H * W, 32, C // 32)
    mean = xg.mean(dim=(1, 3))
    rstd = torch.rsqrt(xg.var(dim=(1, 3), unbiased=False) + eps)
    a = rstd.repeat_interleave(C // 32, dim=1) * gamma[None]
    b = beta[None] - mean.repeat_interleave(C // 32, dim=1) * a
    return torch.stack([a, b], -1).contiguous(), torch.stack([mean, rstd], -1).contiguous()


@pytest.mark.parametrize("N,H,Cin,Cout", [(4, 16, 256, 256), (3, 32, 256, 128)])
def test_ddpm_resblock_backward_composed_from_operators(ops, N, H, Cin, Cout):
    """Groundwork for the U-Net backward (row a9): a whole DDPM ResnetBlock (unet_small.py:117-136: GN-swish-conv3x3 + temb,
    GN-swish-conv3x3, + identity | nin_shortcut) differentiated with the product operators only - GroupNorm backward, dgrad
    GEMMs on transposed weights, tensor-core wgrad, column sums - against torch autograd of the same block in fp32.
    Tolerance 3e-2 per gradient tensor (bf16 activations and activation gradients)."""
    torch.manual_seed(40)
    dev = "cuda"
    eps = 1e-6
    x = torch.randn(N, Cin, H, H, device=dev)
    temb_act = torch.randn(N, 512, device=dev)  # = swish(temb)
    P = {
        "g1": torch.randn(Cin, device=dev) * 0.3 + 1, "b1": torch.randn(Cin, device=dev) * 0.3,
        "w1": torch.randn(Cout, Cin, 3, 3, device=dev) / (3 * Cin**0.5), "c1": torch.randn(Cout, device=dev) * 0.1,
        "wt": torch.randn(Cout, 512, device=dev) / 512**0.5, "ct": torch.randn(Cout, device=dev) * 0.1,
        "g2": torch.randn(Cout, device=dev) * 0.3 + 1, "b2": torch.randn(Cout, device=dev) * 0.3,
        "w2": torch.randn(Cout, Cout, 3, 3, device=dev) / (3 * Cout**0.5), "c2": torch.randn(Cout, device=dev) * 0.1,
    }
    if Cin != Cout:
        P["ws"] = torch.randn(Cout, Cin, 1, 1, device=dev) / Cin**0.5
        P["cs"] = torch.randn(Cout, device=dev) * 0.1
    d_out = torch.randn(N, Cout, H, H, device=dev)

    # ---- reference: fp32 autograd (weights rounded to bf16 like the packed copies)
    R = {k: (v.to(torch.bfloat16).float() if v.dim() == 4 else v.clone()).requires_grad_(True) for k, v in P.items()}
    xr = nhwc(x).float().permute(0, 3, 1, 2).requires_grad_(True)
    h = F.conv2d(F.silu(F.group_norm(xr, 32, R["g1"], R["b1"], eps)), R["w1"], R["c1"], padding=1)
    h = h + F.linear(temb_act, R["wt"], R["ct"])[:, :, None, None]
    h = F.conv2d(F.silu(F.group_norm(h, 32, R["g2"], R["b2"], eps)), R["w2"], R["c2"], padding=1)
    sc = F.conv2d(xr, R["ws"], R["cs"]) if Cin != Cout else xr
    (sc + h).backward(nhwc(d_out).float().permute(0, 3, 1, 2))

    # ---- forward with the product operators, keeping what the backward needs
    xb = nhwc(x)
    g1 = ops.group_norm(xb, P["g1"], P["b1"], eps, 1)
    tproj = (temb_act @ P["wt"].t() + P["ct"]).contiguous()
    h1 = ops.conv_gemm([(g1, Cin, Cin)], [(0, 9)], ops.pack_conv_weight(P["w1"]), N, H, H, bias=P["c1"], rowvec=tproj).view(N, H, H, Cout)
    g2 = ops.group_norm(h1, P["g2"], P["b2"], eps, 1)
    ab1, mr1 = _gn_saved(xb, P["g1"], P["b1"], eps)
    ab2, mr2 = _gn_saved(h1, P["g2"], P["b2"], eps)

    # ---- backward with the product operators
    dO = nhwc(d_out)
    G = {}
    G["c2"] = dO.float().sum((0, 1, 2))
    G["w2"] = ops.conv_wgrad(dO, g2, 3)
    dG2 = ops.conv_gemm([(dO, Cout, Cout)], [(0, 9)], ops.pack_conv_weight_dgrad([P["w2"]]), N, H, H)
    dH1, G["g2"], G["b2"] = ops.group_norm_bwd(h1, dG2.view(N, H * H, Cout), ab2, mr2, 1)
    dH1 = dH1.view(N, H, H, Cout)
    d_tproj = dH1.float().sum((1, 2))  # per-image column sums (host-side here; a tiny reduction kernel in the plan)
    G["wt"], G["ct"] = d_tproj.t() @ temb_act, d_tproj.sum(0)
    G["c1"] = dH1.float().sum((0, 1, 2))
    G["w1"] = ops.conv_wgrad(dH1, g1, 3)
    dG1 = ops.conv_gemm([(dH1, Cout, Cout)], [(0, 9)], ops.pack_conv_weight_dgrad([P["w1"]]), N, H, H)
    dXa, G["g1"], G["b1"] = ops.group_norm_bwd(xb, dG1.view(N, H * H, Cin), ab1, mr1, 1)
    if Cin != Cout:
        G["ws"] = ops.conv_wgrad(dO, xb, 1)
        G["cs"] = G["c2"]
        dX = ops.conv_gemm([(dO, Cout, Cout)], [(0, 1)], ops.pack_conv_weight_dgrad([P["ws"]]), N, H, H, residual=dXa.view(N * H * H, Cin))
    else:
        dX = dXa.float().view(N, H, H, Cin) + dO.float()
    torch.cuda.synchronize()
    worst = 0.0
    for k in P:
        e = rel_l2(G[k], R[k].grad)
        worst = max(worst, e)
        print(f"resblock {Cin}->{Cout} grad {k}: {e:.2e}")
        assert e < 3e-2, (k, e)
    e = rel_l2(dX.float().view(N, H, H, Cin), xr.grad.permute(0, 2, 3, 1))
    print(f"resblock {Cin}->{Cout} dX: {e:.2e} (worst parameter {worst:.2e})")
    assert e < 3e-2
