"""Kernel-level parity: the tcgen05 implicit-GEMM conv / batched GEMM and GroupNorm against torch fp32 math on the
same bf16-rounded inputs.  Tolerances: outputs are bf16 (rel 2^-8 per element); we assert rel-L2 <= 4e-3 which is
the bf16 output rounding floor, i.e. the accumulation itself must be exact to fp32."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def nhwc(x):  # NCHW fp32 -> NHWC bf16
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def ref_conv(x_bf, w, b=None, stride=1, pad=1, asym=False):
    """x_bf: NHWC bf16; w OIHW fp32 (rounded to bf16 like the packed copy)."""
    x = x_bf.float().permute(0, 3, 1, 2)
    wq = w.to(torch.bfloat16).float()
    if asym:
        x = F.pad(x, (0, 1, 0, 1))
        y = F.conv2d(x, wq, b, stride=2, padding=0)
    else:
        y = F.conv2d(x, wq, b, stride=stride, padding=pad)
    return y.permute(0, 2, 3, 1).contiguous()


@pytest.fixture(scope="module", params=[0, 2], ids=["single", "pair"])
def ops(request):
    """Every test of this module runs twice: with the one-CTA-per-tile persistent kernel and with the cta_group::2
    pair kernel forced on for every GEMM it supports (block_n 128 / 192 / 256)."""
    from diffusion_by_maxentirl_b200 import _lib as L
    from diffusion_by_maxentirl_b200 import ops as o

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    L.lib().dxmi_set_option(b"pair", request.param)
    o.pair_mode = request.param
    yield o
    L.lib().dxmi_set_option(b"pair", 0)
    o.pair_mode = 0


@pytest.mark.parametrize(
    "N,H,Cin,Cout,block_n",
    [
        (2, 16, 128, 128, 0),
        (2, 32, 128, 128, 128),
        (4, 16, 256, 256, 256),
        (4, 16, 256, 256, 128),
        (8, 8, 256, 256, 0),
        (3, 8, 256, 256, 128),   # odd image count: tile spans 2 images, tail masked
        (16, 4, 256, 256, 0),
        (5, 4, 512, 256, 0),
        (2, 16, 64, 64, 64),
        (2, 64, 192, 192, 64),   # ImageNet-64 geometry
    ],
)
def test_conv3x3(ops, N, H, Cin, Cout, block_n):
    torch.manual_seed(0)
    dev = "cuda"
    x = nhwc(torch.randn(N, Cin, H, H, device=dev))
    w = torch.randn(Cout, Cin, 3, 3, device=dev) / (3 * Cin**0.5)
    b = torch.randn(Cout, device=dev)
    wp = ops.pack_conv_weight(w)
    y = ops.conv_gemm([(x, Cin, Cin)], [(0, 9)], wp, N, H, H, bias=b, block_n=block_n)
    torch.cuda.synchronize()
    ref = ref_conv(x, w, b)
    assert rel_l2(y.view(N, H, H, Cout), ref) < 4e-3
    yf = ops.conv_gemm([(x, Cin, Cin)], [(0, 9)], wp, N, H, H, bias=b, block_n=block_n, out_fp32=True)
    assert rel_l2(yf.view(N, H, H, Cout), ref) < 2e-5


def test_conv1x1_plain_gemm(ops):
    torch.manual_seed(1)
    dev = "cuda"
    N, H, Cin, Cout = 4, 16, 256, 768
    x = nhwc(torch.randn(N, Cin, H, H, device=dev))
    w = torch.randn(Cout, Cin, 1, 1, device=dev) / Cin**0.5
    b = torch.randn(Cout, device=dev)
    wp = ops.pack_conv_weight(w)
    y = ops.conv_gemm([(x, Cin, Cin)], [(0, 1)], wp, N, H, H, bias=b, out_fp32=True)
    ref = ref_conv(x, w, b, pad=0)
    assert rel_l2(y.view(N, H, H, Cout), ref) < 2e-5


@pytest.mark.parametrize("N,H,C", [(2, 32, 128), (4, 16, 256), (8, 8, 256)])
def test_conv_stride2_asym_pad(ops, N, H, C):
    """Downsample: pad (0,1,0,1) then 3x3 stride 2 (unet_small.py:69-73) through a TMA map with elementStrides=2."""
    torch.manual_seed(2)
    dev = "cuda"
    x = nhwc(torch.randn(N, C, H, H, device=dev))
    w = torch.randn(C, C, 3, 3, device=dev) / (3 * C**0.5)
    b = torch.randn(C, device=dev)
    wp = ops.pack_conv_weight(w)
    y = ops.conv_gemm([(x, C, C)], [(0, 9)], wp, N, H, H, out_H=H // 2, out_W=H // 2, stride=2, bias=b, out_fp32=True)
    ref = ref_conv(x, w, b, asym=True)
    assert rel_l2(y.view(N, H // 2, H // 2, C), ref) < 2e-5


def test_resblock_conv2_fused_shortcut_and_epilogue(ops):
    """conv3x3(g) + nin_shortcut(cat(xa, xb)) in one accumulator, + bias; and conv1-style bias + temb rowvec + SiLU."""
    torch.manual_seed(3)
    dev = "cuda"
    N, H, Ca, Cb, Co = 4, 16, 256, 128, 256
    g = nhwc(torch.randn(N, Co, H, H, device=dev))
    xa = nhwc(torch.randn(N, Ca, H, H, device=dev))
    xb = nhwc(torch.randn(N, Cb, H, H, device=dev))
    w2 = torch.randn(Co, Co, 3, 3, device=dev) / (3 * Co**0.5)
    wn = torch.randn(Co, Ca + Cb, 1, 1, device=dev) / (Ca + Cb) ** 0.5
    b = torch.randn(Co, device=dev)
    wp = ops.pack_conv_weight(None, parts=[(w2, 0, Co), (wn, 0, Ca), (wn, Ca, Cb)])
    y = ops.conv_gemm([(g, Co, Co), (xa, Ca, Ca), (xb, Cb, Cb)], [(0, 9), (1, 1), (2, 1)], wp, N, H, H, bias=b,
                      out_fp32=True)
    ref = ref_conv(g, w2, b) + ref_conv(torch.cat([xa, xb], -1), wn, None, pad=0)
    assert rel_l2(y.view(N, H, H, Co), ref) < 2e-5

    # bias + per-image row vector + residual + leaky relu
    rowvec = torch.randn(N, 3 * Co, device=dev)[:, Co:2 * Co]  # strided view, ld = 3*Co
    res = nhwc(torch.randn(N, Co, H, H, device=dev))
    wp2 = ops.pack_conv_weight(w2)
    y2 = ops.conv_gemm([(g, Co, Co)], [(0, 9)], wp2, N, H, H, bias=b, rowvec=rowvec, residual=res, act=1, out_fp32=True)
    ref2 = F.leaky_relu(ref_conv(g, w2, b) + rowvec[:, None, None, :] + res.float(), 0.2)
    assert rel_l2(y2.view(N, H, H, Co), ref2) < 2e-5


def test_attention_gemms(ops):
    """q k^T with fused row softmax, V^T with weights as the A operand, and P V - the DDPM AttnBlock at 16x16."""
    torch.manual_seed(4)
    dev = "cuda"
    B, S, Cc = 3, 256, 256
    hn = (torch.randn(B, S, Cc, device=dev)).to(torch.bfloat16)
    qk = (torch.randn(B, S, 2 * Cc, device=dev)).to(torch.bfloat16)
    scale = Cc**-0.5
    # S = softmax(scale q k^T)
    P = torch.empty(B, S, S, dtype=torch.bfloat16, device=dev)
    ops.conv_gemm([(qk, Cc, 2 * Cc)], [(0, 1)], qk[:, :, Cc:], B, 1, S, batch=B, a_batched=True, b_batched=True,
                  b_rows=S, b_ld=2 * Cc, b_batch_stride=S * 2 * Cc, alpha=scale, softmax=True, out=P, ldo=S,
                  out_batch_stride=S * S, rows_per_image=1)
    q, k = qk[..., :Cc].float(), qk[..., Cc:].float()
    Pref = torch.softmax(scale * q @ k.transpose(1, 2), dim=-1)
    assert rel_l2(P, Pref) < 4e-3
    # V^T[b] = Wv hn[b]^T + bv (bias along M)
    wv = (torch.randn(Cc, Cc, device=dev) / Cc**0.5).to(torch.bfloat16)
    bv = torch.randn(Cc, device=dev)
    vT = torch.empty(B, Cc, S, dtype=torch.bfloat16, device=dev)
    ops.conv_gemm([(wv, Cc, Cc)], [(0, 1)], hn, 1, 1, Cc, batch=B, b_batched=True, b_rows=S, b_ld=Cc,
                  b_batch_stride=S * Cc, bias=bv, bias_along_m=True, out=vT, ldo=S, out_batch_stride=Cc * S,
                  rows_per_image=1)
    vTref = wv.float() @ hn.float().transpose(1, 2) + bv[None, :, None]
    assert rel_l2(vT, vTref) < 4e-3
    # O = P V
    O = torch.empty(B, S, Cc, dtype=torch.bfloat16, device=dev)
    ops.conv_gemm([(P, S, S)], [(0, 1)], vT, B, 1, S, batch=B, a_batched=True, b_batched=True, b_rows=Cc, b_ld=S,
                  b_batch_stride=Cc * S, out=O, ldo=Cc, out_batch_stride=S * Cc, rows_per_image=1)
    Oref = P.float() @ vT.float().transpose(1, 2)
    assert rel_l2(O, Oref) < 4e-3


@pytest.mark.parametrize("N,H,C1,C2,silu", [(4, 32, 128, 0, 1), (4, 16, 256, 128, 1), (3, 8, 256, 256, 0), (2, 4, 256, 0, 1)])
def test_group_norm(ops, N, H, C1, C2, silu):
    torch.manual_seed(5)
    dev = "cuda"
    x1 = nhwc(torch.randn(N, C1, H, H, device=dev) * 2 + 0.5)
    x2 = nhwc(torch.randn(N, C2, H, H, device=dev)) if C2 else None
    C = C1 + C2
    gamma = torch.randn(C, device=dev)
    beta = torch.randn(C, device=dev)
    y = ops.group_norm(x1, gamma, beta, 1e-6, silu, x2=x2)
    xc = torch.cat([x1, x2], -1) if C2 else x1
    ref = F.group_norm(xc.float().permute(0, 3, 1, 2), 32, gamma, beta, 1e-6)
    if silu:
        ref = F.silu(ref)
    assert rel_l2(y, ref.permute(0, 2, 3, 1)) < 4e-3


@pytest.mark.parametrize("version", [1, 2])
def test_both_gemm_kernel_versions(ops, version):
    """gemm_version 1 (one tile per CTA, direct stores) and 2 (persistent, TMA-store epilogue) agree with torch."""
    from diffusion_by_maxentirl_b200 import _lib as L

    L.lib().dxmi_set_option(b"gemm_version", version)
    try:
        torch.manual_seed(7)
        dev = "cuda"
        N, H, Cin, Cout = 6, 16, 128, 192
        x = nhwc(torch.randn(N, Cin, H, H, device=dev))
        w = torch.randn(Cout, Cin, 3, 3, device=dev) / (3 * Cin**0.5)
        b = torch.randn(Cout, device=dev)
        res = nhwc(torch.randn(N, Cout, H, H, device=dev))
        wp = ops.pack_conv_weight(w)
        ref = F.silu(ref_conv(x, w, b) + res.float())
        for bn in (0, 64, 128, 256):
            y = ops.conv_gemm([(x, Cin, Cin)], [(0, 9)], wp, N, H, H, bias=b, residual=res, act=2, block_n=bn)
            assert rel_l2(y.view(N, H, H, Cout), ref) < 4e-3, bn
    finally:
        L.lib().dxmi_set_option(b"gemm_version", 2)


@pytest.mark.parametrize("N,H,Cin,Cout", [(4, 16, 128, 256), (3, 32, 64, 128), (5, 8, 256, 192), (5, 4, 256, 256), (24, 4, 128, 128)])
def test_fused_groupnorm_partials(ops, N, H, Cin, Cout):
    """The persistent kernel's fused statistics: per 16/32/64/128-row segment and column, (sum, sumsq) of the bf16 outputs
    (16-row segments = one 4x4 image; ragged last tile at N=5)."""
    torch.manual_seed(8)
    dev = "cuda"
    x = nhwc(torch.randn(N, Cin, H, H, device=dev))
    w = torch.randn(Cout, Cin, 3, 3, device=dev) / (3 * Cin**0.5)
    b = torch.randn(Cout, device=dev)
    wp = ops.pack_conv_weight(w)
    M = N * H * H
    from diffusion_by_maxentirl_b200 import _lib as L

    L.lib().dxmi_set_option(b"halo", 1)
    tpi = L.lib().dxmi_op_halo_tiles_per_image(H, H)
    L.lib().dxmi_set_option(b"halo", 0)
    if tpi and not ops.pair_mode:  # (the pair kernel has no halo mode)
        L.lib().dxmi_set_option(b"halo", 1)
        # halo-mode conv (32- / 64-wide maps): one partial per tile of the zero-padded (W+2)-wide position grid
        stats = torch.full((N, tpi, Cout, 2), float("nan"), device=dev)
        try:
            y = ops.conv_gemm([(x, Cin, Cin)], [(0, 9)], wp, N, H, H, bias=b, gn_stats=stats, gn_halo_P=tpi)
        finally:
            L.lib().dxmi_set_option(b"halo", 0)
        torch.cuda.synchronize()
        assert rel_l2(y.view(N, H, H, Cout), ref_conv(x, w, b)) < 4e-3
        yf = y.float().view(N, H * H, Cout)
        assert torch.isfinite(stats).all()
        assert torch.allclose(stats[..., 0].sum(1), yf.sum(1), rtol=1e-4, atol=5e-3)
        assert torch.allclose(stats[..., 1].sum(1), (yf * yf).sum(1), rtol=1e-4, atol=5e-3)
        return
    for seg in (16, 32, 64, 128):
        if (H * H) % seg:
            continue
        stats = torch.full((M // seg, Cout, 2), float("nan"), device=dev)
        y = ops.conv_gemm([(x, Cin, Cin)], [(0, 9)], wp, N, H, H, bias=b, gn_stats=stats, gn_seg=seg)
        torch.cuda.synchronize()
        yf = y.float().view(M // seg, seg, Cout)
        assert torch.isfinite(stats).all(), seg
        assert torch.allclose(stats[..., 0], yf.sum(1), rtol=1e-4, atol=2e-3), seg
        assert torch.allclose(stats[..., 1], (yf * yf).sum(1), rtol=1e-4, atol=2e-3), seg


@pytest.mark.parametrize("N,H,Cin,Cout,bn", [(40, 16, 256, 256, 0), (5, 8, 256, 256, 0), (3, 16, 128, 192, 0), (70, 8, 64, 128, 0), (4, 32, 192, 192, 0),
                                              (40, 16, 256, 256, 128)])
def test_up2_phase_convolution(ops, N, H, Cin, Cout, bn):
    """Nearest-2x upsample + 3x3 conv (unet_small.py:52-64; cm/unet.py:103-118, :186-199) as four 2x2 phase convolutions of the
    low-resolution tensor with pre-summed weights (up2 mode): against torch on the upsampled tensor; packed phase filters
    against their closed form; the GroupNorm partials [image][phase][segment] against the stored bf16 outputs; with and
    without the per-image row vector; pair and one-CTA kernels (ragged batch of 8x8 tiles: two images per tile)."""
    torch.manual_seed(21)
    dev = "cuda"
    x = nhwc(torch.randn(N, Cin, H, H, device=dev))
    w = torch.randn(Cout, Cin, 3, 3, device=dev) / (3 * Cin**0.5)
    b = torch.randn(Cout, device=dev)
    rowvec = torch.randn(N, Cout, device=dev)
    wp = ops.pack_conv_weight_up2(w)
    # closed form of the phase filters: rows {0},{1,2} (py = 0) / {0,1},{2} (py = 1), columns likewise
    sel = {0: ([0], [1, 2]), 1: ([0, 1], [2])}
    for py in (0, 1):
        for px in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    ref_w = w[:, :, sel[py][dy], :][:, :, :, sel[px][dx]].sum((2, 3))
                    got = wp[py * 2 + px].view(Cout, 4, Cin)[:, dy * 2 + dx].float()
                    # (fp32 sums in another order may round to the neighbouring bf16 value)
                    assert torch.allclose(got, ref_w, rtol=2.0**-7, atol=1e-6), (py, px, dy, dx)
                    assert (got == ref_w.to(torch.bfloat16).float()).float().mean() > 0.99
    xu = x.permute(0, 3, 1, 2).float()
    xu = F.interpolate(xu, scale_factor=2, mode="nearest")
    ref = F.conv2d(xu, w, b, padding=1).permute(0, 2, 3, 1)
    seg = 128 if (H * H) % 128 == 0 else 64
    P = 4 * (H * H // seg)
    stats = torch.full((N, P, Cout, 2), float("nan"), device=dev)
    y = ops.conv_up2(x, wp, bias=b, gn_stats=stats, gn_seg=seg, block_n=bn)
    torch.cuda.synchronize()
    assert y.shape == (N, 2 * H, 2 * H, Cout)
    assert rel_l2(y, ref) < 6e-3, rel_l2(y, ref)
    # the same contraction with the bf16 phase filters in torch: only accumulation order differs
    yf = y.float()
    for py in (0, 1):
        for px in (0, 1):
            wq = wp[py * 2 + px].view(Cout, 2, 2, Cin).permute(0, 3, 1, 2).float()
            xp = F.pad(x.permute(0, 3, 1, 2).float(), (1 - px, px, 1 - py, py))
            rp = F.conv2d(xp, wq, b).permute(0, 2, 3, 1)
            assert rel_l2(yf[:, py::2, px::2], rp) < 4e-3, (py, px)
    assert torch.isfinite(stats).all()
    s = stats.view(N, 4, P // 4, Cout, 2)
    for ph in range(4):
        yp = yf[:, ph >> 1::2, ph & 1::2].reshape(N, P // 4, seg, Cout)
        assert torch.allclose(s[:, ph, :, :, 0], yp.sum(2), rtol=1e-4, atol=2e-3), ph
        assert torch.allclose(s[:, ph, :, :, 1], (yp * yp).sum(2), rtol=1e-4, atol=2e-3), ph
    y2 = ops.conv_up2(x, wp, bias=b, rowvec=rowvec, block_n=bn)
    assert rel_l2(y2, ref + rowvec[:, None, None, :]) < 6e-3


@pytest.mark.parametrize("N,H,Ca,Cb,Co", [(3, 32, 128, 64, 128), (2, 64, 192, 192, 192), (5, 32, 256, 0, 256)])
def test_halo_conv_fused_shortcut_residual(ops, N, H, Ca, Cb, Co):
    """Halo-tile 3x3 conv (one TMA load per channel chunk serves all 9 taps) + centre-tap 1x1 segments over two more
    sources + bias + per-image row vector + residual, against torch; and halo on/off give the same answer."""
    from diffusion_by_maxentirl_b200 import _lib as L

    torch.manual_seed(9)
    dev = "cuda"
    g = nhwc(torch.randn(N, Co, H, H, device=dev))
    xa = nhwc(torch.randn(N, Ca, H, H, device=dev))
    w2 = torch.randn(Co, Co, 3, 3, device=dev) / (3 * Co**0.5)
    b = torch.randn(Co, device=dev)
    rowvec = torch.randn(N, Co, device=dev)
    srcs, segs, parts = [(g, Co, Co), (xa, Ca, Ca)], [(0, 9), (1, 1)], None
    wn = torch.randn(Co, Ca + Cb, 1, 1, device=dev) / (Ca + Cb) ** 0.5
    parts = [(w2, 0, Co), (wn, 0, Ca)]
    cat = xa
    if Cb:
        xb = nhwc(torch.randn(N, Cb, H, H, device=dev))
        srcs.append((xb, Cb, Cb))
        segs.append((2, 1))
        parts.append((wn, Ca, Cb))
        cat = torch.cat([xa, xb], -1)
    wp = ops.pack_conv_weight(None, parts=parts)
    ref = ref_conv(g, w2, b) + ref_conv(cat, wn, None, pad=0) + rowvec[:, None, None, :]
    outs = []
    for halo in (1, 0):
        L.lib().dxmi_set_option(b"halo", halo)
        try:
            y = ops.conv_gemm(srcs, segs, wp, N, H, H, bias=b, rowvec=rowvec, out_fp32=True)
        finally:
            L.lib().dxmi_set_option(b"halo", 0)
        assert rel_l2(y.view(N, H, H, Co), ref) < 2e-5, halo
        outs.append(y)
    assert rel_l2(outs[0], outs[1]) < 1e-6
    res = nhwc(torch.randn(N, Co, H, H, device=dev))
    wp2 = ops.pack_conv_weight(w2)
    y2 = ops.conv_gemm([(g, Co, Co)], [(0, 9)], wp2, N, H, H, bias=b, residual=res, act=2)
    ref2 = F.silu(ref_conv(g, w2, b) + res.float())
    assert rel_l2(y2.view(N, H, H, Co), ref2) < 4e-3


def test_pair_kernel_weights_stationary(ops):
    """Large-M 128->128 3x3 conv: the pair kernel keeps its half of the packed weights resident in shared memory for all tiles
    (gemm_tc2p.cu RESB).  Same accumulation order as the streaming variant -> bitwise equal; both match torch."""
    from diffusion_by_maxentirl_b200 import _lib as L

    if not ops.pair_mode:
        pytest.skip("pair kernel only")
    torch.manual_seed(12)
    dev = "cuda"
    N, H, C = 80, 32, 128  # 640 row tiles = 320 pair tiles >= 4 per CTA pair
    x = nhwc(torch.randn(N, C, H, H, device=dev))
    w = torch.randn(C, C, 3, 3, device=dev) / (3 * C**0.5)
    b = torch.randn(C, device=dev)
    rowvec = torch.randn(N, C, device=dev)
    res = nhwc(torch.randn(N, C, H, H, device=dev))
    wp = ops.pack_conv_weight(w)
    outs = []
    L.lib().dxmi_set_option(b"shift3", 0)  # the shift-3 A-reuse mode accumulates the taps in another order (tested below)
    for resb in (1, 0):
        L.lib().dxmi_set_option(b"pair_resident_b", resb)
        try:
            stats = torch.zeros(N * H * H // 128, C, 2, device=dev)
            y = ops.conv_gemm([(x, C, C)], [(0, 9)], wp, N, H, H, bias=b, rowvec=rowvec, residual=res, gn_stats=stats, gn_seg=128)
        finally:
            L.lib().dxmi_set_option(b"pair_resident_b", 0)
        outs.append((y, stats))
    L.lib().dxmi_set_option(b"shift3", 1)
    ref = ref_conv(x, w, b) + rowvec[:, None, None, :] + res.float()
    assert rel_l2(outs[0][0].view(N, H, H, C), ref) < 4e-3
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    yf = outs[0][0].float().view(-1, 128, C)
    assert torch.allclose(outs[0][1][..., 0], yf.sum(1), rtol=1e-4, atol=2e-3)


@pytest.mark.parametrize("H,Cin,extra", [(32, 128, 0), (32, 256, 128), (16, 128, 0), (64, 64, 64)])
def test_pair_kernel_shift3_a_reuse(ops, H, Cin, extra):
    """Shift-3 mode of the pair kernel (BLOCK_N = 128, 3x3 stride-1, tile = full image rows): three [(bh+2) x W x 64] boxes per
    channel chunk serve the nine taps.  Must match torch and equal the nine-box form bit for bit (same accumulation order); with an
    extra 1x1 K segment (conv2 + nin_shortcut) the two stage kinds share the ring."""
    from diffusion_by_maxentirl_b200 import _lib as L

    if not ops.pair_mode:
        pytest.skip("pair kernel only")
    torch.manual_seed(21)
    dev = "cuda"
    N, Co = (160 if H <= 32 else 40), 128
    if H == 16:
        N = 640
    x = nhwc(torch.randn(N, Cin, H, H, device=dev))
    w = torch.randn(Co, Cin, 3, 3, device=dev) / (3 * Cin**0.5)
    b = torch.randn(Co, device=dev)
    srcs, segs, parts = [(x, Cin, Cin)], [(0, 9)], [(w, 0, Cin)]
    ref = ref_conv(x, w, b)
    if extra:
        x2 = nhwc(torch.randn(N, extra, H, H, device=dev))
        w2 = torch.randn(Co, extra, 1, 1, device=dev) / extra**0.5
        srcs.append((x2, extra, extra))
        segs.append((1, 1))
        parts.append((w2, 0, extra))
        ref = ref + F.conv2d(x2.float().permute(0, 3, 1, 2), w2.to(torch.bfloat16).float()).permute(0, 2, 3, 1)
    wp = ops.pack_conv_weight(None, parts=parts)
    outs = []
    for s3 in (1, 0):
        L.lib().dxmi_set_option(b"shift3", s3)
        try:
            seg = 128
            stats = torch.zeros(N * H * H // seg, Co, 2, device=dev)
            y = ops.conv_gemm(srcs, segs, wp, N, H, H, bias=b, gn_stats=stats, gn_seg=seg, block_n=128)
        finally:
            L.lib().dxmi_set_option(b"shift3", 1)
        outs.append((y, stats))
    assert rel_l2(outs[0][0].view(N, H, H, Co), ref) < 4e-3
    # every kernel variant walks K in the same (chunk, column shift, row shift) order: bitwise equal (batch invariance relies on it)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    yf = outs[0][0].float().view(-1, 128, Co)
    assert torch.allclose(outs[0][1][..., 0], yf.sum(1), rtol=1e-4, atol=2e-3)
