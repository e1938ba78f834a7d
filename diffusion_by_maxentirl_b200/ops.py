"""Kernel-level operators of libdxmi_b200.so on torch tensors (used by the parity tests and by profiling scripts).

Layouts: activations are NHWC bf16 (`[N, H, W, C]` contiguous, or channel slices of such a tensor); packed weights are
bf16 `[Cout, K]` with `K = sum_segments taps * C_segment` (tap-major, channel-minor inside a segment).
"""
import ctypes as C

import torch

from . import _lib as L


def pack_conv_weight(w, parts=None):
    """OIHW conv weight(s) -> bf16 [Cout, K].  `parts` = list of (weight, c_off, c_cnt) concatenated along K."""
    if parts is None:
        parts = [(w, 0, w.shape[1])]
    cout = parts[0][0].shape[0]
    K = sum(p[0].shape[2] * p[0].shape[3] * p[2] for p in parts)
    dst = torch.empty(cout, K, dtype=torch.bfloat16, device=parts[0][0].device)
    k_off = 0
    for wt, c_off, c_cnt in parts:
        wt = wt.contiguous()
        dt = L.F16 if wt.dtype == torch.float16 else L.F32
        assert wt.dtype in (torch.float16, torch.float32)
        L.check(
            L.lib().dxmi_op_pack_conv_weight(
                L.ptr(wt), dt, wt.shape[0], wt.shape[1], wt.shape[2], wt.shape[3], c_off, c_cnt, L.ptr(dst), K, k_off,
                L.stream_ptr(),
            ),
            "pack_conv_weight",
        )
        k_off += wt.shape[2] * wt.shape[3] * c_cnt
    return dst


def pack_conv_weight_up2(w):
    """OIHW 3x3 weight -> bf16 [4, Cout, 4*Cin]: the four pre-summed 2x2 phase filters of conv3x3(nearest_upsample2x(x))."""
    w = w.contiguous()
    assert w.dim() == 4 and w.shape[2] == w.shape[3] == 3 and w.dtype in (torch.float16, torch.float32)
    dst = torch.empty(4, w.shape[0], 4 * w.shape[1], dtype=torch.bfloat16, device=w.device)
    dt = L.F16 if w.dtype == torch.float16 else L.F32
    L.check(L.lib().dxmi_op_pack_conv_weight_up2(L.ptr(w), dt, w.shape[0], w.shape[1], L.ptr(dst), L.stream_ptr()), "pack_conv_weight_up2")
    return dst


def conv_up2(x, w_up2, bias=None, rowvec=None, gn_stats=None, gn_seg=32, block_n=0):
    """conv3x3(nearest_upsample2x(x)) as four phase convolutions: x NHWC bf16 [N,H,W,C] -> NHWC bf16 [N,2H,2W,Cout]."""
    N, H, W, Cin = x.shape
    Cout = w_up2.shape[1]
    out = torch.empty(N, 2 * H, 2 * W, Cout, dtype=torch.bfloat16, device=x.device)
    conv_gemm([(x, Cin, Cin)], [(0, 4)], w_up2, N, H, W, bias=bias, rowvec=rowvec, out=out, batch=4, b_batched=True,
              b_batch_stride=Cout * 4 * Cin, gn_stats=gn_stats, gn_seg=gn_seg, block_n=block_n, up2=True)
    return out


def conv_gemm(
    srcs,
    segs,
    w_packed,
    N,
    H,
    W,
    out_H=None,
    out_W=None,
    stride=1,
    bias=None,
    rowvec=None,
    rows_per_image=None,
    residual=None,
    act=0,
    out=None,
    out_fp32=False,
    block_n=0,
    batch=1,
    a_batched=False,
    b_batched=False,
    b_rows=None,
    b_ld=None,
    b_batch_stride=0,
    ldo=None,
    out_batch_stride=0,
    bias_along_m=False,
    alpha=1.0,
    softmax=False,
    res_batch_stride=0,
    gn_stats=None,
    gn_seg=32,
    gn_halo_P=0,
    gate=None,
    up2=False,
):
    """srcs: list of (tensor, C_used, ld) NHWC bf16 sources; segs: list of (src_index, taps)."""
    d = L.GemmDesc()
    for i, (t, c, ld) in enumerate(srcs):
        assert t.dtype == torch.bfloat16
        d.a_ptr[i] = t.data_ptr()
        d.a_C[i] = c
        d.a_ld[i] = ld
    d.N, d.H, d.W = N, H, W
    d.nseg = len(segs)
    for i, (s, taps) in enumerate(segs):
        d.seg_src[i] = s
        d.seg_taps[i] = taps
    d.stride = stride
    d.out_H = out_H if out_H is not None else H
    d.out_W = out_W if out_W is not None else W
    d.b_ptr = w_packed.data_ptr()
    d.b_rows = b_rows if b_rows is not None else w_packed.shape[-2]
    d.b_ld = b_ld if b_ld is not None else w_packed.shape[-1]
    d.b_batch_stride = b_batch_stride
    d.batch = batch
    d.a_batched = int(a_batched)
    d.b_batched = int(b_batched)
    rows = (d.out_H * d.out_W) if a_batched else N * d.out_H * d.out_W
    if out is None:
        shape = (batch, rows, d.b_rows) if batch > 1 else (rows, d.b_rows)
        out = torch.empty(shape, dtype=torch.float32 if out_fp32 else torch.bfloat16, device=w_packed.device)
        if batch > 1:
            out_batch_stride = rows * d.b_rows
    d.out = out.data_ptr()
    d.ldo = ldo if ldo is not None else d.b_rows
    d.out_batch_stride = out_batch_stride
    d.out_fp32 = int(out_fp32)
    if bias is not None:
        assert bias.dtype == torch.float32
        d.bias = bias.data_ptr()
    d.bias_along_m = int(bias_along_m)
    if rowvec is not None:
        assert rowvec.dtype == torch.float32
        d.rowvec = rowvec.data_ptr()
        d.ldrv = rowvec.stride(0)
    d.rows_per_image = rows_per_image if rows_per_image is not None else d.out_H * d.out_W
    if residual is not None:
        assert residual.dtype == torch.bfloat16
        d.residual = residual.data_ptr()
        d.ldr = residual.shape[-1]
        d.res_batch_stride = res_batch_stride
    d.act = act
    d.alpha = alpha
    d.softmax = int(softmax)
    d.block_n = block_n
    d.up2 = int(up2)
    if gn_stats is not None:
        assert gn_stats.dtype == torch.float32
        d.gn_stats = gn_stats.data_ptr()
        d.gn_seg = gn_seg
        d.gn_halo_P = gn_halo_P
    if gate is not None:
        assert gate.dtype == torch.bfloat16
        d.gate = gate.data_ptr()
        d.ldg = gate.shape[-1]
    L.check(L.lib().dxmi_op_conv_gemm(C.byref(d), L.stream_ptr()), "conv_gemm")
    return out


def pack_conv_weight_dgrad(parts):
    """Data-gradient packing of OIHW conv weights: bf16 [Cin, K] with K = sum over parts of taps * Cout, taps flipped and
    O/I transposed, so that dX = conv_gemm(dY sources, these rows).  `parts`: list of OIHW weights sharing Cin."""
    cin = parts[0].shape[1]
    K = sum(w.shape[2] * w.shape[3] * w.shape[0] for w in parts)
    dst = torch.empty(cin, K, dtype=torch.bfloat16, device=parts[0].device)
    k_off = 0
    for w in parts:
        w = w.contiguous()
        assert w.shape[1] == cin and w.dtype in (torch.float16, torch.float32)
        taps = w.shape[2] * w.shape[3]
        dt = L.F16 if w.dtype == torch.float16 else L.F32
        L.check(L.lib().dxmi_op_pack_conv_weight_dgrad(L.ptr(w), dt, w.shape[0], cin, taps, L.ptr(dst), K, k_off, L.stream_ptr()),
                "pack_conv_weight_dgrad")
        k_off += taps * w.shape[0]
    return dst


def conv_wgrad(dy, x, k, scale=1.0, grad=None, ci_off=0):
    """Weight gradient of a k x k (k = 3: pad 1, stride 1; k = 1) convolution: dy NHWC bf16 [N,H,W,Cout], x NHWC bf16
    [N,H,W,Cin] -> fp32 OIHW [Cout, Cin_total, k, k] (input channels ci_off .. ci_off + Cin of `grad` when given)."""
    N, H, W, Cout = dy.shape
    Cin = x.shape[-1]
    assert dy.dtype == torch.bfloat16 and x.dtype == torch.bfloat16 and dy.is_contiguous() and x.is_contiguous()
    if grad is None:
        grad = torch.empty(Cout, Cin, k, k, dtype=torch.float32, device=dy.device)
    n_ws = L.lib().dxmi_op_wgrad_ws_floats(N, H, W, Cout, Cin, k * k)
    if n_ws < 0:
        raise RuntimeError(L.lib().dxmi_last_error().decode())
    ws = torch.empty(n_ws, dtype=torch.float32, device=dy.device)
    L.check(L.lib().dxmi_op_conv_wgrad(L.ptr(dy), L.ptr(x), N, H, W, Cout, Cin, k * k, L.ptr(grad), grad.shape[1], ci_off,
                                        float(scale), L.ptr(ws), L.stream_ptr()), "conv_wgrad")
    return grad


def group_norm(x1, gamma, beta, eps, silu, x2=None, film=None, groups=32):
    """GroupNorm(+SiLU)(+FiLM) over the channel concat of NHWC bf16 x1 (and x2) -> NHWC bf16."""
    N, H, W, C1 = x1.shape
    C2 = x2.shape[-1] if x2 is not None else 0
    out = torch.empty(N, H, W, C1 + C2, dtype=torch.bfloat16, device=x1.device)
    ws = torch.empty(L.lib().dxmi_op_gn_ws_floats(N, H * W, groups), dtype=torch.float32, device=x1.device)
    L.check(
        L.lib().dxmi_op_group_norm(
            L.ptr(x1), C1, C1, L.ptr(x2), C2, C2, N, H * W, groups, eps, L.ptr(gamma), L.ptr(beta), L.ptr(film),
            film.stride(0) if film is not None else 0, int(silu), L.ptr(ws), L.ptr(out), L.stream_ptr(),
        ),
        "group_norm",
    )
    return out


def attention(qk, vt, heads, scale, q_col0=0, k_col0=None):
    """Fused d=64 attention: qk bf16 [B, seq, 2C] (q | k, head-major), vt bf16 [B, C, seq] -> bf16 [B, seq, C]."""
    B, seq, ld = qk.shape
    Cc = heads * 64
    assert vt.shape == (B, Cc, seq) and qk.dtype == vt.dtype == torch.bfloat16
    out = torch.empty(B, seq, Cc, dtype=torch.bfloat16, device=qk.device)
    L.check(L.lib().dxmi_op_attention(L.ptr(qk), ld, q_col0, Cc if k_col0 is None else k_col0, L.ptr(vt), L.ptr(out), Cc, B,
                                      heads, seq, scale, L.stream_ptr()), "attention")
    return out


def group_norm_bwd(x1, dy, ab, mr, silu, x2=None, groups=32, want_param_grads=True):
    """GroupNorm(32)(+SiLU) backward.  x1 (| x2): NHWC bf16, dy: [N, HW, C] bf16, ab: [N, C, 2] fp32 forward affine, mr: [N, 32, 2]
    fp32 (mean, rstd).  Returns (dx [N, HW, C] bf16, dgamma [C], dbeta [C])."""
    N = x1.shape[0]
    C1 = x1.shape[-1]
    C2 = x2.shape[-1] if x2 is not None else 0
    C = C1 + C2
    HW = x1.numel() // (N * C1)
    assert dy.dtype == torch.bfloat16 and x1.dtype == torch.bfloat16 and ab.dtype == torch.float32 and mr.dtype == torch.float32
    ws = torch.empty(L.lib().dxmi_op_gn_bwd_ws_floats(N, HW, C), dtype=torch.float32, device=x1.device)
    dx = torch.empty(N, HW, C, dtype=torch.bfloat16, device=x1.device)
    dg = torch.empty(C, dtype=torch.float32, device=x1.device) if want_param_grads else None
    db = torch.empty(C, dtype=torch.float32, device=x1.device) if want_param_grads else None
    L.check(L.lib().dxmi_op_group_norm_bwd(L.ptr(x1), C1, L.ptr(x2) if x2 is not None else None, C2, L.ptr(dy), L.ptr(ab), L.ptr(mr), N,
                                            HW, groups, int(silu), L.ptr(ws), L.ptr(dx), L.ptr(dg) if dg is not None else None,
                                            L.ptr(db) if db is not None else None, L.stream_ptr()), "group_norm_bwd")
    return dx, dg, db
