"""Microbenchmark of the short-K 1x1 projection GEMMs of the ImageNet-64 attention blocks (proj_out: M = 65536, N = K = 384, residual +
GroupNorm partials; qkv: N = 1152) by epilogue shape and tile width.  Usage (under gpurun): python tools/bench_1x1.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusion_by_maxentirl_b200 import _lib as L, ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


for (B, H, K, N) in [(64, 32, 384, 384), (64, 32, 384, 1152), (64, 16, 576, 576), (64, 8, 768, 768)]:
    M = B * H * H
    x = torch.randn(B, H, H, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) / K**0.5).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    res = torch.randn(B, H, H, N, device=dev).to(torch.bfloat16)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    seg = 128 if (H * H) % 128 == 0 else 64
    stats = torch.zeros(M // seg, N, 2, device=dev)
    for pair in (1, 2):
        L.lib().dxmi_set_option(b"pair", pair)
        for bn in (0, 128, 192, 256):
            if N % (bn or 1):
                continue
            row = []
            for name, kw in [("bias", {}), ("+stats", dict(gn_stats=stats, gn_seg=seg)), ("+res", dict(residual=res)),
                             ("+res+stats", dict(residual=res, gn_stats=stats, gn_seg=seg))]:
                us = timeit(lambda: ops.conv_gemm([(x, K, K)], [(0, 1)], w, B, H, H, bias=b, out=out, block_n=bn, **kw))
                row.append(f"{name} {us:6.1f}")
            print(f"M={M} N={N} K={K} pair_opt={pair} bn={bn}: " + " | ".join(row) + f"   (2MNK = {2.0 * M * N * K / 1e9:.1f} GFLOP)")
    L.lib().dxmi_set_option(b"pair", 1)
