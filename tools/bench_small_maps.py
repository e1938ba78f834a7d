"""Small-map 3x3 convs (B=256: 8x8 -> M=16384, 4x4 -> M=4096; 256 -> 256 and 512 -> 256 channels) at every column-tile width."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusion_by_maxentirl_b200 import _lib as L  # noqa: E402
from diffusion_by_maxentirl_b200 import ops  # noqa: E402

dev = "cuda"
N = 256
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for H, Cin in [(8, 256), (8, 512), (4, 256), (4, 512), (16, 256)]:
    Cout = 256
    x = torch.randn(N, H, H, Cin, device=dev).to(torch.bfloat16)
    w = torch.randn(Cout, Cin, 3, 3, device=dev) / (3 * Cin**0.5)
    b = torch.randn(Cout, device=dev)
    wp = ops.pack_conv_weight(w)
    M = N * H * H
    out = torch.empty(M, Cout, dtype=torch.bfloat16, device=dev)
    res = torch.randn(M, Cout, device=dev).to(torch.bfloat16)
    seg = 128 if (H * H) % 128 == 0 else (64 if (H * H) % 64 == 0 else (32 if (H * H) % 32 == 0 else 16))
    st = torch.empty(M // seg, Cout, 2, device=dev)
    flops = 2.0 * M * Cout * 9 * Cin
    line = f"{H}x{H} {Cin}->{Cout} (M={M}):"
    for bn in (0, 64, 128, 256):
        kw = dict(bias=b, block_n=bn, out=out, residual=res, gn_stats=st, gn_seg=seg)
        for _ in range(3):
            ops.conv_gemm([(x, Cin, Cin)], [(0, 9)], wp, N, H, H, **kw)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()  # cold L2, like inside a rollout where 100+ MB of activations pass between two uses of a weight
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv_gemm([(x, Cin, Cin)], [(0, 9)], wp, N, H, H, **kw)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        line += f"  bn={bn or 'auto'}: {us:6.1f} us ({flops / us / 1e6:5.0f} TF/s)"
    print(line, flush=True)
