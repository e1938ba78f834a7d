"""Phase timeline of the persistent GEMM on the small-map shapes (8x8 / 4x4 at batch 256)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusion_by_maxentirl_b200 import _lib as L  # noqa: E402
from diffusion_by_maxentirl_b200 import ops  # noqa: E402

dev = "cuda"
L.lib().dxmi_set_option(b"pair", 0)
for name, N, H, Cin, Cout, taps, bn, res, stats in [
        ("c4_256_256 bn64 +res+stats", 256, 4, 256, 256, 9, 64, True, True), ("c4 bn128 +res+stats", 256, 4, 256, 256, 9, 128, True, True),
        ("c4 bn64 plain", 256, 4, 256, 256, 9, 64, False, False),
        ("c8 bn256 +res+stats", 256, 8, 256, 256, 9, 256, True, True), ("c8 bn128 +res+stats", 256, 8, 256, 256, 9, 128, True, True),
        ("c8 bn128 plain", 256, 8, 256, 256, 9, 128, False, False)]:
    x = torch.randn(N, H, H, Cin, device=dev).to(torch.bfloat16)
    w = torch.randn(Cout, Cin, 3, 3, device=dev) / (3 * Cin**0.5)
    b = torch.randn(Cout, device=dev)
    wp = ops.pack_conv_weight(w)
    M = N * H * H
    out = torch.empty(M, Cout, dtype=torch.bfloat16, device=dev)
    r = torch.randn(M, Cout, device=dev).to(torch.bfloat16) if res else None
    seg = 128 if (H * H) % 128 == 0 else (64 if (H * H) % 64 == 0 else (32 if (H * H) % 32 == 0 else 16))
    st = torch.empty(M // seg, Cout, 2, device=dev) if stats else None
    kw = dict(bias=b, block_n=bn, out=out, residual=r, gn_stats=st, gn_seg=seg)
    buf = torch.zeros(148 * 8 + 148 * 16, dtype=torch.int64, device=dev)
    for _ in range(3):
        ops.conv_gemm([(x, Cin, Cin)], [(0, taps)], wp, N, H, H, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.conv_gemm([(x, Cin, Cin)], [(0, taps)], wp, N, H, H, **kw)
    e1.record()
    torch.cuda.synchronize()
    L.lib().dxmi_set_debug_buffer(L.ptr(buf))
    ops.conv_gemm([(x, Cin, Cin)], [(0, taps)], wp, N, H, H, **kw)
    torch.cuda.synchronize()
    L.lib().dxmi_set_debug_buffer(None)
    t = buf.cpu()[:148 * 8].view(148, 8).double()
    t = t[t[:, 7] > 0]
    d = lambda a, bb: float((t[:, a] - t[:, bb]).median())  # noqa: E731
    span = float(t[:, 7].max() - t[:, 1].min())
    tiles = (M // 128) * ((Cout + bn - 1) // bn)
    print(f"{name}: {e0.elapsed_time(e1) * 50:.1f} us/launch (back-to-back), {tiles} tiles on {len(t)} CTAs, kernel span {span:.0f} ns | median ns: "
          f"first-data {d(2, 1):.0f}, mainloop0 {d(3, 2):.0f}, mainloop1 {d(4, 3):.0f}, epi0 {d(6, 5):.0f}, total {d(7, 1):.0f}", flush=True)
