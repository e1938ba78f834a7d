"""Per-CTA phase timeline of the persistent conv GEMM kernel (globaltimer stamps)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusion_by_maxentirl_b200 import _lib as L  # noqa: E402
from diffusion_by_maxentirl_b200 import ops  # noqa: E402

dev = "cuda"
import os as _os
L.lib().dxmi_set_option(b"dbg_mode", int(_os.environ.get("DBG_MODE", "0")))
for name, N, H, Cin, Cout, taps, bn, res, stats in [
        ("p16_256_256_1x1", 256, 16, 256, 256, 1, 256, False, False), ("p16_256_256_1x1+res+stats", 256, 16, 256, 256, 1, 256, True, True),
        ("p16_1x1+res", 256, 16, 256, 256, 1, 256, True, False), ("p16_1x1+stats", 256, 16, 256, 256, 1, 256, False, True),
        ("c16_256_256", 256, 16, 256, 256, 9, 256, False, False), ("c16_256_256+res+stats", 256, 16, 256, 256, 9, 256, True, True),
        ("c32_128_128", 256, 32, 128, 128, 9, 128, False, False), ("c4_256_256", 256, 4, 256, 256, 9, 256, False, False)]:
    x = torch.randn(N, H, H, Cin, device=dev).to(torch.bfloat16)
    k = 3 if taps == 9 else 1
    w = torch.randn(Cout, Cin, k, k, device=dev) / (k * Cin**0.5)
    b = torch.randn(Cout, device=dev)
    wp = ops.pack_conv_weight(w)
    M = N * H * H
    out = torch.empty(M, Cout, dtype=torch.bfloat16, device=dev)
    r = torch.randn(M, Cout, device=dev).to(torch.bfloat16) if res else None
    seg = 128 if (H * H) % 128 == 0 else (64 if (H * H) % 64 == 0 else 32)
    st = torch.empty(M // seg, Cout, 2, device=dev) if stats else None
    kw = dict(bias=b, block_n=bn, out=out, residual=r, gn_stats=st, gn_seg=seg)
    buf = torch.zeros(148 * 8 + 148 * 16, dtype=torch.int64, device=dev)
    for _ in range(3):
        ops.conv_gemm([(x, Cin, Cin)], [(0, taps)], wp, N, H, H, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.conv_gemm([(x, Cin, Cin)], [(0, taps)], wp, N, H, H, **kw)
    e1.record()
    torch.cuda.synchronize()
    L.lib().dxmi_set_debug_buffer(L.ptr(buf))
    ops.conv_gemm([(x, Cin, Cin)], [(0, taps)], wp, N, H, H, **kw)
    torch.cuda.synchronize()
    L.lib().dxmi_set_debug_buffer(None)
    allb = buf.cpu()
    cyc = allb[148 * 8:].view(148, 16).double()
    t = allb[:148 * 8].view(148, 8).double()
    cyc = cyc[t[:, 7] > 0]
    t = t[t[:, 7] > 0]
    d = lambda a, bb: float((t[:, a] - t[:, bb]).median())  # noqa: E731
    tiles = (M // 128) * ((Cout + bn - 1) // bn)
    print(f"{name}: {e0.elapsed_time(e1) * 100:.1f} us/launch, {tiles} tiles on {len(t)} CTAs | median ns: setup {d(1, 0):.0f}, "
          f"first-data {d(2, 1):.0f}, mainloop0 {d(3, 2):.0f}, mainloop1 {d(4, 3):.0f}, epi0 wait {d(5, 3):.0f}, "
          f"epi0 {d(6, 5):.0f}, total {d(7, 0):.0f}", flush=True)
