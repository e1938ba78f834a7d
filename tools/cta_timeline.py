"""Per-CTA phase timeline of the conv GEMM kernel (globaltimer stamps): where a CTA's life goes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusion_by_maxentirl_b200 import _lib as L  # noqa: E402
from diffusion_by_maxentirl_b200 import ops  # noqa: E402

dev = "cuda"
for name, N, H, Cin, Cout, taps, bn in [("p16_256_256_1x1", 256, 16, 256, 256, 1, 256), ("c16_256_256", 256, 16, 256, 256, 9, 256),
                                        ("c32_128_128", 256, 32, 128, 128, 9, 128), ("c4_256_256", 256, 4, 256, 256, 9, 64)]:
    x = torch.randn(N, H, H, Cin, device=dev).to(torch.bfloat16)
    k = 3 if taps == 9 else 1
    w = torch.randn(Cout, Cin, k, k, device=dev) / (k * Cin**0.5)
    b = torch.randn(Cout, device=dev)
    wp = ops.pack_conv_weight(w)
    out = torch.empty(N * H * H, Cout, dtype=torch.bfloat16, device=dev)
    n_cta = (N * H * H // 128) * ((Cout + bn - 1) // bn)
    buf = torch.zeros(n_cta, 8, dtype=torch.int64, device=dev)
    for _ in range(3):
        ops.conv_gemm([(x, Cin, Cin)], [(0, taps)], wp, N, H, H, bias=b, block_n=bn, out=out)
    L.lib().dxmi_set_debug_buffer(L.ptr(buf))
    ops.conv_gemm([(x, Cin, Cin)], [(0, taps)], wp, N, H, H, bias=b, block_n=bn, out=out)
    torch.cuda.synchronize()
    L.lib().dxmi_set_debug_buffer(None)
    t = buf.cpu().double()
    t0 = t[:, 0].min()
    d = lambda a, bb: float((t[:, a] - t[:, bb]).median())  # noqa: E731
    print(f"{name} block_n={bn} ctas={n_cta}: kernel span {float(t[:, 5].max() - t0) / 1e3:.1f} us | per-CTA median ns: "
          f"setup {d(1, 0):.0f}, first-data {d(2, 1):.0f}, mainloop {d(3, 2):.0f}, mma-drain {d(4, 3):.0f}, "
          f"epilogue {d(5, 4):.0f}, total {d(5, 0):.0f}", flush=True)
    starts = (t[:, 0] - t0).sort().values
    print("   CTA start times (us) at ranks 0,147,148,295,296,-1:", [round(float(starts[min(i, n_cta - 1)]) / 1e3, 1) for i in (0, 147, 148, 295, 296, n_cta - 1)])
