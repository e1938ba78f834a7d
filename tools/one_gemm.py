"""One GEMM shape, a few launches (ncu target)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusion_by_maxentirl_b200 import ops  # noqa: E402

N, H, Cin, Cout, taps = 256, 16, 256, 256, int(os.environ.get("TAPS", "9"))
res_on = os.environ.get("RES", "1") == "1"
dev = "cuda"
x = torch.randn(N, H, H, Cin, device=dev).to(torch.bfloat16)
k = 3 if taps == 9 else 1
w = torch.randn(Cout, Cin, k, k, device=dev) / (k * Cin**0.5)
b = torch.randn(Cout, device=dev)
wp = ops.pack_conv_weight(w)
M = N * H * H
out = torch.empty(M, Cout, dtype=torch.bfloat16, device=dev)
r = torch.randn(M, Cout, device=dev).to(torch.bfloat16) if res_on else None
st = torch.empty(M // 128, Cout, 2, device=dev) if res_on else None
for _ in range(4):
    ops.conv_gemm([(x, Cin, Cin)], [(0, taps)], wp, N, H, H, bias=b, block_n=256, out=out, residual=r, gn_stats=st, gn_seg=128)
torch.cuda.synchronize()
