"""Cycle budget of the pair kernel's epilogue warps on the 32x32 128->128 3x3 conv (B=256).  Needs a library built with
`make -C diffusion_by_maxentirl_b200/csrc clean all EXTRA=-DDXMI_EPI_PROFILE` (clock64 stamps inside gemm_epi.cuh; never shipped)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusion_by_maxentirl_b200 import _lib as L  # noqa: E402
from diffusion_by_maxentirl_b200 import ops  # noqa: E402

dev = "cuda"
lib = L.lib()
N, H, C = 256, 32, 128
x = torch.randn(N, H, H, C, device=dev).to(torch.bfloat16)
w = torch.randn(C, C, 3, 3, device=dev) / (3 * C**0.5)
b = torch.randn(C, device=dev)
wp = ops.pack_conv_weight(w)
M = N * H * H
out = torch.empty(M, C, dtype=torch.bfloat16, device=dev)
res = torch.randn(M, C, device=dev).to(torch.bfloat16)
rv = torch.randn(N, C, device=dev)
st = torch.empty(M // 128, C, 2, device=dev)
names = ["acc wait", "stage 0 + bar", "stage next", "finish+store", "prefetch+fold", "barrier 2", "publish", "barrier 1"]


def run(tag, **kw):
    buf = torch.zeros(148 * 8 + 148 * 16, dtype=torch.int64, device=dev)
    for _ in range(3):
        ops.conv_gemm([(x, C, C)], [(0, 9)], wp, N, H, H, bias=b, block_n=128, out=out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.conv_gemm([(x, C, C)], [(0, 9)], wp, N, H, H, bias=b, block_n=128, out=out, **kw)
    e1.record()
    torch.cuda.synchronize()
    lib.dxmi_set_debug_buffer(L.ptr(buf))
    ops.conv_gemm([(x, C, C)], [(0, 9)], wp, N, H, H, bias=b, block_n=128, out=out, **kw)
    torch.cuda.synchronize()
    lib.dxmi_set_debug_buffer(None)
    t = buf.cpu()[:148 * 8].view(148, 8).double()
    tiles = 2048 / 148
    med = t.median(0).values / tiles
    print(f"{tag}: {e0.elapsed_time(e1) * 100:.1f} us/launch; cycles per 128x128 tile (median CTA, {tiles:.1f} tiles/CTA): "
          + ", ".join(f"{n} {v:.0f}" for n, v in zip(names, med.tolist())) + f" | sum {med.sum():.0f}", flush=True)


for m2 in (0, 1):
    lib.dxmi_set_option(b"s3_m2", m2)
    print(f"s3_m2 = {m2}")
    run("bias only")
    run("bias + stats", gn_stats=st, gn_seg=128)
    run("rowvec + stats", rowvec=rv, gn_stats=st, gn_seg=128)
    run("residual + stats", residual=res, gn_stats=st, gn_seg=128)
